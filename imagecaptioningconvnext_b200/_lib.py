"""ctypes binding of libccx.so (the C ABI declared in include/ccx.h).

There is NO fallback: if the shared library is missing or a kernel rejects a shape, this raises.  The library is
built in-tree by ``make`` / ``__graft_entry__.build()`` for sm_100a only.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libccx.so")

CCX_F32, CCX_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU, ACT_GELU_GRAD = 0, 1, 2, 3
MAX_BLOCKS = 64

_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t


class LinearDesc(C.Structure):
    _fields_ = [("A", _vp), ("A_lo", _vp), ("W", _vp), ("W_lo", _vp), ("C", _vp), ("C_lo", _vp),
                ("bias", _vp), ("colscale", _vp), ("rowscale", _vp), ("residual", _vp), ("emask", _vp),
                ("lda", _i64), ("ldw", _i64), ("ldc", _i64), ("ldr", _i64), ("ldm", _i64),
                ("M", _i32), ("N", _i32), ("K", _i32), ("rows_per_group", _i32),
                ("act", _i32), ("in_dtype", _i32), ("out_dtype", _i32), ("split", _i32),
                ("a_mn", _i32), ("w_mn", _i32), ("res_mul", _i32)]


class CNBlockWeights(C.Structure):
    _fields_ = [("dw_w", _vp), ("dw_b", _vp), ("ln_g", _vp), ("ln_b", _vp),
                ("w1", _vp), ("w1_lo", _vp), ("b1", _vp),
                ("w2", _vp), ("w2_lo", _vp), ("b2", _vp), ("layer_scale", _vp)]


class DownsampleWeights(C.Structure):
    _fields_ = [("ln_g", _vp), ("ln_b", _vp), ("w", _vp), ("w_lo", _vp), ("b", _vp)]


class EncoderWeights(C.Structure):
    _fields_ = [("stem_w", _vp), ("stem_b", _vp), ("stem_ln_g", _vp), ("stem_ln_b", _vp),
                ("blocks", CNBlockWeights * MAX_BLOCKS), ("down", DownsampleWeights * 3),
                ("depths", _i32 * 4), ("dims", _i32 * 4), ("compute_dtype", _i32)]


class LstmTF(C.Structure):
    _fields_ = [("XH_hi", _vp), ("XH_lo", _vp), ("C_all", _vp), ("HG", _vp), ("G", _vp), ("alphas", _vp),
                ("H_all_hi", _vp), ("H_all_lo", _vp), ("dropmask", _vp), ("att1", _vp), ("enc", _vp),
                ("w_h", _vp), ("w_h_lo", _vp), ("b_h", _vp), ("w_f", _vp), ("b_f", _vp),
                ("w_lstm", _vp), ("w_lstm_lo", _vp), ("b_lstm", _vp), ("bts_host", C.POINTER(_i32)),
                ("B", _i32), ("T", _i32), ("P", _i32), ("E", _i32), ("A", _i32), ("D", _i32), ("Emb", _i32),
                ("compute_dtype", _i32)]


class LstmTFBwd(C.Structure):
    _fields_ = [("dH_all", _vp), ("dalphas", _vp), ("dG_all", _vp), ("dHG_all", _vp), ("dXH_all", _vp),
                ("dh", _vp), ("dc", _vp), ("d_att1", _vp), ("d_enc", _vp), ("d_wf", _vp),
                ("w_lstm_t", _vp), ("w_lstm_t_lo", _vp), ("w_h_t", _vp), ("w_h_t_lo", _vp),
                ("scratch_hi", _vp), ("scratch_lo", _vp), ("scratch2_hi", _vp), ("scratch2_lo", _vp),
                ("dawe_all", _vp), ("dalpha_all", _vp), ("de_all", _vp)]


class LstmPersist(C.Structure):
    _fields_ = [("E_all", _vp), ("w2p", _vp), ("att1_bf", _vp), ("enc_bf", _vp), ("decode_len", _vp),
                ("counters", _vp), ("scratch", _vp), ("awe_all", _vp), ("dbg", _vp)]


class LstmPersistBwd(C.Structure):
    _fields_ = [("wx", _vp), ("wht", _vp), ("dG_bf", _vp), ("scratch", _vp), ("Xp", _vp), ("awe_all", _vp),
                ("att1_bf", _vp), ("enc_bf", _vp), ("decode_len", _vp), ("counters", _vp), ("dbg", _vp)]


class CastSeg(C.Structure):
    """ccx_cast_seg (include/ccx.h): one rectangular piece of a weight refresh."""
    _fields_ = [("src", _vp), ("src2", _vp), ("dst", _vp), ("row_map", _vp), ("src_ld", _i64), ("dst_ld", _i64),
                ("rows", _i32), ("cols", _i32), ("flags", _i32), ("tile0", _i32)]


# name -> (restype, argtypes); must list every symbol include/ccx.h declares (tests/test_abi.py checks it)
SIGNATURES = {
    "ccx_version": (C.c_int, []),
    "ccx_status_string": (C.c_char_p, [C.c_int]),
    "ccx_num_sms": (C.c_int, []),
    "ccx_linear": (C.c_int, [C.POINTER(LinearDesc), _vp]),
    "ccx_set_gemm_pair_mode": (C.c_int, [_i32]),
    "ccx_set_sm_limit": (C.c_int, [_i32]),
    "ccx_split_tf32": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "ccx_cast_bf16": (C.c_int, [_vp, _vp, _i64, _vp]),
    "ccx_stem_ln": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp]),
    "ccx_stem_ln_u8": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp]),
    "ccx_dwconv7_ln": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp]),
    "ccx_ln_rows": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _i32, _i32, _i32, _i32, _vp]),
    "ccx_embed_rows": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _i64,
                                 _i64, _i32, _i32, _vp]),
    "ccx_mean_pixels": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _i32, _i64, _vp]),
    "ccx_bahdanau_attention": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _i64, _i32,
                                         _i32, _i32, _i32, _i32, _i32, _vp]),
    "ccx_lstm_pointwise": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i32, _vp, _i64, _vp,
                                     _i64, _i32, _i32, _vp]),
    "ccx_greedy_next": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _vp]),
    "ccx_mha_small": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _i32, _i64, _i64, _vp,
                                _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _vp]),
    "ccx_avgpool_nhwc": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "ccx_encoder_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ccx_encoder_run": (C.c_int, [C.POINTER(EncoderWeights), _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp,
                                  _sz, _vp]),
    "ccx_attn_head_mean": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _f32,
                                     _i32, _vp]),
    "ccx_mha_decode": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _i32, _i64, _vp, _i64, _i32,
                                 _i32, _i32, _i32, _i32, _f32, _vp]),
    "ccx_beam_topk": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "ccx_beam_update": (C.c_int, [_i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp, _vp, _i64, _vp, _vp]),
    "ccx_gather_rows": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp]),
    "ccx_convert_operand": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _i64, _i32, _f32, _vp, _vp, _i32, _i64, _i32, _i32,
                                      _i32, _i32, _vp]),
    "ccx_colsum_acc": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _f32, _vp, _i32, _i32, _vp]),
    "ccx_cast_segments": (C.c_int, [_vp, _i32, _i32, C.c_double, _vp]),
    "ccx_stream_capture_status": (C.c_int, [_vp]),
    "ccx_convert_colsum": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _f32, _vp, _i64, _vp, _i32, _i32, _vp]),
    "ccx_ln_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _i32, _i32, _i32, _vp]),
    "ccx_mha_bwd": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp,
                              _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32,
                              _f32, _vp]),
    "ccx_mha_bwd_tc": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp,
                              _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32,
                              _f32, _vp]),
    "ccx_softmax_ce": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _f32, _vp, _vp, _i64, _vp, _i32, _vp]),
    "ccx_softmax_ce_dev": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _i32, _vp]),
    "ccx_free_running_targets": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _i64, _i64, _vp]),
    "ccx_embedding_bwd": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ccx_lstm_pointwise_bwd": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32,
                                         _vp]),
    "ccx_bahdanau_attention_bwd": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64,
                                             _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ccx_bcast_add_rows": (C.c_int, [_vp, _vp, _f32, _i32, _i32, _i32, _vp]),
    "ccx_dwconv7_plain": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ccx_scale_rows_cols": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _i64, _i32, _vp]),
    "ccx_gelu_bwd": (C.c_int, [_vp, _vp, _i64, _vp]),
    "ccx_cnblock_param_grads": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "ccx_dwconv7_wgrad": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ccx_avgpool_nhwc_bwd": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "ccx_lstm_tf_forward": (C.c_int, [C.POINTER(LstmTF), _vp]),
    "ccx_lstm_tf_backward": (C.c_int, [C.POINTER(LstmTF), C.POINTER(LstmTFBwd), _vp]),
    "ccx_lstm_persist_supported": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "ccx_lstm_tf_forward_persist": (C.c_int, [C.POINTER(LstmTF), C.POINTER(LstmPersist), _vp]),
    "ccx_lstm_tf_backward_persist": (C.c_int, [C.POINTER(LstmTF), C.POINTER(LstmTFBwd), C.POINTER(LstmPersistBwd),
                                                _vp]),
    "ccx_adam_clamp": (C.c_int, [_vp, _vp, _vp, _i32, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _i32, C.c_double,
                                 _vp]),
    "ccx_adam_clamp_dev": (C.c_int, [_vp, _vp, _vp, _i32, _f32, _f32, _f32, _f32, _vp, _f32, _i32, C.c_double, _vp]),
    "ccx_prof_begin": (C.c_int, []),
    "ccx_prof_spans": (C.c_int, [C.POINTER(_i32), C.POINTER(C.c_double), C.POINTER(C.c_double), _i32]),
    "ccx_prof_end": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64), _i32]),
}
PROF_KINDS = ("gemm", "dwconv_ln", "stem", "ln_rows", "pool", "elementwise", "attention", "lstm", "loss", "optimizer",
              "gemm_skinny")

_lib = None


def lib():
    """The loaded library.  Raises if it has not been built — there is no Python/CPU substitute."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()). "
                "imagecaptioningconvnext_b200 has no CPU / eager fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


_DEBUG_CAPTURE = os.environ.get("CCX_DEBUG_CAPTURE", "0") != "0"
_capture_seen = [0]


def capture_probe(label):
    """CCX_DEBUG_CAPTURE=1: report (stderr) the first point at which the current stream's capture is found invalidated —
    the call just before `label` broke it (an API that is illegal while a stream captures)."""
    if not _DEBUG_CAPTURE:
        return
    import sys
    st = lib().ccx_stream_capture_status(stream_ptr())
    if st != _capture_seen[0]:
        print(f"[ccx capture] status {_capture_seen[0]} -> {st} at: {label}", file=sys.stderr, flush=True)
        _capture_seen[0] = st


def check(rc, what=""):
    if _DEBUG_CAPTURE:
        capture_probe("after " + what)
    if rc != 0:
        msg = lib().ccx_status_string(rc).decode()
        raise RuntimeError(f"libccx {what} failed: {msg} (status {rc})")


def prof_begin():
    check(lib().ccx_prof_begin(), "prof_begin")


def prof_spans(max_spans=100000):
    """Per-launch [(kind, ms, work)] since prof_begin (synchronises); call before prof_end."""
    kind, ms, work = (_i32 * max_spans)(), (C.c_double * max_spans)(), (C.c_double * max_spans)()
    n = lib().ccx_prof_spans(kind, ms, work, max_spans)
    if n < 0:
        check(n, "prof_spans")
    return [(PROF_KINDS[kind[i]], ms[i], work[i]) for i in range(min(n, max_spans))]


def prof_end():
    """-> {kind: {"ms": total ms, "work": FLOPs (gemm) or bytes, "launches": n}}; synchronises the device."""
    n = len(PROF_KINDS)
    ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (_i64 * n)()
    check(lib().ccx_prof_end(ms, work, cnt, n), "prof_end")
    return {k: {"ms": ms[i], "work": work[i], "launches": int(cnt[i])} for i, k in enumerate(PROF_KINDS)}


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (the raw getter: torch.cuda.current_stream()
    costs ~15 us of Python per call, and a train step asks ~130 times)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def dt_code(dtype):
    if dtype == torch.float32:
        return CCX_F32
    if dtype == torch.bfloat16:
        return CCX_BF16
    raise ValueError(f"unsupported compute dtype {dtype}; libccx computes in float32 (3xTF32) or bfloat16")


def require_cuda(t, name):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor: imagecaptioningconvnext_b200 has no CPU path")


# ------------------------------------------------------------------------------------------------
# thin op wrappers (allocate outputs with torch, launch on the current stream)
# ------------------------------------------------------------------------------------------------
def split_tf32(x):
    """fp32 tensor -> (hi, lo) fp32 tensors, hi exactly tf32-representable."""
    x = x.contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    check(lib().ccx_split_tf32(ptr(x), ptr(hi), ptr(lo), x.numel(), stream_ptr()), "split_tf32")
    return hi, lo


def cast_bf16(x):
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(lib().ccx_cast_bf16(ptr(x), ptr(y), x.numel(), stream_ptr()), "cast_bf16")
    return y


_PREPARE_LOG = [None]     # {"made": [Operand, ...], "reuse": [Operand, ...] or None} while a PreparedCache prepares


class prepare_log:
    """Context around ``owner._prepare()``: records the Operands ``Operand.prepare`` hands out, in call order, and —
    given the record of the previous preparation — makes it refresh those in place instead of allocating."""

    def __init__(self, reuse=None):
        self.state = {"made": [], "reuse": list(reuse) if reuse else None}

    def __enter__(self):
        self._old = _PREPARE_LOG[0]
        _PREPARE_LOG[0] = self.state
        return self.state

    def __exit__(self, *exc):
        _PREPARE_LOG[0] = self._old


class Operand:
    """A GEMM operand prepared for the chosen compute dtype: bf16 tensor, or (hi, lo) fp32 pair."""
    __slots__ = ("hi", "lo", "dtype")

    def __init__(self, hi, lo, dtype):
        self.hi, self.lo, self.dtype = hi, lo, dtype

    def map(self, fn):
        """Apply a view-producing function (slice / reshape) to both halves."""
        return Operand(fn(self.hi), None if self.lo is None else fn(self.lo), self.dtype)

    @property
    def lo_ptr(self):
        return None if self.lo is None else self.lo.data_ptr()

    @staticmethod
    def zeros(shape, compute_dtype, device):
        op = Operand.empty(shape, compute_dtype, device)
        op.hi.zero_()
        if op.lo is not None:
            op.lo.zero_()
        return op

    def refresh(self, x_fp32):
        """Re-convert a new fp32 value INTO this operand's buffers (same pointers: weight tables and captured CUDA
        graphs that reference them stay valid)."""
        x = x_fp32.contiguous()
        if x.shape != self.hi.shape:
            raise ValueError("Operand.refresh: shape changed")
        if self.dtype == torch.bfloat16:
            check(lib().ccx_cast_bf16(ptr(x), ptr(self.hi), x.numel(), stream_ptr()), "cast_bf16")
        else:
            check(lib().ccx_split_tf32(ptr(x), ptr(self.hi), ptr(self.lo), x.numel(), stream_ptr()), "split_tf32")
        return self

    @staticmethod
    def prepare(x_fp32, compute_dtype):
        if _PREPARE_LOG[0] is not None:
            # PreparedCache (see _host.py): a weight refresh converts straight INTO the operand the previous preparation
            # created at this position of the call sequence (no second buffer, no device-to-device copy) ...
            reuse = _PREPARE_LOG[0].get("reuse")
            if reuse:
                old = reuse.pop(0)
                if old.dtype == (torch.bfloat16 if compute_dtype == torch.bfloat16 else torch.float32) and \
                        tuple(old.hi.shape) == tuple(x_fp32.shape) and old.hi.is_contiguous():
                    _PREPARE_LOG[0]["made"].append(old)
                    return old.refresh(x_fp32)
                reuse.clear()                      # the sequence changed: allocate from here on
        if compute_dtype == torch.bfloat16:
            op = Operand(cast_bf16(x_fp32), None, torch.bfloat16)
        else:
            hi, lo = split_tf32(x_fp32)
            op = Operand(hi, lo, torch.float32)
        if _PREPARE_LOG[0] is not None:           # ... and a first preparation records that sequence
            _PREPARE_LOG[0]["made"].append(op)
        return op

    @staticmethod
    def empty(shape, compute_dtype, device):
        if compute_dtype == torch.bfloat16:
            return Operand(torch.empty(shape, dtype=torch.bfloat16, device=device), None, torch.bfloat16)
        return Operand(torch.empty(shape, dtype=torch.float32, device=device),
                       torch.empty(shape, dtype=torch.float32, device=device), torch.float32)


def linear(a, w, bias=None, act=ACT_NONE, colscale=None, rowscale=None, rows_per_group=1, residual=None,
           out=None, out_dtype=torch.float32, split=False, emask=None, k=None, n=None, a_mn=False, w_mn=False,
           res_mul=False):
    """C = epilogue(A . W^T).  a, w: Operand (2-D, row-major, unit inner stride).  Returns a tensor, or an
    Operand when split=True (fp32 compute only).  a_mn / w_mn (bf16 only): the operand is handed over TRANSPOSED,
    as a [K, M] / [K, N] row-major array, and read in place (see ccx_linear_desc)."""
    M, K = a.hi.shape[::-1] if a_mn else a.hi.shape
    N, K2 = w.hi.shape[::-1] if w_mn else w.hi.shape
    if k is not None:      # logical contraction length when the operands carry zero-padded columns
        K = K2 = k
    if n is not None:
        N = n
    if K != K2 or a.dtype != w.dtype:
        raise ValueError(f"linear: operand mismatch A{tuple(a.hi.shape)} {a.dtype} W{tuple(w.hi.shape)} {w.dtype}")
    dev = a.hi.device
    d = LinearDesc()
    d.A, d.A_lo, d.W, d.W_lo = ptr(a.hi), ptr(a.lo), ptr(w.hi), ptr(w.lo)
    res = None
    if split:
        res = Operand.empty((M, N), torch.float32, dev) if out is None else out
        d.C, d.C_lo, d.ldc = ptr(res.hi), ptr(res.lo), res.hi.stride(0)
        out_dtype = torch.float32
    else:
        res = torch.empty((M, N), dtype=out_dtype, device=dev) if out is None else out
        d.C, d.C_lo, d.ldc = ptr(res), None, res.stride(0)
        out_dtype = res.dtype
    d.bias, d.colscale, d.rowscale = ptr(bias), ptr(colscale), ptr(rowscale)
    d.residual = ptr(residual)
    d.emask = ptr(emask)
    d.lda, d.ldw = a.hi.stride(0), w.hi.stride(0)
    d.ldr = residual.stride(0) if residual is not None else 0
    d.ldm = emask.stride(0) if emask is not None else 0
    d.M, d.N, d.K = M, N, K
    d.rows_per_group = rows_per_group
    d.act = act
    d.in_dtype, d.out_dtype = dt_code(a.dtype), dt_code(out_dtype)
    d.split = 1 if split else 0
    d.a_mn, d.w_mn = (1 if a_mn else 0), (1 if w_mn else 0)
    d.res_mul = 1 if res_mul else 0
    if _DEBUG_CAPTURE:
        capture_probe(f"before linear M={M} N={N} K={K} (host-side preparation / torch allocations since the last call)")
    check(lib().ccx_linear(C.byref(d), stream_ptr()), f"linear M={M} N={N} K={K}")
    return res
