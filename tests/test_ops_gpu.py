"""GPU parity of the individual C-ABI entry points against plain torch fp32 ops on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _g(seed=0):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("C,H,W,B", [(128, 64, 64, 2), (256, 32, 32, 2), (512, 16, 16, 3), (1024, 8, 8, 3),
                                     (128, 13, 21, 1), (384, 5, 9, 2),
                                     # more tiles than resident clusters: the persistent kernel's multi-tile loop
                                     (128, 64, 64, 5), (256, 32, 32, 11), (512, 16, 16, 41), (512, 24, 40, 7)])
@pytest.mark.parametrize("mode", ["bf16", "split", "plain"])
def test_dwconv7_ln(C, H, W, B, mode):
    from imagecaptioningconvnext_b200 import _lib
    g = _g(C + H)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(C, 1, 7, 7, generator=g) * 0.2
    b = torch.randn(C, generator=g)
    gam = 1 + 0.3 * torch.randn(C, generator=g)
    bet = 0.3 * torch.randn(C, generator=g)
    ref = F.layer_norm(F.conv2d(x, w, b, padding=3, groups=C).permute(0, 2, 3, 1), (C,), gam, bet, 1e-6)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    wd = w.reshape(C, 49).t().contiguous().cuda()
    M = B * H * W
    if mode == "bf16":
        out = torch.empty(M, C, dtype=torch.bfloat16, device="cuda")
        lo = None
        dt = _lib.CCX_BF16
    else:
        out = torch.empty(M, C, dtype=torch.float32, device="cuda")
        lo = torch.empty_like(out) if mode == "split" else None
        dt = _lib.CCX_F32
    bd, gd, ed = b.cuda(), gam.cuda(), bet.cuda()  # keep alive: data_ptr() of a temporary dangles
    rc = _lib.lib().ccx_dwconv7_ln(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), gd.data_ptr(),
                                   ed.data_ptr(), out.data_ptr(), _lib.ptr(lo), B, H, W, C, 1e-6, dt,
                                   _lib.stream_ptr())
    _lib.check(rc)
    y = out.float() + (lo if lo is not None else 0)
    tol = 1e-2 if mode == "bf16" else 2e-5
    assert rel_err(y.view(B, H, W, C), ref) < tol
    if mode == "split":  # hi must be exactly tf32-representable
        assert int((out.view(torch.int32) & 0x1FFF).abs().max()) == 0


@pytest.mark.parametrize("C,H,W,B", [(1024, 8, 8, 32), (512, 16, 16, 5), (128, 13, 21, 3), (96, 5, 9, 1), (1024, 8, 8, 1)])
def test_dwconv7_wgrad_accumulates_the_filter_gradient(C, H, W, B):
    """ccx_dwconv7_wgrad: dw[tap][c] += sum du * shifted x (torchvision convnext.py:52, the depthwise Conv2d's weight
    gradient) against torch autograd; the buffer is accumulated into, not overwritten."""
    from imagecaptioningconvnext_b200 import _lib
    g = _g(C + H + B)
    x = torch.randn(B, C, H, W, generator=g)
    du = torch.randn(B, C, H, W, generator=g)
    w = torch.zeros(C, 1, 7, 7, requires_grad=True)
    F.conv2d(x, w, None, padding=3, groups=C).backward(du)
    ref = w.grad.reshape(C, 49).t()                                    # tap-major [49][C]
    xd, dd = x.permute(0, 2, 3, 1).contiguous().cuda(), du.permute(0, 2, 3, 1).contiguous().cuda()
    base = torch.randn(49, C, generator=g)
    dw = base.clone().cuda()
    _lib.check(_lib.lib().ccx_dwconv7_wgrad(xd.data_ptr(), dd.data_ptr(), dw.data_ptr(), B, H, W, C, _lib.stream_ptr()))
    assert rel_err(dw.cpu() - base, ref) < 2e-5


@pytest.mark.parametrize("B,H,W", [(2, 256, 256), (1, 64, 96), (3, 32, 32)])
def test_stem_ln(B, H, W):
    from imagecaptioningconvnext_b200 import _lib
    g = _g(B)
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(128, 3, 4, 4, generator=g) * 0.1
    b = torch.randn(128, generator=g)
    gam, bet = 1 + 0.2 * torch.randn(128, generator=g), 0.2 * torch.randn(128, generator=g)
    ref = F.layer_norm(F.conv2d(x, w, b, stride=4).permute(0, 2, 3, 1), (128,), gam, bet, 1e-6)
    out = torch.empty(B, H // 4, W // 4, 128, device="cuda")
    wk = w.reshape(128, 48).t().contiguous().cuda()
    xd, bd, gd, ed = x.cuda(), b.cuda(), gam.cuda(), bet.cuda()
    _lib.check(_lib.lib().ccx_stem_ln(xd.data_ptr(), wk.data_ptr(), bd.data_ptr(), gd.data_ptr(),
                                      ed.data_ptr(), out.data_ptr(), B, H, W, 1e-6, _lib.stream_ptr()))
    assert rel_err(out, ref) < 2e-5


@pytest.mark.parametrize("C", [128, 256, 512])
def test_ln_rows_patchmerge_equals_ln2d_plus_conv_im2col(C):
    from imagecaptioningconvnext_b200 import _lib
    g = _g(C)
    B, H, W = 2, 8, 12
    x = torch.randn(B, H, W, C, generator=g)
    gam, bet = 1 + 0.2 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    wc = torch.randn(2 * C, C, 2, 2, generator=g) * 0.05
    ln = F.layer_norm(x, (C,), gam, bet, 1e-6)
    ref = F.conv2d(ln.permute(0, 3, 1, 2), wc, None, stride=2).permute(0, 2, 3, 1)
    out = torch.empty(B * H * W // 4, 4 * C, device="cuda")
    xd, gd, ed = x.cuda(), gam.cuda(), bet.cuda()
    _lib.check(_lib.lib().ccx_ln_rows(xd.data_ptr(), gd.data_ptr(), ed.data_ptr(),
                                      out.data_ptr(), None, None, B * H * W, C, 1e-6, _lib.CCX_F32, 1, H, W,
                                      _lib.stream_ptr()))
    wm = wc.permute(0, 2, 3, 1).reshape(2 * C, 4 * C)
    got = (out.cpu() @ wm.t()).view(B, H // 2, W // 2, 2 * C)
    assert rel_err(got, ref) < 1e-4
    # plain (no merge) bf16 output
    o2 = torch.empty(B * H * W, C, dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.lib().ccx_ln_rows(xd.data_ptr(), gd.data_ptr(), ed.data_ptr(),
                                      o2.data_ptr(), None, None, B * H * W, C, 1e-6, _lib.CCX_BF16, 0, H, W,
                                      _lib.stream_ptr()))
    assert rel_err(o2.float().view(B, H, W, C), ln) < 1e-2


@pytest.mark.parametrize("H,S", [(8, 7), (8, 14), (8, 1), (2, 7), (16, 7)])
def test_avgpool_nhwc(H, S):
    from imagecaptioningconvnext_b200 import _lib
    x = torch.randn(3, 1024, H, H, generator=_g(H))
    ref = F.adaptive_avg_pool2d(x, (S, S)).permute(0, 2, 3, 1)
    out = torch.empty(3, S, S, 1024, device="cuda")
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    _lib.check(_lib.lib().ccx_avgpool_nhwc(xd.data_ptr(), out.data_ptr(), 3, H, H, 1024, S, _lib.stream_ptr()))
    assert rel_err(out, ref) < 1e-6


# the last three shapes fill the machine with 256x256 pair tiles; pair=True runs them on the cta_group::2 kernel
@pytest.mark.parametrize("M,N,K,pair", [(128, 128, 64, False), (300, 200, 192, False), (77, 9490, 512, False),
                                        (4096, 512, 128, False), (1000, 512, 2048, False), (5, 1536, 512, False),
                                        (1, 512, 1024, False), (16384, 512, 2048, False), (16384, 512, 2048, True),
                                        (9472, 2048, 512, True), (20000, 1024, 96, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_epilogues(M, N, K, pair, dtype):
    from imagecaptioningconvnext_b200 import _lib
    from imagecaptioningconvnext_b200._lib import Operand
    _lib.lib().ccx_set_gemm_pair_mode(1 if pair else 0)
    try:
        _linear_epilogue_checks(M, N, K, dtype)
    finally:
        _lib.lib().ccx_set_gemm_pair_mode(0)


def _linear_epilogue_checks(M, N, K, dtype):
    from imagecaptioningconvnext_b200 import _lib
    from imagecaptioningconvnext_b200._lib import Operand
    g = _g(M + N)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    cs = torch.rand(N, generator=g) + 0.5
    res = torch.randn(M, N, generator=g)
    rpg = 7
    rs = torch.rand((M + rpg - 1) // rpg, generator=g) + 0.5
    A, Wt = Operand.prepare(a.cuda(), dtype), Operand.prepare(w.cuda(), dtype)
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    # bias only
    y = _lib.linear(A, Wt, bias=bias.cuda())
    assert rel_err(y, a @ w.t() + bias) < tol
    # GELU
    y = _lib.linear(A, Wt, bias=bias.cuda(), act=_lib.ACT_GELU)
    assert rel_err(y, F.gelu(a @ w.t() + bias)) < tol
    # ReLU + layer-scale + row-scale + residual
    y = _lib.linear(A, Wt, bias=bias.cuda(), act=_lib.ACT_RELU, colscale=cs.cuda(), rowscale=rs.cuda(),
                    rows_per_group=rpg, residual=res.cuda())
    ref = F.relu(a @ w.t() + bias) * cs * rs.repeat_interleave(rpg)[:M, None] + res
    assert rel_err(y, ref) < tol
    if dtype == torch.float32:
        op = _lib.linear(A, Wt, bias=bias.cuda(), split=True)
        assert rel_err(op.hi + op.lo, a @ w.t() + bias) < tol
    else:
        y = _lib.linear(A, Wt, bias=bias.cuda(), out_dtype=torch.bfloat16)
        assert y.dtype == torch.bfloat16 and rel_err(y.float(), a @ w.t() + bias) < tol


@pytest.mark.parametrize("M,N,K", [(512, 256, 320), (1024, 4096, 2048), (304, 200, 136), (4096, 1024, 1664), (128, 64, 64)])
@pytest.mark.parametrize("a_mn,w_mn", [(True, False), (False, True), (True, True)])
def test_linear_reads_transposed_operands_in_place(M, N, K, a_mn, w_mn):
    """ccx_linear with a_mn / w_mn: the operand is handed over as its transpose ([K, M] / [K, N] row-major, pitch a
    multiple of 8) and read MN-major by the tensor core — what lets dX = dY . W and dW = dY^T . X run on the forward's
    own buffers.  bf16 products are exact in fp32, so the result matches the plain call up to summation order."""
    from imagecaptioningconvnext_b200 import _lib
    from imagecaptioningconvnext_b200._lib import Operand
    g = _g(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)

    def op(x, transposed):
        if not transposed:
            return Operand(x.cuda(), None, torch.bfloat16)
        rows, cols = x.shape[1], x.shape[0]                      # [K, M] storage with the pitch padded to 8
        buf = torch.zeros(rows, (cols + 7) // 8 * 8, dtype=torch.bfloat16, device="cuda")
        buf[:, :cols] = x.t().cuda()
        return Operand(buf[:, :cols], None, torch.bfloat16)
    y = _lib.linear(op(a, a_mn), op(w, w_mn), bias=bias.cuda(), residual=res.cuda(), a_mn=a_mn, w_mn=w_mn)
    ref = a.float() @ w.float().t() + bias + res
    assert rel_err(y, ref) < 1e-5


# The TMA epilogue of gemm_tn_kernel (csrc/ccx_gemm_epilogue.cuh): bf16 operands, output rows 16-byte aligned.
#  (40000,256,128) / (40000,128,256): several tiles per CTA (box prefetch across tiles, the last one through the ring);
#  (8192,512,2048): one tile per CTA (ring only); (300,200,192) / (1000,72,64): ragged M and N, clipped by the tensor map;
#  (2100,1536,512): N tiles of different CTAs share rows.
@pytest.mark.parametrize("M,N,K", [(40000, 256, 128), (40000, 128, 256), (8192, 512, 2048), (300, 200, 192),
                                   (1000, 72, 64), (2100, 1536, 512)])
def test_linear_tma_epilogue_paths(M, N, K):
    from imagecaptioningconvnext_b200 import _lib
    from imagecaptioningconvnext_b200._lib import Operand
    g = _g(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    cs = torch.rand(N, generator=g) + 0.5
    rpg = 49
    rs = torch.rand((M + rpg - 1) // rpg, generator=g) + 0.5
    res = torch.randn(M, N, generator=g)
    A, Wt = Operand(a.cuda(), None, torch.bfloat16), Operand(w.cuda(), None, torch.bfloat16)
    prod = a.float() @ w.float().t()
    tol = 1e-2          # bf16 output rounding (2^-9) plus the tanh-form GELU; the products themselves are exact
    # Linear + GELU -> bf16 (the CNBlock's first Linear)
    y = _lib.linear(A, Wt, bias=bias.cuda(), act=_lib.ACT_GELU, out_dtype=torch.bfloat16)
    assert rel_err(y.float(), F.gelu(prod + bias)) < tol
    # bias + layer-scale x row-scale + fp32 residual, written IN PLACE over the residual (the CNBlock's second Linear)
    x = res.clone().cuda()
    y = _lib.linear(A, Wt, bias=bias.cuda(), colscale=cs.cuda(), rowscale=rs.cuda(), rows_per_group=rpg, residual=x, out=x)
    ref = (prod + bias) * cs * rs.repeat_interleave(rpg)[:M, None] + res
    assert y.data_ptr() == x.data_ptr() and rel_err(y, ref) < 1e-5
    # bf16 output with a bf16 residual and ReLU
    rb = res.to(torch.bfloat16).cuda()
    y = _lib.linear(A, Wt, bias=bias.cuda(), act=_lib.ACT_RELU, residual=rb, out_dtype=torch.bfloat16)
    assert rel_err(y.float(), F.relu(prod + bias) + rb.float().cpu()) < tol
    # GELU' of the recomputed pre-activation times an upstream gradient, in place (CNBlock backward)
    dh = res.clone().cuda()
    _lib.linear(A, Wt, bias=bias.cuda(), act=_lib.ACT_GELU_GRAD, residual=dh, out=dh, res_mul=True)
    # (bf16 operands: the derivative of the tanh-form GELU the bf16 forward evaluates, with tanh.approx ~5e-4 abs;
    #  the erf form of the fp32 path is covered by the fp32 encoder-backward parity tests)
    xr = (prod + bias).double()
    u = xr * (0.7978845608 + 0.035677408136 * xr * xr)
    t = torch.tanh(u)
    gp = 0.5 * (1 + t) + 0.5 * xr * (1 - t * t) * (0.7978845608 + 3 * 0.035677408136 * xr * xr)
    assert rel_err(dh, (gp * res.double()).float()) < 2e-3
    exact = 0.5 * (1 + torch.erf(xr / 2 ** 0.5)) + xr * torch.exp(-0.5 * xr * xr) / (2 * torch.pi) ** 0.5
    assert float((gp - exact).abs().max()) < 2e-3          # the two forms of GELU' agree to 1e-3
    # a strided output view (row pitch 2N): the tensor map carries the pitch
    wide = torch.zeros(M, 2 * N, device="cuda")
    _lib.linear(A, Wt, bias=bias.cuda(), out=wide[:, N:])
    assert rel_err(wide[:, N:], prod + bias) < 1e-5 and float(wide[:, :N].abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K", [(32, 2048, 2048), (17, 2048, 2048), (1, 1536, 512), (32, 512, 1536), (9, 72, 64),
                                   (32, 9490, 512)])
def test_skinny_gemm_matches_exact_bf16_product(M, N, K):
    """M <= 32 bf16 GEMMs take the mma.sync kernel (csrc/gemm_skinny.cu).  Against the exact product of the
    bf16-rounded operands (fp64), with bias, with a strided residual and a strided output: the tolerance only
    leaves room for fp32 summation order — a wrong fragment / k-permutation mapping fails by O(1)."""
    from imagecaptioningconvnext_b200 import _lib
    from imagecaptioningconvnext_b200._lib import Operand
    g = _g(M * 7 + N)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    A, Wt = Operand.prepare(a.cuda(), torch.bfloat16), Operand.prepare(w.cuda(), torch.bfloat16)
    exact = (a.bfloat16().double() @ w.bfloat16().double().t())
    y = _lib.linear(A, Wt, bias=bias.cuda())
    assert y.shape == (M, N) and rel_err(y.double(), exact + bias.double()) < 2e-5
    # residual read with a leading dimension, output written into a column slice of a wider buffer
    wide_res = torch.randn(M, N + 24, generator=g).cuda()
    wide_out = torch.full((M, N + 40), 7.0, device="cuda")
    _lib.linear(A, Wt, residual=wide_res[:, 8:8 + N], out=wide_out[:, 16:16 + N])
    assert rel_err(wide_out[:, 16:16 + N].double(), exact + wide_res[:, 8:8 + N].cpu().double()) < 2e-5
    assert bool((wide_out[:, :16] == 7.0).all()) and bool((wide_out[:, 16 + N:] == 7.0).all())


@pytest.mark.parametrize("R,C", [(1632, 2048), (200, 512), (77, 40), (33, 9490), (64, 24)])
@pytest.mark.parametrize("src_bf16", [False, True])
def test_convert_operand_rows_and_transpose_bf16_paths(R, C, src_bf16):
    """ccx_convert_operand: the vectorised bf16-destination paths (8 elements per thread, 64x64 transpose tiles) and
    the generic fallbacks (C % 8 != 0) against torch: cast, element-wise multiplier, ReLU-mask mode, zero padding of
    the transposed rows."""
    from imagecaptioningconvnext_b200 import _lib
    from imagecaptioningconvnext_b200._lib import Operand
    from imagecaptioningconvnext_b200.train_ops import to_operand
    g = torch.Generator().manual_seed(R * 131 + C)
    x = torch.randn(R, C, generator=g).cuda()
    src = Operand(x.to(torch.bfloat16), None, torch.bfloat16) if src_bf16 else x
    xr = src.hi.float() if src_bf16 else x
    mul = torch.randn(R, C, generator=g).cuda()
    for mode, scale, ref in ((0, 1.0, xr), (1, 1.0, xr * mul), (2, 2.0, torch.where(mul > 0, xr * 2.0, torch.zeros_like(xr)))):
        m = None if mode == 0 else mul
        o = to_operand(src, torch.bfloat16, m, mode, scale)
        assert torch.equal(o.hi.float(), ref.to(torch.bfloat16).float()), (mode, "rows")
        t = to_operand(src, torch.bfloat16, m, mode, scale, transpose=True)
        assert t.hi.shape[0] == C and t.hi.shape[1] >= R
        assert torch.equal(t.hi[:, :R].float(), ref.t().to(torch.bfloat16).float()), (mode, "transpose")
        assert float(t.hi[:, R:].float().abs().max() if t.hi.shape[1] > R else 0.0) == 0.0


@pytest.mark.parametrize("R,C", [(1632, 9490), (8192, 512), (100, 36), (5, 7)])
def test_colsum_acc_matches_torch(R, C):
    from imagecaptioningconvnext_b200.train_ops import colsum_acc
    g = torch.Generator().manual_seed(R + C)
    x = torch.randn(R, C, generator=g).cuda()
    mul = torch.randn(R, C, generator=g).cuda()
    for mode, scale, ref in ((0, 1.0, x), (1, 1.0, x * mul), (2, 2.0, torch.where(mul > 0, x * 2.0, torch.zeros_like(x)))):
        out = torch.ones(C, device="cuda")                       # accumulates
        colsum_acc(x, out, None if mode == 0 else mul, mode, scale)
        assert rel_err(out, 1.0 + ref.double().sum(0).float()) < 1e-5, mode


@pytest.mark.parametrize("R,C", [(1632, 512), (2048, 4096), (37, 24), (5, 8)])
def test_convert_colsum_equals_the_two_separate_passes(R, C):
    """ccx_convert_colsum (one read of dY -> bf16 operand + bias gradient) against ccx_convert_operand and torch sums,
    for the three multiplier modes; the sums accumulate (+=)."""
    from imagecaptioningconvnext_b200 import _lib
    from imagecaptioningconvnext_b200._lib import ptr
    from imagecaptioningconvnext_b200.train_ops import to_operand
    g = torch.Generator().manual_seed(R * 7 + C)
    wide = torch.randn(R, C + 8, generator=g).cuda()
    x = wide[:, 4:4 + C]                                            # leading dimension != C, still 16-byte aligned
    mul = torch.randn(R, C, generator=g).cuda()
    for mode, scale, ref in ((0, 1.0, x), (1, 1.0, x * mul), (2, 2.0, torch.where(mul > 0, x * 2.0, torch.zeros_like(x)))):
        m = None if mode == 0 else mul
        out = torch.full((R, C), 9.0, dtype=torch.bfloat16, device="cuda")
        sums = torch.ones(C, device="cuda")
        _lib.check(_lib.lib().ccx_convert_colsum(ptr(x), x.stride(0), ptr(m), C if m is not None else 0, mode, scale,
                                                 ptr(out), C, ptr(sums), R, C, _lib.stream_ptr()), "convert_colsum")
        assert torch.equal(out, to_operand(x.contiguous(), torch.bfloat16, m, mode, scale).hi), mode
        assert rel_err(sums, 1.0 + ref.double().sum(0).float()) < 1e-5, mode
    bad = torch.randn(8, 6).cuda()                                   # C % 4 != 0: refused, never a silent fallback
    assert _lib.lib().ccx_convert_colsum(ptr(bad), 6, None, 0, 0, 1.0, ptr(out), 6, ptr(sums), 8, 6,
                                         _lib.stream_ptr()) != 0


def test_cast_bf16_vector_and_tail_paths():
    from imagecaptioningconvnext_b200 import _lib
    for n in (8 * 1000, 8 * 1000 + 3, 5):
        x = torch.randn(n, generator=torch.Generator().manual_seed(n)).cuda()
        assert torch.equal(_lib.cast_bf16(x), x.to(torch.bfloat16))


@pytest.mark.parametrize("Tq,Tk,causal,pad,drop", [(52, 52, 1, True, True), (52, 49, 0, False, True), (64, 64, 1, False, False),
                                                   (7, 49, 0, False, False), (33, 17, 0, True, True)])
def test_tensor_core_attention_matches_the_simt_kernels(Tq, Tk, causal, pad, drop):
    """csrc/mha_tc.cu (mma.sync, bf16 operands) against mha_small_kernel / mha_bwd_kernel (fp32 SIMT) on the same inputs:
    context, the softmax probabilities handed to the backward, and dQ / dK / dV.  nn.MultiheadAttention inside
    nn.TransformerDecoderLayer (models/transformerDecoder.py:102-106), B x 8 heads of 64."""
    import math
    from imagecaptioningconvnext_b200 import _lib
    L, st = _lib.lib(), _lib.stream_ptr()
    B, H, hd = 3, 8, 64
    D = H * hd
    g = _g(Tq * 100 + Tk)
    q = torch.randn(B, Tq, D, generator=g).cuda()
    k = torch.randn(B, Tk, D, generator=g).cuda()
    v = torch.randn(B, Tk, D, generator=g).cuda()
    key_pad = None
    if pad:
        key_pad = torch.zeros(B, Tk, dtype=torch.uint8)
        key_pad[0, Tk - 3:] = 1
        key_pad[2, Tk // 2:] = 1
        key_pad = key_pad.cuda()
    pm = ((torch.rand(B, H, Tq, Tk, generator=g) > 0.1).float() / 0.9).cuda() if drop else None
    scale = 1.0 / math.sqrt(hd)
    outs = {}
    for name, dt, tdt in (("simt", _lib.CCX_F32, torch.float32), ("tc", _lib.CCX_BF16, torch.bfloat16)):
        ctx = torch.zeros(B, Tq, D, dtype=tdt, device="cuda")
        probs = torch.zeros(B, H, Tq, Tk, device="cuda")
        _lib.check(L.ccx_mha_small(q.data_ptr(), Tq * D, D, k.data_ptr(), Tk * D, D, v.data_ptr(), Tk * D, D,
                                   ctx.data_ptr(), None, dt, Tq * D, D, _lib.ptr(key_pad), _lib.ptr(pm),
                                   probs.data_ptr(), B, H, Tq, Tk, hd, causal, 0, scale, 1, st))
        outs[name] = (ctx.float(), probs)
    assert rel_err(outs["tc"][0], outs["simt"][0]) < 2e-2
    assert rel_err(outs["tc"][1], outs["simt"][1]) < 2e-2
    dctx = torch.randn(B, Tq, D, generator=g).cuda()
    probs = outs["simt"][1]
    grads = {}
    for name, fn in (("simt", L.ccx_mha_bwd), ("tc", L.ccx_mha_bwd_tc)):
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        _lib.check(fn(q.data_ptr(), Tq * D, D, k.data_ptr(), Tk * D, D, v.data_ptr(), Tk * D, D, dctx.data_ptr(), Tq * D, D,
                      probs.data_ptr(), _lib.ptr(pm), dq.data_ptr(), Tq * D, D, dk.data_ptr(), Tk * D, D,
                      dv.data_ptr(), Tk * D, D, B, H, Tq, Tk, hd, scale, st))
        grads[name] = (dq, dk, dv)
    for a, b in zip(grads["tc"], grads["simt"]):
        assert rel_err(a, b) < 2e-2
