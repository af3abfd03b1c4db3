"""Host-side logic that needs no GPU: caption bookkeeping of the free-running training path, buffer carving,
state_dict key mapping of the attention-map decoder, the host-copy stash."""
import os

import torch


def test_generated_captions_lengths_follow_first_end_token():
    """utils/utils.py:270-276: decode length = index of the first <end> + 1, or maxDecodeLen."""
    from imagecaptioningconvnext_b200.decoder_train import generated_captions
    start, end, T = 98, 99, 6
    seqs = torch.tensor([[5, 6, end, 0, 0, 0],        # finished at step 2
                         [end, 0, 0, 0, 0, 0],        # finished at once
                         [1, 2, 3, 4, 5, 6],          # never finished
                         [7, end, 8, end, 0, 0]])     # first <end> counts
    caps, lens = generated_captions(seqs, start, end, T)
    assert caps.shape == (4, T + 1) and bool((caps[:, 0] == start).all()) and torch.equal(caps[:, 1:], seqs)
    assert lens.view(-1).tolist() == [3 + 1, 1 + 1, 6 + 1, 2 + 1]          # caption_lengths = decode length + 1


def test_zeros_many_is_one_allocation_of_aligned_zero_views():
    from imagecaptioningconvnext_b200.train_ops import zero_grads_like, zeros_many
    shapes = [(3, 5), (7,), (2, 2, 2), (130,)]
    bufs = zeros_many(shapes, torch.device("cpu"))
    assert [tuple(b.shape) for b in bufs] == shapes
    assert len({b.untyped_storage().data_ptr() for b in bufs}) == 1
    offs = [b.storage_offset() for b in bufs]
    assert all(o % 64 == 0 for o in offs) and offs == sorted(offs)
    for b in bufs:
        assert b.is_contiguous() and float(b.abs().sum()) == 0.0
    bufs[0].fill_(1.0)
    assert float(bufs[1].abs().sum()) == 0.0                              # no overlap
    p = [("a", torch.nn.Parameter(torch.ones(4, 4))), ("b", torch.nn.Parameter(torch.ones(3), requires_grad=False))]
    g = zero_grads_like(p)
    assert set(g) == {"a"} and g["a"].shape == (4, 4)


def test_attention_viz_decoder_uses_the_reference_key_names(golden_dir):
    from imagecaptioningconvnext_b200 import TransformerDecoder, TransformerDecoderForAttentionViz
    keys = torch.load(os.path.join(golden_dir, "attvis.pt"))["state_dict_keys"]
    m = TransformerDecoderForAttentionViz(512, 512, 9490, 52, torch.device("cpu"))
    assert sorted(m.state_dict().keys()) == keys
    # round trip through the plain decoder's naming
    plain = TransformerDecoder(512, 512, 9490, 52, torch.device("cpu"), None, None, True)
    renamed = {k.replace("transformer_decoder.layers.", "decoder_layers."): v for k, v in plain.state_dict().items()}
    m.load_state_dict(renamed)
    for (ka, a), (kb, b) in zip(sorted(m.state_dict().items()), sorted(renamed.items())):
        assert ka == kb and torch.equal(a, b)
    assert len(m.decoder_layers) == 6


def test_host_copy_stash_is_bound_to_the_tensor_object_and_version():
    from imagecaptioningconvnext_b200 import _host

    class FakeCuda(torch.Tensor):            # a CPU tensor that claims to live on the GPU (no GPU in this suite)
        @property
        def is_cuda(self):
            return True

    t = torch.arange(4).as_subclass(FakeCuda)
    assert _host.host_copy(torch.arange(4)) is not None                    # CPU tensors pass through
    _host._HOST_COPIES.clear()
    h = torch.tensor([9, 9, 9, 9])
    _host.stash_host_copy(t, h)
    assert _host.host_copy(t).data_ptr() == h.data_ptr()                   # same object, same version
    t.add_(1)                                                              # in-place update -> stale
    assert _host.host_copy(t).data_ptr() != h.data_ptr()
    u = torch.arange(4).as_subclass(FakeCuda)                              # another tensor, equal content
    assert _host.host_copy(u).data_ptr() != h.data_ptr()
    try:
        _host.stash_host_copy(t, torch.zeros(3))
        raise AssertionError("shape mismatch accepted")
    except ValueError:
        pass


def test_encoder_fine_tune_marks_the_same_parameters_as_the_reference(golden_dir):
    """models/encoder.py:15-21,29-34: constructor default (only child 7 trainable) and every startingLayer."""
    from imagecaptioningconvnext_b200 import Encoder
    g = torch.load(os.path.join(golden_dir, "fine_tune.pt"))
    enc = Encoder()
    trainable = lambda: sorted(n for n, p in enc.named_parameters() if p.requires_grad)
    assert sorted(enc.state_dict().keys()) == g["keys"]
    assert sum(p.numel() for p in enc.parameters()) == g["n_params"] == 87564416
    assert trainable() == g["default"]
    for L in range(0, 9):
        enc.fine_tune(True, L)
        assert trainable() == g[L], L
    enc.fine_tune(False)
    assert trainable() == g["off"] == []


def test_clamp_adam_loads_a_reference_adam_state_dict_and_survives_pickling():
    """The reference resumes with optimizer.load_state_dict(checkpoint[...]) of a torch.optim.Adam state
    (trainMultiGPU.py:218,223): no 'grad_clip' key there; pickled ClampAdam objects (torch.save of the optimizer)
    must come back with their pointer-table cache; weight decay / amsgrad states are refused, not silently ignored."""
    import copy
    import pickle

    import pytest

    from imagecaptioningconvnext_b200.optim import ClampAdam
    w = [torch.nn.Parameter(torch.ones(4, 3)), torch.nn.Parameter(torch.ones(5))]
    ref = torch.optim.Adam(w, lr=1e-4)
    for p in w:
        p.grad = torch.ones_like(p)
    ref.step()
    sd = copy.deepcopy(ref.state_dict())
    assert "grad_clip" not in sd["param_groups"][0]
    opt = ClampAdam(w, lr=3e-4, grad_clip=5.0)
    opt._tables[0] = ("stale",)
    opt.load_state_dict(sd)
    g = opt.param_groups[0]
    assert g["grad_clip"] == 5.0 and g["lr"] == 1e-4 and opt._tables == {}
    assert torch.equal(opt.state[w[0]]["exp_avg"], ref.state[w[0]]["exp_avg"])
    assert float(opt.state[w[1]]["step"]) == 1.0
    clone = pickle.loads(pickle.dumps(opt))
    assert clone._tables == {} and clone.param_groups[0]["grad_clip"] == 5.0
    bad = copy.deepcopy(sd)
    bad["param_groups"][0]["weight_decay"] = 0.01
    with pytest.raises(ValueError):
        ClampAdam(w, lr=1e-4).load_state_dict(bad)
    bad = copy.deepcopy(sd)
    bad["param_groups"][0]["amsgrad"] = True
    with pytest.raises(ValueError):
        ClampAdam(w, lr=1e-4).load_state_dict(bad)


def test_refresh_plan_table_layout_and_shape_checks():
    """_host.RefreshPlan (the table behind ccx_cast_segments): tile offsets accumulate per piece, flags encode
    transpose / fp32 destination, 1-D tensors become one-row pieces, and a mismatched destination is refused when the
    plan is built (not discovered on the device)."""
    import ctypes

    import pytest

    from imagecaptioningconvnext_b200 import _lib
    from imagecaptioningconvnext_b200._host import RefreshPlan
    assert ctypes.sizeof(_lib.CastSeg) == 64                       # mirrors `ccx_cast_seg` (include/ccx.h)
    w = torch.zeros(130, 70)
    b = torch.zeros(70)
    plan = RefreshPlan()
    plan.add(torch.zeros(130, 70, dtype=torch.bfloat16), w)                                   # 3 x 2 tiles
    plan.add(torch.zeros(70, 200, dtype=torch.bfloat16)[:, 10:140], w, transpose=True)        # 3 x 2 tiles
    plan.add(torch.zeros(5, 70, dtype=torch.bfloat16), w, row_map=torch.tensor([4, 3, 2, 1, 0]))   # 1 x 2 tiles
    plan.add(torch.zeros(70), b, src2=b)                                                      # 1 x 2 tiles, fp32
    assert [s.tile0 for s in plan.segs] == [0, 6, 12, 14] and plan.tiles == 16
    assert [s.flags for s in plan.segs] == [0, 1, 0, 2]
    assert [(s.rows, s.cols) for s in plan.segs] == [(130, 70), (130, 70), (5, 70), (1, 70)]
    assert plan.segs[1].dst_ld == 200 and plan.segs[1].src_ld == 70
    assert plan.segs[3].src2 == b.data_ptr() and plan.segs[0].src2 is None
    for bad_dst, kw in ((torch.zeros(70, 130, dtype=torch.bfloat16), {}),                     # not transposed: wrong shape
                        (torch.zeros(130, 70, dtype=torch.bfloat16), {"transpose": True}),
                        (torch.zeros(130, 70, dtype=torch.float16), {}),                      # unsupported destination type
                        (torch.zeros(130, 140, dtype=torch.bfloat16)[:, ::2], {})):           # inner stride != 1
        with pytest.raises(ValueError):
            RefreshPlan().add(bad_dst, w, **kw)
    with pytest.raises(ValueError):
        RefreshPlan().add(torch.zeros(130, 70, dtype=torch.bfloat16), w, src2=torch.zeros(130, 71)[:, :70])


def test_no_gc_collects_first_and_restores_the_collector():
    """_host.no_gc (wrapped around every CUDA-graph capture): cycles that died earlier are collected BEFORE the block,
    nothing is collected inside it, and the collector's state is restored afterwards."""
    import gc
    import weakref

    from imagecaptioningconvnext_b200._host import no_gc

    class Node:
        pass

    def dead_cycle():
        a, b = Node(), Node()
        a.other, b.other = b, a
        return weakref.ref(a)

    assert gc.isenabled()
    early = dead_cycle()
    with no_gc():
        assert early() is None and not gc.isenabled()
        inside = dead_cycle()
        junk = [[] for _ in range(5000)]                  # enough allocations to trip an automatic collection
        del junk
        assert inside() is not None
    assert gc.isenabled()
    gc.collect()
    assert inside() is None
    gc.disable()
    try:
        with no_gc():
            pass
        assert not gc.isenabled()                         # was off before: stays off
    finally:
        gc.enable()
