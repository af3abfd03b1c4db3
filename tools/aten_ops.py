"""Which ATen ops does one CapturedTrainStep body still call, and from which source lines?  A TorchDispatchMode logs
every aten op of one eager execution of the step body (the launches the CUDA graph replays) with the innermost frame of
this package on the Python stack — autograd Function.backward bodies included, which torch.profiler's stacks miss.
View-only ops (no kernel) are listed separately.

    python tools/aten_ops.py [lstm|transformer]
"""
import collections
import os
import sys
import traceback

import torch
from torch.utils._python_dispatch import TorchDispatchMode

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder, TransformerDecoder  # noqa: E402
from imagecaptioningconvnext_b200.train_step import CapturedTrainStep, make_optimizers  # noqa: E402
from synthetic import (random_encoder_state, random_lstm_decoder_state, random_transformer_decoder_state,  # noqa: E402
                       synthetic_captions, synthetic_images)

NO_KERNEL = {"view", "reshape", "detach", "alias", "as_strided", "expand", "permute", "transpose", "t", "slice", "select",
             "unsqueeze", "squeeze", "narrow", "unbind", "split", "split_with_sizes", "_unsafe_view", "empty", "empty_like",
             "empty_strided", "new_empty", "new_empty_strided", "_local_scalar_dense", "lift_fresh", "set_", "record_stream",
             "unfold", "chunk", "resize_", "_reshape_alias", "view_as", "expand_as", "unflatten", "flatten", "diagonal",
             "is_pinned", "is_same_size", "sym_size", "sym_stride", "sym_numel", "sym_storage_offset"}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "imagecaptioningconvnext_b200")


class Log(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.ops = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func).replace("aten.", "")
        where = "(autograd engine / no package frame)"
        for fr in reversed(traceback.extract_stack()):
            if fr.filename.startswith(PKG):
                where = f"{os.path.basename(fr.filename)}:{fr.lineno} {fr.name}"
                break
        self.ops[(name, where)] += 1
        return func(*args, **(kwargs or {}))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "lstm"
    V, dev, bf16 = 9490, torch.device("cuda"), torch.bfloat16
    enc = Encoder(compute_dtype=bf16)
    enc.load_state_dict(random_encoder_state(seed=0, layer_scale=1.0))
    enc = enc.to(dev).train()
    if mode == "lstm":
        enc.fine_tune(True, 7)
        dec = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=bf16)
        dec.load_state_dict(random_lstm_decoder_state(0, V))
    else:
        enc.fine_tune(False)
        dec = TransformerDecoder(512, 512, V, 52, dev, None, None, True, compute_dtype=bf16)
        dec.load_state_dict(random_transformer_decoder_state(0, V))
    dec = dec.to(dev).train()
    d_opt, e_opt = make_optimizers(enc, dec)
    step = CapturedTrainStep(enc, dec, d_opt, e_opt)
    imgs = synthetic_images(32, 1).to(dev)
    caps, lens = synthetic_captions(32, 7, V)
    caps, lens = caps.to(dev), lens.to(dev)
    for _ in range(3):
        step.eager_step(imgs, caps, lens)
    torch.cuda.synchronize()
    log = Log()
    with log:
        step.eager_step(imgs, caps, lens)
    torch.cuda.synchronize()
    kern = {k: n for k, n in log.ops.items() if k[0].split(".")[0] not in NO_KERNEL}
    views = sum(log.ops.values()) - sum(kern.values())
    print(f"{mode}: {sum(kern.values())} aten ops that may launch a kernel, {views} view-only ops")
    for (name, where), n in sorted(kern.items(), key=lambda kv: (-kv[1], kv[0])):
        print(f"{n:4d} x {name:34s} {where}")


if __name__ == "__main__":
    main()
