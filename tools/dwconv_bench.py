"""ccx_dwconv7_ln on the four ConvNeXt stage shapes at batch 32 (256x256 images), CUDA-event timed back to back
(inputs L2-warm as inside the encoder).  Kernel choice through CCX_DWCONV_V2 / CCX_DWCONV_CH128 / CCX_DWCONV_V1.
FP32-pipe floor = MACs / (148 SMs x 128 lanes) cycles; HBM floor = (4 + 2) bytes per element at the measured peak."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecaptioningconvnext_b200 import _lib
B = 32
print("variant:", {k: v for k, v in os.environ.items() if k.startswith("CCX_DWCONV")})
for C, H in ((128, 64), (256, 32), (512, 16), (1024, 8)):
    x = torch.randn(B, H, H, C, device="cuda")
    w = torch.randn(49, C, device="cuda") * 0.2
    b, g, e = (torch.randn(C, device="cuda") for _ in range(3))
    out = torch.empty(B * H * H, C, dtype=torch.bfloat16, device="cuda")
    call = lambda: _lib.check(_lib.lib().ccx_dwconv7_ln(x.data_ptr(), w.data_ptr(), b.data_ptr(), g.data_ptr(), e.data_ptr(),
                                                        out.data_ptr(), None, B, H, H, C, 1e-6, _lib.CCX_BF16, _lib.stream_ptr()))
    for _ in range(5): call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n): call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    macs = B * H * H * C * 49
    print(f"C={C:5d} {H}x{H}: {us:7.1f} us   fp32-pipe floor {macs / (148 * 128) / 1.9e3:6.1f} us   hbm floor {B*H*H*C*6/6.5e6:6.1f} us")
