"""The encoder's GEMM shapes at batch 32, back to back (operands L2-warm as inside the step), CUDA-event timed:
libccx (CCX_GEMM_2CTA / CCX_GEMM_EPI select the kernel / epilogue) next to torch.matmul (cuBLASLt, no epilogue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecaptioningconvnext_b200 import _lib
from imagecaptioningconvnext_b200._lib import Operand
dev = torch.device("cuda")
def timeit(fn, n=40):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
print("variant:", {k: v for k, v in os.environ.items() if k.startswith("CCX_GEMM")})
tot_f = tot_c = 0.0
for (M, N, K, kind, count) in [(131072, 512, 128, "gelu", 3), (131072, 128, 512, "res", 3), (32768, 1024, 256, "gelu", 3),
                               (32768, 256, 1024, "res", 3), (8192, 2048, 512, "gelu", 27), (8192, 512, 2048, "res", 27),
                               (2048, 4096, 1024, "gelu", 3), (2048, 1024, 4096, "res", 3)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    A, W = Operand(a, None, torch.bfloat16), Operand(w, None, torch.bfloat16)
    bias = torch.randn(N, device=dev); cs = torch.rand(N, device=dev); res = torch.randn(M, N, device=dev)
    out_b = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if kind == "gelu":
        fused = timeit(lambda: _lib.linear(A, W, bias=bias, act=_lib.ACT_GELU, out=out_b))
    else:
        fused = timeit(lambda: _lib.linear(A, W, bias=bias, colscale=cs, residual=res, out=res))     # in place, fp32
    cub = timeit(lambda: torch.matmul(a, w.t(), out=out_b))
    fl = 2.0 * M * N * K
    tot_f += fused * count; tot_c += cub * count
    print(f"{str((M, N, K)):>22s} [{kind:4s}] x{count:2d}: ours {fused:6.1f} us {fl / fused / 1e6:6.0f} TF/s   cuBLAS(plain) {cub:6.1f} us {fl / cub / 1e6:6.0f} TF/s")
print(f"encoder forward GEMMs per step: ours {tot_f / 1e3:.3f} ms, cuBLAS plain {tot_c / 1e3:.3f} ms")
