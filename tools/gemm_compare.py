"""libccx tcgen05 GEMM vs torch.matmul (cuBLASLt) on the ConvNeXt stage shapes, bf16, CUDA-event timed, L2 flushed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecaptioningconvnext_b200 import _lib
from imagecaptioningconvnext_b200._lib import Operand
dev = torch.device("cuda")
flush = torch.empty(256 * 2**20, dtype=torch.uint8, device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
print(f"{'shape (M,N,K)':>22s} {'ours plain':>11s} {'ours fused':>11s} {'cuBLAS':>9s}   TFLOP/s ours-fused / cuBLAS")
for (M, N, K, kind) in [(262144, 512, 128, "gelu"), (262144, 128, 512, "res"), (65536, 1024, 256, "gelu"), (65536, 256, 1024, "res"),
                        (16384, 2048, 512, "gelu"), (16384, 512, 2048, "res"), (4096, 4096, 1024, "gelu"), (4096, 1024, 4096, "res")]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    A, W = Operand(a, None, torch.bfloat16), Operand(w, None, torch.bfloat16)
    bias = torch.randn(N, device=dev); cs = torch.rand(N, device=dev); res = torch.randn(M, N, device=dev)
    out_b = torch.empty(M, N, device=dev, dtype=torch.bfloat16); out_f = torch.empty(M, N, device=dev)
    plain = timeit(lambda: _lib.linear(A, W, out=out_b))
    if kind == "gelu":
        fused = timeit(lambda: _lib.linear(A, W, bias=bias, act=_lib.ACT_GELU, out=out_b))
    else:
        fused = timeit(lambda: _lib.linear(A, W, bias=bias, colscale=cs, residual=res, out=out_f))
    cub = timeit(lambda: torch.matmul(a, w.t(), out=out_b))
    fl = 2.0 * M * N * K
    print(f"{str((M, N, K)):>22s} {plain:9.1f}us {fused:9.1f}us {cub:7.1f}us   {fl / fused / 1e6:7.0f} / {fl / cub / 1e6:7.0f}   [{kind}]")
