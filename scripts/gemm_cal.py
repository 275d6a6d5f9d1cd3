"""Profiling helper: the dominant contraction (cond_transform, [14336 x 8192 x 920], bf16x3, through lfi_gemm) and a cuBLAS bf16
8192^3 product in ONE ncu capture, so that the tensor-pipe counters of both are read with the same metric."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lets_face_it_b200 import _cabi as cabi
L = cabi.lib()
dev = "cuda"
M, N, K = 14336, 8192, 920
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev); bias = torch.zeros(N, device=dev)
mode = cabi.GEMM_BF16X3
ws = torch.empty(max(int(L.lfi_gemm_ws_bytes(mode, 0, 1, M, N, K, 1)), 256), dtype=torch.uint8, device=dev)
def ours():
    cabi.check(L.lfi_gemm(mode, 0, 1, M, N, K, A.data_ptr(), K, 0, W.data_ptr(), K, 0, C.data_ptr(), N, 0, bias.data_ptr(), 0, None, 0, 0, 1,
                          cabi.EPI_BIAS | cabi.EPI_LRELU, ws.data_ptr(), ws.numel(), cabi.stream_ptr()), "lfi_gemm")
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ours(); torch.matmul(a, b)
torch.cuda.synchronize(); torch.cuda.profiler.start()
ours(); torch.matmul(a, b)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
