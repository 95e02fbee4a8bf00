"""SRNN_TRACE_GEMM=1 python tools/gemm_trace.py [M N K bn rows] -> in-kernel clock64 trace of CTA (0,0,0) of k_gemm_umma."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srnn_b200 as S

M, N, K, bn, rows = [int(x) for x in (sys.argv[1:6] + ["8192", "1024", "1024", "256", "1"][len(sys.argv) - 1:])]
L = S._lib
A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
out = torch.empty(M, N, device="cuda")
mode = S.MODE_BF16 | (128 << 8) | (bn << 16) | (rows << 28)
for _ in range(3):
    L.check(L.load().srnn_gemm(M, N, K, A.data_ptr(), B.data_ptr(), None, None, 0, out.data_ptr(), mode,
                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
os.environ.pop("SRNN_TRACE_GEMM", None)
