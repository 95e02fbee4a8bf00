"""Development aid: time/trace the tcgen05 GEMM hook at the tier shapes (SRNN_TRACE_GEMM=1 prints in-kernel stamps)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srnn_b200 as S
L = S._lib
lib = L.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
for (M, N, K, bm, bn) in [(256, 3072, 1024, 128, 256), (256, 3072, 1024, 128, 64), (256, 20480, 1024, 128, 256), (256, 20480, 1024, 128, 128)]:
    A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    out = torch.empty(M, N, device="cuda")
    mode = S.MODE_BF16 | (bm << 8) | (bn << 16)
    for i in range(3):
        L.check(lib.srnn_gemm(M, N, K, A.data_ptr(), B.data_ptr(), None, None, 0, out.data_ptr(), mode, st()))
    torch.cuda.synchronize()
