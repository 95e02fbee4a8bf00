"""tcgen05/TMA GEMM building block (SRNN_MODE_BF16) against an fp32 torch reference on bf16-rounded operands."""
import ctypes as C

import numpy as np
import pytest
import torch

import srnn_b200 as S

pytestmark = pytest.mark.gpu
L = S._lib


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("bm,bn", [(128, 64), (128, 256), (128, 128), (128, 32), (64, 32), (64, 64)])
@pytest.mark.parametrize("shape", [(256, 1024, 1024), (100, 300, 200), (32, 128, 64)])
def test_umma_gemm(bm, bn, shape):
    M, N, K = shape                       # rows, features, K
    g = torch.Generator().manual_seed(M + N + K + bm + bn)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    bias, add = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    Ab, Bb = A.bfloat16().double(), B.bfloat16().double()
    ref = torch.relu(Ab @ Bb.t() + bias.double() + add.double()).float()
    dA, dB, db, da = A.cuda(), B.cuda(), bias.cuda(), add.cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    mode = S.MODE_BF16 | (bm << 8) | (bn << 16)
    L.check(L.load().srnn_gemm(M, N, K, dA.data_ptr(), dB.data_ptr(), db.data_ptr(), da.data_ptr(), 1,
                               out.data_ptr(), mode, stream()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    err = np.abs(got - ref.numpy()).max()
    assert np.isfinite(got).all(), "non-finite output"
    assert err < 2e-3 * K ** 0.5, "max err %g" % err


@pytest.mark.parametrize("bn", [256, 128])
@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("shape", [(256, 1024, 1024), (100, 304, 200), (300, 256, 64), (1000, 40, 640)])
def test_umma_gemm_rows_orientation(bn, split, shape):
    """ROWS orientation (activation rows on the TMEM lanes, vector epilogue) with and without split-K."""
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K + bn + split)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    bias, add = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    Ab, Bb = A.bfloat16().double(), B.bfloat16().double()
    ref = Ab @ Bb.t()
    if not split:
        ref = torch.relu(ref + bias.double() + add.double())
    dA, dB, db, da = A.cuda(), B.cuda(), bias.cuda(), add.cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    mode = S.MODE_BF16 | (128 << 8) | (bn << 16) | (1 << 28) | (split << 29)
    L.check(L.load().srnn_gemm(M, N, K, dA.data_ptr(), dB.data_ptr(), None if split else db.data_ptr(),
                               None if split else da.data_ptr(), 0 if split else 1, out.data_ptr(), mode, stream()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.isfinite(got).all(), "non-finite output"
    err = np.abs(got - ref.float().numpy()).max()
    assert err < 2e-3 * K ** 0.5, "max err %g" % err


@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("shape", [(256, 1024, 1024), (100, 304, 200), (300, 256, 64), (1000, 40, 650), (128, 128, 133120 // 8)])
def test_umma_gemm_mn_major_operands(split, shape):
    """Both operands MN-major (given as A^T, B^T in memory): the transposed-copy-free weight-gradient form dW = dOut^T . In."""
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K + split)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    ref = (A.bfloat16().double() @ B.bfloat16().double().t()).float()
    dA, dB = A.cuda(), B.cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    mode = S.MODE_BF16 | (1 << 30) | (split << 29)
    L.check(L.load().srnn_gemm(M, N, K, dA.data_ptr(), dB.data_ptr(), None, None, 0, out.data_ptr(), mode, stream()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.isfinite(got).all(), "non-finite output"
    err = np.abs(got - ref.numpy()).max()
    assert err < 2e-3 * K ** 0.5, "max err %g" % err


@pytest.mark.parametrize("kind", ["kmajor", "mnmajor", "mnmajor_split", "swap", "wide", "wide_tma"])
@pytest.mark.parametrize("shape", [(256, 1024, 1024), (1000, 296, 640), (9600, 512, 256), (300, 256, 64), (256, 20480, 1024),
                                   (200, 20480 - 40, 128)])
def test_umma_gemm_cta_pair(shape, kind):
    """cta_group::2 kernel (two SMs, one M=256 UMMA, operand halves shared through the pair's shared memory), forced on
    for every shape through SRNN_GEMM_PAIR=2 in a fresh process so that the library's cached mode is not affected."""
    import subprocess, sys, os, textwrap
    M, N, K = shape
    code = textwrap.dedent(f"""
        import ctypes as C, numpy as np, torch, sys
        sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
        import srnn_b200 as S
        M, N, K = {M}, {N}, {K}
        g = torch.Generator().manual_seed(M + N + K)
        A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
        bias, add = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
        kind = {kind!r}
        plain = A.bfloat16().double() @ B.bfloat16().double().t()
        ref = (torch.relu(plain + bias.double() + add.double()) if kind in ("kmajor", "swap", "wide") else plain).float()
        if kind == "wide_tma":      # bias + ReLU only: the TMA-store epilogue
            ref = torch.relu(plain + bias.double()).float()
        out = torch.full((M, N), float("nan"), device="cuda")
        dA, dB, db, da = A.cuda(), B.cuda(), bias.cuda(), add.cuda()
        if kind in ("kmajor", "wide"):
            mode = S.MODE_BF16 | (128 << 8) | (256 << 16) | (1 << 28)
            args = (db.data_ptr(), da.data_ptr(), 1)
        elif kind == "swap":
            mode = S.MODE_BF16 | (128 << 8) | (256 << 16)
            args = (db.data_ptr(), da.data_ptr(), 1)
        elif kind == "wide_tma":
            mode = S.MODE_BF16 | (128 << 8) | (256 << 16) | (1 << 28)
            args = (db.data_ptr(), None, 1)
        else:
            mode = S.MODE_BF16 | (1 << 30) | ((1 << 29) if kind == "mnmajor_split" else 0)
            args = (None, None, 0)
        S._lib.check(S._lib.load().srnn_gemm(M, N, K, dA.data_ptr(), dB.data_ptr(), args[0], args[1], args[2],
                                             out.data_ptr(), mode, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        err = float((out.cpu() - ref).abs().max())
        assert np.isfinite(out.cpu().numpy()).all() and err < 2e-3 * K ** 0.5, err
        print("ok", err)
    """)
    env = dict(os.environ, SRNN_GEMM_PAIR="2")
    if N > 4096 and not kind.startswith("wide"):
        pytest.skip("upsampling-sized problems are the wide kernel's cases")
    if kind.startswith("wide"):         # one-wave (256 + 32)-feature pair tiles: at most 74 tiles
        if ((M + 255) // 256) * ((N + 287) // 288) > 74:
            pytest.skip("more tiles than CTA pairs")
        env["SRNN_GEMM_HOOK_WIDE_PAIR"] = "1"
    if kind == "swap":
        if M > 256:
            pytest.skip("swap-AB pair form takes at most 256 batch rows")
        env["SRNN_GEMM_HOOK_SWAP_PAIR"] = "1"
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_bf16_paths_with_cta_pair_forced():
    """Re-run the bf16 teacher-forcing / backward parity tests with SRNN_GEMM_PAIR=2, so that the CTA-pair kernel (bf16 outputs
    with ReLU / mask epilogues, MN-major weight gradients) serves every eligible GEMM at the small test sizes too."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SRNN_GEMM_PAIR="2")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(root, "tests", "test_gpu_parity.py"),
                        os.path.join(root, "tests", "test_gpu_training.py"), "-k",
                        "predict_bf16 or bf16_backward or other_architectures or mlp_fwd"],
                       env=env, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout, r.stdout[-500:]
