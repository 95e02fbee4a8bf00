"""GRU layer over a frame sequence through the C-ABI hooks srnn_gru_seq_fwd / srnn_gru_seq_bwd: the frame-by-frame fp32
schedule and the persistent tcgen05 recurrence (gru_persist.cu) against torch autograd on the nn.GRU equations
(model.py:133-159,244) in float64."""
import ctypes as C

import numpy as np
import pytest
import torch

import srnn_b200 as S

pytestmark = pytest.mark.gpu
L = S._lib


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def reference(gi, w_hh, b_hh, h0, dy, round_bf16):
    """gi (B,F,3H) -> y, gh, and the gradients of sum(y*dy) w.r.t. gi, gh, h0 (float64 autograd)."""
    B, F, H3 = gi.shape
    H = H3 // 3
    gi = gi.double().requires_grad_(True)
    h0 = h0.double().requires_grad_(True)
    w = (w_hh.bfloat16() if round_bf16 else w_hh).double()
    h, ys, ghs = h0, [], []
    for f in range(F):
        hin = h
        if round_bf16:                      # the recurrent operand is exchanged in bf16; the fp32 master state is kept
            hin = h + (h.detach().float().bfloat16().double() - h.detach())
        gh = hin @ w.t() + b_hh.double()
        gh.retain_grad()
        ghs.append(gh)
        r = torch.sigmoid(gi[:, f, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, f, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, f, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        ys.append(h)
    y = torch.stack(ys, 1)
    (y * dy.double()).sum().backward()
    return (y.detach(), torch.stack([g.detach() for g in ghs], 1), gi.grad, torch.stack([g.grad for g in ghs], 1), h0.grad)


@pytest.mark.parametrize("mode", [S.MODE_FP32, S.MODE_BF16])
@pytest.mark.parametrize("B,F,H", [(5, 7, 64), (128, 13, 128), (37, 52, 256), (128, 4, 1024), (1, 1, 192)])
def test_gru_sequence_forward_backward(B, F, H, mode):
    g = torch.Generator().manual_seed(B * 1000 + F * 10 + H)
    gi = torch.randn(B, F, 3 * H, generator=g)
    w_hh = torch.randn(3 * H, H, generator=g) / H ** 0.5
    b_hh = 0.1 * torch.randn(3 * H, generator=g)
    h0 = 0.5 * torch.randn(B, H, generator=g)
    dy = torch.randn(B, F, H, generator=g)
    bf16 = mode == S.MODE_BF16
    y_r, gh_r, dgi_r, dgh_r, dh0_r = reference(gi, w_hh, b_hh, h0, dy, bf16)

    d = lambda t: t.contiguous().cuda()
    gi_d, w_d, b_d, h0_d, dy_d = d(gi), d(w_hh), d(b_hh), d(h0), d(dy)
    y = torch.full((B, F, H), float("nan"), device="cuda")
    gh = torch.full((B, F, 3 * H), float("nan"), device="cuda")
    hl = torch.full((B, H), float("nan"), device="cuda")
    lib = L.load()
    L.check(lib.srnn_gru_seq_fwd(B, F, H, gi_d.data_ptr(), w_d.data_ptr(), b_d.data_ptr(), h0_d.data_ptr(), y.data_ptr(),
                                 gh.data_ptr(), hl.data_ptr(), mode, stream()))
    dgi = torch.full((B, F, 3 * H), float("nan"), device="cuda")
    dgh = torch.full((B, F, 3 * H), float("nan"), device="cuda")
    dh0 = torch.full((B, H), float("nan"), device="cuda")
    L.check(lib.srnn_gru_seq_bwd(B, F, H, gi_d.data_ptr(), gh.data_ptr(), y.data_ptr(), h0_d.data_ptr(), w_d.data_ptr(),
                                 dy_d.data_ptr(), dgi.data_ptr(), dgh.data_ptr(), dh0.data_ptr(), mode, stream()))
    torch.cuda.synchronize()
    # fp32: 1e-3-class parity; bf16: operand rounding of the recurrent GEMMs (weights rounded identically in the reference,
    # remaining difference = accumulation order + the bf16 rounding of dGH in the backward recurrence)
    tol_f = 2e-4 if not bf16 else 2e-3
    tol_b = 5e-4 if not bf16 else 3e-2
    for name, got, ref, tol in [("y", y, y_r, tol_f), ("gh", gh, gh_r, tol_f), ("h_last", hl, y_r[:, -1], tol_f),
                                ("dgi", dgi, dgi_r, tol_b), ("dgh", dgh, dgh_r, tol_b), ("dh0", dh0, dh0_r, tol_b)]:
        a = got.cpu().double().numpy()
        r = ref.numpy()
        assert np.isfinite(a).all(), name
        err = np.abs(a - r).max() / max(1.0, np.abs(r).max())
        assert err < tol, "%s: max err %g (tol %g)" % (name, err, tol)
