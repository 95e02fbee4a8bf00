"""north_star equivalence #3: "loss curves overlay over the first 1k training steps".

tests/golden/loss_curve.npz holds the loss of the UNMODIFIED reference (trainer closure + gradient_clipping(Adam), lr 1e-3)
over 1000 consecutive TBPTT steps on a fixed synthetic stream (make_loss_curve.py).  The CPU test pins the oracle on the
first steps; the GPU tests run all 1000 steps through Predictor / ClampAdam (C-ABI kernels) in both arithmetic modes."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import loss_curve_inputs as I     # noqa: E402

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_curve.npz")


def load():
    z = np.load(FIX)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    data = torch.from_numpy(z["data"].astype(np.int64))
    return sd, data, torch.from_numpy(I.conditioners()), torch.from_numpy(I.speakers()), z["losses"]


def smooth(x, k=25):
    return np.convolve(x, np.ones(k) / k, mode="valid")


def test_fixture_shape_and_learning_signal():
    sd, data, cond, spk, ref = load()
    assert ref.shape == (I.STEPS,) and data.shape == (I.B, I.LOOKBACK + I.STEPS * I.T)
    assert ref[0] > 8.0 and smooth(ref)[-1] < 4.5          # the stream is learnable: the curve has a shape to overlay


def test_oracle_overlays_reference_first_steps():
    from oracle import srnn_oracle as O
    sd, data, cond, spk, ref = load()
    cfg = O.Config(**I.CONFIG)
    st, hidden = O.AdamState(), None
    for i in range(12):
        x, y, c = I.chunk(data, cond, i)
        loss, grads, hidden, _ = O.loss_and_grads(sd, cfg, hidden, x, i == 0, c, spk, y)
        assert abs(float(loss) - ref[i]) < 2e-3 * max(1, i), (i, float(loss), ref[i])
        sd = O.clamp_adam_step(sd, grads, st, lr=I.LR)


@pytest.mark.gpu
@pytest.mark.parametrize("mode_name", ["fp32", "bf16"])
def test_gpu_loss_curve_overlays_reference(mode_name):
    import srnn_b200 as S
    mode = S.MODE_FP32 if mode_name == "fp32" else S.MODE_BF16
    sd, data, cond, spk, ref = load()
    m = S.SampleRNN(**I.CONFIG)
    p = S.Predictor(m, mode=mode)
    p.load_state_dict(sd)
    p.cuda()
    opt = S.ClampAdam(p.parameters(), lr=I.LR, model=m)
    data_d, cond_d, spk_d = data.cuda(), cond.cuda(), spk.cuda()
    losses = []
    for i in range(I.STEPS):
        x, y, c = I.chunk(data_d, cond_d, i)
        x, y, c = x.contiguous(), y.contiguous(), c.contiguous()

        def closure():
            out = p(x, i == 0, c, spk_d, None, None)
            loss = S.sequence_nll_loss_bits(out, y)
            loss.backward()
            return loss.detach()

        opt.zero_grad()
        losses.append(opt.step(closure))
    got = torch.stack(losses).double().cpu().numpy()
    assert np.isfinite(got).all()
    d = np.abs(got - ref)
    ds = np.abs(smooth(got) - smooth(ref))
    if mode_name == "fp32":
        # identical arithmetic up to summation order: the first steps agree to ~1e-4 bits; later the two fp32 trajectories
        # separate slowly (Adam sign flips at |g| ~ eps), which shows as per-step jitter but not in the smoothed curve
        assert d[:20].max() < 2e-3, d[:20].max()
        assert d.max() < 0.25 and ds.max() < 0.06, (d.max(), ds.max())
    else:
        # bf16 tensor-core arithmetic: per-step loss within the bf16 logit tolerance, smoothed curves overlay
        assert d[:20].max() < 0.05, d[:20].max()
        assert d.max() < 0.5 and ds.max() < 0.12, (d.max(), ds.max())
    assert abs(smooth(got)[-1] - smooth(ref)[-1]) < 0.1
    print("loss-curve overlay %s: max|d|=%.4f  max smoothed |d|=%.4f  final %.3f vs %.3f bits" % (
        mode_name, d.max(), ds.max(), smooth(got)[-1], smooth(ref)[-1]))
