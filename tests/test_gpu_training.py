"""Teacher-forced training step on the GPU (forward, loss, backward, fused clamp+Adam) against the reference's own
gradients / updated parameters (golden fixtures) and the oracle."""
import numpy as np
import pytest
import torch

import srnn_b200 as S
from oracle import srnn_oracle as O

pytestmark = pytest.mark.gpu


def build(golden):
    c = golden.c
    m = S.SampleRNN(c["frame_sizes"], c["n_rnn"], c["dim"], c["learn_h0"], c["q_levels"], c["ulaw"], c["weight_norm"],
                    c["cond_dim"], c["spk_dim"])
    p = S.Predictor(m)
    p.load_state_dict(golden.state_dict())
    p.cuda()
    return m, p


def test_backward_matches_reference_gradients(golden):
    m, p = build(golden)
    x, y, c = golden.chunk(0)
    spk = torch.from_numpy(golden["spk"])
    logp = p(x, True, c, spk, None, None)
    assert logp.requires_grad
    loss = S.sequence_nll_loss_bits(logp, y)
    loss.backward()
    assert abs(float(loss) - float(golden["train/loss0"])) < 1e-4
    for k, q in p.named_parameters():
        ref = golden["train/grad0/" + k]
        got = q.grad.detach().cpu().numpy()
        np.testing.assert_allclose(got, ref, atol=3e-6 + 2e-4 * np.abs(ref).max(), err_msg=k)


def test_three_training_steps_match_reference(golden):
    m, p = build(golden)
    spk = torch.from_numpy(golden["spk"])
    opt = S.ClampAdam(p.parameters(), lr=1e-3, model=m)
    for i in range(3):
        x, y, c = golden.chunk(i)

        def closure():
            out = p(x, i == 0, c, spk, None, None)
            loss = S.sequence_nll_loss_bits(out, y)
            loss.backward()
            return loss

        opt.zero_grad()
        loss = opt.step(closure)
        assert abs(float(loss) - float(golden[f"train/loss{i}"])) < 3e-4, i
    sd = p.state_dict()
    for k in sd:
        d = np.abs(sd[k].cpu().numpy() - golden["train/sd3/" + k])
        assert d.max() <= 3.1e-3, k              # a few Adam elements with |grad| ~ eps may flip (see oracle test)
        assert (d > 1e-4).mean() <= 2e-3, k


def test_clamp_adam_against_oracle_formula():
    torch.manual_seed(0)
    ps = [torch.randn(n, device="cuda").requires_grad_(True) for n in (5, 4097, 70000)]
    ref = {str(i): q.detach().cpu().clone() for i, q in enumerate(ps)}
    opt = S.ClampAdam(ps, lr=1e-2)
    st = O.AdamState()
    for step in range(4):
        grads = {}
        for i, q in enumerate(ps):
            g = 3 * torch.randn_like(q)           # |g| > 1 exercises the clamp
            q.grad = g
            grads[str(i)] = g.cpu()
        opt.step()
        ref = O.clamp_adam_step(ref, grads, st, lr=1e-2)
    for i, q in enumerate(ps):
        np.testing.assert_allclose(q.detach().cpu().numpy(), ref[str(i)].numpy(), atol=2e-6)


# B = 100: both 64-row halves of the persistent GRU kernels; B = 130: beyond them, the frame-by-frame fallback schedule
# (1024, 8, 1040) and (1024, 72, 1040) are BASELINE.json configs[2] (C3) at its own width and sequence length: 13 / 52-frame
# persistent GRU launches, the CTA-pair GEMMs and the 20-tap table gradient at H = 1024; B = 72 runs both 64-row halves
@pytest.mark.parametrize("dim,B,T", [(64, 3, 160), (128, 5, 80), (64, 100, 80), (64, 130, 80), (1024, 8, 1040), (1024, 72, 1040)])
def test_bf16_backward_against_oracle(dim, B, T):
    """tcgen05 training path: gradients within bf16 accuracy of the fp32 oracle (relative L2 per tensor)."""
    torch.manual_seed(dim + 7)
    c = dict(frame_sizes=[20, 4], n_rnn=2, dim=dim, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True,
             cond_dim=86, spk_dim=6)
    m = S.SampleRNN(**c)
    p = S.Predictor(m, mode=S.MODE_BF16)
    with torch.no_grad():
        for k, v in p.state_dict().items():
            if "bias" in k or k.endswith("h0"):
                v.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in p.state_dict().items()}
    p.cuda()
    x = torch.randint(0, 256, (B, 80 + T - 1))
    y = torch.randint(0, 256, (B, T))
    cond = torch.rand(B, T // 80, 86, dtype=torch.float64)
    spk = torch.randint(0, 6, (B, 1))
    loss_ref, grads, _, _ = O.loss_and_grads(sd, O.Config(**c), None, x, True, cond, spk, y)
    out = p(x, True, cond, spk, None, None)
    loss = S.sequence_nll_loss_bits(out, y)
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_ref)) < 0.02
    worst = 0.0
    report = []
    for k, q in p.named_parameters():
        ref = grads[k].numpy().astype(np.float64)
        got = q.grad.detach().cpu().numpy().astype(np.float64)
        assert np.isfinite(got).all(), k
        nref = np.linalg.norm(ref)
        if nref < 1e-7:
            continue
        rel = np.linalg.norm(got - ref) / nref
        worst = max(worst, rel)
        report.append((round(float(rel), 4), k))
    assert worst > 0          # something was actually compared
    # Output-layer gradients involve no ReLU mask: tight.  Everything behind a ReLU inherits the mask flips of units whose
    # pre-activation is within bf16 rounding of zero (sqrt(fraction flipped) ~ 5 % at these widths; the fp32 path of
    # the same code is exact, see test_backward_matches_reference_gradients), so those are gated loosely.
    rep = dict((k, r) for r, k in report)
    assert rep["model.sample_level_mlp.output.weight_v"] < 0.01 and rep["model.sample_level_mlp.output.bias"] < 0.01, report
    assert worst < 0.12, sorted(report, reverse=True)[:12]
    assert float(np.median([r for r, _ in report])) < 0.08, sorted(report, reverse=True)[:12]


@pytest.mark.parametrize("cfg", [
    dict(frame_sizes=[16], n_rnn=1, dim=128, learn_h0=True, q_levels=256, ulaw=True, weight_norm=False, cond_dim=43, spk_dim=6),   # C1 shape
    dict(frame_sizes=[4, 2, 2], n_rnn=1, dim=64, learn_h0=False, q_levels=256, ulaw=False, weight_norm=True, cond_dim=5, spk_dim=6),
    dict(frame_sizes=[20, 4], n_rnn=3, dim=256, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86, spk_dim=6),
])
def test_bf16_backward_other_architectures(cfg):
    """tcgen05 training path on single-tier / three-tier / three-layer models: loss and gradients against the fp32 oracle."""
    torch.manual_seed(17)
    m = S.SampleRNN(**cfg)
    p = S.Predictor(m, mode=S.MODE_BF16)
    with torch.no_grad():
        for k, v in p.state_dict().items():
            if "bias" in k or k.endswith("h0"):
                v.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in p.state_dict().items()}
    p.cuda()
    lb = m.lookback
    B, T = 6, 2 * lb
    x = torch.randint(0, 256, (B, lb + T - 1))
    y = torch.randint(0, 256, (B, T))
    cond = torch.rand(B, T // lb, cfg["cond_dim"], dtype=torch.float64)
    spk = torch.randint(0, 6, (B, 1))
    loss_ref, grads, _, _ = O.loss_and_grads(sd, O.Config(**cfg), None, x, True, cond, spk, y)
    out = p(x, True, cond, spk, None, None)
    loss = S.sequence_nll_loss_bits(out, y)
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_ref)) < 0.02
    rels = []
    for k, q in p.named_parameters():
        ref = grads[k].numpy().astype(np.float64)
        got = q.grad.detach().cpu().numpy().astype(np.float64)
        assert np.isfinite(got).all(), k
        n = np.linalg.norm(ref)
        if n > 1e-7:
            rels.append((np.linalg.norm(got - ref) / n, k))
    assert rels and max(rels)[0] < 0.15, sorted(rels, reverse=True)[:8]
    assert float(np.median([r for r, _ in rels])) < 0.08, sorted(rels, reverse=True)[:8]


@pytest.mark.parametrize("dim,B,T", [(128, 5, 160), (1024, 8, 160)])
def test_bf16_gradient_error_is_relu_mask_flips(dim, B, T):
    """The loose per-tensor gate of test_bf16_backward_against_oracle (0.12) is explained by ReLU-mask flips of units whose
    pre-activation lies within bf16 rounding of zero.  Asserted here rather than assumed: with both MLP ReLUs held open
    (large positive tier-0 upsampling bias and hidden bias: every pre-activation far from zero, masks all one) the same
    kernels on the same shapes must agree with the fp32 oracle to plain bf16 operand accuracy on EVERY tensor."""
    torch.manual_seed(dim + 3)
    c = dict(frame_sizes=[20, 4], n_rnn=2, dim=dim, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True,
             cond_dim=86, spk_dim=6)
    m = S.SampleRNN(**c)
    p = S.Predictor(m, mode=S.MODE_BF16)
    with torch.no_grad():
        for k, v in p.state_dict().items():
            if "bias" in k or k.endswith("h0"):
                v.normal_(0, 0.1)
        sd0 = p.state_dict()
        sd0["model.frame_level_rnns.0.upsampling.bias"].add_(6.0)
        sd0["model.sample_level_mlp.hidden.bias"].add_(6.0)
    sd = {k: v.clone() for k, v in p.state_dict().items()}
    p.cuda()
    x = torch.randint(0, 256, (B, 80 + T - 1))
    y = torch.randint(0, 256, (B, T))
    cond = torch.rand(B, T // 80, 86, dtype=torch.float64)
    spk = torch.randint(0, 6, (B, 1))
    loss_ref, grads, _, _ = O.loss_and_grads(sd, O.Config(**c), None, x, True, cond, spk, y)
    out = p(x, True, cond, spk, None, None)
    loss = S.sequence_nll_loss_bits(out, y)
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_ref)) < 0.02 + 2e-3 * float(loss_ref)
    report = []
    for k, q in p.named_parameters():
        ref = grads[k].numpy().astype(np.float64)
        got = q.grad.detach().cpu().numpy().astype(np.float64)
        nref = np.linalg.norm(ref)
        if nref < 1e-7:
            continue
        report.append((round(float(np.linalg.norm(got - ref) / nref), 4), k))
    print("no-mask-flip gradient errors:", sorted(report, reverse=True)[:6])
    assert max(report)[0] < 0.05, sorted(report, reverse=True)[:12]       # measured 0.033 (dim 128); regular gate 0.12


def test_table_foldback_on_tensor_cores_matches_fp32_form(monkeypatch):
    """bf16 training path: the fold-back of dTbl onto mlp.input (H,Q,FS) and the embedding (Q,Q) runs as split-bf16 tcgen05
    products; with SRNN_FOLDBACK_F32=1 the same step uses the FFMA GEMMs.  Same inputs to both, so the two gradients agree
    to the split product's accuracy (~2^-16 relative per term)."""
    torch.manual_seed(21)
    c = dict(frame_sizes=[20, 4], n_rnn=2, dim=128, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True,
             cond_dim=86, spk_dim=6)
    m = S.SampleRNN(**c).cuda()
    p = S.Predictor(m, mode=S.MODE_BF16)
    B, T = 6, 160
    x, y = torch.randint(0, 256, (B, 80 + T - 1)), torch.randint(0, 256, (B, T))
    cond, spk = torch.rand(B, T // 80, 86), torch.randint(0, 6, (B, 1))
    res = []
    for f32 in (False, True):
        if f32:
            monkeypatch.setenv("SRNN_FOLDBACK_F32", "1")
        else:
            monkeypatch.delenv("SRNN_FOLDBACK_F32", raising=False)
        for q in p.parameters():
            q.grad = None
        S.sequence_nll_loss_bits(p(x, True, cond, spk, None, None), y).backward()
        res.append({k: q.grad.detach().clone() for k, q in p.named_parameters() if "sample_level_mlp.input" in k or "embedding" in k})
    assert len(res[0]) >= 2
    for k in res[0]:
        a, b = res[0][k], res[1][k]
        assert float((a - b).norm()) <= 2e-4 * float(b.norm()) + 1e-9, (k, float((a - b).norm()), float(b.norm()))
