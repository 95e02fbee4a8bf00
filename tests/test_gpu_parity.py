"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle and the committed golden vectors."""
import ctypes as C

import numpy as np
import pytest
import torch

import srnn_b200 as S
from oracle import srnn_oracle as O

pytestmark = pytest.mark.gpu
L = S._lib


def build(golden_or_cfg, sd=None, mode=S.MODE_FP32):
    c = golden_or_cfg.c if hasattr(golden_or_cfg, "c") else golden_or_cfg
    m = S.SampleRNN(c["frame_sizes"], c["n_rnn"], c["dim"], c["learn_h0"], c["q_levels"], c["ulaw"],
                    c["weight_norm"], c["cond_dim"], c["spk_dim"])
    p = S.Predictor(m, mode=mode)
    p.load_state_dict(sd if sd is not None else golden_or_cfg.state_dict())
    p.cuda()
    return m, p


def logp_gate(got, ref, rel=1e-3):
    """north_star fp32 gate: |dlogp| <= 1e-3 * max(1, |ref|) element-wise."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    viol = np.abs(got - ref) > rel * np.maximum(1.0, np.abs(ref))
    assert not viol.any(), "max |d| %.3e (%d violations)" % (np.abs(got - ref).max(), int(viol.sum()))


def audio_to_index(audio, lut2):
    """Invert the (monotonic) dequantiser: the reference returns audio (model.py:520), parity is on the indices."""
    tab = np.asarray(lut2, np.float64) / 2
    return np.abs(np.asarray(audio, np.float64)[..., None] - tab).argmin(-1)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---- sampler in isolation: bit-exact on identical probabilities + uniforms --------------------------------------
def test_sampler_bit_exact():
    rng = np.random.default_rng(7)
    rows = 4096
    p = rng.random((rows, 256), dtype=np.float32)
    p[:512] = np.exp(8 * rng.standard_normal((512, 256))).astype(np.float32)       # peaky rows
    p[512:600] *= (rng.random((88, 256)) < 0.1)                                      # sparse rows with zeros
    p[600:700] = 0
    p[np.arange(600, 700), rng.integers(0, 256, 100)] = 1                            # one-hot rows
    p[700:800] = np.float32(1e-30) * rng.random((100, 256), dtype=np.float32)       # denormal-range mass
    u = rng.random(rows, dtype=np.float32)
    u[:16] = 0.0
    u[16:32] = np.float32(1) - np.float32(2 ** -24)
    dp, du = torch.from_numpy(p).cuda(), torch.from_numpy(u).cuda()
    idx = torch.empty(rows, dtype=torch.int32, device="cuda")
    L.check(L.load().srnn_sample_rows(dp.data_ptr(), du.data_ptr(), rows, idx.data_ptr(), stream()))
    np.testing.assert_array_equal(idx.cpu().numpy().astype(np.int64), O.sample_rows(p, u))


def test_gemm_hook_fp32():
    g = torch.Generator().manual_seed(0)
    for (M, N, K) in [(1, 7, 5), (3, 96, 32), (130, 70, 166), (256, 256, 1024)]:
        A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
        bias, add = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
        ref = torch.relu(A.double() @ B.double().t() + bias.double() + add.double()).float()
        dA, dB, db, da = A.cuda(), B.cuda(), bias.cuda(), add.cuda()
        out = torch.empty(M, N, device="cuda")
        L.check(L.load().srnn_gemm(M, N, K, dA.data_ptr(), dB.data_ptr(), db.data_ptr(), da.data_ptr(), 1,
                                   out.data_ptr(), S.MODE_FP32, stream()))
        np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=2e-4 * K ** 0.5, rtol=1e-5)


def test_dequant_lut(golden):
    m, p = build(golden)
    h = m._ensure_packed()
    out = torch.empty(256, device="cuda")
    L.check(L.load().srnn_dequant_lut(h, out.data_ptr(), stream()))
    np.testing.assert_allclose(out.cpu().numpy(), golden["lut"], atol=1e-6)


# ---- Predictor.forward against the reference's own outputs -------------------------------------------------------
def test_predict_golden_chunks_with_carry(golden):
    m, p = build(golden)
    spk = torch.from_numpy(golden["spk"])
    with torch.no_grad():
        for i in range(3):
            x, y, c = golden.chunk(i)
            logp = p(x, i == 0, c, spk, None, None)
            assert logp.is_cuda and tuple(logp.shape) == (x.shape[0], y.shape[1], 256)
            logp_gate(logp.cpu().numpy(), golden[f"tf/logp{i}"])
            assert abs(float(O.nll_bits(logp.cpu(), y)) - float(golden[f"tf/loss{i}"])) < 1e-4
            for t, rnn in enumerate(m.frame_level_rnns):
                np.testing.assert_allclose(p.hidden_states[rnn].cpu().numpy(), golden[f"tf/hidden{i}_{t}"], atol=2e-5)


def test_predict_matches_oracle_seeded_wider():
    torch.manual_seed(5)
    c = dict(frame_sizes=[20, 4], n_rnn=2, dim=96, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True,
             cond_dim=86, spk_dim=6)
    m = S.SampleRNN(**c)
    p = S.Predictor(m)
    with torch.no_grad():
        for k, v in p.state_dict().items():
            if "bias" in k or k.endswith("h0"):
                v.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in p.state_dict().items()}
    p.cuda()
    B, T = 5, 240
    x = torch.randint(0, 256, (B, 80 + T - 1))
    cond = torch.rand(B, T // 80, 86, dtype=torch.float64)
    spk = torch.randint(0, 6, (B, 1))
    w = O.unpack_state_dict(sd, O.Config(**c))
    with torch.no_grad():
        ref = O.Predictor(w).forward(x, True, cond, spk)
        got = p(x, True, cond, spk, None, None)
    logp_gate(got.cpu().numpy(), ref.numpy())


# ---- Generator ------------------------------------------------------------------------------------------------------
def test_generate_golden_shared_conditioner(golden):
    m, p = build(golden)
    gen = S.Generator(m, cuda=True)
    audio, samples, logp = gen(3, 0, golden["gen/cond"], int(golden["gen/spk"]), uniforms=golden["gen/uniforms"],
                               return_samples=True, return_logp=True)
    assert audio.device.type == "cpu" and audio.dtype == torch.float32
    logp_gate(logp.numpy(), golden["gen/logp"])
    # bit-exact sampled indices on the same uniforms (the device LUT's expf may differ from the CPU's by 1 ulp)
    np.testing.assert_array_equal(samples.numpy(), audio_to_index(golden["gen/audio"], golden["lut"]))
    np.testing.assert_allclose(audio.numpy(), golden["gen/audio"], atol=2e-7)


def test_generate_golden_batched_conditioner(golden):
    m, p = build(golden)
    B = golden["genb/cond"].shape[0]
    audio, samples = S.Generator(m, cuda=True)(B, 0, golden["genb/cond"], golden["genb/spk"],
                                               uniforms=golden["gen/uniforms"][:, :B], return_samples=True)
    np.testing.assert_array_equal(samples.numpy(), audio_to_index(golden["genb/audio"], golden["lut"]))
    np.testing.assert_allclose(audio.numpy(), golden["genb/audio"], atol=2e-7)


def test_teacher_forced_equals_autoregressive(golden):
    """Self-consistency: the log-probs the generator sampled from == Predictor on the generated sequence."""
    m, p = build(golden)
    cfg = golden.cfg()
    B, n_cond = 4, 3
    g = torch.Generator().manual_seed(11)
    cond = torch.rand(B, n_cond, cfg.cond_dim, generator=g)
    spk = torch.randint(0, cfg.spk_dim, (B,), generator=g)
    audio, samples, logp = S.Generator(m, cuda=True)(B, 0, cond, spk, seed=3, return_samples=True, return_logp=True)
    seq = torch.cat([torch.full((B, cfg.lookback), 128, dtype=torch.long), samples.long()], 1)
    with torch.no_grad():
        tf = p(seq[:, :-1], True, cond, spk.reshape(B, 1), None, None)
    np.testing.assert_allclose(tf.cpu().numpy(), logp.numpy(), atol=2e-4)


# ---- bf16 tensor-core mode (tcgen05): stated tolerance = 1.5x the reference's own bf16-autocast drift ----
BF16_MAX_ABS, BF16_MEAN_ABS = 0.08, 0.015


@pytest.mark.parametrize("mode", [S.MODE_BF16, S.MODE_BF16_GRAPH])
# (1024, 256) is BASELINE.json configs[1] itself: the default schedule of the headline run (persistent / cluster sample kernel,
# shadow GEMMs, programmatic launches, split input expansion all on) against the oracle, teacher-forced on the generated sequence
@pytest.mark.parametrize("dim,B", [(64, 5), (128, 37), (128, 256), (1024, 40), (1024, 256)])
def test_generate_bf16_mode_against_oracle(dim, B, mode):
    torch.manual_seed(dim)
    c = dict(frame_sizes=[20, 4], n_rnn=2, dim=dim, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True,
             cond_dim=86, spk_dim=6)
    m = S.SampleRNN(**c)
    p = S.Predictor(m)
    with torch.no_grad():
        for k, v in p.state_dict().items():
            if "bias" in k or k.endswith("h0"):
                v.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in p.state_dict().items()}
    p.cuda()
    n_cond = 2
    cond = torch.rand(B, n_cond, 86)
    spk = torch.randint(0, 6, (B,))
    audio, samples, logp = S.Generator(m, cuda=True, mode=mode)(B, 0, cond, spk, seed=5, return_samples=True,
                                                                return_logp=True)
    assert int(samples.min()) >= 0 and int(samples.max()) <= 255
    seq = torch.cat([torch.full((B, 80), 128, dtype=torch.long), samples.long()], 1)
    w = O.unpack_state_dict(sd, O.Config(**c))
    with torch.no_grad():
        ref = O.Predictor(w).forward(seq[:, :-1], True, cond, spk.reshape(B, 1))
    d = (ref - logp).abs()
    assert float(d.max()) <= BF16_MAX_ABS and float(d.mean()) <= BF16_MEAN_ABS, (float(d.max()), float(d.mean()))
    # and the fp32-mode run of the same generator on the same sequence stays inside the fp32 gate
    with torch.no_grad():
        tf32 = p(seq[:, :-1], True, cond, spk.reshape(B, 1), None, None)
    logp_gate(tf32.cpu().numpy(), ref.numpy())


@pytest.mark.parametrize("mode", [S.MODE_BF16, S.MODE_BF16_GRAPH])
@pytest.mark.parametrize("cfg", [
    dict(frame_sizes=[16], n_rnn=1, dim=128, learn_h0=True, q_levels=256, ulaw=True, weight_norm=False, cond_dim=43, spk_dim=6),   # C1 shape
    dict(frame_sizes=[4, 2, 2], n_rnn=1, dim=64, learn_h0=False, q_levels=256, ulaw=False, weight_norm=True, cond_dim=5, spk_dim=6),
    dict(frame_sizes=[20, 4], n_rnn=3, dim=256, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86, spk_dim=6),
    # the folded-input schedule (dim % 256 == 0) on a single-tier model (top tier only: C1 shape) and on one GRU layer per tier
    # (the folded upsampling then contracts the lite cell's own output)
    dict(frame_sizes=[16], n_rnn=1, dim=256, learn_h0=True, q_levels=256, ulaw=True, weight_norm=False, cond_dim=43, spk_dim=6),
    dict(frame_sizes=[20, 4], n_rnn=1, dim=256, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86, spk_dim=6),
])
def test_generate_bf16_other_architectures(cfg, mode):
    """Single-tier (C1), three-tier / linear-quantised / buffer-h0 and three-GRU-layer models through the tcgen05 generator."""
    torch.manual_seed(len(cfg["frame_sizes"]) * 10 + cfg["n_rnn"])
    m = S.SampleRNN(**cfg)
    p = S.Predictor(m)
    with torch.no_grad():
        for k, v in p.state_dict().items():
            if "bias" in k or k.endswith("h0"):
                v.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in p.state_dict().items()}
    p.cuda()
    B, n_cond = 9, 3
    cond = torch.rand(B, n_cond, cfg["cond_dim"])
    spk = torch.randint(0, 6, (B,))
    audio, samples, logp = S.Generator(m, cuda=True, mode=mode)(B, 0, cond, spk, seed=7, return_samples=True, return_logp=True)
    lb = m.lookback
    seq = torch.cat([torch.full((B, lb), 128, dtype=torch.long), samples.long()], 1)
    w = O.unpack_state_dict(sd, O.Config(**cfg))
    with torch.no_grad():
        ref = O.Predictor(w).forward(seq[:, :-1], True, cond, spk.reshape(B, 1))
    d = (ref - logp).abs()
    assert float(d.max()) <= BF16_MAX_ABS and float(d.mean()) <= BF16_MEAN_ABS, (float(d.max()), float(d.mean()))
    # teacher-forced tcgen05 forward on the same sequence
    pb = S.Predictor(m, mode=S.MODE_BF16)
    with torch.no_grad():
        tf = pb(seq[:, :-1], True, cond, spk.reshape(B, 1), None, None)
    d2 = (ref - tf.cpu()).abs()
    assert float(d2.max()) <= BF16_MAX_ABS and float(d2.mean()) <= BF16_MEAN_ABS, (float(d2.max()), float(d2.mean()))


def test_generate_large_batch_runs_in_independent_chunks():
    """B = 300 at dim 1024 exceeds what the persistent sample kernel keeps co-resident (288): srnn_generate then runs balanced
    utterance chunks (160 + 140) back to back.  Utterances are independent, so the result must be bit-identical to generating
    each block by its own call on the sliced inputs."""
    torch.manual_seed(3)
    c = dict(frame_sizes=[20, 4], n_rnn=2, dim=1024, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True,
             cond_dim=86, spk_dim=6)
    m = S.SampleRNN(**c).cuda()
    gen = S.Generator(m, cuda=True, mode=S.MODE_BF16)
    B, n_cond = 300, 1
    g = torch.Generator().manual_seed(9)
    cond, spk, uni = torch.rand(B, n_cond, 86, generator=g), torch.randint(0, 6, (B,), generator=g), torch.rand(80, B, generator=g)
    _, whole, lp = gen(B, 0, cond, spk, uniforms=uni, return_samples=True, return_logp=True)
    for lo, hi in ((0, 160), (160, 300)):
        _, part, lpp = gen(hi - lo, 0, cond[lo:hi], spk[lo:hi], uniforms=uni[:, lo:hi].contiguous(), return_samples=True,
                           return_logp=True)
        assert torch.equal(whole[lo:hi], part)
        assert torch.equal(lp[lo:hi], lpp)
    # shared-conditioner (reference) form through the chunked path
    _, shared = gen(B, 0, cond[0].numpy(), int(spk[0]), uniforms=uni, return_samples=True)
    _, one = gen(160, 0, cond[0].numpy(), int(spk[0]), uniforms=uni[:, :160].contiguous(), return_samples=True)
    assert torch.equal(shared[:160], one)


def test_generate_one_launch_per_contraction_tiles_at_large_batch():
    """MODE_BF16_GRAPH (one GEMM launch per contraction) picks its MLP tile from the batch size (128 x 64 at 1100 utterances,
    64 x 32 at 256): each output element is the same K-ordered tcgen05 accumulation whatever the tile, so the first 256
    utterances of a 1100-utterance call must equal a 256-utterance call bit for bit."""
    torch.manual_seed(6)
    c = dict(frame_sizes=[20, 4], n_rnn=2, dim=1024, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True,
             cond_dim=86, spk_dim=6)
    m = S.SampleRNN(**c).cuda()
    gen = S.Generator(m, cuda=True, mode=S.MODE_BF16_GRAPH)
    B, n_cond = 1100, 1
    g = torch.Generator().manual_seed(12)
    cond, spk, uni = torch.rand(B, n_cond, 86, generator=g), torch.randint(0, 6, (B,), generator=g), torch.rand(80, B, generator=g)
    _, whole, lp = gen(B, 0, cond, spk, uniforms=uni, return_samples=True, return_logp=True)
    _, part, lpp = gen(256, 0, cond[:256], spk[:256], uniforms=uni[:, :256].contiguous(), return_samples=True, return_logp=True)
    assert torch.equal(whole[:256], part)
    assert torch.equal(lp[:256], lpp)


def test_generate_schedule_variants_are_bit_identical(monkeypatch):
    """The generation schedule devices that only reorder / overlap launches -- recurrent projections in the shadow of the sample
    kernel, programmatic dependent launches of the tier kernels and of the sample kernel -- must leave samples and
    log-probabilities bit-identical when switched off (dim 1024 x 256 utterances is the configuration that enables them all;
    5 periods).  Two devices re-associate fp32 sums and are checked separately with the others held fixed:
    * the split top-tier input expansion (known columns accumulated beside the previous period's last sample launch): same
      log-probabilities to 3e-2 at the first sample of every period on utterances that have not diverged, > 80 % sample agreement;
    * the input expansion folded into the first GRU layer (gi_0 = G a + W_ih0 upper + b_gi0, different bf16 rounding points):
      trajectories part quickly, so it is gated where the histories still agree -- the first samples of the call -- and, against
      the oracle, by test_generate_bf16_mode_against_oracle[1024-256] (tools/fold_error.py: max |dlogp| 0.0213 / mean 0.0032
      with the fold, 0.0226 / 0.0033 without)."""
    torch.manual_seed(5)
    c = dict(frame_sizes=[20, 4], n_rnn=2, dim=1024, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True,
             cond_dim=86, spk_dim=6)
    m = S.SampleRNN(**c).cuda()
    gen = S.Generator(m, cuda=True, mode=S.MODE_BF16)
    B, n_cond = 256, 5
    g = torch.Generator().manual_seed(11)
    cond, spk = torch.rand(B, n_cond, 86, generator=g), torch.randint(0, 6, (B,), generator=g)
    uni = torch.rand(80 * n_cond, B, generator=g)
    switches = ("SRNN_NO_SHADOW_GH", "SRNN_NO_PDL", "SRNN_NO_SHADOW_IN", "SRNN_NO_GI_FOLD", "SRNN_NO_PDL_SAMPLE")
    for k in switches:
        monkeypatch.delenv(k, raising=False)
    run = lambda: gen(B, 0, cond, spk, uniforms=uni, return_samples=True, return_logp=True)[1:]
    fold, lp_fold = run()                                    # default schedule: everything on
    monkeypatch.setenv("SRNN_NO_GI_FOLD", "1")
    full, lpf = run()                                        # split input expansion, separate input kernels + cell GEMMs
    for k in ("SRNN_NO_PDL_SAMPLE",):
        monkeypatch.setenv(k, "1")
        out, lp2 = run()
        monkeypatch.delenv(k)
        assert torch.equal(full, out) and torch.equal(lpf, lp2), k
    monkeypatch.setenv("SRNN_NO_SHADOW_IN", "1")
    ref, lp = run()
    for k in ("SRNN_NO_SHADOW_GH", "SRNN_NO_PDL"):
        monkeypatch.setenv(k, "1")
        out, lp2 = run()
        monkeypatch.delenv(k)
        assert torch.equal(ref, out), k
        assert torch.equal(lp, lp2), k
    monkeypatch.delenv("SRNN_NO_SHADOW_IN")
    monkeypatch.delenv("SRNN_NO_GI_FOLD")
    # Split input expansion.  Measured: 92 % of all samples agree (a re-associated fp32 sum flips a few bf16 roundings of x,
    # which moves logits by ~1e-4..1e-3; with 7-bit-entropy distributions that flips a draw now and then, and the utterance
    # diverges from there).  A wrong conditioner frame or sample window in the shadow part would move every log-probability by O(0.1).
    agree = (full == ref).float().mean().item()
    assert agree > 0.8, agree
    worst = 0.0
    for p in range(n_cond):                      # first sample of each period, utterances still on the same trajectory
        same = (full[:, :80 * p] == ref[:, :80 * p]).all(dim=1) if p else torch.ones(B, dtype=torch.bool, device=full.device)
        assert same.float().mean().item() > 0.5, (p, same.float().mean().item())
        d = (lpf[:, 80 * p] - lp[:, 80 * p]).abs().amax(dim=-1)[same]
        worst = max(worst, float(d.max()))
    print("split input expansion: max |dlogp| at period starts on undiverged utterances = %.2e" % worst)
    assert worst < 0.03, worst
    # Folded first layer: log-probs of the first 20 samples on utterances whose histories still agree with the unfolded run
    same_hist = torch.cumprod(torch.cat([torch.ones_like(fold[:, :1], dtype=torch.long), (fold[:, :19] == full[:, :19]).long()], 1), 1).bool()
    d = (lp_fold[:, :20] - lpf[:, :20]).abs().amax(dim=-1)
    assert same_hist[:, 0].all() and float(same_hist.float().mean()) > 0.5
    print("folded first layer: max / mean |dlogp| over shared histories = %.3e / %.3e" % (float(d[same_hist].max()), float(d[same_hist].mean())))
    assert float(d[same_hist].max()) < 0.03 and float(d[same_hist].mean()) < 0.01     # measured 1.1e-2 / 6.1e-3


@pytest.mark.parametrize("dim,B,T", [(64, 3, 160), (128, 5, 240), (256, 130, 80)])
def test_predict_bf16_mode_against_oracle(dim, B, T):
    torch.manual_seed(dim + 1)
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
    x = torch.randint(0, 256, (B, 80 + 2 * T - 1))
    cond = torch.rand(B, 2 * T // 80, 86, dtype=torch.float64)
    spk = torch.randint(0, 6, (B, 1))
    w = O.unpack_state_dict(sd, O.Config(**c))
    ref_p = O.Predictor(w)
    with torch.no_grad():
        for i in range(2):                          # two consecutive chunks: the bf16 path carries hidden state too
            xs = x[:, i * T: i * T + 80 + T - 1].contiguous()
            cs = cond[:, i * (T // 80): (i + 1) * (T // 80)].contiguous()
            ref = ref_p.forward(xs, i == 0, cs, spk)
            got = p(xs, i == 0, cs, spk, None, None).cpu()
            d = (ref - got).abs()
            assert float(d.max()) <= BF16_MAX_ABS and float(d.mean()) <= BF16_MEAN_ABS, (i, float(d.max()), float(d.mean()))


def test_nll_loss_bits_hook(golden):
    m, p = build(golden)
    x, y, c = golden.chunk(0)
    spk = torch.from_numpy(golden["spk"])
    with torch.no_grad():
        logp = p(x, True, c, spk, None, None)
    loss = torch.empty(1, device="cuda")
    tgt = y.cuda().contiguous()
    L.check(L.load().srnn_nll_loss_bits(m._ensure_packed(), logp.data_ptr(), tgt.data_ptr(), int(tgt.numel()),
                                        loss.data_ptr(), stream()))
    assert abs(float(loss.item()) - float(golden["tf/loss0"])) < 1e-4


@pytest.mark.parametrize("mode", [S.MODE_FP32, S.MODE_BF16])
def test_mlp_fwd_hook(mode):
    """srnn_mlp_fwd = SampleLevelMLP.forward (model.py:308-325) against the oracle's dense embedding + k=FS conv form."""
    import ctypes as C
    torch.manual_seed(4)
    c = dict(frame_sizes=[20, 4], n_rnn=1, dim=128, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86,
             spk_dim=6)
    m = S.SampleRNN(**c)
    p = S.Predictor(m)
    with torch.no_grad():
        for k, v in p.state_dict().items():
            if "bias" in k:
                v.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in p.state_dict().items()}
    p.cuda()
    h = m._ensure_packed()
    B, T = 3, 50
    prev = torch.randint(0, 256, (B, T + 19))
    upper = torch.randn(B, T, 128)
    w = O.unpack_state_dict(sd, O.Config(**c))
    with torch.no_grad():
        ref = O.mlp_forward(w, prev, upper)
    out = torch.full((B, T, 256), float("nan"), device="cuda")
    pd, ud = prev.cuda(), upper.cuda().contiguous()
    S._lib.check(S._lib.load().srnn_mlp_fwd(h, B, T, pd.data_ptr(), ud.data_ptr(), out.data_ptr(), mode,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    got = out.cpu()
    if mode == S.MODE_FP32:
        logp_gate(got.numpy(), ref.numpy())
    else:
        d = (got - ref).abs()
        assert float(d.max()) <= BF16_MAX_ABS and float(d.mean()) <= BF16_MEAN_ABS, (float(d.max()), float(d.mean()))


# ---- tensor-core parity mode (SRNN_MODE_BF16X3): the fp32 gates on tcgen05 ------------------------------------------
def test_gemm_hook_split_bf16():
    """W.x as Wh.xh + Wl.xh + Wh.xl in one tcgen05 GEMM over K' = 3K: error ~2^-16 per product (plain bf16: 2^-8)."""
    g = torch.Generator().manual_seed(1)
    for (M, N, K) in [(5, 128, 64), (130, 256, 1024), (256, 1024, 1024), (300, 3072, 1024), (2000, 256, 1024)]:
        A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
        bias = torch.randn(N, generator=g)
        ref = torch.relu(A.double() @ B.double().t() + bias.double())
        dA, dB, db = A.cuda(), B.cuda(), bias.cuda()
        out = torch.empty(M, N, device="cuda")
        L.check(L.load().srnn_gemm(M, N, K, dA.data_ptr(), dB.data_ptr(), db.data_ptr(), None, 1, out.data_ptr(),
                                   S.MODE_BF16X3, stream()))
        err = (out.cpu().double() - ref).abs().max().item()
        assert err <= 6e-5 * K ** 0.5, (M, N, K, err)
        # the plain bf16 product of the same operands is two orders of magnitude further away
        out16 = torch.empty(M, N, device="cuda")
        L.check(L.load().srnn_gemm(M, N, K, dA.data_ptr(), dB.data_ptr(), db.data_ptr(), None, 1, out16.data_ptr(),
                                   S.MODE_BF16, stream()))
        assert (out16.cpu().double() - ref).abs().max().item() > 20 * err


def _seeded(c, seed):
    torch.manual_seed(seed)
    m = S.SampleRNN(**c)
    p = S.Predictor(m, mode=S.MODE_BF16X3)
    with torch.no_grad():
        for k, v in p.state_dict().items():
            if "bias" in k or k.endswith("h0"):
                v.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in p.state_dict().items()}
    p.cuda()
    return m, p, O.unpack_state_dict(sd, O.Config(**c))


@pytest.mark.parametrize("cfg,B,T", [
    (dict(frame_sizes=[20, 4], n_rnn=2, dim=128, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86, spk_dim=6), 5, 240),
    (dict(frame_sizes=[20, 4], n_rnn=2, dim=1024, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86, spk_dim=6), 8, 160),
    (dict(frame_sizes=[16], n_rnn=1, dim=64, learn_h0=True, q_levels=256, ulaw=True, weight_norm=False, cond_dim=43, spk_dim=6), 3, 64),
    (dict(frame_sizes=[4, 2, 2], n_rnn=1, dim=64, learn_h0=False, q_levels=256, ulaw=False, weight_norm=True, cond_dim=5, spk_dim=6), 7, 64),
])
def test_predict_split_bf16_mode_holds_the_fp32_gate(cfg, B, T):
    """Teacher-forced log-probs of SRNN_MODE_BF16X3 against the oracle: the 1e-3 relative gate of the fp32 mode; hidden-state
    carry over two chunks; and a backward pass after it (fp32 kernels on the saved activations) matches the fp32 mode's."""
    m, p, w = _seeded(cfg, 5)
    lb = m.lookback
    x = torch.randint(0, 256, (B, lb + 2 * T - 1))
    cond = torch.rand(B, 2 * T // lb, cfg["cond_dim"], dtype=torch.float64)
    spk = torch.randint(0, 6, (B, 1))
    ref_p = O.Predictor(w)
    with torch.no_grad():
        for i in range(2):
            xi, ci = x[:, i * T: i * T + lb + T - 1], cond[:, i * T // lb: (i + 1) * T // lb]
            ref = ref_p.forward(xi, i == 0, ci, spk)
            got = p(xi, i == 0, ci, spk, None, None)
            logp_gate(got.cpu().numpy(), ref.numpy())
    p32 = S.Predictor(m, mode=S.MODE_FP32)
    grads = []
    for pred in (p, p32):
        for q in m.parameters():
            q.grad = None
        out = pred(x[:, :lb + T - 1], True, cond[:, :T // lb], spk, None, None)
        S.sequence_nll_loss_bits(out, x[:, lb:lb + T].cuda()).backward()
        grads.append([q.grad.detach().clone() for q in m.parameters() if q.grad is not None])
    assert len(grads[0]) == len(grads[1]) > 0
    for a, b in zip(*grads):
        # a unit within ~1e-5 of zero may land on the other side of its ReLU: isolated elements move, the tensor does not
        assert float((a - b).norm()) <= 1e-7 + 3e-3 * float(b.norm())
        assert float((a - b).abs().max()) <= 1e-5 + 0.1 * float(b.abs().max())


@pytest.mark.parametrize("dim,B", [(128, 37), (1024, 40), (1024, 256)])
def test_generate_split_bf16_mode_against_oracle(dim, B):
    """SRNN_MODE_BF16X3 generation: log-probs within the fp32 gate (1e-3 relative) of the oracle teacher-forced on the generated
    sequence, and the sampled indices equal to the fp32 mode's on the same uniforms up to isolated CDF-boundary ties."""
    c = dict(frame_sizes=[20, 4], n_rnn=2, dim=dim, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True,
             cond_dim=86, spk_dim=6)
    m, p, w = _seeded(c, dim)
    n_cond = 2
    g = torch.Generator().manual_seed(4)
    cond, spk, uni = torch.rand(B, n_cond, 86, generator=g), torch.randint(0, 6, (B,), generator=g), torch.rand(n_cond * 80, B, generator=g)
    _, samples, logp = S.Generator(m, cuda=True, mode=S.MODE_BF16X3)(B, 0, cond, spk, uniforms=uni, return_samples=True,
                                                                      return_logp=True)
    seq = torch.cat([torch.full((B, 80), 128, dtype=torch.long), samples.long()], 1)
    with torch.no_grad():
        ref = O.Predictor(w).forward(seq[:, :-1], True, cond, spk.reshape(B, 1))
    logp_gate(logp.numpy(), ref.numpy())
    _, s32 = S.Generator(m, cuda=True, mode=S.MODE_FP32)(B, 0, cond, spk, uniforms=uni, return_samples=True)
    same = (s32 == samples)
    shared = torch.cumprod(torch.cat([torch.ones_like(same[:, :1]), same[:, :-1]], 1).long(), 1).bool()
    # while the two runs share a history they may differ only where u falls within rounding of a CDF boundary
    assert float(same[shared].float().mean()) >= 0.999, float(same[shared].float().mean())


@pytest.mark.parametrize("mode,dim,B", [(S.MODE_BF16, 1024, 24), (S.MODE_FP32, 64, 5)])
def test_generate_reuses_the_instantiated_graph(mode, dim, B, monkeypatch):
    """A second srnn_generate call of the same shape on the same buffers replays the cached graph (no capture, no
    instantiation); new input VALUES in those buffers are picked up, and the result equals an uncached call bit for bit."""
    torch.manual_seed(2)
    c = dict(frame_sizes=[20, 4], n_rnn=2, dim=dim, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86, spk_dim=6)
    m = S.SampleRNN(**c).cuda()
    gen = S.Generator(m, cuda=True, mode=mode)
    n_cond = 3
    g = torch.Generator().manual_seed(8)
    cond = torch.rand(B, n_cond, 86, generator=g).cuda()
    spk = torch.randint(0, 6, (B,), generator=g).cuda()
    uni = torch.rand(80 * n_cond, B, generator=g).cuda()
    monkeypatch.delenv("SRNN_NO_GRAPH_CACHE", raising=False)
    reuse = lambda: L.load().srnn_graph_reuse_count(m._ctx)
    outs = []
    r0 = None
    for i in range(4):
        if i == 2:
            uni.copy_(torch.rand(80 * n_cond, B, generator=g))          # same buffer, new values
        res = gen(B, 0, cond, spk, uniforms=uni, device_output=True, return_samples=True, return_logp=True)
        outs.append([t.clone() for t in res])
        del res                                                          # output blocks go back to the caching allocator
        if i == 0:
            r0 = reuse()
    assert reuse() - r0 >= 2, "the generation graph was re-captured on every call"
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))
    assert all(torch.equal(a, b) for a, b in zip(outs[2], outs[3]))
    assert not torch.equal(outs[0][1], outs[2][1])
    monkeypatch.setenv("SRNN_NO_GRAPH_CACHE", "1")
    before = reuse()
    res = gen(B, 0, cond, spk, uniforms=uni, device_output=True, return_samples=True, return_logp=True)
    assert reuse() == before
    assert all(torch.equal(a, b) for a, b in zip(outs[3], res))


def test_teacher_forced_table_gather_from_shared_memory_is_bit_identical(monkeypatch):
    """Above 32 768 tokens the bf16 teacher-forced pass gathers from 16-feature table slices held in shared memory
    (k_mlp_gather_slice) instead of re-reading table rows from L2 (k_mlp_gather_bf16v, SRNN_GATHER_V1=1).  Both sum the
    conditioning and the taps in the same order, so every log-probability must be bit-identical."""
    torch.manual_seed(31)
    c = dict(frame_sizes=[20, 4], n_rnn=1, dim=128, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86, spk_dim=6)
    m = S.SampleRNN(**c).cuda()
    p = S.Predictor(m, mode=S.MODE_BF16)
    B, T = 33, 1040
    x = torch.randint(0, 256, (B, 80 + T - 1))
    cond, spk = torch.rand(B, T // 80, 86), torch.randint(0, 6, (B, 1))
    outs = []
    for v1 in (False, True):
        if v1:
            monkeypatch.setenv("SRNN_GATHER_V1", "1")
        else:
            monkeypatch.delenv("SRNN_GATHER_V1", raising=False)
        with torch.no_grad():
            outs.append(p(x, True, cond, spk, None, None).clone())
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("mode", [S.MODE_FP32, S.MODE_BF16X3])
@pytest.mark.parametrize("cfg", [
    dict(frame_sizes=[20, 4], n_rnn=2, dim=64, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86, spk_dim=6),
    dict(frame_sizes=[4, 2, 2], n_rnn=1, dim=64, learn_h0=False, q_levels=256, ulaw=False, weight_norm=True, cond_dim=5, spk_dim=6),
])
def test_per_module_forward_api_composes_to_predictor(cfg, mode):
    """The reference's per-module calls -- Runner.run_rnn over FrameLevelRNN.forward, then SampleLevelMLP.forward, in the order
    and with the tensor shapes of Predictor.forward (model.py:357-436) -- reproduce the fused srnn_predict_fwd and the oracle,
    including the hidden-state carry over two chunks."""
    m, p, w = _seeded(cfg, 9)
    p.mode = S.MODE_FP32
    m.module_mode = mode
    lb = m.lookback
    B, T = 3, 2 * lb
    x = torch.randint(0, 256, (B, lb + 2 * T - 1))
    cond = torch.rand(B, 2 * T // lb, cfg["cond_dim"])
    spk = torch.randint(0, 6, (B, 1))
    runner = S.Runner(m)
    ref_p = O.Predictor(w)
    with torch.no_grad():
        for i in range(2):
            xi, ci = x[:, i * T: i * T + lb + T - 1], cond[:, i * T // lb: (i + 1) * T // lb]
            if i == 0:
                runner.reset_hidden_states()
            upper = None
            for rnn in reversed(m.frame_level_rnns):
                n = rnn.n_frame_samples
                win = xi[:, lb - n: xi.shape[1] - n + 1]
                prev = (2 * m.dequantize(win, m.q_levels)).reshape(B, -1, n)
                upper = runner.run_rnn(rnn, prev, upper, ci if upper is None else None, spk if upper is None else None)
            fs0 = m.frame_level_rnns[0].frame_size
            got = m.sample_level_mlp(xi[:, lb - fs0:], upper)
            fused = p(xi, i == 0, ci.double(), spk, None, None)
            ref = ref_p.forward(xi, i == 0, ci.double(), spk)
            logp_gate(got.cpu().numpy(), ref.numpy())
            np.testing.assert_allclose(got.cpu().numpy(), fused.cpu().numpy(), atol=2e-4)
            for rnn in m.frame_level_rnns:
                np.testing.assert_allclose(runner.hidden_states[rnn].cpu().numpy(), p.hidden_states[rnn].cpu().numpy(), atol=1e-4)
