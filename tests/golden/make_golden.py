#!/usr/bin/env python
"""Generate the golden fixtures by EXECUTING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*.npz

The reference (`/root/reference/model.py`, `nn.py`, `utils.py`, `optim.py`) is imported as is.
Work-arounds for its quirks are applied from the OUTSIDE:
  * `Predictor.forward` only runs at B=1 (model.py:209) -> every batch row is a separate call with
    its own `hidden_states`; the batch loss is the mean of the row losses (same T per row).
  * it writes `<spk>.txt` into cwd (model.py:210-214) -> cwd is a temp dir; stdout is silenced.
  * `Tensor.multinomial` (model.py:517) is replaced during generation by the defined sampler
    (`oracle.srnn_oracle.sample_rows`) fed with pre-drawn uniforms u[t, b].
  * `optimizer.zero_grad(set_to_none=False)` reproduces the torch-0.4 semantics the reference's
    `optim.py:13` relies on (grads of unused h0 are zeros, not None).
`/root/reference` does not exist on the GPU box, so these vectors are committed.
"""
import contextlib
import io
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("SRNN_REFERENCE", "/root/reference")

CONFIGS = {
    # C2-shaped (3-tier "master best": look-ahead cond 86, weight-norm, 2 GRU layers) at small width
    "c2s": dict(frame_sizes=[20, 4], n_rnn=2, dim=32, learn_h0=True, q_levels=256, ulaw=True,
                weight_norm=True, cond_dim=86, spk_dim=6, B=2, n_cond=2),
    # C1-shaped (2-tier, frame 16, defaults of train.py:31-47: no weight-norm, cond 43)
    "c1s": dict(frame_sizes=[16], n_rnn=1, dim=32, learn_h0=True, q_levels=256, ulaw=True,
                weight_norm=False, cond_dim=43, spk_dim=6, B=2, n_cond=4),
    # generality: three frame tiers, linear quantisation, h0 as a buffer
    "c3s": dict(frame_sizes=[4, 2, 2], n_rnn=1, dim=16, learn_h0=False, q_levels=256, ulaw=False,
                weight_norm=True, cond_dim=5, spk_dim=6, B=3, n_cond=5),
}


def import_reference():
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import model as ref_model      # noqa
    import nn as ref_nn            # noqa
    import optim as ref_optim      # noqa
    sys.path.pop(0)
    return ref_model, ref_nn, ref_optim


@contextlib.contextmanager
def quiet_tmp_cwd():
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                yield
        finally:
            os.chdir(old)


def build(ref_model, c, seed=77977):
    torch.manual_seed(seed)                                   # train.py:62 default seed
    m = ref_model.SampleRNN(c["frame_sizes"], c["n_rnn"], c["dim"], c["learn_h0"], c["q_levels"], c["ulaw"],
                            c["weight_norm"], c["cond_dim"], c["spk_dim"])
    p = ref_model.Predictor(m)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():                                     # exercise h0 / biases / g (all trivial at init)
        for k, v in p.state_dict().items():
            if k.endswith(".h0"):
                v.copy_(0.1 * torch.randn(v.shape, generator=g))
            elif "bias" in k:
                v.copy_(0.05 * torch.randn(v.shape, generator=g))
            elif k.endswith("weight_g"):
                v.mul_(0.5 + torch.rand(v.shape, generator=g))
    return m, p


def make_inputs(c, lookback, seed=3):
    g = torch.Generator().manual_seed(seed)
    B, n_cond = c["B"], c["n_cond"]
    T = n_cond * lookback
    n_chunks = 3
    total = lookback + n_chunks * T
    data = torch.randint(0, c["q_levels"], (B, total), generator=g)
    cond = torch.rand(B, n_chunks * n_cond + 1, c["cond_dim"], generator=g, dtype=torch.float64)  # dataset.py:274 f64
    spk = torch.randint(0, c["spk_dim"], (B, 1), generator=g)
    return data, cond, spk, T, n_chunks


def chunk(data, cond, T, lookback, n_cond, i):
    """dataset.py:241-266: input = data[s : s+lookback+T-1], target = data[s+lookback : s+lookback+T], cond offset +1."""
    s = i * T
    x = data[:, s: s + lookback + T - 1].contiguous()
    y = data[:, s + lookback: s + lookback + T].contiguous()
    c = cond[:, i * n_cond + 1: i * n_cond + 1 + n_cond].contiguous()
    return x, y, c


def run_predictor_rows(predictor, hs, x, reset, cond, spk):
    """B=1-looped reference forward; hs = per-row hidden_states dicts (None = fresh)."""
    outs = []
    for b in range(x.shape[0]):
        if hs[b] is not None:
            predictor.hidden_states = hs[b]
        out = predictor(x[b:b + 1], reset, cond[b:b + 1], spk[b:b + 1], None, None)
        hs[b] = dict(predictor.hidden_states)
        outs.append(out)
    return torch.cat(outs, 0)


def hidden_to_np(predictor_model, hs):
    """-> list over tiers of (n_rnn, B, H)"""
    res = []
    for rnn in predictor_model.frame_level_rnns:
        res.append(torch.cat([h[rnn] for h in hs], dim=1).detach().numpy())
    return res


def main():
    from oracle import srnn_oracle as O
    ref_model, ref_nn, ref_optim = import_reference()
    for name, c in CONFIGS.items():
        out = {}
        m, p = build(ref_model, c)
        lookback = m.lookback
        sd0 = {k: v.detach().clone() for k, v in p.state_dict().items()}
        for k, v in sd0.items():
            out["sd/" + k] = v.numpy()
        data, cond, spk, T, n_chunks = make_inputs(c, lookback)
        out["data"], out["cond"], out["spk"] = data.numpy(), cond.numpy(), spk.numpy()
        B = c["B"]

        # ---- teacher-forced forward over consecutive chunks (TBPTT carry; reset only on chunk 0) ----
        with quiet_tmp_cwd(), torch.no_grad():
            hs = [None] * B
            for i in range(n_chunks):
                x, y, cc = chunk(data, cond, T, lookback, c["n_cond"], i)
                logp = run_predictor_rows(p, hs, x, i == 0, cc, spk)
                out[f"tf/logp{i}"] = logp.numpy()
                out[f"tf/loss{i}"] = np.float64(ref_nn.sequence_nll_loss_bits(logp, y).item())
                for t, h in enumerate(hidden_to_np(m, hs)):
                    out[f"tf/hidden{i}_{t}"] = h

        # ---- dequantiser LUT straight from the reference function (utils.py) ----
        out["lut"] = (2 * m.dequantize(torch.arange(c["q_levels"]), c["q_levels"])).numpy()

        # ---- generation, shared conditioner (reference form) ----
        n_seqs, n_cond_g = 3, c["n_cond"]
        g = torch.Generator().manual_seed(1234)
        uni = torch.rand(n_cond_g * lookback, n_seqs, generator=g).numpy().astype(np.float32)
        gcond = cond[0, :n_cond_g].numpy()
        gspk = int(spk[0, 0])
        out["gen/uniforms"], out["gen/cond"], out["gen/spk"] = uni, gcond, np.int64(gspk)

        def run_gen(n, cnd, sp, u):
            step = {"i": 0}
            rec = []
            orig = torch.Tensor.multinomial

            def patched(self, num_samples, *a, **k):
                assert num_samples == 1
                idx = O.sample_rows(self.detach().numpy(), u[step["i"]])
                step["i"] += 1
                return torch.from_numpy(idx).reshape(-1, 1)

            hook = m.sample_level_mlp.register_forward_hook(lambda mod, i, o: rec.append(o.detach().clone()))
            torch.Tensor.multinomial = patched
            try:
                with quiet_tmp_cwd(), torch.no_grad():
                    audio = ref_model.Generator(m, cuda=False)(n, 0, cnd, sp)
            finally:
                torch.Tensor.multinomial = orig
                hook.remove()
            return audio, torch.cat(rec, dim=1)

        audio, glogp = run_gen(n_seqs, gcond, gspk, uni)
        out["gen/audio"] = audio.numpy()
        out["gen/logp"] = glogp.numpy()
        # ---- generation, per-utterance conditioners (extension): reference run once per utterance ----
        bcond = cond[:, :n_cond_g].numpy()
        auds, lps = [], []
        for b in range(B):
            a, lp = run_gen(1, bcond[b], int(spk[b, 0]), uni[:, b:b + 1])
            auds.append(a)
            lps.append(lp)
        out["genb/cond"], out["genb/spk"] = bcond, spk[:, 0].numpy()
        out["genb/audio"] = torch.cat(auds, 0).numpy()
        out["genb/logp"] = torch.cat(lps, 0).numpy()

        # ---- training: 3 steps of the reference closure + gradient_clipping(Adam) ----
        m, p = build(ref_model, c)
        params = [q for q in p.parameters()]
        base = torch.optim.Adam(params, lr=1e-3)                           # train.py:56,238
        opt = ref_optim.gradient_clipping(base)                            # train.py:241
        hs = [None] * B
        grads0 = {}
        with quiet_tmp_cwd():
            for i in range(n_chunks):
                x, y, cc = chunk(data, cond, T, lookback, c["n_cond"], i)

                def closure():
                    total = 0.0
                    for b in range(B):
                        if hs[b] is not None:
                            p.hidden_states = hs[b]
                        o = p(x[b:b + 1], i == 0, cc[b:b + 1], spk[b:b + 1], None, None)
                        hs[b] = dict(p.hidden_states)
                        loss = ref_nn.sequence_nll_loss_bits(o, y[b:b + 1]) / B
                        loss.backward()
                        total += loss.item()
                    if i == 0:
                        for k, q in p.named_parameters():
                            grads0[k] = q.grad.detach().clone()
                    return torch.tensor(total)

                base.zero_grad(set_to_none=False)                         # trainer/__init__.py:111
                loss = opt.step(closure)                                  # trainer/__init__.py:112
                out[f"train/loss{i}"] = np.float64(float(loss))
        for k, gq in grads0.items():
            out["train/grad0/" + k] = gq.numpy()
        for k, v in p.state_dict().items():
            out["train/sd3/" + k] = v.detach().numpy()
        out["cfg"] = np.array(repr({k: v for k, v in c.items()}))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
