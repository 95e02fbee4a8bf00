#!/usr/bin/env python
"""Loss curve of the UNMODIFIED reference over the first 1000 training steps (north_star: "loss curves overlay over the
first 1k training steps").  Build container only (needs /root/reference); writes tests/golden/loss_curve.npz.

The reference closure is trainer/__init__.py:99-112: zero_grad, forward, sequence_nll_loss_bits, backward,
gradient_clipping(Adam).step.  Work-arounds as in make_golden.py (B=1-looped Predictor, temp cwd, zero_grad semantics).
The quantised stream is produced by the reference's own utils.uquantize."""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import loss_curve_inputs as I          # noqa: E402
import make_golden as G                # noqa: E402


def main():
    ref_model, ref_nn, ref_optim = G.import_reference()
    sys.path.insert(0, G.REF)
    import utils as ref_utils
    sys.path.pop(0)
    c = dict(I.CONFIG)
    m, p = G.build(ref_model, c)
    out = {"sd/" + k: v.detach().clone().numpy() for k, v in p.state_dict().items()}
    data = ref_utils.uquantize(torch.from_numpy(I.audio()).float(), c["q_levels"]).long()      # utils.py:33-36
    cond = torch.from_numpy(I.conditioners())
    spk = torch.from_numpy(I.speakers())
    out["data"] = data.numpy().astype(np.uint8)
    base = torch.optim.Adam(list(p.parameters()), lr=I.LR)
    opt = ref_optim.gradient_clipping(base)
    hs = [None] * I.B
    losses = []
    t0 = time.time()
    with G.quiet_tmp_cwd():
        for i in range(I.STEPS):
            x, y, cc = I.chunk(data, cond, i)

            def closure():
                total = 0.0
                for b in range(I.B):
                    if hs[b] is not None:
                        p.hidden_states = hs[b]
                    o = p(x[b:b + 1], i == 0, cc[b:b + 1], spk[b:b + 1], None, None)
                    hs[b] = dict(p.hidden_states)
                    loss = ref_nn.sequence_nll_loss_bits(o, y[b:b + 1]) / I.B
                    loss.backward()
                    total += loss.item()
                return torch.tensor(total)

            base.zero_grad(set_to_none=False)
            losses.append(float(opt.step(closure)))
    out["losses"] = np.asarray(losses, dtype=np.float64)
    path = os.path.join(HERE, "loss_curve.npz")
    np.savez_compressed(path, **out)
    sys.stderr.write("%d steps in %.0f s: loss %.3f -> %.3f bits; %s %d KiB\n" % (
        I.STEPS, time.time() - t0, losses[0], np.mean(losses[-50:]), path, os.path.getsize(path) // 1024))


if __name__ == "__main__":
    main()
