"""Development aid: per-tensor relative error of the bf16 backward against the fp32 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import srnn_b200 as S
from oracle import srnn_oracle as O
dim, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(dim + 7)
c = dict(frame_sizes=[20, 4], n_rnn=2, dim=dim, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86, spk_dim=6)
m = S.SampleRNN(**c); p = S.Predictor(m, mode=S.MODE_BF16)
with torch.no_grad():
    for k, v in p.state_dict().items():
        if "bias" in k or k.endswith("h0"): v.normal_(0, 0.1)
sd = {k: v.clone() for k, v in p.state_dict().items()}
p.cuda()
x = torch.randint(0, 256, (B, 80 + T - 1)); y = torch.randint(0, 256, (B, T))
cond = torch.rand(B, T // 80, 86, dtype=torch.float64); spk = torch.randint(0, 6, (B, 1))
loss_ref, grads, _, _ = O.loss_and_grads(sd, O.Config(**c), None, x, True, cond, spk, y)
for mode in (S.MODE_BF16, S.MODE_FP32):
    p.mode = mode
    for q in p.parameters(): q.grad = None
    out = p(x, True, cond, spk, None, None)
    loss = S.sequence_nll_loss_bits(out, y); loss.backward()
    print("mode", mode, "loss", float(loss.detach()), "ref", float(loss_ref))
    for k, q in p.named_parameters():
        ref = grads[k].numpy().astype(np.float64); got = q.grad.detach().cpu().numpy().astype(np.float64)
        n = np.linalg.norm(ref)
        print("   %-55s |ref| %.3e  rel %.4f" % (k[6:], n, np.linalg.norm(got - ref) / (n + 1e-30)))
