"""Development aid: bf16-mode log-prob error against the oracle with and without the folded first layer (SRNN_NO_GI_FOLD)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srnn_b200 as S
from oracle import srnn_oracle as O
dim, B = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(dim)
c = dict(frame_sizes=[20, 4], n_rnn=2, dim=dim, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86, spk_dim=6)
m = S.SampleRNN(**c); p = S.Predictor(m)
with torch.no_grad():
    for k, v in p.state_dict().items():
        if "bias" in k or k.endswith("h0"): v.normal_(0, 0.1)
sd = {k: v.clone() for k, v in p.state_dict().items()}
p.cuda()
w = O.unpack_state_dict(sd, O.Config(**c))
n_cond = 2
cond, spk = torch.rand(B, n_cond, 86), torch.randint(0, 6, (B,))
for fold in (True, False):
    if fold: os.environ.pop("SRNN_NO_GI_FOLD", None)
    else: os.environ["SRNN_NO_GI_FOLD"] = "1"
    _, samples, logp = S.Generator(m, cuda=True, mode=S.MODE_BF16)(B, 0, cond, spk, seed=5, return_samples=True, return_logp=True)
    seq = torch.cat([torch.full((B, 80), 128, dtype=torch.long), samples.long()], 1)
    with torch.no_grad():
        ref = O.Predictor(w).forward(seq[:, :-1], True, cond, spk.reshape(B, 1))
    d = (ref - logp).abs()
    print("fold" if fold else "no fold", "max %.4f mean %.5f  | first period max %.4f mean %.5f | second period max %.4f mean %.5f" % (
        float(d.max()), float(d.mean()), float(d[:, :80].max()), float(d[:, :80].mean()), float(d[:, 80:].max()), float(d[:, 80:].mean())))
