"""Training data path: TBPTT chunking + quantisers.  The golden file holds what the UNMODIFIED
`FolderDataset.__getitem__` / `utils.uquantize` / `utils.linear_quantize` return (tests/golden/make_data_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as D

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_path.npz"))
BS, N, SEQ, OVERLAP, COND_LEN, COND_DIM, Q = [int(v) for v in Z["meta"]]
KW = dict(overlap_len=OVERLAP, q_levels=Q, seq_len=SEQ, batch_size=BS, cond_len=COND_LEN)


def test_oracle_quantisers_match_reference():
    assert np.array_equal(D.uquantize(Z["q/x"], Q), Z["q/ulaw"])
    assert int(Z["q/ulaw"].max()) == 256                    # the reference's overflow at x == 1.0
    assert np.array_equal(D.linear_quantize(Z["q/rows"], Q), Z["q/linear"])


@pytest.mark.parametrize("tag", ["ulaw", "lin"])
def test_oracle_items_match_reference(tag):
    ulaw = tag == "ulaw"
    data = Z["data"] if ulaw else np.stack([D.linear_quantize(r, Q) for r in Z["data"]])
    for idx in range(int(Z[f"{tag}/n_items"])):
        d, reset, t, c, s = D.get_item(data, Z["cond"], Z["global_spk"], idx, ulaw=ulaw, **KW)
        assert np.array_equal(d, Z[f"{tag}/{idx}/data"]) and np.array_equal(t, Z[f"{tag}/{idx}/target"])
        assert np.array_equal(c, Z[f"{tag}/{idx}/cond"]) and np.array_equal(s, Z[f"{tag}/{idx}/spk"])
        assert bool(reset) == bool(Z[f"{tag}/{idx}/reset"])


@pytest.mark.gpu
def test_gpu_quantisers_against_reference():
    import srnn_b200 as S
    x = torch.from_numpy(Z["q/x"]).cuda()
    q = S.quantize(x, Q, True).cpu().numpy()
    ref = np.minimum(Z["q/ulaw"], Q - 1)                    # documented deviation: index 256 is clamped to 255
    d = np.abs(q - ref)
    # identical formula and operation order; logf on the GPU and log on the host may differ in the last bit, which moves a
    # sample sitting exactly on a bin edge by one level
    assert d.max() <= 1 and (d != 0).mean() <= 1e-4, (d.max(), (d != 0).mean())
    assert q.min() >= 0 and q.max() <= Q - 1
    rows = torch.from_numpy(Z["q/rows"]).cuda()
    ql = S.quantize(rows, Q, False).cpu().numpy()
    dl = np.abs(ql - Z["q/linear"])
    assert dl.max() <= 1 and (dl != 0).mean() <= 1e-4, (dl.max(), (dl != 0).mean())


@pytest.mark.gpu
def test_gpu_tbptt_linear_batches_against_reference():
    """ulaw=False: the stored data is already quantised per file (dataset.py:129-130); items are exact slices cast to int64
    (dataset.py:249-251), input and target of one sample always agree."""
    import srnn_b200 as S
    stored = np.stack([D.linear_quantize(r, Q) for r in Z["data"]])
    bt = S.TBPTTBatcher(stored, Z["cond"], Z["global_spk"], OVERLAP, Q, False, SEQ, BS, COND_LEN)
    n_items = int(Z["lin/n_items"])
    for k in range(n_items // BS):
        data, reset, target, cond, spk = bt.batch(k)
        assert data.dtype == torch.int64 and target.dtype == torch.int64
        for r in range(BS):
            idx = k * BS + r
            assert np.array_equal(data[r].cpu().numpy(), Z[f"lin/{idx}/data"])
            assert np.array_equal(target[r].cpu().numpy(), Z[f"lin/{idx}/target"])
            assert np.array_equal(cond[r].cpu().numpy(), Z[f"lin/{idx}/cond"])
            assert int(spk[r, 0]) == int(Z[f"lin/{idx}/spk"][0]) and bool(reset) == bool(Z[f"lin/{idx}/reset"])
        # the same sample seen as an input and as a target is the same level
        assert torch.equal(data[:, OVERLAP:], target[:, :-1])


@pytest.mark.gpu
def test_gpu_tbptt_batches_against_reference():
    import srnn_b200 as S
    bt = S.TBPTTBatcher(Z["data"], Z["cond"], Z["global_spk"], OVERLAP, Q, True, SEQ, BS, COND_LEN)
    n_items = int(Z["ulaw/n_items"])
    assert len(bt) >= n_items // BS
    mism = 0
    for k in range(n_items // BS):
        data, reset, target, cond, spk = bt.batch(k)
        assert data.dtype == torch.int64 and data.shape == (BS, OVERLAP + SEQ - 1) and target.shape == (BS, SEQ)
        for r in range(BS):
            idx = k * BS + r
            ref_d, ref_t = np.minimum(Z[f"ulaw/{idx}/data"], Q - 1), np.minimum(Z[f"ulaw/{idx}/target"], Q - 1)
            mism += int((data[r].cpu().numpy() != ref_d).sum() + (target[r].cpu().numpy() != ref_t).sum())
            assert np.array_equal(cond[r].cpu().numpy(), Z[f"ulaw/{idx}/cond"])          # float64, exact slices
            assert int(spk[r, 0]) == int(Z[f"ulaw/{idx}/spk"][0])
            assert bool(reset) == bool(Z[f"ulaw/{idx}/reset"])
    assert mism <= 2, mism                                   # bin-edge samples only (see the quantiser test)


@pytest.mark.gpu
def test_gpu_batcher_feeds_predictor():
    import srnn_b200 as S
    c = dict(frame_sizes=[20, 4], n_rnn=1, dim=64, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=COND_DIM,
             spk_dim=6)
    torch.manual_seed(0)
    m = S.SampleRNN(**c)
    p = S.Predictor(m).cuda()
    bt = S.TBPTTBatcher(Z["data"], Z["cond"], Z["global_spk"], m.lookback, Q, True, SEQ, BS, COND_LEN)
    opt = S.ClampAdam(p.parameters(), lr=1e-3, model=m)
    losses = []
    for data, reset, target, cond, spk in bt:
        def closure():
            out = p(data, reset, cond, spk, None, None)
            loss = S.sequence_nll_loss_bits(out, target)
            loss.backward()
            return loss.detach()
        opt.zero_grad()
        losses.append(float(opt.step(closure)))
    assert len(losses) == len(bt) and all(np.isfinite(losses)) and losses[0] > 7.5
