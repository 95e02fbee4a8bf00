"""Reference-compatible checkpoints / tags / log lines."""
import os
import re

import torch

import srnn_b200 as S

C = dict(frame_sizes=[4, 2], n_rnn=1, dim=16, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=5, spk_dim=6)


def test_saver_patterns_best_and_natural_order(tmp_path):
    d = str(tmp_path / "results" / "tag" / "checkpoints")
    m = S.SampleRNN(**C)
    p = S.Predictor(m)
    sv = S.CheckpointSaver(d, keep_old_checkpoints=True)
    for ep, it, vl in [(1, 100, 5.0), (2, 200, 4.0), (9, 900, 4.5), (10, 1000, 4.2)]:
        last, best = sv.epoch(ep, it, p, vl)
        assert os.path.basename(last) == "ep%d-it%d" % (ep, it)
        assert (best is not None) == (vl < 4.9 and ep == 2 or ep == 1)
    names = sorted(os.listdir(d))
    assert "best-ep2-it200" in names and "best-ep1-it100" not in names          # trainer/plugins.py:139-150
    sd, ep, it = S.load_last_checkpoint(d)
    assert (ep, it) == (10, 1000)                                              # natural order: ep10 after ep9 (train.py:112)
    q = S.Predictor(S.SampleRNN(**C))
    q.load_state_dict(sd, strict=True)
    for k, v in p.state_dict().items():
        assert torch.equal(v, q.state_dict()[k])
    assert set(sd) == set(p.state_dict()) and all(k.startswith("model.") for k in sd)
    assert S.parse_checkpoint_name("results/x/checkpoints/best-ep7-it12345") == (7, 12345)   # generate.py:66-83
    sv2 = S.CheckpointSaver(str(tmp_path / "c2"))
    sv2.epoch(1, 10, p, 3.0)
    sv2.epoch(2, 20, p, 3.5)
    assert sorted(os.listdir(str(tmp_path / "c2"))) == ["best-ep1-it10", "ep2-it20"]          # old 'last' cleared


def test_reference_checkpoint_layout_loads(golden):
    """A state_dict in the reference's layout (golden fixture written from the unmodified reference) loads strictly."""
    c = golden.c
    p = S.Predictor(S.SampleRNN(c["frame_sizes"], c["n_rnn"], c["dim"], c["learn_h0"], c["q_levels"], c["ulaw"],
                                c["weight_norm"], c["cond_dim"], c["spk_dim"]))
    p.load_state_dict(golden.state_dict(), strict=True)


def test_tag_and_log_line():
    defaults = dict(n_rnn=2, dim=1024, learn_h0=True, q_levels=256, seq_len=1040, weight_norm=True, batch_size=128, seed=77977,
                    ulaw=True, look_ahead=False, norm_ind=False, static_spk=False, qrnn=False, scheduler=False,
                    learning_rate=0.001)
    params = dict(defaults, exp="master", frame_sizes=[20, 4], look_ahead=True, norm_ind=True, dataset="tcstar", cond_set="cond")
    tag = S.make_tag(params, defaults)
    assert tag == "exp:master~frame_sizes:20,4~look_ahead:T~norm_ind:T~dataset:tcstar~cond_set:cond"   # train.py:66-84
    line = S.log_line(3, 1200, 3.71234, 95.2, validation_loss=3.9, test_loss=4.0)
    assert re.search("training_loss:.*time:", line)                                                  # plotlog.py:23
    assert float(re.search("training_loss: ([-0-9.]+)", line).group(1)) == 3.7123                     # plotlog.py:24
    assert float(re.search("validation_loss: ([-0-9.]+)", line).group(1)) == 3.9
    assert float(re.search("test_loss: ([-0-9.]+)", line).group(1)) == 4.0
