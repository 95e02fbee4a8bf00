"""Generation front-end: Ahocoder feature preparation against the reference's own interpolation outputs,
ragged batching against solo runs, float32 WAV files."""
import os

import numpy as np
import pytest
import torch

import srnn_b200 as S
from oracle import frontend_oracle as FO

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend.npz"))


def test_interpolation_matches_reference_outputs():
    for k in range(int(Z["n"])):
        sig, sym = Z[f"{k}/signal"], float(Z[f"{k}/sym"])
        for fn in (FO.interpolation, S.interpolation):
            o, uv = fn(sig.copy(), sym)
            assert np.array_equal(o, Z[f"{k}/out"]) and np.array_equal(uv, Z[f"{k}/uv"]), (k, fn.__module__)


def features(rs, n):
    cc = rs.randn(n, 40)
    lf0 = rs.uniform(4, 6, size=n)
    lf0[rs.rand(n) < 0.3] = -1e10
    gv = rs.uniform(2e3, 6e3, size=n)
    gv[rs.rand(n) < 0.3] = 0.0
    return cc, lf0, gv


@pytest.mark.parametrize("norm_ind,look_ahead", [(True, True), (False, False)])
def test_conditioner_matches_oracle(norm_ind, look_ahead):
    rs = np.random.RandomState(3)
    cc, lf0, gv = features(rs, 57)
    lo, hi = (rs.randn(6, 43) - 3, rs.randn(6, 43) + 3) if norm_ind else (rs.randn(43) - 3, rs.randn(43) + 3)
    a = S.build_conditioner(cc, lf0, gv, 4, lo, hi, norm_ind, look_ahead)
    b = FO.conditioner(cc, lf0, gv, 4, lo, hi, norm_ind, look_ahead)
    assert a.shape == (57, 43 * (1 + look_ahead)) and np.array_equal(a, b)


def test_wav_float32_roundtrip(tmp_path):
    from scipy.io import wavfile
    a = np.linspace(-1, 1, 1601, dtype=np.float32)
    p = str(tmp_path / "a.wav")
    S.write_wav_f32(p, a, 16000)
    sr, b = wavfile.read(p)
    assert sr == 16000 and b.dtype == np.float32 and np.array_equal(a, b)


@pytest.mark.gpu
def test_ragged_batch_equals_solo_runs_and_writes_files(tmp_path):
    torch.manual_seed(2)
    c = dict(frame_sizes=[20, 4], n_rnn=2, dim=128, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=86,
             spk_dim=6)
    m = S.SampleRNN(**c).cuda()
    rs = np.random.RandomState(5)
    min_max = (rs.randn(6, 43) - 3, rs.randn(6, 43) + 3)
    gen = S.BatchedFileGenerator(m, min_max, ["s%d" % i for i in range(6)], norm_ind=True, look_ahead=True, max_batch=4)
    lens = [3, 1, 4, 2, 4, 1, 2]
    conds, spks, unis = [], [], []
    for i, n in enumerate(lens):
        cc, lf0, gv = features(rs, n)
        conds.append(S.build_conditioner(cc, lf0, gv, i % 6, min_max[0], min_max[1], True, True))
        spks.append(i % 6)
        unis.append(rs.rand(n * 80).astype(np.float32))
    batched = gen(conds, spks, uniforms=unis)
    for i, n in enumerate(lens):
        assert batched[i].shape == (n * 80,) and batched[i].dtype == np.float32
        solo = gen([conds[i]], [spks[i]], uniforms=[unis[i]])[0]
        assert np.array_equal(batched[i], solo), i                       # padding and batch order change nothing
    # file-to-file form
    bases = []
    for i in range(2):
        cc, lf0, gv = features(rs, 2 + i)
        b = str(tmp_path / ("utt%d" % i))
        np.savetxt(b + ".cc", cc)
        np.savetxt(b + ".lf0", lf0)
        np.savetxt(b + ".gv", gv)
        bases.append(b)
    paths = gen.generate_files(bases, ["s1", "s3"], str(tmp_path / "samples"), tag="ep1-it10", seed=1)
    from scipy.io import wavfile
    for i, p in enumerate(paths):
        sr, a = wavfile.read(p)
        assert sr == 16000 and a.dtype == np.float32 and a.shape == ((2 + i) * 80,) and np.abs(a).max() <= 1.0
    assert os.path.basename(paths[1]) == "ep1-it10_file-utt1_spk-s3.wav"
