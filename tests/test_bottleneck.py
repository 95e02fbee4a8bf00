"""Bottle-neck voice-conversion variant (BASELINE.json configs[4]).  PARITY UNPINNED: the reference tree has no source for it
(run_sampleneck.sh:2); these tests pin the CUDA chain kernel to the thesis-derived restatement in oracle/bottleneck_oracle.py
and check that the chain composes with the unchanged hot path."""
import numpy as np
import pytest
import torch

import srnn_b200 as S
from oracle import bottleneck_oracle as BO
from oracle import srnn_oracle as O

CFG = dict(frame_sizes=[20, 4], n_rnn=2, dim=64, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=43,
           spk_dim=6)


def _layers(cnd):
    sd = cnd.state_dict()
    return [{k.split(".", 2)[2]: v.cpu().numpy() for k, v in sd.items() if k.startswith("layers.%d." % i)}
            for i in range(len(cnd.layers))]


def test_chain_shapes_and_names():
    m = S.BottleneckSampleRNN(ind_cond_dim=30, **CFG)
    assert m.conditioner.dims == [43, 40, 30, 20, 30]
    assert m.core.cond_dim == 30 and m.core.frame_level_rnns[-1].cond_expand.weight_v.shape == (64, 30, 1)
    assert "conditioner.layers.0.weight_g" in m.state_dict() and m.lookback == 80
    with pytest.raises(S.SrnnError):
        m.conditioner(torch.rand(3, 43))                       # parameters on the CPU: no fallback


def test_oracle_chain_is_relu_of_affine_maps():
    torch.manual_seed(0)
    c = S.BottleneckConditioner(43, 10)
    x = np.random.default_rng(0).random((5, 43), dtype=np.float32)
    y = BO.chain_forward(_layers(c), x)
    assert y.shape == (5, 10) and (y >= 0).all()
    w0, b0 = BO.fold(_layers(c)[0])
    np.testing.assert_allclose(np.linalg.norm(w0, axis=1), c.layers[0].weight_g.detach().numpy().reshape(-1), rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("ind", [10, 30])
def test_gpu_chain_against_oracle(ind):
    torch.manual_seed(ind)
    c = S.BottleneckConditioner(43, ind)
    with torch.no_grad():
        for p in c.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.3)
    layers = _layers(c)
    c.cuda()
    x = torch.rand(7, 11, 43)
    got = c(x).cpu().numpy()
    ref = BO.chain_forward(layers, x.numpy())
    np.testing.assert_allclose(got, ref, atol=2e-6, rtol=1e-5)


@pytest.mark.gpu
def test_gpu_bottleneck_generation_equals_core_on_chain_output():
    """The chain in front changes nothing else: generating through BottleneckGenerator == generating with the core model on the
    oracle's chain output (fp32 mode, same uniforms), and the log-probs stay inside the fp32 gate against the oracle."""
    torch.manual_seed(3)
    m = S.BottleneckSampleRNN(ind_cond_dim=30, **CFG)
    layers = _layers(m.conditioner)
    core_cfg = dict(CFG, cond_dim=30)
    sd = {"model." + k: v.clone() for k, v in m.core.state_dict().items()}
    m.cuda()
    B, n_cond = 5, 2
    cond = torch.rand(B, n_cond, 43)
    spk = torch.randint(0, 6, (B,))
    u = torch.rand(n_cond * 80, B)
    audio, samples, logp = S.BottleneckGenerator(m, cuda=True)(B, 0, cond, spk, uniforms=u, return_samples=True, return_logp=True)
    c30 = torch.from_numpy(BO.chain_forward(layers, cond.numpy()))
    w = O.unpack_state_dict(sd, O.Config(**core_cfg))
    ref = O.Generator(w)(B, c30.numpy(), spk.numpy(), u.numpy())
    assert (ref.numpy() == samples.long().numpy()).mean() > 0.99
    seq = torch.cat([torch.full((B, 80), 128, dtype=torch.long), samples.long()], 1)
    with torch.no_grad():
        tf = O.Predictor(w).forward(seq[:, :-1], True, c30, spk.reshape(B, 1))
    assert float((tf - logp).abs().max()) < 2e-3
    # teacher-forced pass through the wrapper
    p = S.BottleneckPredictor(m)
    with torch.no_grad():
        got = p(seq[:, :-1], True, cond, spk.reshape(B, 1))
    assert float((got.cpu() - tf).abs().max()) < 2e-3
