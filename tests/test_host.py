"""CPU-side checks: the C-ABI library loads and exports every symbol include/srnn_b200.h declares, the host mirror
keeps the reference's state_dict layout, and the product path refuses to run without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import srnn_b200 as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "srnn_b200.h")).read()
    return sorted(set(re.findall(r"SRNN_API[^;(]*?\b(srnn_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(S._lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(S._lib.exported_symbols()) == names          # the ctypes table covers the whole header
    assert S._lib.load().srnn_version() >= 100


def test_state_dict_layout_matches_reference(golden):
    c = golden.c
    m = S.SampleRNN(c["frame_sizes"], c["n_rnn"], c["dim"], c["learn_h0"], c["q_levels"], c["ulaw"],
                    c["weight_norm"], c["cond_dim"], c["spk_dim"])
    p = S.Predictor(m)
    ref = golden.state_dict()
    mine = p.state_dict()
    assert sorted(mine) == sorted(ref)
    for k in ref:
        assert tuple(mine[k].shape) == tuple(ref[k].shape), k
    p.load_state_dict(ref)                                      # strict
    assert m.lookback == int(np.prod(c["frame_sizes"]))
    names = {n for n, _ in p.named_parameters()}
    assert ("model.frame_level_rnns.0.h0" in names) == bool(c["learn_h0"])   # buffer when learn_h0=False


def test_init_statistics():
    torch.manual_seed(0)
    m = S.SampleRNN([16], 1, 256, True, 256, True, True, 43, 6)
    r = m.frame_level_rnns[0]
    assert float(r.h0.abs().max()) == 0.0
    w = r.rnn.weight_hh_l0[512:]                                # third chunk is orthogonal (model.py:160-164)
    np.testing.assert_allclose((w @ w.t()).detach().numpy(), np.eye(256), atol=1e-4)
    g, v = r.upsampling.conv_t.weight_g, r.upsampling.conv_t.weight_v
    np.testing.assert_allclose(g.detach().reshape(-1).numpy(), v.detach().reshape(256, -1).norm(dim=1).numpy(), rtol=1e-5)
    assert float(v.abs().max()) <= (6 / 256) ** 0.5 + 1e-6


def test_no_cpu_fallback():
    m = S.SampleRNN([4, 2], 1, 16, True, 256, True, False, 5, 6)
    with pytest.raises(S.SrnnError):
        S.Generator(m)(1, 0, np.zeros((2, 5)), 0)
    with pytest.raises(S.SrnnError):
        S.Predictor(m)(torch.zeros(1, 8 + 8 - 1, dtype=torch.long), True, torch.zeros(1, 1, 5), torch.zeros(1, 1))
    if not torch.cuda.is_available():
        cfg = S._lib.Config()
        cfg.n_tiers, cfg.n_rnn, cfg.dim, cfg.q_levels, cfg.cond_dim, cfg.spk_dim, cfg.ulaw = 1, 1, 16, 256, 5, 6, 1
        cfg.frame_sizes[0] = 4
        h = C.c_void_p()
        assert S._lib.load().srnn_create(C.byref(cfg), C.byref(h)) == -2      # SRNN_ERR_CUDA
        assert b"no CPU fallback" in S._lib.load().srnn_last_error()


def test_argument_validation():
    lib = S._lib.load()
    cfg = S._lib.Config()
    cfg.n_tiers, cfg.n_rnn, cfg.dim, cfg.q_levels, cfg.cond_dim, cfg.spk_dim, cfg.ulaw = 1, 1, 16, 128, 5, 6, 1
    cfg.frame_sizes[0] = 4
    h = C.c_void_p()
    assert lib.srnn_create(C.byref(cfg), C.byref(h)) == -4                   # q_levels != 256 unsupported
    cfg.q_levels, cfg.n_tiers = 256, 9
    assert lib.srnn_create(C.byref(cfg), C.byref(h)) == -1
    with pytest.raises(NotImplementedError):
        S.SampleRNN([4], 1, 16, True, 256, True, False, 5, 6, qrnn=True)


def test_gradient_clipping_wrapper_mirrors_the_reference_call():
    """train.py:238-241: ``optimizer = gradient_clipping(torch.optim.Adam(predictor.parameters()))``."""
    import torch
    import srnn_b200 as S
    ps = [torch.zeros(3, requires_grad=True), torch.zeros(2, 2, requires_grad=True)]
    opt = S.gradient_clipping(torch.optim.Adam(ps, lr=3e-4, betas=(0.8, 0.9), eps=1e-6))
    assert isinstance(opt, S.ClampAdam) and opt.lr == 3e-4 and opt.betas == (0.8, 0.9) and opt.eps == 1e-6 and opt.clamp == 1.0
    assert [id(p) for p in opt.params] == [id(p) for p in ps]
    assert S.gradient_clipping(opt) is opt
    with pytest.raises(Exception):
        S.gradient_clipping(torch.optim.SGD(ps, lr=0.1))


def test_per_module_forward_has_no_cpu_fallback_and_survives_copies():
    """FrameLevelRNN.forward / SampleLevelMLP.forward run on the owning SampleRNN's CUDA context: on a CPU model they raise
    (no fallback); the weak owner registry does not leak into state_dict / pickling / deepcopy of the modules."""
    import copy
    import pickle
    import torch
    import srnn_b200 as S
    c = dict(frame_sizes=[4, 2], n_rnn=1, dim=16, learn_h0=True, q_levels=256, ulaw=True, weight_norm=True, cond_dim=5, spk_dim=3)
    m = S.SampleRNN(**c)
    top, mlp = m.frame_level_rnns[-1], m.sample_level_mlp
    with pytest.raises(S.SrnnError):
        top(torch.zeros(2, 3, 8), None, None, torch.zeros(2, 3, 5), torch.zeros(2, 1, dtype=torch.long))
    with pytest.raises(S.SrnnError):
        mlp(torch.zeros(2, 7, dtype=torch.long), torch.zeros(2, 4, 16))
    keys = set(m.state_dict().keys())
    m2 = copy.deepcopy(m)
    assert set(m2.state_dict().keys()) == keys and not any("_owner" in k for k in keys)
    pickle.loads(pickle.dumps(m.state_dict()))
    # a copied tier is not registered until its own model binds it (then it resolves to the copy, not the original)
    m2._bind_modules()
    from importlib import import_module
    M = import_module("jalil-saboorizadeh-multi-speaker-neural-vocoder_b200.model")
    assert M._owner_of(m2.frame_level_rnns[0])[0] is m2 and M._owner_of(m.frame_level_rnns[0])[0] is m
