import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """One committed fixture file produced by tests/golden/make_golden.py from the unmodified reference."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.c = ast.literal_eval(str(self.z["cfg"]))

    def __getitem__(self, k):
        return self.z[k]

    def state_dict(self, prefix="sd/"):
        return {k[len(prefix):]: torch.from_numpy(self.z[k]) for k in self.z.files if k.startswith(prefix)}

    def cfg(self):
        from oracle.srnn_oracle import Config
        c = self.c
        return Config(c["frame_sizes"], c["n_rnn"], c["dim"], c["learn_h0"], c["q_levels"], c["ulaw"],
                      c["weight_norm"], c["cond_dim"], c["spk_dim"])

    def chunk(self, i):
        """dataset.py:241-266 chunking, identical to make_golden.chunk."""
        cfg = self.cfg()
        lookback, n_cond = cfg.lookback, self.c["n_cond"]
        T = n_cond * lookback
        data, cond = torch.from_numpy(self.z["data"]), torch.from_numpy(self.z["cond"])
        s = i * T
        x = data[:, s: s + lookback + T - 1].contiguous()
        y = data[:, s + lookback: s + lookback + T].contiguous()
        c = cond[:, i * n_cond + 1: i * n_cond + 1 + n_cond].contiguous()
        return x, y, c


@pytest.fixture(params=["c2s", "c1s", "c3s"])
def golden(request):
    return Golden(request.param)
