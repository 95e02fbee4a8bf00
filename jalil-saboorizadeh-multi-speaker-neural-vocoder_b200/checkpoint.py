"""Checkpoints and log lines in the reference's layout, so that reference-trained checkpoints load here and
checkpoints / logs written here are readable by the reference's tooling.

  * `results/<tag>/checkpoints/ep{E}-it{I}` and `best-ep{E}-it{I}` hold `torch.save(predictor.state_dict())`
    (trainer/plugins.py:113-155); the keys are the `model.*` names of SURVEY App. A, which `Predictor.state_dict()` of this
    package reproduces exactly.
  * `<tag>` = `make_tag(params)` (train.py:66-84); `load_last_checkpoint` picks the naturally-sorted last `ep*-it*`
    (train.py:108-124); `generate.py:66-83` parses epoch / iteration out of any `.*ep{E}-it{I}` name.
  * `plotlog.py:23-26` greps `training_loss: <x>`, `validation_loss: <x>`, `test_loss: <x>` and `training_loss:.*time:`.
"""
import os
import re
from glob import glob

import torch

LAST_PATTERN = "ep{}-it{}"            # trainer/plugins.py:115
BEST_PATTERN = "best-ep{}-it{}"       # trainer/plugins.py:116

TAG_PARAMS = ["exp", "frame_sizes", "n_rnn", "dim", "learn_h0", "ulaw", "q_levels", "seq_len", "look_ahead", "norm_ind",
              "batch_size", "dataset", "cond_set", "static_spk", "seed", "weight_norm", "qrnn", "scheduler",
              "learning_rate"]        # train.py:61-64


def make_tag(params, default_params):
    """train.py:66-84: `key:value` of every tag parameter that differs from its default, joined by '~'."""
    def to_string(v):
        if isinstance(v, bool):
            return "T" if v else "F"
        if isinstance(v, list):
            return ",".join(map(to_string, v))
        return str(v)
    return "~".join(k + ":" + to_string(params[k]) for k in TAG_PARAMS
                    if k in params and (k not in default_params or params[k] != default_params[k]))


def _natural_key(s):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]


def parse_checkpoint_name(path):
    """generate.py:66-83 -> (epoch, iteration), (0, 0) when the name carries none."""
    m = re.match(".*" + LAST_PATTERN.format(r"(\d+)", r"(\d+)"), os.path.basename(path))
    return (int(m.group(1)), int(m.group(2))) if m else (0, 0)


def load_last_checkpoint(checkpoints_path, map_location=None):
    """train.py:108-124 -> (state_dict, epoch, iteration) of the last `ep*-it*` file (natural order), or None."""
    paths = sorted(glob(os.path.join(checkpoints_path, LAST_PATTERN.format("*", "*"))), key=_natural_key)
    paths = [p for p in paths if re.match(LAST_PATTERN.format(r"(\d+)", r"(\d+)") + "$", os.path.basename(p))]
    if not paths:
        return None
    epoch, iteration = parse_checkpoint_name(paths[-1])
    return torch.load(paths[-1], map_location=map_location), epoch, iteration


class CheckpointSaver:
    """trainer/plugins.py:113-155 (`SaverPlugin`) without the plugin machinery: call `epoch()` once per epoch."""

    def __init__(self, checkpoints_path, keep_old_checkpoints=False):
        self.checkpoints_path, self.keep_old_checkpoints = checkpoints_path, keep_old_checkpoints
        self._best_val_loss = float("+inf")
        os.makedirs(checkpoints_path, exist_ok=True)

    def _clear(self, pattern):
        for f in glob(os.path.join(self.checkpoints_path, pattern)):
            os.remove(f)

    def epoch(self, epoch_index, iterations, predictor, validation_loss):
        sd = {k: v.detach().cpu() for k, v in predictor.state_dict().items()}
        if not self.keep_old_checkpoints:
            self._clear(LAST_PATTERN.format("*", "*"))
        last = os.path.join(self.checkpoints_path, LAST_PATTERN.format(epoch_index, iterations))
        torch.save(sd, last)
        best = None
        if validation_loss < self._best_val_loss:
            self._clear(BEST_PATTERN.format("*", "*"))
            best = os.path.join(self.checkpoints_path, BEST_PATTERN.format(epoch_index, iterations))
            torch.save(sd, best)
            self._best_val_loss = validation_loss
        return last, best


def log_line(epoch, iteration, training_loss, time_s, validation_loss=None, test_loss=None):
    """One log line in the shape `plotlog.py:23-26` parses (the reference's Logger prints `name: value` fields in this order)."""
    s = "epoch: %d\titeration: %d\ttraining_loss: %.4f" % (epoch, iteration, training_loss)
    if validation_loss is not None:
        s += "\tvalidation_loss: %.4f" % validation_loss
    if test_loss is not None:
        s += "\ttest_loss: %.4f" % test_loss
    return s + "\ttime: %ds" % int(time_s)
