"""Batched generation front-end keeping ``generate.py``'s file contract.

The reference rebuilds the model and reloads the checkpoint for every file and generates ONE utterance per call
(generate.py:131-252).  Here the model is bound once and many utterances of different lengths go through one
``srnn_generate`` call: conditioners are padded to the longest utterance (generation is causal and utterances are
independent, so a padded utterance's first samples equal a solo run bit for bit) and the audio is trimmed afterwards.

Per utterance (generate.py:146-196): Ahocoder text features ``<name>.cc`` (40 cepstra), ``<name>.lf0`` and ``<name>.gv``
-> unvoiced-segment interpolation (interpolate.py:45-72) -> [cc | f0 | fv | uv] (43) -> min-max normalisation with the
training partition's ``min_max_{ind,joint}[_static].npy`` (per speaker or jointly, dataset.py:189-198) -> optional
look-ahead concatenation of the next frame (86, dataset.py:213-221 / generate.py:190-194) -> float32 WAV.
The feature preparation is host-side numpy like the reference; only the synthesis runs on the GPU.
"""
import os
import struct

import numpy as np
import torch

from . import _lib as L
from .model import Generator

F0_UNVOICED, FV_UNVOICED = -10000000000, 1e3          # generate.py:152,157


def interpolation(signal, unvoiced_symbol):
    """interpolate.py:45-72 vectorised per unvoiced run: leading run <- first voiced value, inner runs <- straight line
    between the neighbouring voiced frames, trailing run <- last voiced value.  Returns (interpolated, uv mask)."""
    signal = np.asarray(signal, dtype=np.float64)
    n = signal.shape[0]
    out = signal.copy()
    uv = np.ones(n, dtype=np.int8)
    voiced = signal > unvoiced_symbol
    if n == 0 or voiced.all():
        return out, uv
    if not voiced.any():                                 # nothing to anchor on: the reference leaves such a signal unchanged
        return out, uv
    idx = np.flatnonzero(voiced)
    first, last = idx[0], idx[-1]
    if first > 0:                                        # interpolate.py:53-56
        out[:first] = signal[first]
        uv[:first] = 0
    # inner unvoiced runs (interpolate.py:57-67): the line starts AT the last voiced frame t0 (rewritten with its own value)
    gaps = np.flatnonzero(np.diff(idx) > 1)
    for g in gaps:
        t0, t1 = idx[g], idx[g + 1]
        f0, f1 = signal[t0], signal[t1]
        t = np.arange(t0, t1)
        out[t0:t1] = f0 + (t - t0) * ((f1 - f0) / (t1 - t0))
        uv[t0:t1] = 0
    if last < n - 1:                                     # interpolate.py:69-71 (from the last voiced frame on)
        out[last:] = signal[last]
        uv[last:] = 0
    return out, uv


def build_conditioner(cc, lf0, gv, speaker, min_cond, max_cond, norm_ind, look_ahead):
    """generate.py:146-194 for one utterance -> (n_frames, 43 * (1 + look_ahead)) float64."""
    cc = np.asarray(cc, dtype=np.float64)
    f0, _ = interpolation(lf0, F0_UNVOICED)
    fv, uv = interpolation(gv, FV_UNVOICED)
    cond = np.concatenate((cc, f0.reshape(-1, 1), fv.reshape(-1, 1), uv.reshape(-1, 1).astype(np.float64)), axis=1)
    min_cond, max_cond = np.asarray(min_cond), np.asarray(max_cond)
    if norm_ind:
        cond = (cond - min_cond[speaker]) / (max_cond[speaker] - min_cond[speaker])
    else:
        cond = (cond - min_cond) / (max_cond - min_cond)
    if look_ahead:
        delayed = np.copy(cond)
        delayed[:-1, :] = delayed[1:, :]
        cond = np.concatenate((cond, delayed), axis=1)
    return cond


def write_wav_f32(path, audio, sample_rate):
    """IEEE-float32 mono WAV, what ``librosa.output.write_wav`` produces for float input (generate.py:103-109)."""
    a = np.ascontiguousarray(np.asarray(audio, dtype="<f4"))
    data = a.tobytes()
    fmt = struct.pack("<HHIIHH", 3, 1, int(sample_rate), int(sample_rate) * 4, 4, 32)
    fact = struct.pack("<I", a.size)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact + b"data" + \
        struct.pack("<I", len(data)) + data
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


class BatchedFileGenerator:
    """``gen = BatchedFileGenerator(model, min_max, spk_ids, norm_ind, look_ahead); audio = gen(conds, speakers)``."""

    def __init__(self, model, min_max, spk_ids, norm_ind=True, look_ahead=True, sample_rate=16000, mode=L.MODE_BF16,
                 max_batch=256):
        self.model, self.sample_rate, self.norm_ind, self.look_ahead = model, sample_rate, bool(norm_ind), bool(look_ahead)
        self.min_cond, self.max_cond = np.asarray(min_max[0]), np.asarray(min_max[1])
        self.spk_ids = [str(s) for s in spk_ids]
        self.max_batch = int(max_batch)
        self.generator = Generator(model, cuda=True, mode=mode if model.dim % 64 == 0 else L.MODE_FP32)

    def speaker_index(self, name):                        # generate.py:163
        return self.spk_ids.index(str(name))

    def load(self, base, speaker_name):
        """Read ``base + '.cc' / '.lf0' / '.gv'`` (generate.py:146-160) -> (conditioner, speaker index)."""
        spk = self.speaker_index(speaker_name)
        cond = build_conditioner(np.loadtxt(base + ".cc"), np.loadtxt(base + ".lf0"), np.loadtxt(base + ".gv"), spk,
                                 self.min_cond, self.max_cond, self.norm_ind, self.look_ahead)
        return cond, spk

    @torch.no_grad()
    def __call__(self, conds, speakers, seed=None, uniforms=None):
        """conds: list of (n_frames_u, cond_dim) arrays (ragged); speakers: list of speaker indices.  Returns a list of
        float32 audio arrays of n_frames_u * lookback samples.  ``uniforms``: optional list of (n_frames_u * lookback,)
        arrays of pre-drawn U[0,1) for the defined sampler (per utterance, so results do not depend on the batching)."""
        if len(conds) != len(speakers):
            raise ValueError("one speaker per conditioner file (generate.py:139-141)")
        lookback, out = self.model.lookback, [None] * len(conds)
        order = sorted(range(len(conds)), key=lambda i: -len(conds[i]))          # similar lengths share a batch
        for lo in range(0, len(order), self.max_batch):
            ids = order[lo:lo + self.max_batch]
            n_max = max(len(conds[i]) for i in ids)
            B = len(ids)
            cond = np.zeros((B, n_max, self.model.cond_dim), dtype=np.float32)
            uni = None if uniforms is None else np.zeros((n_max * lookback, B), dtype=np.float32)
            for r, i in enumerate(ids):
                c = np.asarray(conds[i])
                if c.shape[1] != self.model.cond_dim:
                    raise ValueError("conditioner width %d, model expects %d" % (c.shape[1], self.model.cond_dim))
                cond[r, :len(c)] = c
                if uni is not None:
                    uni[:len(c) * lookback, r] = np.asarray(uniforms[i], dtype=np.float32)[:len(c) * lookback]
            spk = torch.tensor([int(speakers[i]) for i in ids], dtype=torch.int64)
            audio = self.generator(B, 0, torch.from_numpy(cond), spk, seed=seed,
                                   uniforms=None if uni is None else torch.from_numpy(uni))
            for r, i in enumerate(ids):
                out[i] = audio[r, :len(conds[i]) * lookback].numpy().astype(np.float32)
        return out

    def generate_files(self, bases, speaker_names, out_dir, tag="gen", seed=None):
        """File-to-file form of generate.py's main loop: returns the written paths
        (``<out_dir>/<tag>_file-<name>_spk-<speaker>.wav``, generate.py:96-98)."""
        loaded = [self.load(b, s) for b, s in zip(bases, speaker_names)]
        audio = self([c for c, _ in loaded], [s for _, s in loaded], seed=seed)
        os.makedirs(out_dir, exist_ok=True)
        paths = []
        for b, s, a in zip(bases, speaker_names, audio):
            p = os.path.join(out_dir, "%s_file-%s_spk-%s.wav" % (tag, os.path.basename(b), s))
            write_wav_f32(p, a, self.sample_rate)
            paths.append(p)
        return paths
