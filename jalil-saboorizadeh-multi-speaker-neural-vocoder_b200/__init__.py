"""B200-native SampleRNN hot path (generation + teacher-forced step) behind the reference ``model.py`` API.

The directory name carries hyphens (it mirrors the reference repository's name), so import it with
``importlib.import_module("jalil-saboorizadeh-multi-speaker-neural-vocoder_b200")`` or through the root-level
alias module ``srnn_b200``.
"""
from . import _lib
from ._lib import MODE_BF16, MODE_BF16_GRAPH, MODE_BF16X3, MODE_FP32, SrnnError
from .optim import ClampAdam, gradient_clipping, sequence_nll_loss_bits
from .sharding import shard_batch, shard_range
from .data import TBPTTBatcher, quantize
from .checkpoint import CheckpointSaver, load_last_checkpoint, log_line, make_tag, parse_checkpoint_name
from .frontend import BatchedFileGenerator, build_conditioner, interpolation, write_wav_f32
from .bottleneck import BottleneckConditioner, BottleneckGenerator, BottleneckPredictor, BottleneckSampleRNN
from .model import FrameLevelRNN, Generator, LearnedUpsampling1d, Predictor, Runner, SampleLevelMLP, SampleRNN

HAS_BF16 = True      # the library carries the tcgen05 (SRNN_MODE_BF16) path; it needs dim % 64 == 0

__all__ = ["SampleRNN", "Predictor", "Generator", "Runner", "FrameLevelRNN", "SampleLevelMLP",
           "LearnedUpsampling1d", "BottleneckConditioner", "BottleneckSampleRNN", "BottleneckPredictor", "BottleneckGenerator", "shard_range", "shard_batch", "TBPTTBatcher", "quantize", "CheckpointSaver", "load_last_checkpoint", "log_line", "make_tag", "parse_checkpoint_name", "BatchedFileGenerator", "build_conditioner", "interpolation", "write_wav_f32", "ClampAdam", "gradient_clipping", "sequence_nll_loss_bits", "MODE_FP32", "MODE_BF16", "MODE_BF16_GRAPH", "MODE_BF16X3", "SrnnError", "_lib"]
