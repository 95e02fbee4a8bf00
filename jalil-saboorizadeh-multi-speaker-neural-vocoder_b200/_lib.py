"""ctypes binding of include/srnn_b200.h.  Loading fails loudly: there is no CPU or PyTorch fallback."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libsrnn_b200.so")

MAX_TIERS, MAX_RNN, Q, MAX_CHAIN = 4, 4, 256, 8
MODE_FP32, MODE_BF16, MODE_BF16_GRAPH, MODE_BF16X3 = 0, 1, 2, 3

f32p = C.POINTER(C.c_float)


class Config(C.Structure):
    _fields_ = [("n_tiers", C.c_int32), ("frame_sizes", C.c_int32 * MAX_TIERS), ("n_rnn", C.c_int32),
                ("dim", C.c_int32), ("q_levels", C.c_int32), ("cond_dim", C.c_int32), ("spk_dim", C.c_int32),
                ("ulaw", C.c_int32)]


class ConvParams(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("weight_g", C.c_void_p), ("weight_v", C.c_void_p), ("bias", C.c_void_p)]


class TierParams(C.Structure):
    _fields_ = [("h0", C.c_void_p), ("input_expand", ConvParams), ("cond_expand", ConvParams),
                ("spk_embedding", C.c_void_p), ("spk_expand", ConvParams),
                ("weight_ih", C.c_void_p * MAX_RNN), ("weight_hh", C.c_void_p * MAX_RNN),
                ("bias_ih", C.c_void_p * MAX_RNN), ("bias_hh", C.c_void_p * MAX_RNN),
                ("upsampling", ConvParams)]


class Params(C.Structure):
    _fields_ = [("tiers", TierParams * MAX_TIERS), ("embedding", C.c_void_p), ("mlp_input", ConvParams),
                ("mlp_hidden", ConvParams), ("mlp_output", ConvParams)]


class CondChain(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (MAX_CHAIN + 1)), ("layers", ConvParams * MAX_CHAIN)]


_SIGNATURES = {
    "srnn_cond_chain_fwd": (C.c_int, [C.POINTER(CondChain), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srnn_last_error": (C.c_char_p, []),
    "srnn_version": (C.c_int, []),
    "srnn_sample_kernel_name": (C.c_char_p, []),
    "srnn_launch_count": (C.c_int64, []),
    "srnn_graph_reuse_count": (C.c_int64, [C.c_void_p]),
    "srnn_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "srnn_destroy": (C.c_int, [C.c_void_p]),
    "srnn_lookback": (C.c_int, [C.c_void_p]),
    "srnn_pack_weights": (C.c_int, [C.c_void_p, C.POINTER(Params), C.c_void_p]),
    "srnn_predict_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "srnn_predict_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Params), C.POINTER(Params), C.c_void_p]),
    "srnn_clamp_adam_step": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_int32, C.c_float, C.c_void_p]),
    "srnn_predict_bwd_nll": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Params), C.POINTER(Params),
                                       C.c_void_p]),
    "srnn_clamp_adam_step_scaled": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                              C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_float, C.c_float, C.c_float,
                                              C.c_float, C.c_int32, C.c_float, C.c_float, C.c_void_p]),
    "srnn_mlp_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "srnn_tier_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "srnn_gru_seq_fwd": (C.c_int, [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 7 + [C.c_int32, C.c_void_p]),
    "srnn_gru_seq_bwd": (C.c_int, [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 9 + [C.c_int32, C.c_void_p]),
    "srnn_quantize": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "srnn_bwd_wait_early": (C.c_int, [C.c_void_p, C.c_void_p]),
    "srnn_bwd_wait_stage": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "srnn_timed_kernel": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "srnn_nll_loss_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "srnn_generate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "srnn_sample_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "srnn_dequant_lut": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "srnn_gemm": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
}

_lib = None


class SrnnError(RuntimeError):
    pass


def load():
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SrnnError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise SrnnError(f"srnn error {rc}: {load().srnn_last_error().decode()}")


def exported_symbols():
    return sorted(_SIGNATURES)
