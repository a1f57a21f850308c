"""ctypes binding of libnca_b200.so (the C ABI in include/nca_b200.h).

There is no CPU fallback and no pure-PyTorch fallback: if the library is missing or no CUDA device is
present every compute call raises ``NcaError``.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NCA_PAD = {"constant": 0, "zeros": 0, "circular": 1, "replicate": 2, "reflect": 3}
NCA_COND_NONE, NCA_COND_CPE, NCA_COND_TENSOR = 0, 1, 2
NCA_PREC = {"fp32": 0, "bf16": 1, "f16x3": 2}
NCA_MASK_SUPPLIED, NCA_MASK_PHILOX = 0, 1


class NcaError(RuntimeError):
    pass


class DyncaDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("fc", C.c_int32),
                ("cond_kind", C.c_int32), ("cc", C.c_int32), ("pad_mode", C.c_int32), ("n_scales", C.c_int32),
                ("precision", C.c_int32), ("mask_mode", C.c_int32), ("update_rate", C.c_float)]


class DyncaWeights(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p)]


class EncDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("hid", C.c_int32),
                ("living_dim", C.c_int32), ("mask_mode", C.c_int32), ("alive_thr", C.c_float),
                ("fire_rate", C.c_float), ("clamp", C.c_float), ("precision", C.c_int32)]


class EncWeights(C.Structure):
    _fields_ = [("wp", C.c_void_p), ("wa", C.c_void_p), ("ba", C.c_void_p), ("wb", C.c_void_p), ("bb", C.c_void_p),
                ("wc", C.c_void_p)]


# every symbol include/nca_b200.h declares (tests check the library exports all of them)
SYMBOLS = ["nca_last_error", "nca_abi_version", "nca_launch_count", "nca_launch_count_reset", "nca_dynca_perceive",
           "nca_edge_extract", "nca_dynca_forward", "nca_dynca_backward", "nca_dynca_workspace_bytes", "nca_dynca_op_hist_bytes",
           "nca_dynca_kernel_variant",
           "nca_philox_mask", "nca_philox_mask_at", "nca_enc_forward", "nca_enc_backward", "nca_enc_workspace_bytes",
           "nca_encoder_forward", "nca_encoder_backward",
           "nca_pool_gather", "nca_pool_dead_flags", "nca_pool_scatter", "nca_normalized_adam_step", "nca_overflow_workspace_bytes", "nca_overflow_loss",
           "nca_frame_to_cond_channel", "nca_state_to_rgb8"]


def lib_path():
    return os.path.join(HERE, "libnca_b200.so")


def load_library():
    """Load (once) the in-tree CUDA library.  Raises NcaError when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise NcaError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). There is no CPU / PyTorch fallback for the NCA step.")
    lib = C.CDLL(path)
    P, I, U64, F, SZ = C.c_void_p, C.c_int32, C.c_uint64, C.c_float, C.c_size_t
    lib.nca_last_error.restype = C.c_char_p
    lib.nca_abi_version.restype = C.c_int
    lib.nca_launch_count.restype = C.c_longlong
    lib.nca_launch_count_reset.restype = None
    lib.nca_dynca_workspace_bytes.restype = SZ
    lib.nca_dynca_workspace_bytes.argtypes = [C.POINTER(DyncaDesc), I]
    lib.nca_dynca_perceive.argtypes = [C.POINTER(DyncaDesc), P, P, P, P]
    lib.nca_edge_extract.argtypes = [C.c_int, C.c_int, C.c_int, P, C.c_int, P, P]
    lib.nca_philox_mask.argtypes = [I, I, I, F, I, U64, I, I, P, P]
    lib.nca_philox_mask_at.argtypes = [I, I, I, F, I, U64, I, P, I, P, P]
    lib.nca_dynca_op_hist_bytes.restype = SZ
    lib.nca_dynca_op_hist_bytes.argtypes = [C.POINTER(DyncaDesc), I]
    lib.nca_dynca_forward.argtypes = [C.POINTER(DyncaDesc), C.POINTER(DyncaWeights), P, P, U64, I, I, I, P, P, P, P, SZ, P]
    lib.nca_dynca_kernel_variant.argtypes = [C.POINTER(DyncaDesc), I]
    lib.nca_dynca_backward.argtypes = [C.POINTER(DyncaDesc), C.POINTER(DyncaWeights), P, P, U64, I, I, P, P, P, P,
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_int32), I, I, F, P,
                                       C.POINTER(DyncaWeights), P, SZ, P]
    lib.nca_enc_workspace_bytes.restype = SZ
    lib.nca_enc_workspace_bytes.argtypes = [C.POINTER(EncDesc), I]
    lib.nca_enc_forward.argtypes = [C.POINTER(EncDesc), C.POINTER(EncWeights), P, P, U64, I, I, I, P, P, P, SZ, P]
    lib.nca_enc_backward.argtypes = [C.POINTER(EncDesc), C.POINTER(EncWeights), P, P, U64, I, I, P, P, P, P, P,
                                     C.POINTER(EncWeights), P, SZ, P]
    lib.nca_encoder_forward.argtypes = [I, I, I, I, I, P, P, P, P, P, P, P, I, P]
    lib.nca_encoder_backward.argtypes = [I, I, I, I, I, P, P, P, P, I, P, P, P, P]
    PP = C.POINTER(C.c_void_p)
    lib.nca_pool_gather.argtypes = [I, I, I, I, P, P, I, P, I, P, I, P, P, P]
    lib.nca_pool_dead_flags.argtypes = [I, I, I, I, P, P, I, I, F, P, P]
    lib.nca_pool_scatter.argtypes = [I, I, I, I, P, P, I, P, I, P]
    lib.nca_normalized_adam_step.argtypes = [I, PP, PP, PP, PP, C.POINTER(C.c_int64), I, F, F, F, F, F, I, I, P]
    lib.nca_overflow_workspace_bytes.restype = SZ
    lib.nca_overflow_workspace_bytes.argtypes = []
    lib.nca_overflow_loss.argtypes = [P, SZ, P, P, F, P, I, P, SZ, P]
    lib.nca_frame_to_cond_channel.argtypes = [I, I, I, I, P, P, I, P]
    lib.nca_state_to_rgb8.argtypes = [I, I, I, I, P, F, P, P]
    if lib.nca_abi_version() != 6:
        raise NcaError("libnca_b200.so ABI version mismatch")
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load_library().nca_last_error().decode("utf-8", "replace")
        raise NcaError(f"libnca_b200 error {rc}: {msg}")
