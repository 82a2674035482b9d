"""ctypes binding of libmjhmc_b200.so (include/mjhmc_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load, every
entry point of the package raises.  ``load()`` builds the library in-tree with nvcc
when it is absent and nvcc is available.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MJHMC_B200_LIB") or os.path.join(HERE, "libmjhmc_b200.so")

F32, F64 = 0, 1
DIST_TEST_GAUSSIAN, DIST_DIAG_GAUSSIAN, DIST_ROUGH_WELL, DIST_FUNNEL, DIST_FUNNEL_LITERAL, \
    DIST_DENSE_GAUSSIAN, DIST_PRODUCT_OF_T, DIST_MULTIMODAL = range(8)
SAMPLER_DISCRETE, SAMPLER_CONTINUOUS_TIME, SAMPLER_MARKOV_JUMP = range(3)
RNG_PHILOX, RNG_INJECT = 0, 1
RNG_FLAG_LITERAL_RACE, RNG_FLAG_REGISTER_STATE = 1, 2
CNT_L, CNT_F, CNT_FL, CNT_R, CNT_E, CNT_DEDX, CNT_FAIL, CNT_EXEC = range(8)
N_COUNTERS = 8
COUNTER_STRIPES = 32
COUNTER_ROWS = COUNTER_STRIPES + 1
INT64_MAX = (1 << 63) - 1
ABI_VERSION = 2


class Dist(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dtype", C.c_int32), ("ndims", C.c_int32), ("nbasis", C.c_int32),
                ("p", C.c_double * 4), ("a0", C.c_void_p), ("a1", C.c_void_p), ("a2", C.c_void_p),
                ("ws", C.c_void_p)]


class HP(C.Structure):
    _fields_ = [("sampler", C.c_int32), ("num_leapfrog_steps", C.c_int32), ("epsilon", C.c_double),
                ("beta", C.c_double), ("p_flip", C.c_double), ("p_r", C.c_double)]


class RNG(C.Structure):
    _fields_ = [("mode", C.c_int32), ("flags", C.c_int32), ("seed", C.c_uint64), ("attempt0", C.c_uint64),
                ("particle0", C.c_uint64), ("Z", C.c_void_p), ("U", C.c_void_p), ("U0", C.c_void_p),
                ("inj_ld", C.c_int64)]


class State(C.Structure):
    _fields_ = [("X", C.c_void_p), ("V", C.c_void_p), ("H_cache", C.c_void_p), ("cache_active", C.c_void_p),
                ("n", C.c_int64), ("ld", C.c_int64)]


class FullState(C.Structure):
    _fields_ = [("X", C.c_void_p), ("V", C.c_void_p), ("G", C.c_void_p), ("EX", C.c_void_p), ("EV", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [("samples", C.c_void_p), ("stride_k", C.c_int64), ("stride_it", C.c_int64),
                ("dwell", C.c_void_p), ("dwell_last", C.c_void_p), ("choice", C.c_void_p),
                ("counters", C.c_void_p), ("energy", C.c_void_p)]


# every symbol include/mjhmc_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "mjhmc_last_error": (C.c_char_p, []),
    "mjhmc_abi_version": (C.c_int, []),
    "mjhmc_fused_supported": (C.c_int, [_P(Dist)]),
    "mjhmc_sample_fused": (C.c_int, [_P(Dist), _P(HP), _P(RNG), _P(State), _P(State), C.c_int32, _P(Outputs),
                                     C.c_void_p]),
    "mjhmc_stream_supported": (C.c_int, [_P(Dist)]),
    "mjhmc_sample_stream": (C.c_int, [_P(Dist), _P(HP), _P(RNG), _P(State), _P(State), C.c_int32, _P(Outputs),
                                      C.c_void_p]),
    "mjhmc_stream_set_tma": (None, [C.c_int32]),
    "mjhmc_stream_last_launch": (None, [_P(C.c_int64)]),
    "mjhmc_stream_probe_blocks": (C.c_int, [C.c_int64]),
    "mjhmc_energy": (C.c_int, [_P(Dist), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "mjhmc_gradient": (C.c_int, [_P(Dist), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "mjhmc_kinetic": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "mjhmc_kick_drift": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                   C.c_double, C.c_void_p]),
    "mjhmc_kick": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double,
                             C.c_void_p]),
    "mjhmc_transition": (C.c_int, [C.c_int32, C.c_int32, _P(HP), _P(RNG), C.c_int64, C.c_int64, _P(FullState),
                                   _P(FullState), C.c_void_p, C.c_void_p, C.c_void_p, _P(Outputs), C.c_void_p]),
    "mjhmc_dense_tc_workspace_bytes": (C.c_int64, [_P(Dist)]),
    "mjhmc_dense_tc_prepare": (C.c_int, [_P(Dist), C.c_void_p]),
    "mjhmc_counters_read": (C.c_int, [C.c_void_p, _P(C.c_int64), C.c_void_p]),
    "mjhmc_counters_reset": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mjhmc_resample_scratch_bytes": (C.c_int64, [C.c_int64]),
    "mjhmc_resample": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mjhmc_autocorr": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mjhmc_autocorr_fft_scratch_bytes": (C.c_int64, [C.c_int32]),
    "mjhmc_autocorr_fft": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                     C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mjhmc_ladder_visits": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mjhmc_moments": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load(build_if_missing=True):
    """Returns the loaded library; raises NativeLibraryError if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise NativeLibraryError("%s is missing (run `python -m mjhmc_b200.build`)" % LIB_PATH)
        from . import build as _build
        try:
            _build.build(verbose=False)
        except Exception as exc:   # noqa: BLE001 - surfaced verbatim
            raise NativeLibraryError("libmjhmc_b200.so is missing and could not be built: %s" % exc)
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:
        raise NativeLibraryError("cannot load %s: %s (there is no CPU fallback)" % (LIB_PATH, exc))
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise NativeLibraryError("%s does not export %s" % (LIB_PATH, name))
        fn.restype = res
        fn.argtypes = args
    if lib.mjhmc_abi_version() != ABI_VERSION:
        raise NativeLibraryError("ABI version mismatch: library %d, binding %d" % (lib.mjhmc_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().mjhmc_last_error()
        raise RuntimeError("mjhmc_b200 native call failed%s: %s" % (" in " + what if what else "",
                                                                   msg.decode() if msg else rc))
