"""ctypes binding of libeig.so (include/eig.h).  No torch types cross this boundary: only pointers and sizes.

`get_library()` is the product entry: it loads the in-tree `libeig.so` built by csrc/build.sh and raises if the
library is missing - there is no CPU fallback.  (`EigLibrary(path)` with an explicit path exists so the
`-m "not gpu"` tests can bind the g++-built kernel-source emulator under tests/emu; the product never does.)
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeig.so")

EIG_OK, EIG_E_INVALID, EIG_E_CUDA, EIG_E_STATE, EIG_E_CAPACITY, EIG_E_NODEVICE, EIG_E_RANGE = 0, -1, -2, -3, -4, -5, -6
CONV_SIMT, CONV_TC = 0, 1
PAIR_POPULATION, PAIR_SINGLE_IMAGE = 0, 1

# every symbol include/eig.h declares: (name, restype, argtypes)
_P, _I, _D = C.c_void_p, C.c_int, C.c_double
SYMBOLS = [
    ("eig_last_error", C.c_char_p, []),
    ("eig_error", C.c_char_p, [_P]),
    ("eig_set_option", _I, [_P, C.c_char_p, _I]),
    ("eig_version", _I, []),
    ("eig_launch_count", C.c_int64, []),
    ("eig_create", _I, [C.POINTER(_P), _I, _I, _I, _I, C.POINTER(_I), _I]),
    ("eig_create_render", _I, [C.POINTER(_P), _I, _I, _I, _I, _I]),
    ("eig_destroy", None, [_P]),
    ("eig_set_conv_mode", _I, [_P, _I]),
    ("eig_load_weights", _I, [_P, _I, C.POINTER(C.c_char_p), C.POINTER(_P), C.POINTER(C.c_int64)]),
    ("eig_set_grid", _I, [_P, _P, _P]),
    ("eig_cppn_render", _I, [_P, _P, _P, _I, _I, _I, _I, _D, _P, _P, _P]),
    ("eig_prednet_run", _I, [_P, _P, _I, _I, _I, _P, _P]),
    ("eig_prednet_reset", _I, [_P, _I, _P]),
    ("eig_prednet_forward", _I, [_P, _P, _I, _P, _P, _P]),
    ("eig_flow", _I, [_P, _P, _P, _I, _P, _P, _P, _P, _P]),
    ("eig_score", _I, [_P, _P, _P, _I, _I, _P, _P]),
    ("eig_eval", _I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    ("eig_eval_host", _I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    ("eig_range_check", _I, [_P, _P]),
    ("eig_debug_buffers", _I, [_P] + [C.POINTER(_P)] * 6),
    ("eig_memcpy_d2h", _I, [_P, _P, C.c_int64]),
    ("eig_profile_begin", _I, [_P]),
    ("eig_profile_end", _I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
]


class EigError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libeig error %d: %s" % (code, msg))
        self.code = code


class EigLibrary:
    def __init__(self, path):
        if not os.path.isfile(path):
            raise FileNotFoundError(
                "%s not found: build it with evolutionary_illusion_generator_b200/csrc/build.sh "
                "(or __graft_entry__.build()); the engine has no CPU fallback" % path)
        self.path = path
        self.dll = C.CDLL(path)
        for name, res, args in SYMBOLS:
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, rc):
        if rc != 0:
            raise EigError(rc, self.eig_last_error().decode("utf-8", "replace"))


_lib = None


def get_library():
    global _lib
    if _lib is None:
        _lib = EigLibrary(LIB_PATH)
    return _lib
