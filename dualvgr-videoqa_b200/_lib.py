"""ctypes binding of libdualvgr_b200.so (C ABI: include/dualvgr_b200.h).

The library is the product: if it is missing this module raises at import time — there is no CPU or PyTorch
fallback for any op it exports."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdualvgr_b200.so")


class DvgrError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or tools/build_lib.sh). There is no fallback path.")

lib = ctypes.CDLL(LIB_PATH)

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


class Operand(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("major", c_int), ("ndim", c_int), ("dims", c_ll * 4), ("strides", c_ll * 4)]


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", Operand), ("B", Operand),
        ("M", c_int), ("N", c_int), ("K", c_int), ("batch", c_int),
        ("a_c0", c_int * 4), ("a_c2", c_int * 4), ("a_c3", c_int * 4),
        ("b_c0", c_int * 4), ("b_c2", c_int * 4), ("b_c3", c_int * 4),
        ("k_inner", c_int), ("a_c2_step", c_int * 4), ("b_c2_step", c_int * 4),
        ("C", c_void_p), ("ldc", c_ll), ("c_batch", c_ll),
        ("out_f32", c_int), ("act", c_int), ("beta", c_int),
        ("bias", c_void_p), ("bias_batch", c_ll), ("row_map", c_void_p),
        ("bn", c_int), ("max_ctas", c_int),
    ]


class LstmArgs(ctypes.Structure):
    _fields_ = [
        ("S", c_int), ("H", c_int), ("T", c_int), ("ndir", c_int), ("s", c_int),
        ("gates", c_void_p), ("whh", c_void_p), ("h_hist", c_void_p), ("c_hist", c_void_p),
        ("h_last", c_void_p), ("h_last_ld", c_ll), ("seq_len", c_void_p),
        ("seq_out", c_void_p), ("seq_out_ld", c_ll),
        ("dc", c_void_p), ("dh_last", c_void_p), ("dh_last_ld", c_ll), ("dh_seq", c_void_p),
    ]


lib.dvgr_last_error.restype = ctypes.c_char_p
lib.dvgr_abi_version.restype = c_int
lib.dvgr_launch_count.restype = c_ll


def check(rc, what=""):
    if rc != 0:
        raise DvgrError(f"{what}: {lib.dvgr_last_error().decode()}")


def launch_count():
    return int(lib.dvgr_launch_count())


def _sig(name, argtypes):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = c_int
    return fn


gemm = _sig("dvgr_gemm", [ctypes.POINTER(GemmArgs), c_void_p])
gemm_reference = _sig("dvgr_gemm_reference",
                      [c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p])
lstm_step_fwd = _sig("dvgr_lstm_step_fwd", [ctypes.POINTER(LstmArgs), c_void_p])
lstm_step_bwd = _sig("dvgr_lstm_step_bwd", [ctypes.POINTER(LstmArgs), c_void_p])
