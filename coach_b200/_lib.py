"""ctypes binding of libcoach_b200.so (include/coach_b200.h).

There is NO fallback: if the CUDA library is missing or does not load, every compute path of ``coach_b200``
raises.  Build it with ``python -m coach_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcoach_b200.so")

c_void_p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_double = ctypes.c_double
c_float = ctypes.c_float

CB200_MAX_COLUMNS = 8


class Column(ctypes.Structure):
    """struct cb200_column"""
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("row_bytes", c_i64)]


# name -> (restype, argtypes); every symbol of include/coach_b200.h (tests/test_cabi.py checks the two stay in sync)
PROTOTYPES = {
    "cb200_abi_version": (c_int, []),
    "cb200_last_error": (ctypes.c_char_p, []),
    "cb200_launch_count": (c_i64, []),
    "cb200_device_info": (c_int, [ctypes.POINTER(c_int)] * 3),
    "cb200_tune": (c_int, [ctypes.c_char_p, c_int]),
    "cb200_per_init": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "cb200_per_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64,
                                 c_void_p, c_void_p]),
    "cb200_per_priorities_device": (c_int, [c_void_p, c_i64, c_double, c_double, c_void_p, c_void_p, c_void_p,
                                            c_void_p]),
    "cb200_host_priorities": (c_int, [c_void_p, c_i64, c_double, c_double, c_void_p, c_void_p]),
    "cb200_per_store": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_double, c_double,
                                c_void_p]),
    "cb200_per_sample": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i64, c_double, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "cb200_gather": (c_int, [ctypes.POINTER(Column), c_int, c_void_p, c_i64, c_void_p]),
    "cb200_per_sample_gather": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i64, c_double, c_void_p,
                                        c_void_p, c_void_p, ctypes.POINTER(Column), c_int, c_void_p]),
    "cb200_scatter_ring": (c_int, [ctypes.POINTER(Column), c_int, c_i64, c_i64, c_i64, c_void_p]),
}

_lib = None


class CoachB200Error(RuntimeError):
    pass


def load():
    """Loads the library once; raises ImportError with build instructions when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "coach_b200: %s is missing. The CUDA library is mandatory (there is no CPU path); build it with "
            "`python -m coach_b200.build`." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError here = header / library mismatch: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.cb200_abi_version() != 1:
        raise ImportError("coach_b200: ABI version mismatch between _lib.py and libcoach_b200.so")
    _lib = lib
    return lib


def check(rc):
    """Maps a C-ABI return code to the exception the reference would raise for the same condition."""
    if rc == 0:
        return
    msg = load().cb200_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg)
    raise CoachB200Error(msg)


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def make_columns(pairs):
    """pairs: iterable of (src_ptr, dst_ptr, row_bytes) -> (ctypes array, count)."""
    pairs = list(pairs)
    if len(pairs) > CB200_MAX_COLUMNS:
        raise ValueError("at most %d columns per call" % CB200_MAX_COLUMNS)
    arr = (Column * len(pairs))()
    for k, (s, d, rb) in enumerate(pairs):
        arr[k].src, arr[k].dst, arr[k].row_bytes = s, d, rb
    return arr, len(pairs)
