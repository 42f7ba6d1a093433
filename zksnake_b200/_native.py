"""ctypes binding of libzkb200.so (the C ABI in include/zkb200.h).

There is no CPU fallback: importing works anywhere (so that host-only logic can be tested), but every compute
entry point raises unless a B200 is present and `ensure_init()` succeeded.  The library is looked up in-tree
(zksnake_b200/libzkb200.so, built by `__graft_entry__.build()` / `make -C zksnake_b200/csrc`).
"""
import ctypes
import operator
import os

import numpy as np

BN254, BLS12_381 = 0, 1
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzkb200.so")

ERR_CUDA, ERR_ARG, ERR_MISMATCH, ERR_DOMAIN, ERR_NOT_DIVISIBLE, ERR_NOINIT, ERR_POINT = -1, -2, -3, -4, -5, -6, -7


class ZkbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(zksnake_b200 has no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    c_sz, c_vp, c_int, c_u32 = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32
    sig = {
        "zkb_device_count": (c_int, []),
        "zkb_init": (c_int, [c_int]),
        "zkb_shutdown": (None, []),
        "zkb_last_error": (ctypes.c_char_p, []),
        "zkb_stream": (c_vp, []),
        "zkb_sync": (c_int, []),
        "zkb_launch_count": (ctypes.c_ulonglong, []),
        "zkb_transfer_count": (None, [ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)]),
        "zkb_timer_start": (c_int, []),
        "zkb_timer_stop": (c_int, [ctypes.POINTER(ctypes.c_float)]),
        "zkb_prof_enable": (c_int, [c_int]),
        "zkb_prof_read": (c_int, [c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_ulonglong)]),
        "zkb_imad_peak": (c_int, [c_int, ctypes.POINTER(ctypes.c_double)]),
        "zkb_msm_kernel_info": (c_int, [c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
        "zkb_dev_alloc": (c_int, [c_sz, ctypes.POINTER(c_vp)]),
        "zkb_dev_free": (c_int, [c_vp]),
        "zkb_host_alloc": (c_int, [c_sz, ctypes.POINTER(c_vp)]),
        "zkb_host_free": (c_int, [c_vp]),
        "zkb_h2d": (c_int, [c_vp, c_vp, c_sz]),
        "zkb_h2d_async": (c_int, [c_vp, c_vp, c_sz]),
        "zkb_d2h": (c_int, [c_vp, c_vp, c_sz]),
        "zkb_d2d": (c_int, [c_vp, c_vp, c_sz]),
        "zkb_memset": (c_int, [c_vp, c_int, c_sz]),
        "zkb_ntt": (c_int, [c_int, c_int, c_int, c_u32, c_vp, c_sz, c_vp]),
        "zkb_ntt_dev": (c_int, [c_int, c_int, c_int, c_u32, c_vp, c_sz, c_vp]),
        "zkb_vec_op": (c_int, [c_int, c_int, c_sz, c_vp, c_sz, c_vp, c_sz, c_vp]),
        "zkb_vec_op_dev": (c_int, [c_int, c_int, c_sz, c_vp, c_sz, c_vp, c_sz, c_vp]),
        "zkb_fr_reduce": (c_int, [c_int, c_sz, c_vp]),
        "zkb_fr_reduce_dev": (c_int, [c_int, c_sz, c_vp]),
        "zkb_fr_powers_dev": (c_int, [c_int, c_vp, c_vp, c_sz, c_vp]),
        "zkb_fr_axpy_dev": (c_int, [c_int, c_sz, c_vp, c_vp, c_sz, c_vp, c_sz, c_vp]),
        "zkb_fr_mul_powers_dev": (c_int, [c_int, c_sz, c_vp, c_vp, c_vp, c_vp]),
        "zkb_fr_inverse_dev": (c_int, [c_int, c_sz, c_vp, c_vp]),
        "zkb_fr_scan_dev": (c_int, [c_int, c_int, c_sz, c_vp, c_vp]),
        "zkb_fr_gather_dev": (c_int, [c_int, c_sz, c_vp, c_sz, c_sz, c_vp]),
        "zkb_fr_gather_index_dev": (c_int, [c_int, c_sz, c_vp, c_vp, c_vp]),
        "zkb_fr_eval_dev": (c_int, [c_int, c_sz, c_vp, c_vp, c_vp]),
        "zkb_fr_trim_dev": (c_int, [c_int, c_sz, c_vp, ctypes.POINTER(c_sz)]),
        "zkb_fr_div_vanishing_dev": (c_int, [c_int, c_sz, c_sz, c_vp, c_vp, ctypes.POINTER(c_int)]),
        "zkb_plonk_quotient_dev": (c_int, [c_int, c_sz, c_sz, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "zkb_fr_add_sparse_dev": (c_int, [c_int, c_vp, c_sz, c_vp, c_vp, c_int]),
        "zkb_affine_bytes": (c_sz, [c_int, c_int]),
        "zkb_points_upload": (c_int, [c_int, c_int, c_vp, c_sz, c_vp]),
        "zkb_points_download": (c_int, [c_int, c_int, c_vp, c_sz, c_vp]),
        "zkb_compressed_bytes": (c_sz, [c_int, c_int]),
        "zkb_points_compress": (c_int, [c_int, c_int, c_vp, c_sz, c_vp]),
        "zkb_points_decompress": (c_int, [c_int, c_int, c_vp, c_sz, c_int, c_vp, c_vp, c_vp]),
        "zkb_msm": (c_int, [c_int, c_int, c_vp, c_sz, c_vp, c_sz, c_vp, ctypes.POINTER(c_int)]),
        "zkb_msm_dev": (c_int, [c_int, c_int, c_vp, c_vp, c_sz, c_vp, ctypes.POINTER(c_int)]),
        "zkb_msm_table_create": (c_int, [c_int, c_int, c_vp, c_sz, c_u32, c_u32, ctypes.POINTER(c_vp)]),
        "zkb_msm_table_free": (None, [c_vp]),
        "zkb_msm_table_info": (c_int, [c_vp, ctypes.POINTER(c_u32), ctypes.POINTER(c_u32), ctypes.POINTER(c_sz)]),
        "zkb_msm_table_dev": (c_int, [c_vp, c_vp, c_sz, c_u32, c_u32, c_vp, ctypes.POINTER(c_int)]),
        "zkb_msm_table_batch_dev": (c_int, [c_vp, c_int, c_vp, c_vp, c_u32, c_u32, c_vp, c_vp]),
        "zkb_groth16_pk_build_tables": (c_int, [c_vp, c_u32]),
        "zkb_msm_dev_windows": (c_int, [c_int, c_int, c_vp, c_vp, c_sz, c_u32, c_u32, c_vp, ctypes.POINTER(c_int)]),
        "zkb_groth16_pk_set_window_shard": (c_int, [c_vp, c_u32, c_u32]),
        "zkb_groth16_pk_set_kw_windows": (c_int, [c_vp, c_u32, c_u32, c_int]),
        "zkb_groth16_pk_msm_info": (c_int, [c_vp, c_int, c_vp, c_vp]),
        "zkb_groth16_precompute": (c_int, [c_vp, c_vp, c_vp]),
        "zkb_batch_mul_dev": (c_int, [c_int, c_int, c_vp, c_int, c_vp, c_sz, c_vp]),
        "zkb_msm_set_tuning": (None, [c_int, c_int, c_int]),
        "zkb_groth16_h": (c_int, [c_int, c_u32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "zkb_groth16_h_dev": (c_int, [c_int, c_u32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int]),
        "zkb_groth16_pk_create": (c_int, [c_int, c_u32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp, c_vp,
                                          ctypes.POINTER(c_vp)]),
        "zkb_groth16_pk_create_sharded": (c_int, [c_int, c_u32, c_vp, c_vp, c_vp, c_sz, c_sz, c_vp, c_sz, c_sz, c_sz, c_vp, c_vp,
                                                  c_vp, c_vp, c_vp, ctypes.POINTER(c_vp)]),
        "zkb_groth16_pk_free": (None, [c_vp]),
        "zkb_groth16_prove": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "zkb_groth16_prove_dev": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "zkb_r1cs_create": (c_int, [c_int, c_sz, c_sz, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp)]),
        "zkb_r1cs_free": (None, [c_vp]),
        "zkb_r1cs_eval": (c_int, [c_vp, c_vp, c_sz, c_vp, c_vp, c_vp]),
        "zkb_r1cs_eval_dev": (c_int, [c_vp, c_vp, c_sz, c_vp, c_vp, c_vp]),
        "zkb_groth16_prove_witness": (c_int, [c_vp, c_vp, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "zkb_groth16_prove_witness_dev": (c_int, [c_vp, c_vp, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "zkb_groth16_partial": (c_int, [c_vp, c_vp, c_vp, c_int, c_sz, c_vp, c_vp]),
        "zkb_groth16_spread_begin": (c_int, [c_vp, c_vp, c_vp, c_int, c_sz, ctypes.c_uint, c_vp, c_vp]),
        "zkb_groth16_spread_quotient": (c_int, [c_vp, c_vp, c_vp]),
        "zkb_groth16_spread_finish": (c_int, [c_vp, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "zkb_groth16_assemble": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "zkb_groth16_assemble_partials": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "zkb_groth16_last_poly": (c_int, [c_vp, c_int, c_vp]),
        "zkb_groth16_last_msm": (c_int, [c_vp, c_int, c_vp, ctypes.POINTER(c_int)]),
        "zkb_test_field_op_host": (c_int, [c_int, c_int, c_sz, c_vp, c_vp, c_vp]),
        "zkb_test_field_op_dev": (c_int, [c_int, c_int, c_sz, c_vp, c_vp, c_vp]),
        "zkb_point_lincomb": (c_int, [c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.POINTER(c_int)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTED = _load()
_initialised = False


def last_error():
    return (lib.zkb_last_error() or b"").decode()


def check(rc):
    """Map C-ABI codes to the exception types the reference raises."""
    if rc == 0:
        return
    msg = last_error()
    if rc in (ERR_ARG, ERR_MISMATCH, ERR_DOMAIN, ERR_NOT_DIVISIBLE, ERR_POINT):
        raise ValueError(msg)
    raise ZkbError(msg or f"libzkb200 error {rc}")


def ensure_init(device=None):
    """Bind this process to one GPU (LOCAL_RANK by default).  Raises when no B200 is visible."""
    global _initialised
    if _initialised and device is None:
        return
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
        n = lib.zkb_device_count()
        if n > 0:
            device %= n
    check(lib.zkb_init(device))
    _initialised = True


def gpu_available():
    return lib.zkb_device_count() > 0


# ---- numpy <-> Python int helpers (little-endian limbs) --------------------------------------------------------
try:
    from . import _marshal   # CPython extension built next to libzkb200.so (csrc/pymarshal.cpp)
except ImportError as exc:   # fail loudly: there is no pure-Python marshalling path
    raise ImportError(f"zksnake_b200._marshal is missing ({exc}): build it with "
                      "`python -c 'import __graft_entry__ as g; g.build()'`") from exc


def ints_to_limbs(values, nbytes=32, modulus=None, out=None, item=-1, allow_negative=True):
    """list[int] -> uint64 array of shape (len, nbytes/8), read straight from the PyLong digits by host threads
    (csrc/pymarshal.cpp; the reference does this one BigUint at a time: src/bn254/polynomial.rs:537-540).  Values must be in
    [0, 2^(8 nbytes)) unless `modulus` is given, in which case negative / oversized values are reduced mod it.  `out`: an
    existing C-contiguous uint64 array of that shape (e.g. a view of pinned host memory) to fill instead of a new one.
    item >= 0: the elements are tuples and the integer is element[item] (the (coeff, terms) pairs the reference's Polynomial
    constructor passes, python/zksnake/polynomial.py:40-43).  allow_negative=False: a negative value raises OverflowError (what
    pyo3's BigUint extraction does) instead of being reduced."""
    if not isinstance(values, (list, tuple)):
        values = list(values)
    k = nbytes // 8
    arr = out if out is not None else np.empty((len(values), k), dtype=np.uint64)
    assert arr.dtype == np.uint64 and arr.flags["C_CONTIGUOUS"] and arr.size == len(values) * k
    try:
        _marshal.ints_to_limbs(values, arr.ctypes.data, k, modulus, item, allow_negative)
    except TypeError:
        if item >= 0:
            raise
        # numpy integers and other __index__ types (floats and strings still raise TypeError)
        _marshal.ints_to_limbs([operator.index(v) for v in values], arr.ctypes.data, k, modulus, -1, allow_negative)
    return arr


def limbs_to_ints(arr, nbytes=32):
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    k = nbytes // 8
    return _marshal.limbs_to_ints(arr.ctypes.data, arr.size // k, k)


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class DeviceBuffer:
    """Owning handle of a raw device allocation."""

    def __init__(self, nbytes):
        ensure_init()
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        check(lib.zkb_dev_alloc(self.nbytes, ctypes.byref(p)))
        self.ptr = p

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(lib.zkb_h2d(self.ptr, ptr(arr), arr.nbytes))
        return self

    def download(self, dtype=np.uint64, count=None):
        nbytes = self.nbytes if count is None else count
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        check(lib.zkb_d2h(ptr(out), self.ptr, out.nbytes))
        return out

    def at(self, byte_offset):
        return ctypes.c_void_p(self.ptr.value + byte_offset)

    def free(self):
        if self.ptr is not None and self.ptr.value:
            lib.zkb_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


PROF_TAGS = {"ntt": 0, "msm_sort": 1, "msm_accum_g1": 2, "msm_accum_g2": 3, "msm_reduce": 4, "spmv": 5, "vec": 6, "misc": 7}


def prof_read():
    """{family: (total_ms, launches)} since the last zkb_prof_enable(1)."""
    out = {}
    for name, tag in PROF_TAGS.items():
        ms, cnt = ctypes.c_float(), ctypes.c_ulonglong()
        check(lib.zkb_prof_read(tag, ctypes.byref(ms), ctypes.byref(cnt)))
        out[name] = (ms.value, cnt.value)
    return out


class Timer:
    """CUDA-event timer on the library stream."""

    def __enter__(self):
        check(lib.zkb_timer_start())
        return self

    def __exit__(self, *exc):
        ms = ctypes.c_float()
        check(lib.zkb_timer_stop(ctypes.byref(ms)))
        self.ms = ms.value
        return False
