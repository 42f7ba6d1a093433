"""ctypes loader of oracle/cport/libzkcpu.so -- the C++ CPU restatement (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline, --impl reference) may import this.  See zkcpu.cpp for the
reference call sites each function restates and for the parity status ("parity unpinned": pinned to the pure-Python oracle and
algebraic identities, not to reference-produced vectors).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzkcpu.so")


def build(force=False):
    src = os.path.join(_HERE, "zkcpu.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        vp, ci, sz, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_uint64
        L.zkcpu_threads.restype = ci
        L.zkcpu_fft.argtypes = [ci, ci, ci, ci, vp, sz, vp, ci]
        L.zkcpu_msm.argtypes = [ci, ci, vp, sz, vp, sz, vp, ctypes.POINTER(ci), ci]
        L.zkcpu_msm_window.argtypes = [sz]
        L.zkcpu_chain_points.argtypes = [ci, ci, u64, sz, vp, ci]
        L.zkcpu_point_lincomb.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp, ctypes.POINTER(ci)]
        L.zkcpu_groth16_h.argtypes = [ci, ci, vp, vp, vp, vp, vp, vp, vp, ci]
        L.zkcpu_spmv.argtypes = [ci, sz, sz, vp, vp, vp, vp, sz, vp, ci]
        L.zkcpu_groth16_prove.argtypes = [ci, ci, sz, sz, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                          vp, vp, vp, vp, ci]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def fq_limbs(curve):
    return 4 if curve == 0 else 6


def affine_limbs(curve, group):
    return fq_limbs(curve) * 2 * (2 if group == 2 else 1)


def pack(values, nbytes=32):
    buf = b"".join(int(v).to_bytes(nbytes, "little") for v in values)
    return np.frombuffer(buf, dtype=np.uint64).reshape(len(values), nbytes // 8).copy()


def unpack(arr, nbytes=32):
    raw = np.ascontiguousarray(arr).tobytes()
    return [int.from_bytes(raw[i:i + nbytes], "little") for i in range(0, len(raw), nbytes)]


def threads():
    return lib().zkcpu_threads()


def fft(curve, values, log_n, inverse=False, coset=False, nthreads=0):
    """values: (len, 4) uint64 canonical array (or list of ints) -> (2^log_n, 4) uint64."""
    a = values if isinstance(values, np.ndarray) else (pack(values) if len(values) else np.zeros((0, 4), np.uint64))
    if log_n > (28 if curve == 0 else 32):   # EvaluationDomain::new -> None past the field's two-adicity
        raise ValueError("Domain size is too large")
    out = np.zeros((1 << log_n, 4), dtype=np.uint64)
    rc = lib().zkcpu_fft(curve, int(inverse), int(coset), log_n, _p(a) if len(a) else None, len(a), _p(out), nthreads)
    if rc:
        raise ValueError("Domain size is too large")
    return out


def msm(curve, group, points, scalars, nthreads=0):
    """points: (n, affine_limbs) uint64 canonical; scalars: (n, 4) uint64.  Returns (coords uint64 array, is_infinity)."""
    out = np.zeros(affine_limbs(curve, group), dtype=np.uint64)
    inf = ctypes.c_int()
    rc = lib().zkcpu_msm(curve, group, _p(points), len(points), _p(scalars), len(scalars), _p(out), ctypes.byref(inf), nthreads)
    if rc:
        raise ValueError("Number of points and scalars mismatch")
    return out, bool(inf.value)


def chain_points(curve, group, k0, n, nthreads=0):
    """[(k0 + i) * G for i < n] as an (n, affine_limbs) canonical array."""
    out = np.zeros((n, affine_limbs(curve, group)), dtype=np.uint64)
    lib().zkcpu_chain_points(curve, group, k0, n, _p(out), nthreads)
    return out


def lincomb(curve, group, points, scalars):
    """sum_i k_i P_i ; points: list of flat coordinate arrays or None (infinity); scalars: list of int or None (= 1)."""
    al = affine_limbs(curve, group)
    n = len(points)
    pts = np.zeros((n, al), dtype=np.uint64)
    infs = np.zeros(n, dtype=np.int32)
    for i, p in enumerate(points):
        if p is None:
            infs[i] = 1
        else:
            pts[i] = p
    sc = pack([s or 0 for s in scalars])
    has = np.array([0 if s is None else 1 for s in scalars], dtype=np.int32)
    out = np.zeros(al, dtype=np.uint64)
    inf = ctypes.c_int()
    lib().zkcpu_point_lincomb(curve, group, n, _p(pts), _p(infs), _p(sc), _p(has), _p(out), ctypes.byref(inf))
    return out, bool(inf.value)


def groth16_h(curve, log_n, a, b, c, nthreads=0):
    """(U, V, W, H) as (n, 4) arrays from the evaluation vectors; ValueError on a non-zero remainder."""
    n = 1 << log_n
    outs = [np.zeros((n, 4), dtype=np.uint64) for _ in range(4)]
    rc = lib().zkcpu_groth16_h(curve, log_n, _p(a), _p(b), _p(c), *[_p(o) for o in outs], nthreads)
    if rc == -5:
        raise ValueError("(U * V - W) did not divided by Z to zero")
    if rc:
        raise ValueError("Domain size is too large")
    return outs


def spmv(curve, n_out, csr, witness, nthreads=0):
    row_ptr, col, val = csr
    out = np.zeros((n_out, 4), dtype=np.uint64)
    lib().zkcpu_spmv(curve, n_out, len(row_ptr) - 1, _p(row_ptr), _p(col) if len(col) else None, _p(val) if len(val) else None,
                     _p(witness), len(witness), _p(out), nthreads)
    return out


def groth16_prove(curve, log_n, csrs, n_cols, n_public, witness, key, r, s, nthreads=0, want_h=False):
    """key: dict with canonical arrays tau1, tau2, target1, kdelta1, alpha1, beta1, beta2, delta1, delta2.
    Returns (A, B, C coordinate arrays, [infA, infB, infC], H or None)."""
    n = 1 << log_n
    n_rows = len(csrs[0][0]) - 1
    vp3 = ctypes.c_void_p * 3
    rp = vp3(*[c[0].ctypes.data for c in csrs])
    col = vp3(*[c[1].ctypes.data if len(c[1]) else None for c in csrs])
    val = vp3(*[c[2].ctypes.data if len(c[2]) else None for c in csrs])
    oa = np.zeros(affine_limbs(curve, 1), np.uint64)
    ob = np.zeros(affine_limbs(curve, 2), np.uint64)
    oc = np.zeros(affine_limbs(curve, 1), np.uint64)
    inf = (ctypes.c_int * 3)()
    h = np.zeros((n, 4), np.uint64) if want_h else None
    rr, ss = pack([r]), pack([s])
    rc = lib().zkcpu_groth16_prove(curve, log_n, n_rows, n_cols, n_public, rp, col, val, _p(witness), _p(key["tau1"]),
                                   _p(key["tau2"]), _p(key["target1"]), _p(key["kdelta1"]), _p(key["alpha1"]), _p(key["beta1"]),
                                   _p(key["beta2"]), _p(key["delta1"]), _p(key["delta2"]), _p(rr), _p(ss), _p(oa), _p(ob), _p(oc),
                                   inf, _p(h), nthreads)
    if rc == -5:
        raise ValueError("(U * V - W) did not divided by Z to zero")
    if rc:
        raise ValueError(f"zkcpu_groth16_prove failed: {rc}")
    return oa, ob, oc, [inf[0], inf[1], inf[2]], h
