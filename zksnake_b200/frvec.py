"""Device-resident Fr vectors: the B200-side replacement of the Python lists the reference's prover glue passes between its
NTT / MSM calls (python/zksnake/plonk/protocol.py:270-466, polynomial.py:126-165, utils.py:42-62).  An FrVec is a canonical Fr
vector in HBM plus a length; every method is one or a few kernel launches on the library stream (include/zkb200.h:
zkb_ntt_dev, zkb_vec_op_dev, zkb_fr_*_dev), nothing is marshalled through Python ints unless asked for (`to_ints`).

Buffers come from a size-keyed pool: all work is on one stream, so a buffer released by the garbage collector can be handed to
the next allocation without synchronising."""
import ctypes

import numpy as np

from . import _native as nat

_pool = {}
MUL, ADD, SUB = 0, 1, 2


def _alloc(nbytes):
    size = 256
    while size < nbytes:
        size *= 2
    if size > (1 << 22):                      # above 4 MiB: 1/8-octave granularity instead of powers of two
        step = 1 << (size.bit_length() - 5)
        size = (nbytes + step - 1) // step * step
    free = _pool.get(size)
    if free:
        return free.pop(), size
    return nat.DeviceBuffer(size), size


def release_pool():
    for bufs in _pool.values():
        for b in bufs:
            b.free()
    _pool.clear()


def _words(x):
    return nat.ints_to_limbs([int(x)])


class FrVec:
    __slots__ = ("curve", "n", "buf", "_size", "__weakref__")

    def __init__(self, curve, n):
        nat.ensure_init()
        self.curve, self.n = curve, int(n)
        self.buf, self._size = _alloc(max(self.n, 1) * 32)

    def __del__(self):
        try:
            if self.buf is not None:
                _pool.setdefault(self._size, []).append(self.buf)
                self.buf = None
        except Exception:
            pass

    # ---- construction / extraction ----
    @property
    def ptr(self):
        return self.buf.ptr

    def at(self, index):
        return self.buf.at(index * 32)

    def __len__(self):
        return self.n

    @classmethod
    def zeros(cls, curve, n):
        v = cls(curve, n)
        nat.check(nat.lib.zkb_memset(v.ptr, 0, max(n, 1) * 32))
        return v

    @classmethod
    def from_limbs(cls, curve, arr, reduce=True):
        if not (isinstance(arr, np.ndarray) and arr.dtype == np.uint64 and arr.flags["C_CONTIGUOUS"]):
            arr = np.ascontiguousarray(arr, dtype=np.uint64)   # (a pinned array passes through untouched)
        arr = arr.reshape(-1, 4)
        v = cls(curve, len(arr))
        if len(arr):
            nat.check(nat.lib.zkb_h2d(v.ptr, nat.ptr(arr), arr.nbytes))
            if reduce:
                nat.check(nat.lib.zkb_fr_reduce_dev(curve, len(arr), v.ptr))
        return v

    @classmethod
    def from_ints(cls, curve, values):
        vals = [int(x) if 0 <= int(x) < (1 << 256) else int(x) % (1 << 256) for x in values]
        return cls.from_limbs(curve, nat.ints_to_limbs(vals) if vals else np.zeros((0, 4), dtype=np.uint64))

    def to_limbs(self, count=None, offset=0):
        count = self.n - offset if count is None else count
        out = np.zeros((count, 4), dtype=np.uint64)
        if count:
            nat.check(nat.lib.zkb_d2h(nat.ptr(out), self.at(offset), count * 32))
        return out

    def to_ints(self, count=None, offset=0):
        return nat.limbs_to_ints(self.to_limbs(count, offset))

    def item(self, index):
        return self.to_ints(1, index)[0]

    def copy(self, lo=0, hi=None, n=None):
        """elements [lo, hi) as a new vector of length n >= hi - lo (zero-filled beyond)"""
        hi = self.n if hi is None else hi
        cnt = max(hi - lo, 0)
        n = cnt if n is None else n
        out = FrVec.zeros(self.curve, n) if n > cnt else FrVec(self.curve, n)
        if cnt:
            nat.check(nat.lib.zkb_d2d(out.ptr, self.at(lo), cnt * 32))
        return out

    # ---- transforms ----
    def ntt(self, size, inverse=False, coset=0):
        """N = next power of two >= size; input zero-padded / truncated to N (polynomial.rs:536-571).  coset: 0 plain, 1 the
        reference's coset (offset = the domain's own generator), 2 the coset g<w> of the field's multiplicative generator."""
        log_n = max(int(size) - 1, 0).bit_length()
        out = FrVec(self.curve, 1 << log_n)
        nat.check(nat.lib.zkb_ntt_dev(self.curve, 1 if inverse else 0, coset, log_n, self.ptr, self.n, out.ptr))
        return out

    def intt(self, size=None, coset=0):
        return self.ntt(self.n if size is None else size, inverse=True, coset=coset)

    # ---- element-wise ----
    def _binary(self, op, other, n=None):
        n = max(self.n, other.n) if n is None else n
        out = FrVec(self.curve, n)
        nat.check(nat.lib.zkb_vec_op_dev(self.curve, op, n, self.ptr, min(self.n, n), other.ptr, min(other.n, n), out.ptr))
        return out

    def mul(self, other, n=None):
        return self._binary(MUL, other, n)

    def add(self, other, n=None):
        return self._binary(ADD, other, n)

    def sub(self, other, n=None):
        return self._binary(SUB, other, n)

    def axpy(self, s, y=None, n=None):
        """s * self + y  (zero-extended to n)"""
        n = max(self.n, y.n if y is not None else 0) if n is None else n
        out = FrVec(self.curve, n)
        nat.check(nat.lib.zkb_fr_axpy_dev(self.curve, n, nat.ptr(_words(s)), self.ptr, min(self.n, n),
                                          y.ptr if y is not None else None, min(y.n, n) if y is not None else 0, out.ptr))
        return out

    def scale(self, s):
        return self.axpy(s)

    def mul_powers(self, base, scale=1, offset=0, count=None):
        """out[i] = self[offset + i] * scale * base^i"""
        count = self.n - offset if count is None else count
        out = FrVec(self.curve, count)
        nat.check(nat.lib.zkb_fr_mul_powers_dev(self.curve, count, nat.ptr(_words(base)), nat.ptr(_words(scale)), self.at(offset),
                                                out.ptr))
        return out

    @classmethod
    def powers(cls, curve, n, base, scale=1):
        """[scale * base^i]"""
        out = cls(curve, n)
        nat.check(nat.lib.zkb_fr_mul_powers_dev(curve, n, nat.ptr(_words(base)), nat.ptr(_words(scale)), None, out.ptr))
        return out

    def add_sparse(self, entries, subtract=False):
        """in place: self[idx] +=/-= value for (idx, value) in entries (dict or list of pairs)"""
        items = list(entries.items()) if isinstance(entries, dict) else list(entries)
        for k in range(0, len(items), 64):
            part = items[k:k + 64]
            idx = np.array([i for i, _ in part], dtype=np.uint64)
            assert int(idx.max()) < self.n
            vals = nat.ints_to_limbs([int(v) % (1 << 256) for _, v in part])
            nat.check(nat.lib.zkb_fr_add_sparse_dev(self.curve, self.ptr, len(part), nat.ptr(idx), nat.ptr(vals), 1 if subtract else 0))
        return self

    def inverse(self):
        out = FrVec(self.curve, self.n)
        nat.check(nat.lib.zkb_fr_inverse_dev(self.curve, self.n, self.ptr, out.ptr))
        return out

    def prefix_product(self):
        """n + 1 entries: out[0] = 1, out[i] = self[0] * ... * self[i-1]"""
        out = FrVec(self.curve, self.n + 1)
        nat.check(nat.lib.zkb_fr_scan_dev(self.curve, 0, self.n, self.ptr, out.ptr))
        return out

    def suffix_sum(self):
        out = FrVec(self.curve, self.n)
        nat.check(nat.lib.zkb_fr_scan_dev(self.curve, 1, self.n, self.ptr, out.ptr))
        return out

    def gather(self, count, stride, offset=0):
        out = FrVec(self.curve, count)
        nat.check(nat.lib.zkb_fr_gather_dev(self.curve, count, self.ptr, stride, offset, out.ptr))
        return out

    def gather_index(self, d_idx_u32, count):
        out = FrVec(self.curve, count)
        nat.check(nat.lib.zkb_fr_gather_index_dev(self.curve, count, self.ptr, d_idx_u32.ptr, out.ptr))
        return out

    # ---- polynomial operations on coefficient vectors ----
    def eval(self, z, count=None):
        out = np.zeros(4, dtype=np.uint64)
        nat.check(nat.lib.zkb_fr_eval_dev(self.curve, self.n if count is None else count, self.ptr, nat.ptr(_words(z)), nat.ptr(out)))
        return nat.limbs_to_ints(out.reshape(1, 4))[0]

    def div_vanishing(self, d):
        """(quotient of self / (X^d - 1), exact?)  -- polynomial.rs:466-489"""
        q = FrVec.zeros(self.curve, max(self.n - d, 1))
        exact = ctypes.c_int(1)
        nat.check(nat.lib.zkb_fr_div_vanishing_dev(self.curve, self.n, d, self.ptr, q.ptr, ctypes.byref(exact)))
        return q, bool(exact.value)

    def div_linear(self, z, p):
        """(quotient, remainder) of self / (X - z) -- the long division of polynomial.rs:404-438 for a linear divisor, as two
        power scalings around a suffix sum: q_i = z^-(i+1) * sum_{j > i} c_j z^j, remainder = sum_j c_j z^j."""
        z %= p
        if self.n <= 1:
            return FrVec.zeros(self.curve, 1), (self.item(0) if self.n else 0)
        if z == 0:
            return self.copy(1, self.n), self.item(0)
        s = self.mul_powers(z).suffix_sum()
        zi = pow(z, -1, p)
        return s.mul_powers(zi, zi, offset=1, count=self.n - 1), s.item(0)
