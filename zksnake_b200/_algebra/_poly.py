"""Builder of the `polynomial_bn254` / `polynomial_bls12_381` modules: the surface the reference registers in
/root/reference/src/lib.rs:70-166 (Polynomial, fft, ifft, coset_fft, coset_ifft, add/mul_over_evaluation_domain,
get_evaluation_point(s), evaluate_vanishing_polynomial, evaluate_lagrange_coefficients) over libzkb200.so.

Everything that touches field elements runs on the GPU.  The list-taking functions marshal through the C extension
`zksnake_b200._marshal` (host threads reading the PyLong digits; the reference converts one BigUint at a time under the GIL,
src/bn254/polynomial.rs:537-544).  `Polynomial` is DEVICE-RESIDENT: the object owns a canonical Fr
vector in HBM (an `FrVec`), its operators (+ - * neg, vanishing-polynomial multiplication / division, division by a linear
factor, evaluation) are kernel launches on the library stream, and a Python list exists only when `coeffs()` is called.  The
ark-poly invariant the reference exposes -- trailing zeros stripped, `coeffs()` of the zero polynomial is `[]` -- is kept by a
lazily computed stripped length (zkb_fr_trim_dev), so chains of operators never synchronise.

Univariate only -- the multivariate variant and MultilinearPolynomial serve sumcheck/GKR, outside the proving hot path
(SURVEY.md section 2.1 N6).
"""
import ctypes

import numpy as np

from .. import _native as nat
from ..frvec import FrVec

_R = {
    0: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    1: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
}
_GEN = {0: (5, 28), 1: (7, 32)}


def _domain_log(size):
    return 0 if size <= 1 else (size - 1).bit_length()


def build(curve):
    r = _R[curve]
    gen, two_adicity = _GEN[curve]
    two_adic_root = pow(gen, (r - 1) >> two_adicity, r)
    lib = nat.lib

    def _omega(log_n):
        if log_n > two_adicity:
            # the reference unwraps None here (pyo3 PanicException); evaluate_* map it to ValueError
            raise ValueError("Domain size is too large")
        return pow(two_adic_root, 1 << (two_adicity - log_n), r)

    def _pack(values, item=-1):
        """list of non-negative ints (any size: Fr::from(BigUint) reduces) -> (n, 4) uint64 limbs; negatives raise like pyo3's
        BigUint extraction."""
        if not isinstance(values, (list, tuple)):
            values = list(values)
        return nat.ints_to_limbs(values, 32, modulus=r, item=item, allow_negative=False)

    def _word(x):
        return nat.ptr(nat.ints_to_limbs([int(x) % r]))

    def _ntt(values, size, inverse, coset):
        nat.ensure_init()
        log_n = _domain_log(size)
        if log_n > two_adicity:
            raise ValueError("Domain size is too large")
        n = 1 << log_n
        vals = values if len(values) <= n else values[:n]
        a = _pack(vals) if len(vals) else np.zeros((1, 4), dtype=np.uint64)
        out = np.empty((n, 4), dtype=np.uint64)
        nat.check(lib.zkb_ntt(curve, int(inverse), int(coset), log_n, nat.ptr(a), len(vals), nat.ptr(out)))
        return nat.limbs_to_ints(out)

    def fft(coeffs, size):
        """polynomial.rs:536-545."""
        return _ntt(coeffs, size, False, False)

    def coset_fft(coeffs, size):
        """polynomial.rs:548-559."""
        return _ntt(coeffs, size, False, True)

    def ifft(evals, size):
        """polynomial.rs:562-571."""
        return _ntt(evals, size, True, False)

    def coset_ifft(evals, size):
        """polynomial.rs:574-585."""
        return _ntt(evals, size, True, True)

    def _vec(op, size, a, b):
        nat.ensure_init()
        if size == 0:
            return []
        pa = _pack(a[:size]) if len(a) else np.zeros((1, 4), dtype=np.uint64)
        pb = _pack(b[:size]) if len(b) else np.zeros((1, 4), dtype=np.uint64)
        out = np.empty((size, 4), dtype=np.uint64)
        nat.check(lib.zkb_vec_op(curve, op, size, nat.ptr(pa), min(len(a), size), nat.ptr(pb), min(len(b), size), nat.ptr(out)))
        return nat.limbs_to_ints(out)

    def mul_over_evaluation_domain(size, a, b):
        """polynomial.rs:610-634 (short inputs zero-extended)."""
        return _vec(0, size, a, b)

    def add_over_evaluation_domain(size, a, b):
        """polynomial.rs:588-607 (indexes a[i], b[i] for i < size: short input panics in the reference)."""
        if len(a) < size or len(b) < size:
            raise IndexError("index out of bounds")
        return _vec(1, size, a, b)

    def get_evaluation_point(domain, i):
        """polynomial.rs:519-524."""
        log_n = _domain_log(domain)
        return pow(_omega(log_n), i, r)

    def get_all_evaluation_points(domain):
        """polynomial.rs:526-533: [w^i] -- powers on the device, one download."""
        log_n = _domain_log(domain)
        w = _omega(log_n)
        return FrVec.powers(curve, 1 << log_n, w).to_ints()

    def evaluate_vanishing_polynomial(n, tau):
        """polynomial.rs:637-643: tau^N - 1."""
        log_n = _domain_log(n)
        _omega(log_n)
        return (pow(int(tau) % r, 1 << log_n, r) - 1) % r

    def evaluate_lagrange_coefficients(n, tau):
        """polynomial.rs:646-652: L_i(tau) for all i.  L_i(tau) = (1/N) sum_k tau^k w^(-ik) = ifft([tau^k])[i] -- the powers of
        tau and ONE inverse NTT, both on the device (it also covers tau inside the domain, where it yields the unit vector)."""
        log_n = _domain_log(n)
        _omega(log_n)
        return FrVec.powers(curve, 1 << log_n, int(tau) % r).intt().to_ints()

    class Polynomial:
        """Dense univariate polynomial over Fr with an attached evaluation-domain size (polynomial.rs:17-60), resident in HBM.

        _v: FrVec holding at least _n coefficients (None for the zero polynomial); _n: number of coefficients in use -- an
        upper bound of the stripped length until `_trim()` has run (_exact)."""
        __slots__ = ("_v", "_n", "_exact", "_log")

        def __init__(self, num_vars, coeffs, size):
            if num_vars > 1:
                raise NotImplementedError("multivariate polynomials are outside the proving hot path")
            self._log = _domain_log(size)
            if self._log > two_adicity:
                raise ValueError("Domain size is too large")
            n = len(coeffs)
            self._n, self._exact, self._v = n, n == 0, None
            if n:
                # (coeff, [(0, 0)]) terms as the reference's factory builds them (python/zksnake/polynomial.py:40-43), or bare ints
                limbs = _pack(coeffs, item=0 if isinstance(coeffs[0], tuple) else -1)
                self._v = FrVec.from_limbs(curve, limbs)     # reduced mod r on the device (Fr::from(BigUint))

        # ---- construction helpers ----
        @classmethod
        def _wrap(cls, vec, n, log, exact=False):
            o = cls.__new__(cls)
            o._v, o._n, o._log = (vec if n else None), n, log
            o._exact = exact or n == 0
            return o

        @classmethod
        def _from_device(cls, vec, size=None):
            """Adopt a device-resident coefficient vector (an FrVec) without copying -- used by the device provers."""
            return cls._wrap(vec, vec.n, _domain_log(size if size else vec.n))

        def _trim(self):
            """stripped length (ark's DensePolynomial invariant); one tiny kernel + a 8-byte read the first time it is needed"""
            if not self._exact:
                ln = ctypes.c_size_t(0)
                nat.check(lib.zkb_fr_trim_dev(curve, self._n, self._v.ptr, ctypes.byref(ln)))
                self._n = ln.value
                if self._n == 0:
                    self._v = None
                self._exact = True
            return self._n

        def _ptr(self):
            return self._v.ptr if self._v is not None else None

        def device_vector(self):
            """(FrVec, length): the coefficient vector where it lives; the FrVec may be longer than `length`."""
            return self._v, self._n

        # ---- reference API ----
        def coeffs(self):
            """polynomial.rs:132-140: the stripped coefficient list (the only place a Python list is materialised)."""
            n = self._trim()
            return self._v.to_ints(n) if n else []

        def degree(self):
            return max(self._trim() - 1, 0)

        def is_zero(self):
            return self._trim() == 0

        def __eq__(self, other):
            if not isinstance(other, Polynomial):
                return False
            if self._trim() != other._trim():
                return False
            if self._n == 0:
                return True
            return (self - other).is_zero()

        __hash__ = None

        def __str__(self):
            c = self.coeffs()
            terms = []
            for e in range(len(c) - 1, -1, -1):
                if c[e]:
                    terms.append(f"{c[e]}x^{e}" if e > 1 else (f"{c[e]}x" if e == 1 else f"{c[e]}"))
            return " + ".join(terms)

        __repr__ = __str__

        @staticmethod
        def _scalar(other, what):
            # pyo3 extracts BigUint: negative ints (and bools excluded by nothing) fall through to the TypeError branch
            if isinstance(other, int) and other >= 0:
                return other % r
            raise TypeError(f"Unsupported type for {what}")

        def _binary(self, other, op, what):
            if isinstance(other, Polynomial):
                n = max(self._n, other._n)
                if n == 0:
                    return self._wrap(None, 0, self._log)
                out = FrVec(curve, n)
                nat.check(lib.zkb_vec_op_dev(curve, op, n, self._ptr(), self._n, other._ptr(), other._n, out.ptr))
                return self._wrap(out, n, self._log)
            k = self._scalar(other, what)
            n = max(self._n, 1)
            out = self._v.copy(0, self._n, n=n) if self._n else FrVec.zeros(curve, 1)
            if k:
                out.add_sparse([(0, k)], subtract=(op == 2))
            return self._wrap(out, n, self._log)

        def __add__(self, other):
            return self._binary(other, 1, "addition")

        __radd__ = __add__

        def __sub__(self, other):
            return self._binary(other, 2, "subtraction")

        def __rsub__(self, other):
            return (-self) + other

        def _scaled(self, k):
            if self._n == 0 or k == 0:
                return self._wrap(None, 0, self._log)
            out = FrVec(curve, self._n)
            nat.check(lib.zkb_fr_axpy_dev(curve, self._n, _word(k), self._v.ptr, self._n, None, 0, out.ptr))
            return self._wrap(out, self._n, self._log, exact=self._exact)   # k != 0: the leading coefficient stays non-zero

        def __neg__(self):
            return self._scaled(r - 1)

        def __mul__(self, other):
            if isinstance(other, Polynomial):
                # the reference multiplies naively (polynomial.rs:354-358); the product is the same polynomial through one
                # NTT round trip on the smallest domain that holds it
                la, lb = self._trim(), other._trim()
                if la == 0 or lb == 0:
                    return self._wrap(None, 0, self._log)
                size = la + lb - 1
                if _domain_log(size) > two_adicity:
                    raise ValueError("Domain size is too large")
                fa = self._v.copy(0, la).ntt(size)
                fb = other._v.copy(0, lb).ntt(size)
                prod = fa.mul(fb).intt()
                return self._wrap(prod, size, self._log)
            return self._scaled(self._scalar(other, "multiplication"))

        __rmul__ = __mul__

        def __truediv__(self, other):
            """polynomial.rs:404-438 -> [quotient, remainder].  Constant and linear divisors (everything the provers use: KZG
            openings, PlonK round 5) run on the device (power scalings around a suffix-sum scan); a divisor of higher degree
            takes the schoolbook long division over the downloaded coefficients."""
            if not isinstance(other, Polynomial):
                raise TypeError("Can only divide same n-variate polynomial")
            lb = other._trim()
            if lb == 0:
                raise RuntimeError("Polynomial division error")
            la = self._trim()
            if la < lb:
                return [self._wrap(None, 0, 0), self._wrap(self._v, la, _domain_log(la), exact=True)]
            den = other._v.to_ints(lb)
            if lb <= 2:
                lead_inv = pow(den[-1], -1, r)
                if lb == 1:
                    q = self._scaled(lead_inv)
                    q._log = _domain_log(q._trim())
                    return [q, self._wrap(None, 0, 0)]
                z = (-den[0]) * lead_inv % r                      # a1 X + a0 = a1 (X - z)
                qv, rem = self._v.copy(0, la).div_linear(z, r)
                q = self._wrap(qv, la - 1, 0)
                if lead_inv != 1:
                    q = q._scaled(lead_inv)
                q._log = _domain_log(q._trim())
                rp = self._wrap(FrVec.from_ints(curve, [rem]), 1, 0) if rem else self._wrap(None, 0, 0)
                return [q, rp]
            num = self._v.to_ints(la)
            inv_lead = pow(den[-1], -1, r)
            quo = [0] * (la - lb + 1)
            for i in range(len(quo) - 1, -1, -1):
                c = num[i + lb - 1] * inv_lead % r
                quo[i] = c
                if c:
                    for j, d in enumerate(den):
                        num[i + j] = (num[i + j] - c * d) % r
            q = Polynomial(1, quo, 1)
            q._log = _domain_log(q._trim())
            rp = Polynomial(1, num[:lb - 1], 1)
            rp._log = _domain_log(rp._trim())
            return [q, rp]

        def multiply_by_vanishing_poly(self):
            """polynomial.rs:447-464: p * (X^d - 1) = (p shifted up by d) - p."""
            n = self._n
            if n == 0:
                return self._wrap(None, 0, self._log)
            d = 1 << self._log
            shifted = FrVec.zeros(curve, n + d)
            nat.check(lib.zkb_d2d(shifted.at(d), self._v.ptr, n * 32))
            out = FrVec(curve, n + d)
            nat.check(lib.zkb_vec_op_dev(curve, 2, n + d, shifted.ptr, n + d, self._v.ptr, n, out.ptr))
            return self._wrap(out, n + d, self._log, exact=self._exact)

        def divide_by_vanishing_poly(self):
            """polynomial.rs:466-489: (q, rem) with q[j] = sum_{k>=1} p[j + k d], rem = p[:d] + q[:d] -- additions only."""
            d = 1 << self._log
            n = self._n
            if n <= d:
                return [self._wrap(None, 0, self._log), self._wrap(self._v, n, self._log, exact=self._exact)]
            q = FrVec(curve, n - d)
            exact = ctypes.c_int(1)
            nat.check(lib.zkb_fr_div_vanishing_dev(curve, n, d, self._v.ptr, q.ptr, ctypes.byref(exact)))
            if exact.value:
                rem = self._wrap(None, 0, self._log)
            else:
                rv = FrVec(curve, d)
                nat.check(lib.zkb_vec_op_dev(curve, 1, d, self._v.ptr, d, q.ptr, min(d, n - d), rv.ptr))
                rem = self._wrap(rv, d, self._log)
            return [self._wrap(q, n - d, self._log), rem]

        def __call__(self, point):
            if not isinstance(point, int) or point < 0:     # (pyo3's BigUint extraction rejects negatives too)
                raise TypeError("Univariate polynomial evaluation only accept int")
            return self._v.eval(point % r, self._n) if self._n else 0

    return {
        "Polynomial": Polynomial, "fft": fft, "ifft": ifft, "coset_fft": coset_fft, "coset_ifft": coset_ifft,
        "add_over_evaluation_domain": add_over_evaluation_domain, "mul_over_evaluation_domain": mul_over_evaluation_domain,
        "get_evaluation_point": get_evaluation_point, "get_all_evaluation_points": get_all_evaluation_points,
        "evaluate_vanishing_polynomial": evaluate_vanishing_polynomial,
        "evaluate_lagrange_coefficients": evaluate_lagrange_coefficients,
        "CURVE_ID": curve, "MODULUS": r,
    }
