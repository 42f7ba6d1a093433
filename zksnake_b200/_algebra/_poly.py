"""Builder of the `polynomial_bn254` / `polynomial_bls12_381` modules: the surface the reference registers in
/root/reference/src/lib.rs:70-166 (Polynomial, fft, ifft, coset_fft, coset_ifft, add/mul_over_evaluation_domain,
get_evaluation_point(s), evaluate_vanishing_polynomial, evaluate_lagrange_coefficients) over libzkb200.so.

Transforms and element-wise vector products run on the GPU; the `Polynomial` object keeps its (stripped) coefficient list on
the host as Python ints, exactly what `coeffs()` has to return, and does its cheap structural operations (add, sub, vanishing
division = adds only) there.  Univariate only -- the multivariate variant and MultilinearPolynomial serve sumcheck/GKR, outside
the proving hot path.
"""
import numpy as np

from .. import _native as nat

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

    def _omega(log_n):
        if log_n > two_adicity:
            # the reference unwraps None here (pyo3 PanicException); evaluate_* map it to ValueError
            raise ValueError("Domain size is too large")
        return pow(two_adic_root, 1 << (two_adicity - log_n), r)

    def _pack(values):
        vals = [int(v) for v in values]
        for v in vals:
            if v < 0:
                raise OverflowError("can't convert negative int to unsigned")  # BigUint extraction fails in the reference
        return nat.ints_to_limbs([v if v < (1 << 256) else v % r for v in vals], 32)

    def _ntt(values, size, inverse, coset):
        nat.ensure_init()
        log_n = _domain_log(size)
        if log_n > two_adicity:
            raise ValueError("Domain size is too large")
        n = 1 << log_n
        vals = list(values)[:n]
        a = _pack(vals) if vals else np.zeros((1, 4), dtype=np.uint64)
        out = np.zeros((n, 4), dtype=np.uint64)
        nat.check(nat.lib.zkb_ntt(curve, int(inverse), int(coset), log_n, nat.ptr(a), len(vals), nat.ptr(out)))
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
        out = np.zeros((size, 4), dtype=np.uint64)
        nat.check(nat.lib.zkb_vec_op(curve, op, size, nat.ptr(pa), min(len(a), size), nat.ptr(pb), min(len(b), size),
                                     nat.ptr(out)))
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
        """polynomial.rs:526-533."""
        log_n = _domain_log(domain)
        w = _omega(log_n)
        out, t = [], 1
        for _ in range(1 << log_n):
            out.append(t)
            t = t * w % r
        return out

    def evaluate_vanishing_polynomial(n, tau):
        """polynomial.rs:637-643: tau^N - 1."""
        log_n = _domain_log(n)
        _omega(log_n)
        return (pow(int(tau) % r, 1 << log_n, r) - 1) % r

    def evaluate_lagrange_coefficients(n, tau):
        """polynomial.rs:646-652: L_i(tau) for all i.  L_i(tau) = (1/N) sum_k tau^k w^(-ik) = ifft([tau^k])[i] -- one
        inverse NTT of the powers of tau (and it also covers tau inside the domain, where it yields the unit vector)."""
        log_n = _domain_log(n)
        _omega(log_n)
        size = 1 << log_n
        tau = int(tau) % r
        powers, t = [], 1
        for _ in range(size):
            powers.append(t)
            t = t * tau % r
        return ifft(powers, size)

    def _strip(c):
        c = list(c)
        while c and c[-1] == 0:
            c.pop()
        return c

    class Polynomial:
        """Dense univariate polynomial over Fr with an attached evaluation-domain size (polynomial.rs:17-60)."""

        def __init__(self, num_vars, coeffs, size):
            if num_vars > 1:
                raise NotImplementedError("multivariate polynomials are outside the proving hot path")
            vals = []
            for c in coeffs:
                v = c[0] if isinstance(c, tuple) else c
                if v < 0:
                    raise OverflowError("can't convert negative int to unsigned")
                vals.append(int(v) % r)
            self._c = _strip(vals)
            self._log = _domain_log(size)
            if self._log > two_adicity:
                raise ValueError("Domain size is too large")

        @classmethod
        def _make(cls, coeffs, log):
            o = cls.__new__(cls)
            o._c = _strip(coeffs)
            o._log = log
            return o

        def coeffs(self):
            return list(self._c)

        def degree(self):
            return max(len(self._c) - 1, 0)

        def is_zero(self):
            return not self._c

        def __eq__(self, other):
            return isinstance(other, Polynomial) and self._c == other._c

        def __str__(self):
            terms = []
            for e in range(len(self._c) - 1, -1, -1):
                c = self._c[e]
                if c:
                    terms.append(f"{c}x^{e}" if e > 1 else (f"{c}x" if e == 1 else f"{c}"))
            return " + ".join(terms)

        __repr__ = __str__

        def _coerce(self, other, what):
            if isinstance(other, Polynomial):
                return other._c
            if isinstance(other, int):
                if other < 0:
                    raise TypeError(f"Unsupported type for {what}")
                return _strip([other % r])
            raise TypeError(f"Unsupported type for {what}")

        def __add__(self, other):
            o = self._coerce(other, "addition")
            n = max(len(self._c), len(o))
            a = self._c + [0] * (n - len(self._c))
            b = o + [0] * (n - len(o))
            return self._make([(x + y) % r for x, y in zip(a, b)], self._log)

        __radd__ = __add__

        def __sub__(self, other):
            o = self._coerce(other, "subtraction")
            n = max(len(self._c), len(o))
            a = self._c + [0] * (n - len(self._c))
            b = o + [0] * (n - len(o))
            return self._make([(x - y) % r for x, y in zip(a, b)], self._log)

        def __rsub__(self, other):
            return (-self) + other

        def __neg__(self):
            return self._make([(-x) % r for x in self._c], self._log)

        def __mul__(self, other):
            if isinstance(other, int):
                if other < 0:
                    raise TypeError("Unsupported type for multiplication")
                k = other % r
                return self._make([x * k % r for x in self._c], self._log)
            if not isinstance(other, Polynomial):
                raise TypeError("Unsupported type for multiplication")
            if not self._c or not other._c:
                return self._make([], self._log)
            out = [0] * (len(self._c) + len(other._c) - 1)   # naive_mul, polynomial.rs:354-358 (not on the prove path)
            for i, x in enumerate(self._c):
                if x:
                    for j, y in enumerate(other._c):
                        out[i + j] += x * y
            return self._make([v % r for v in out], self._log)

        __rmul__ = __mul__

        def __truediv__(self, other):
            """polynomial.rs:404-438 -> [quotient, remainder] by long division."""
            if not isinstance(other, Polynomial):
                raise TypeError("Can only divide same n-variate polynomial")
            if not other._c:
                raise RuntimeError("Polynomial division error")
            num = list(self._c)
            den = other._c
            if len(num) < len(den):
                return [self._make([], 0), self._make(num, _domain_log(len(num)))]
            inv_lead = pow(den[-1], -1, r)
            quo = [0] * (len(num) - len(den) + 1)
            for i in range(len(quo) - 1, -1, -1):
                c = num[i + len(den) - 1] * inv_lead % r
                quo[i] = c
                if c:
                    for j, d in enumerate(den):
                        num[i + j] = (num[i + j] - c * d) % r
            qs, rs = _strip(quo), _strip(num[:len(den) - 1])
            return [self._make(qs, _domain_log(len(qs))), self._make(rs, _domain_log(len(rs)))]

        def multiply_by_vanishing_poly(self):
            """polynomial.rs:447-464: p * (X^d - 1)."""
            d = 1 << self._log
            out = [0] * d + self._c
            for i, c in enumerate(self._c):
                out[i] = (out[i] - c) % r
            return self._make(out, self._log)

        def divide_by_vanishing_poly(self):
            """polynomial.rs:466-489: (q, rem) with q[j] = sum_{k>=1} p[j + k d], rem = p[:d] + q[:d]."""
            d = 1 << self._log
            p = self._c
            if len(p) < d:
                return [self._make([], self._log), self._make(p, self._log)]
            qv = [sum(p[j + d::d]) % r for j in range(len(p) - d)]
            rem = list(p[:d])
            for j in range(min(d, len(qv))):
                rem[j] = (rem[j] + qv[j]) % r
            return [self._make(qv, self._log), self._make(rem, self._log)]

        def __call__(self, point):
            if not isinstance(point, int):
                raise TypeError("Univariate polynomial evaluation only accept int")
            x = point % r
            acc = 0
            for c in reversed(self._c):
                acc = (acc * x + c) % r
            return acc

    return {
        "Polynomial": Polynomial, "fft": fft, "ifft": ifft, "coset_fft": coset_fft, "coset_ifft": coset_ifft,
        "add_over_evaluation_domain": add_over_evaluation_domain, "mul_over_evaluation_domain": mul_over_evaluation_domain,
        "get_evaluation_point": get_evaluation_point, "get_all_evaluation_points": get_all_evaluation_points,
        "evaluate_vanishing_polynomial": evaluate_vanishing_polynomial,
        "evaluate_lagrange_coefficients": evaluate_lagrange_coefficients,
        "CURVE_ID": curve, "MODULUS": r,
    }
