"""Fr polynomial / NTT oracle (pure Python ints; test infrastructure only).

Restates the semantics of the reference bindings in /root/reference/src/bn254/polynomial.rs
(bls12_381 twin is identical): fft :536-545, coset_fft :548-559, ifft :562-571, coset_ifft :574-585,
add/mul_over_evaluation_domain :588-634, divide_by_vanishing_poly :466-489,
multiply_by_vanishing_poly :447-464, and the Python glue _pad_coeffs / mul_over_fft in
/root/reference/python/zksnake/polynomial.py:126-165.  The arithmetic itself lives in ark-poly 0.4.2
(not vendored); its published radix-2 semantics are restated here:
  out[i] = sum_j c_j * (offset * w^i)^j , natural order, input resized (zero-pad / truncate) to N.
"""
from .fields import PARAMS, domain_log, next_power_of_two


def ntt_definition(c, log_n, w, r):
    """O(N^2) transform straight from the definition (used to pin the fast one)."""
    n = 1 << log_n
    out = []
    for i in range(n):
        wi = pow(w, i, r)
        acc = 0
        for cj in reversed(c):
            acc = (acc * wi + cj) % r
        out.append(acc)
    return out


def _ntt_rec(c, w, r):
    n = len(c)
    if n == 1:
        return c
    e = _ntt_rec(c[0::2], w * w % r, r)
    o = _ntt_rec(c[1::2], w * w % r, r)
    out = [0] * n
    t = 1
    h = n // 2
    for i in range(h):
        x = o[i] * t % r
        out[i] = (e[i] + x) % r
        out[i + h] = (e[i] - x) % r
        t = t * w % r
    return out


def _resize(v, n, r):
    v = [x % r for x in v[:n]]  # Fr::from(BigUint) reduces mod r (polynomial.rs:537-540)
    return v + [0] * (n - len(v))


def fft(curve, coeffs, size=None, coset=False):
    """polynomial.rs:536-545 (coset=False) / :548-559 (coset=True, offset = group_gen)."""
    P = PARAMS[curve]
    size = len(coeffs) if size is None else size
    log_n = domain_log(size)
    n = 1 << log_n
    w = P.omega(log_n)
    c = _resize(coeffs, n, P.r)
    if coset:
        g, t = w, 1
        for j in range(n):
            c[j] = c[j] * t % P.r
            t = t * g % P.r
    return _ntt_rec(c, w, P.r)


def ifft(curve, evals, size=None, coset=False):
    """polynomial.rs:562-571 / :574-585."""
    P = PARAMS[curve]
    size = len(evals) if size is None else size
    log_n = domain_log(size)
    n = 1 << log_n
    w = P.omega(log_n)
    e = _resize(evals, n, P.r)
    winv = pow(w, -1, P.r)
    ninv = pow(n, -1, P.r)
    c = [x * ninv % P.r for x in _ntt_rec(e, winv, P.r)]
    if coset:
        t = 1
        for j in range(n):
            c[j] = c[j] * t % P.r
            t = t * winv % P.r
    return c


def strip(c):
    """DensePolynomial::from_coefficients_vec drops trailing zeros (zero poly = [])."""
    c = list(c)
    while c and c[-1] == 0:
        c.pop()
    return c


def mul_over_evaluation_domain(curve, size, a, b):
    """polynomial.rs:610-634 (short inputs are zero padded)."""
    r = PARAMS[curve].r
    a = _resize(a, size, r)
    b = _resize(b, size, r)
    return [x * y % r for x, y in zip(a, b)]


def add_over_evaluation_domain(curve, size, a, b):
    """polynomial.rs:588-607 (indexes a[i], b[i] for i < size: short input is an error)."""
    r = PARAMS[curve].r
    if len(a) < size or len(b) < size:
        raise IndexError("index out of range")
    return [(a[i] + b[i]) % r for i in range(size)]


def divide_by_vanishing_poly(curve, p, domain_size):
    """(q, rem) of p / (X^d - 1); d = 2^ceil(log2(domain_size))  (polynomial.rs:466-489)."""
    r = PARAMS[curve].r
    d = 1 << domain_log(domain_size)
    p = strip(x % r for x in p)
    if len(p) < d:
        return [], p
    q = list(p[d:])  # q[j] = sum_{k>=1} p[j + k*d]
    for j in range(len(q)):
        q[j] = sum(p[j + d::d]) % r
    rem = list(p[:d])
    for j in range(min(d, len(q))):
        rem[j] = (rem[j] + q[j]) % r
    return strip(q), strip(rem)


def multiply_by_vanishing_poly(curve, p, domain_size):
    """p * (X^d - 1)  (polynomial.rs:447-464)."""
    r = PARAMS[curve].r
    d = 1 << domain_log(domain_size)
    p = strip(x % r for x in p)
    out = [0] * d + p
    for i, c in enumerate(p):
        out[i] = (out[i] - c) % r
    return strip(out)


def poly_sub(curve, a, b):
    r = PARAMS[curve].r
    n = max(len(a), len(b))
    a = list(a) + [0] * (n - len(a))
    b = list(b) + [0] * (n - len(b))
    return strip((x - y) % r for x, y in zip(a, b))


def poly_eval(curve, c, x):
    r = PARAMS[curve].r
    acc = 0
    for cj in reversed(c):
        acc = (acc * x + cj) % r
    return acc


def pad_coeffs(a, b):
    """/root/reference/python/zksnake/polynomial.py:126-148 (quirks kept)."""
    a_degree, b_degree = len(a) - 1, len(b) - 1
    if a_degree != b_degree:
        max_pad = max(a_degree, b_degree)
        length = next_power_of_two(max_pad)
        if a_degree > b_degree:
            pad_a, pad_b = [0] * length, [0] * (a_degree + length - b_degree)
        else:
            pad_b, pad_a = [0] * length, [0] * (b_degree + length - a_degree)
    else:
        pad_a = [0] * next_power_of_two(a_degree)
        pad_b = [0] * next_power_of_two(a_degree)
    return list(a) + pad_a, list(b) + pad_b


def mul_over_fft(curve, a, b):
    """polynomial.py:151-165 -> stripped coefficient list of a*b."""
    a, b = pad_coeffs(a, b)
    fa, fb = fft(curve, a), fft(curve, b)
    ab = mul_over_evaluation_domain(curve, len(fa), fa, fb)
    return strip(ifft(curve, ab))


def evaluate_witness_evals(curve, a_ev, b_ev, c_ev):
    """QAP.evaluate_witness (/root/reference/python/zksnake/groth16/qap.py:42-71) from the
    already-computed A.w, B.w, C.w vectors (length n = power of two).  Returns U, V, W, H
    stripped coefficient lists; raises ValueError on a non-zero remainder (qap.py:68-69)."""
    n = len(a_ev)
    u = strip(ifft(curve, a_ev))
    v = strip(ifft(curve, b_ev))
    w = strip(ifft(curve, c_ev))
    uv = mul_over_fft(curve, u, v)
    hz = poly_sub(curve, uv, w)
    h, rem = divide_by_vanishing_poly(curve, hz, n)
    if rem:
        raise ValueError("(U * V - W) did not divided by Z to zero")
    return u, v, w, h
