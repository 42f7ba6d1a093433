"""Groth16 oracle (pure Python ints; test infrastructure only).

Two independent routes to the proof of /root/reference/python/zksnake/groth16/protocol.py:115-165:
  * `prove_literal`  -- the reference's own sequence: SparseArray.dot (array.py:36-43), QAP.evaluate_witness (qap.py:42-71,
    via oracle.poly), five term-by-term MSMs over an explicitly built proving key (protocol.py:82-109), proof assembly
    (protocol.py:133-165).  Only feasible for small circuits.
  * `prove_closed_form` -- with the toxic waste known, the exponents of A, B, C are closed-form field expressions
    (SURVEY.md section 8c); three scalar multiplications pin the proof bit-exactly at any size.
Both take r, s explicitly (the reference draws them from SystemRandom, utils.py:6-9).
"""
from .curve import group
from .fields import PARAMS, next_power_of_two
from . import poly


def dot(triplets, n_row, witness, p):
    out = [0] * n_row
    for row, col, value in triplets:
        out[row] += witness[col] * value
    return [x % p for x in out]


def lagrange_coeffs(curve, n, tau):
    """L_i(tau) = (tau^n - 1)/n * w^i / (tau - w^i)  (ark evaluate_all_lagrange_coefficients, polynomial.rs:646-652)."""
    P = PARAMS[curve]
    r = P.r
    log_n = n.bit_length() - 1
    w = P.omega(log_n)
    z = (pow(tau, n, r) - 1) % r
    pts, t = [], 1
    for _ in range(n):
        pts.append(t)
        t = t * w % r
    if z == 0:  # tau in the domain
        return [1 if x == tau % r else 0 for x in pts]
    dens = [(tau - x) % r for x in pts]
    # batch inversion
    pref, acc = [], 1
    for d in dens:
        pref.append(acc)
        acc = acc * d % r
    inv_all = pow(acc, -1, r)
    invs = [0] * n
    for i in range(n - 1, -1, -1):
        invs[i] = inv_all * pref[i] % r
        inv_all = inv_all * dens[i] % r
    c = z * pow(n, -1, r) % r
    return [c * x % r * iv % r for x, iv in zip(pts, invs)]


class Setup:
    """Everything protocol.py:32-113 derives from the toxic waste, as scalars."""

    def __init__(self, curve, A, B, C, n_rows, n_cols, n_public, toxic):
        P = PARAMS[curve]
        r = P.r
        self.curve, self.n_public, self.m = curve, n_public, n_cols
        self.n = next_power_of_two(n_rows)
        self.A, self.B, self.C = A, B, C
        self.tau, self.alpha, self.beta, self.gamma, self.delta = toxic
        self.lag = lagrange_coeffs(curve, self.n, self.tau)
        L, R, O = [0] * n_cols, [0] * n_cols, [0] * n_cols
        for trip, acc in ((A, L), (B, R), (C, O)):
            for row, col, value in trip:
                acc[col] = (acc[col] + self.lag[row] * value) % r
        self.K = [(L[i] * self.beta + R[i] * self.alpha + O[i]) % r for i in range(n_cols)]
        self.t = (pow(self.tau, self.n, r) - 1) % r


def prove_closed_form(setup, witness, r_rand, s_rand):
    """(A, B, C) as affine points from closed-form exponents."""
    curve = setup.curve
    P = PARAMS[curve]
    r = P.r
    n = setup.n
    a = dot(setup.A, n, witness, r)
    b = dot(setup.B, n, witness, r)
    c = dot(setup.C, n, witness, r)
    U = sum(x * l for x, l in zip(a, setup.lag)) % r
    V = sum(x * l for x, l in zip(b, setup.lag)) % r
    W = sum(x * l for x, l in zip(c, setup.lag)) % r
    inv_delta = pow(setup.delta, -1, r)
    ht = (U * V - W) % r                        # H(tau) * t(tau)
    kw = sum(w * k for w, k in zip(witness[setup.n_public:], setup.K[setup.n_public:])) % r
    ea = (setup.alpha + U + r_rand * setup.delta) % r
    eb = (setup.beta + V + s_rand * setup.delta) % r
    ec = ((ht + kw) * inv_delta + s_rand * ea + r_rand * eb - r_rand * s_rand % r * setup.delta) % r
    G1, G2 = group(curve, False), group(curve, True)
    return G1.mul(G1.gen, ea), G2.mul(G2.gen, eb), G1.mul(G1.gen, ec)


def msm_exponents(setup, witness):
    """Discrete logs of the five raw MSMs of protocol.py:133-155 (tau_1.U, tau_1.V, tau_2.V, target_1.H, kdelta_1.w_priv)."""
    r = PARAMS[setup.curve].r
    n = setup.n
    a = dot(setup.A, n, witness, r)
    b = dot(setup.B, n, witness, r)
    c = dot(setup.C, n, witness, r)
    U = sum(x * l for x, l in zip(a, setup.lag)) % r
    V = sum(x * l for x, l in zip(b, setup.lag)) % r
    W = sum(x * l for x, l in zip(c, setup.lag)) % r
    inv_delta = pow(setup.delta, -1, r)
    kw = sum(w * k for w, k in zip(witness[setup.n_public:], setup.K[setup.n_public:])) % r
    return U, V, V, (U * V - W) * inv_delta % r, kw * inv_delta % r


def prove_literal(setup, witness, r_rand, s_rand):
    """The reference's sequence, term by term (small circuits only).  Returns (A, B, C, U, V, W, H)."""
    curve = setup.curve
    P = PARAMS[curve]
    r = P.r
    n = setup.n
    G1, G2 = group(curve, False), group(curve, True)
    inv_delta = pow(setup.delta, -1, r)
    pw = [pow(setup.tau, i, r) for i in range(n)]
    tau_1 = [G1.mul(G1.gen, x) for x in pw]
    tau_2 = [G2.mul(G2.gen, x) for x in pw]
    target_1 = [G1.mul(G1.gen, x * setup.t % r * inv_delta % r) for x in pw]
    kdelta_1 = [G1.mul(G1.gen, k * inv_delta % r) for k in setup.K[setup.n_public:]]
    alpha_1, beta_1, beta_2 = G1.mul(G1.gen, setup.alpha), G1.mul(G1.gen, setup.beta), G2.mul(G2.gen, setup.beta)
    delta_1, delta_2 = G1.mul(G1.gen, setup.delta), G2.mul(G2.gen, setup.delta)

    a = dot(setup.A, n, witness, r)
    b = dot(setup.B, n, witness, r)
    c = dot(setup.C, n, witness, r)
    U, V, W, H = poly.evaluate_witness_evals(curve, a, b, c)

    def multiexp(g, s):  # ecc.py:107-126
        if len(s) == 0:
            return None
        return G1.msm(g[:len(s)], s) if not isinstance(g[0][0], tuple) else G2.msm(g[:len(s)], s)

    A = G1.add(G1.add(multiexp(tau_1, U), alpha_1), G1.mul(delta_1, r_rand))
    B1 = G1.add(G1.add(multiexp(tau_1, V), beta_1), G1.mul(delta_1, s_rand))
    B2 = G2.add(G2.add(multiexp(tau_2, V), beta_2), G2.mul(delta_2, s_rand))
    HZ = multiexp(target_1, H)
    priv = witness[setup.n_public:]
    KW = multiexp(kdelta_1, priv) if priv else None
    Cp = G1.add(HZ, KW)
    Cp = G1.add(Cp, G1.mul(A, s_rand))
    Cp = G1.add(Cp, G1.mul(B1, r_rand))
    Cp = G1.add(Cp, G1.mul(G1.neg(delta_1), r_rand * s_rand % r))
    return A, B2, Cp, U, V, W, H


def proof_bytes(curve, A, B, C):
    """Proof.to_bytes, serialization.py:40-42."""
    G1, G2 = group(curve, False), group(curve, True)
    return G1.to_bytes(A) + G2.to_bytes(B) + G1.to_bytes(C)
