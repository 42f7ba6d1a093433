"""PlonK prover oracle (pure Python ints; TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Restates what /root/reference/python/zksnake/plonk/protocol.py:157-484 (prove), :39-155 (setup) and
plonk/serialization.py:102-125 (Proof.to_bytes) compute, by a different route than the product so that the two only agree if
both are right:
  * every polynomial identity is evaluated with plain coefficient arithmetic (schoolbook products, synthetic division) instead
    of the reference's NTT pipeline on the 4n / 8n domains -- the polynomials are unique, so the coefficients must agree;
  * every commitment is the closed form [P(tau)]G1 with the toxic waste tau known, instead of an MSM over the SRS;
  * the grand product is built from the witness values (the blinding terms vanish on the domain).
What IS shared with the product is the reading of the reference: transcript byte conventions (transcript.py:42-71: blake2b,
ints big-endian in `bit_length()` bytes), the order of the 11 blinding draws and the proof layout.  PARITY STATUS: unpinned
against reference output (the reference cannot be built here and ships no PlonK vectors); pinned to the protocol's own
verification equation through an independent pairing (tests/test_gpu_plonk.py) and to this restatement.
"""
import hashlib

from .curve import group
from .fields import PARAMS

K1, K2 = 2, 3


# ---- coefficient-list polynomials over Z_r -----------------------------------------------------------------------------
def strip(c):
    c = list(c)
    while c and c[-1] == 0:
        c.pop()
    return c


def padd(a, b, r):
    n = max(len(a), len(b))
    return strip([((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % r for i in range(n)])


def psub(a, b, r):
    n = max(len(a), len(b))
    return strip([((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % r for i in range(n)])


def pscale(a, k, r):
    return strip([x * k % r for x in a])


def pmul(a, b, r):
    if not a or not b:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] += x * y
    return strip([v % r for v in out])


def peval(a, x, r):
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % r
    return acc


def pdiv_linear(a, z, r):
    """(q, rem) of a / (X - z) by synthetic division."""
    q = [0] * max(len(a) - 1, 0)
    carry = 0
    for i in range(len(a) - 1, 0, -1):
        carry = (a[i] + carry * z) % r
        q[i - 1] = carry
    rem = ((a[0] if a else 0) + carry * z) % r
    return strip(q), rem


def interpolate(evals, omega, r):
    """coefficients of the polynomial with value evals[i] at omega^i (O(n^2) inverse DFT straight from the definition)."""
    n = len(evals)
    inv_n = pow(n, -1, r)
    winv = pow(omega, -1, r)
    out = []
    for j in range(n):
        wj = pow(winv, j, r)
        acc, t = 0, 1
        for e in evals:
            acc = (acc + e * t) % r
            t = t * wj % r
        out.append(acc * inv_n % r)
    return strip(out)


class Transcript:
    def __init__(self, field):
        self.h = hashlib.blake2b(b"")
        self.field = field

    def append_int(self, v):
        self.h.update(v.to_bytes(v.bit_length(), "big"))

    def append_bytes(self, b):
        self.h.update(b)

    def challenge(self):
        d = self.h.digest()
        self.h = hashlib.blake2b(d)
        return int.from_bytes(d, "big") % self.field


class Circuit:
    """selector vectors (length n), permutation (length 3n) -- the fields of zksnake_b200.plonkish.Plonkish."""

    def __init__(self, curve, qL, qR, qO, qM, qC, permutation):
        self.curve = curve
        self.n = len(qL)
        self.q = {"L": qL, "R": qR, "O": qO, "M": qM, "C": qC}
        self.permutation = permutation


def prove(circ, tau, public_witness, private_witness, randoms):
    """-> (proof bytes, dict of intermediate values).  `randoms`: the 11 blinding scalars in draw order."""
    P = PARAMS[circ.curve]
    r, n = P.r, circ.n
    G1 = group(circ.curve, False)
    log_n = n.bit_length() - 1
    omega = P.omega(log_n)
    roots = [pow(omega, i, r) for i in range(n)]
    rnd = iter(randoms)
    commit = lambda poly: G1.mul(G1.gen, peval(poly, tau, r))  # noqa: E731  closed form
    zh = [r - 1] + [0] * (n - 1) + [1]

    ids = roots + [K1 * w % r for w in roots] + [K2 * w % r for w in roots]
    sigma_ev = [[ids[circ.permutation[i + k * n]] for i in range(n)] for k in range(3)]
    sel = {k: interpolate(v, omega, r) for k, v in circ.q.items()}
    sig = [interpolate(s, omega, r) for s in sigma_ev]
    idp = [interpolate(ids[k * n:(k + 1) * n], omega, r) for k in range(3)]
    tr = Transcript(r)
    for k in "LROMC":
        tr.append_bytes(G1.to_bytes(commit(sel[k])))
    for s in sig:
        tr.append_bytes(G1.to_bytes(commit(s)))
    for _, v in public_witness.items():
        tr.append_int(v)

    wires = [list(private_witness[k::3]) for k in range(3)]
    wires = [[x % r for x in w] + [0] * (n - len(w)) for w in wires]
    pi_ev = [0] * n
    for k, v in public_witness.items():
        pi_ev[k] = v % r
    # round 1
    W = [interpolate(w, omega, r) for w in wires]
    PI = interpolate(pi_ev, omega, r)
    for k in range(3):
        b = [next(rnd), next(rnd)]
        W[k] = padd(W[k], pmul(b, zh, r), r)
    A, B, C = W
    Gp = padd(padd(padd(pmul(A, sel["L"], r), pmul(B, sel["R"], r), r), padd(pmul(pmul(A, B, r), sel["M"], r),
                                                                               pmul(C, sel["O"], r), r), r), padd(sel["C"], PI, r), r)
    cm_a, cm_b, cm_c = commit(A), commit(B), commit(C)
    for c in (cm_a, cm_b, cm_c):
        tr.append_bytes(G1.to_bytes(c))
    # round 2
    beta, gamma = tr.challenge(), tr.challenge()
    bz = [next(rnd), next(rnd), next(rnd)]
    acc = [1]
    for i in range(n):
        num = den = 1
        for k in range(3):
            num = num * (wires[k][i] + beta * ids[k * n + i] + gamma) % r
            den = den * (wires[k][i] + beta * sigma_ev[k][i] + gamma) % r
        acc.append(acc[-1] * num * pow(den, -1, r) % r)
    if acc.pop() != 1:
        raise AssertionError("Copy constraints are not satisfied")
    Z = padd(pmul(bz, zh, r), interpolate(acc, omega, r), r)
    cm_z = commit(Z)
    tr.append_bytes(G1.to_bytes(cm_z))
    # round 3
    alpha = tr.challenge()
    Zw = strip([c * pow(omega, i, r) % r for i, c in enumerate(Z)])
    lin = lambda w, s: padd(padd(w, pscale(s, beta, r), r), [gamma], r)  # noqa: E731
    nom = pmul(pmul(lin(A, idp[0]), lin(B, idp[1]), r), lin(C, idp[2]), r)
    den = pmul(pmul(lin(A, sig[0]), lin(B, sig[1]), r), lin(C, sig[2]), r)
    L1 = interpolate([1] + [0] * (n - 1), omega, r)
    numer = padd(padd(Gp, pscale(psub(pmul(nom, Z, r), pmul(den, Zw, r), r), alpha, r), r),
                 pscale(pmul(psub(Z, [1], r), L1, r), alpha * alpha % r, r), r)
    # exact division by X^n - 1
    T = [0] * max(len(numer) - n, 0)
    rem = list(numer)
    for i in range(len(numer) - 1, n - 1, -1):
        c = rem[i]
        T[i - n] = c
        rem[i] = 0
        rem[i - n] = (rem[i - n] + c) % r
    if strip(rem):
        raise AssertionError("quotient has a remainder: gate or copy constraints violated")
    b10, b11 = next(rnd), next(rnd)
    xn = [0] * n + [1]
    T_lo = padd(strip(T[:n]), pscale(xn, b10, r), r)
    T_mid = padd(psub(strip(T[n:2 * n]), [b10], r), pscale(xn, b11, r), r)
    T_hi = psub(strip(T[2 * n:]), [b11], r)
    cm_t = [commit(x) for x in (T_lo, T_mid, T_hi)]
    for c in cm_t:
        tr.append_bytes(G1.to_bytes(c))
    # round 4
    zeta = tr.challenge()
    za, zb, zc = peval(A, zeta, r), peval(B, zeta, r), peval(C, zeta, r)
    zs1, zs2, zzw = peval(sig[0], zeta, r), peval(sig[1], zeta, r), peval(Zw, zeta, r)
    L1z = peval(L1, zeta, r)
    gate = padd(padd(padd(pscale(sel["L"], za, r), pscale(sel["R"], zb, r), r),
                     padd(pscale(sel["O"], zc, r), pscale(sel["M"], za * zb % r, r), r), r),
                padd(sel["C"], [peval(PI, zeta, r)], r), r)
    perm_id = (za + beta * zeta + gamma) * (zb + beta * K1 * zeta + gamma) * (zc + beta * K2 * zeta + gamma) % r
    perm_sig = (za + beta * zs1 + gamma) * (zb + beta * zs2 + gamma) % r
    third = padd(pscale(sig[2], beta, r), [(zc + gamma) % r], r)
    R = padd(gate, pscale(psub(pscale(Z, perm_id, r), pscale(third, perm_sig * zzw % r, r), r), alpha, r), r)
    R = padd(R, pscale(psub(Z, [1], r), alpha * alpha % r * L1z % r, r), r)
    tsum = padd(padd(T_lo, pscale(T_mid, pow(zeta, n, r), r), r), pscale(T_hi, pow(zeta, 2 * n, r), r), r)
    R = psub(R, pscale(tsum, peval(zh, zeta, r), r), r)
    for v in (za, zb, zc, zs1, zs2, zzw):
        tr.append_int(v)
    # round 5
    v = tr.challenge()
    Wp = R
    for k, (poly, val) in enumerate(((A, za), (B, zb), (C, zc), (sig[0], zs1), (sig[1], zs2)), start=1):
        Wp = padd(Wp, pscale(psub(poly, [val], r), pow(v, k, r), r), r)
    Wz, rem1 = pdiv_linear(Wp, zeta, r)
    Wzw, rem2 = pdiv_linear(psub(Z, [zzw], r), zeta * omega % r, r)
    if rem1 or rem2:
        raise AssertionError("opening polynomial not divisible")
    pts = [cm_a, cm_b, cm_c, cm_z] + cm_t + [commit(Wz), commit(Wzw)]
    scal = [za, zb, zc, zs1, zs2, zzw]
    blob = b"".join(G1.to_bytes(p) for p in pts) + b"".join(s.to_bytes(32, "little") for s in scal)
    return blob, {"A": A, "B": B, "C": C, "Z": Z, "T": T, "beta": beta, "gamma": gamma, "alpha": alpha, "zeta": zeta, "v": v,
                  "points": pts, "scalars": scal}
