"""The G2 subgroup criteria of csrc/codec.cu (validate=1) on the CPU: the psi constants compiled into the kernel are the ones the
oracle derives (psi(P) = [p mod r] P on G2), and both criteria agree with the definition r * P = infinity on members, random
non-members of E'(Fq2), cofactor-torsion points and member + torsion sums (oracle curve arithmetic only; no GPU)."""
import os
import random
import re

import pytest

from oracle.curve import group
from oracle.fields import PARAMS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [(0, (9, 1), 4965661367192848881, "PSI_BN", 4), (1, (1, 1), 0xd201000000010000, "PSI_BLS", 6)]


def _psi_constants(cid, xi):
    G = group(cid, True)
    F, q = G.F, PARAMS[cid].q

    def fpow(a, e):
        res, base = (1, 0), a
        while e:
            if e & 1:
                res = F.mul(res, base)
            base = F.mul(base, base)
            e >>= 1
        return res

    g3, g2 = fpow(xi, (q - 1) // 3), fpow(xi, (q - 1) // 2)
    return (g3, g2) if cid == 0 else (F.inv(g3), F.inv(g2))   # D-type twist (BN254) / M-type twist (BLS12-381)


@pytest.mark.parametrize("cid,xi,x_param,table,nlimbs", CASES)
def test_compiled_psi_constants_and_criteria(cid, xi, x_param, table, nlimbs):
    G = group(cid, True)
    F, q, r = G.F, PARAMS[cid].q, PARAMS[cid].r
    cx, cy = _psi_constants(cid, xi)
    # 1. the tables in codec.cu
    src = open(os.path.join(ROOT, "zksnake_b200", "csrc", "codec.cu")).read()
    body = src[src.index(f"static const uint64_t {table}["):]
    body = body[:body.index("};")]
    words = [int(w, 16) for w in re.findall(r"0x([0-9a-fA-F]+)ull", body)]
    assert len(words) == 4 * nlimbs
    vals = [sum(words[k * nlimbs + i] << (64 * i) for i in range(nlimbs)) for k in range(4)]
    assert vals == [cx[0], cx[1], cy[0], cy[1]]
    # the curve parameter the kernel multiplies by
    lo, hi = x_param & 0xFFFFFFFF, x_param >> 32
    assert f"0x{lo:08X}u".lower() in src.lower() and f"0x{hi:08X}u".lower() in src.lower()

    conj = lambda a: (a[0], (-a[1]) % q)                                                   # noqa: E731
    psi = lambda pt: None if pt is None else (F.mul(conj(pt[0]), cx), F.mul(conj(pt[1]), cy))  # noqa: E731
    # 2. psi is the p-power endomorphism on G2
    P = G.mul(G.gen, 987654321)
    assert psi(P) == G.mul(P, q % r) and G.on_curve(psi(P))

    # 3. the criterion of the kernel against the definition
    def member_fast(pt):
        if pt is None:
            return True
        xp = G.mul_raw(pt, x_param)
        if cid == 0:
            lhs = G.add(G.add(G.add(xp, pt), psi(xp)), psi(psi(xp)))
            return lhs == psi(psi(psi(G.add(xp, xp))))
        return G.add(psi(pt), xp) is None

    rnd = random.Random(77 + cid)
    seen_in = seen_out = 0
    while seen_out < 24 or seen_in < 8:
        x = (rnd.randint(0, q - 1), rnd.randint(0, q - 1))
        y = F.sqrt(F.add(F.mul(F.mul(x, x), x), G.b))
        if y is None:
            continue
        p0 = (x, y)
        t = G.mul_raw(p0, r)
        m = G.mul(G.gen, rnd.randint(1, r - 1))
        for c in (p0, t, m, G.add(t, m) if t is not None else None):
            if c is None:
                continue
            slow = G.mul_raw(c, r) is None
            assert member_fast(c) == slow
            seen_in += slow
            seen_out += not slow
