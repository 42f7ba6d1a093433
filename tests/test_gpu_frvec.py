"""GPU parity of the device-resident Fr vector primitives (zksnake_b200/frvec.py over zkb_fr_*_dev) against Python big-int
arithmetic -- the operations the reference's PlonK prover does with Python loops and list marshalling
(python/zksnake/plonk/protocol.py:270-466, utils.py:42-62, src/bn254/polynomial.rs:404-489)."""
import random

import numpy as np
import pytest

from oracle import plonk as op
from oracle.fields import PARAMS

pytestmark = pytest.mark.gpu
CURVES = [0, 1]
SIZES = [1, 7, 2048, 2049, 5000, (1 << 13) + 3]


def rand_vec(r, n, seed, zeros=False):
    rnd = random.Random(seed)
    v = [rnd.randrange(r) for _ in range(n)]
    if zeros:
        for i in range(0, n, 5):
            v[i] = 0
    return v


@pytest.mark.parametrize("curve", CURVES)
def test_elementwise_and_powers(gpu, curve):
    from zksnake_b200.frvec import FrVec
    r = PARAMS[curve].r
    for n in SIZES:
        x, y = rand_vec(r, n, n), rand_vec(r, max(n - 3, 1), n + 1)
        X, Y = FrVec.from_ints(curve, x), FrVec.from_ints(curve, y)
        s = random.Random(n).randrange(r) + r          # >= r: reduced inside
        ypad = y + [0] * (n - len(y))
        assert X.axpy(s, Y).to_ints() == [(s * a + b) % r for a, b in zip(x, ypad)]
        assert X.scale(s).to_ints() == [s * a % r for a in x]
        assert X.axpy(3, Y, n=n + 2).to_ints() == [(3 * a + b) % r for a, b in zip(x, ypad)] + [0, 0]
        base, sc = random.Random(n + 2).randrange(r), random.Random(n + 3).randrange(r)
        assert X.mul_powers(base, sc).to_ints() == [a * sc * pow(base, i, r) % r for i, a in enumerate(x)]
        assert FrVec.powers(curve, n, base, sc).to_ints() == [sc * pow(base, i, r) % r for i in range(n)]
        assert X.mul(Y).to_ints() == [a * b % r for a, b in zip(x, ypad)]
        assert X.sub(Y).to_ints() == [(a - b) % r for a, b in zip(x, ypad)]
        if n > 4:
            assert X.mul_powers(base, 1, offset=2, count=n - 3).to_ints() == [x[2 + i] * pow(base, i, r) % r for i in range(n - 3)]
            Xc = X.copy()
            Xc.add_sparse({0: 5, n - 1: r - 1, 3: s})
            want = list(x)
            want[0] = (want[0] + 5) % r
            want[n - 1] = (want[n - 1] - 1) % r
            want[3] = (want[3] + s) % r
            assert Xc.to_ints() == want
            Xc.add_sparse([(0, 5)], subtract=True)
            want[0] = (want[0] - 5) % r
            assert Xc.to_ints() == want
            assert X.copy(2, 5, n=6).to_ints() == x[2:5] + [0, 0, 0]


@pytest.mark.parametrize("curve", CURVES)
def test_inverse_scans_and_gathers(gpu, curve):
    from zksnake_b200 import _native as nat
    from zksnake_b200.frvec import FrVec
    r = PARAMS[curve].r
    for n in SIZES:
        x = rand_vec(r, n, 7 * n, zeros=True)
        X = FrVec.from_ints(curve, x)
        assert X.inverse().to_ints() == [pow(a, -1, r) if a else 0 for a in x]
        nz = [a or 1 for a in x]
        pref, acc = [1], 1
        for a in nz:
            acc = acc * a % r
            pref.append(acc)
        assert FrVec.from_ints(curve, nz).prefix_product().to_ints() == pref
        suf, acc = [0] * n, 0
        for i in range(n - 1, -1, -1):
            acc = (acc + x[i]) % r
            suf[i] = acc
        assert X.suffix_sum().to_ints() == suf
        if n >= 8:
            assert X.gather(n // 4, 4, 1).to_ints() == x[1::4][:n // 4]
            idx = np.array(random.Random(n).sample(range(n), n), dtype=np.uint32)
            d_idx = nat.DeviceBuffer(idx.nbytes).upload(idx)
            assert X.gather_index(d_idx, n).to_ints() == [x[i] for i in idx]
            d_idx.free()


@pytest.mark.parametrize("curve", CURVES)
def test_polynomial_ops(gpu, curve):
    from zksnake_b200.frvec import FrVec
    r = PARAMS[curve].r
    for n in SIZES + [8192 * 3 + 1]:
        c = rand_vec(r, n, 11 * n)
        C = FrVec.from_ints(curve, c)
        for z in (0, 1, random.Random(n).randrange(r)):
            assert C.eval(z) == op.peval(c, z, r)
            q, rem = C.div_linear(z, r)
            wq, wrem = op.pdiv_linear(c, z, r)
            assert rem == wrem
            got = q.to_ints()[:max(n - 1, 0)] if n > 1 else []
            assert op.strip(got) == wq
    # division by X^d - 1: build p = t * (X^d - 1) (+ a remainder)
    for d, tlen in ((8, 21), (64, 200), (1024, 3 * 1024 + 6)):
        t = rand_vec(r, tlen, d)
        p = [0] * (tlen + d)
        for i, v in enumerate(t):
            p[i] = (p[i] - v) % r
            p[i + d] = (p[i + d] + v) % r
        P = FrVec.from_ints(curve, p + [0] * 5)          # trailing zeros are harmless
        q, exact = P.div_vanishing(d)
        assert exact and op.strip(q.to_ints()) == op.strip(t)
        p[3] = (p[3] + 1) % r
        _, exact = FrVec.from_ints(curve, p).div_vanishing(d)
        assert not exact


@pytest.mark.parametrize("curve", CURVES)
def test_ntt_roundtrip_and_padding(gpu, curve):
    from oracle import poly
    from zksnake_b200.frvec import FrVec
    r = PARAMS[curve].r
    c = rand_vec(r, 50, 5)
    C = FrVec.from_ints(curve, c)
    ev = C.ntt(256)
    assert ev.to_ints() == poly.fft(curve, c, 256)
    assert ev.intt().to_ints() == c + [0] * (256 - 50)
    assert FrVec.from_ints(curve, c).ntt(33).n == 64
