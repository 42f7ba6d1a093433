"""Pins the CPU oracle (oracle/*.py and the C++ restatement oracle/cport) WITHOUT a GPU.

The reference's own tests hold no numeric vectors for this path, so the pins are:
  * the six polynomial KATs of /root/reference/tests/test_algebra.py:6-26,
  * published constants (roots of unity, Montgomery constants, the standard compressed encodings of the generators),
  * the O(N^2) definition of the transform, the coset identity coset_fft(c)[i] = fft(c)[i+1], H*Z == U*V - W,
  * MSM == discrete-log closed form, Groth16 closed-form exponents == the literal reference sequence,
  * the committed derived fixtures tests/golden/derived_kats.json (regression), which oracle, C++ port and GPU all share.
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import cport
from oracle import groth16 as og
from oracle import poly
from oracle.curve import group
from oracle.fields import BN254, BLS12_381, PARAMS

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "derived_kats.json")))["curves"]
NAMES = {BN254: "BN254", BLS12_381: "BLS12_381"}
CURVES = [BN254, BLS12_381]


def H(xs):
    return [int(x, 16) for x in xs]


# ------------------------------------------------------------------------------------------------ published constants
def test_published_field_constants():
    """SURVEY.md section 8c table (equal to the halo2curves / blst constants)."""
    bn, bls = PARAMS[BN254], PARAMS[BLS12_381]
    assert bn.two_adic_root == 0x2a3c09f0a58a7e8500e0a7eb8ef62abc402d111e41112ed49bd61b6e725b19f0
    assert bls.two_adic_root == 0x16a2a19edfe81f20d09b681922c813b4b63683508c2280b93829971f439f0d2b
    assert bn.omega(16) == 421743594562400382753388642386256516545992082196004333756405989743524594615
    assert bn.omega(20) == 17220337697351015657950521176323262483320249231368149235373741788599650842711
    assert bls.omega(16) == 15076889834420168339092859836519192632846122361203618639585008852351569017005
    assert bls.omega(20) == 1755840822790712607783180844474754741366353396308200820563736496551326485835
    assert bn.omega(1) == bn.r - 1 and bls.omega(1) == bls.r - 1
    assert (-pow(bn.r, -1, 1 << 64)) % (1 << 64) == 0xc2e1f593efffffff
    assert (-pow(bls.r, -1, 1 << 64)) % (1 << 64) == 0xfffffffeffffffff
    assert (1 << 256) % bn.r == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    assert (1 << 256) % bls.r == 0x1824b159acc5056f998c4fefecbc4ff55884b7fa0003480200000001fffffffe
    for P in (bn, bls):
        assert pow(P.two_adic_root, 1 << P.two_adicity, P.r) == 1
        assert pow(P.two_adic_root, 1 << (P.two_adicity - 1), P.r) == P.r - 1
        with pytest.raises(ValueError):
            P.omega(P.two_adicity + 1)


def test_generator_encodings():
    """ark-serialize compressed encodings of g1()/g2() and the identity (the BLS12-381 ones are
    the standard Zcash/IETF generators)."""
    g = group(BN254, False)
    assert g.to_bytes(g.gen).hex() == "01" + "00" * 31
    assert g.to_bytes(None).hex() == "00" * 31 + "40"
    g = group(BN254, True)
    assert g.to_bytes(g.gen).hex() == ("edf692d95cbdde46ddda5ef7d422436779445c5e66006a42761e1f12efde0018"
                                       "c212f3aeb785e49712e7a9353349aaf1255dfb31b7bf60723a480d9293938e19")
    g = group(BLS12_381, False)
    assert g.to_bytes(g.gen).hex() == ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
                                       "6c55e83ff97a1aeffb3af00adb22c6bb")
    assert g.to_bytes(None).hex() == "c0" + "00" * 47
    g = group(BLS12_381, True)
    assert g.to_bytes(g.gen).hex() == ("93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
                                       "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")
    for curve in CURVES:
        for g2 in (False, True):
            G = group(curve, g2)
            assert G.on_curve(G.gen) and G.mul(G.gen, G.r) is None
            for k in (1, 2, 3, 0xdeadbeef, G.r - 1):
                p = G.mul(G.gen, k)
                assert G.from_bytes(G.to_bytes(p)) == p


# ------------------------------------------------------------------------------------------------ reference KATs
@pytest.mark.parametrize("curve", CURVES)
def test_reference_polynomial_kats(curve):
    """/root/reference/tests/test_algebra.py:6-26: (1+2x+3x^2)(2+3x+4x^2), +, -, scalar *, evaluation, division."""
    r = PARAMS[curve].r
    a, b = [1, 2, 3], [2, 3, 4]
    assert poly.mul_over_fft(curve, a, b) == [2, 7, 16, 17, 12]
    assert poly.poly_sub(curve, a, b) == [r - 1, r - 1, r - 1]
    assert poly.poly_eval(curve, a, 2) == 17
    q, rem = poly.divide_by_vanishing_poly(curve, [r - 1, 0, 0, 0, 1], 4)   # (x^4 - 1)/(x^4 - 1)
    assert (q, rem) == ([1], [])
    assert poly.multiply_by_vanishing_poly(curve, [1, 1], 4) == [r - 1, r - 1, 0, 0, 1, 1]


# ------------------------------------------------------------------------------------------------ definitions
@pytest.mark.parametrize("curve", CURVES)
def test_fft_is_the_definition(curve):
    P = PARAMS[curve]
    rnd = random.Random(curve)
    for log_n in range(0, 8):
        n = 1 << log_n
        c = [rnd.randrange(P.r) for _ in range(n)]
        w = P.omega(log_n)
        f = poly.fft(curve, c)
        assert f == poly.ntt_definition(c, log_n, w, P.r)
        assert poly.ifft(curve, f) == c
        cf = poly.fft(curve, c, coset=True)
        assert cf == f[1:] + f[:1]                         # offset = group_gen => a rotation by one
        assert poly.ifft(curve, cf, coset=True) == c
    # ragged inputs: zero-pad, truncate, reduce
    c = [rnd.randrange(P.r) for _ in range(5)]
    assert poly.fft(curve, c, 8) == poly.fft(curve, c + [0, 0, 0])
    assert poly.fft(curve, c + [7] * 9, 4) == poly.fft(curve, c[:4])
    assert poly.fft(curve, [P.r + 3, 1]) == poly.fft(curve, [3, 1])


@pytest.mark.parametrize("curve", CURVES)
def test_quotient_identity_and_survey_vectors(curve):
    P = PARAMS[curve]
    r = P.r
    rnd = random.Random(7)
    n = 16
    a = [rnd.randrange(r) for _ in range(n)]
    b = [rnd.randrange(r) for _ in range(n)]
    c = [x * y % r for x, y in zip(a, b)]
    U, V, W, Hq = poly.evaluate_witness_evals(curve, a, b, c)
    z = rnd.randrange(r)
    lhs = (poly.poly_eval(curve, U, z) * poly.poly_eval(curve, V, z) - poly.poly_eval(curve, W, z)) % r
    assert lhs == poly.poly_eval(curve, Hq, z) * (pow(z, n, r) - 1) % r
    c[3] = (c[3] + 1) % r
    with pytest.raises(ValueError):
        poly.evaluate_witness_evals(curve, a, b, c)
    if curve == BN254:
        # chain circuit n_power=4, inp=2
        U, V, W, Hq = poly.evaluate_witness_evals(curve, [2, 4, 8, 16], [2, 2, 2, 1], [4, 8, 16, 16])
        assert Hq == [2736030358979909402780800718157159386068545550052004292962275523321976061950,
                      5472060717959818811622492770471654055631397811449933516338059605094277952887,
                      2736030358979909404433771082018250827021538289509983819438686948353883106221]
        # README circuit: H is the zero polynomial
        U, V, W, Hq = poly.evaluate_witness_evals(curve, [3, 9], [3, 3], [9, 27])
        assert (U, V, W, Hq) == ([6, r - 3], [3], [18, r - 9], [])


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("g2", [False, True])
def test_msm_closed_form(curve, g2):
    G = group(curve, g2)
    rnd = random.Random(3)
    ks = [rnd.randrange(G.r) for _ in range(5)]
    ss = [rnd.randrange(G.r) for _ in range(5)]
    pts = [G.mul(G.gen, k) for k in ks]
    assert G.msm(pts, ss) == G.mul(G.gen, sum(k * s for k, s in zip(ks, ss)) % G.r)
    with pytest.raises(ValueError, match="mismatch"):
        G.msm(pts, ss[:-1])


# ------------------------------------------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize("curve", CURVES)
def test_oracle_matches_golden(curve):
    gold = GOLD[NAMES[curve]]
    for v in gold["ntt"]:
        n = 1 << v["log_n"]
        c = H(v["input"])
        assert poly.fft(curve, c, n) == H(v["fft"])
        assert poly.ifft(curve, c, n) == H(v["ifft"])
        assert poly.fft(curve, c, n, coset=True) == H(v["coset_fft"])
        assert poly.ifft(curve, c, n, coset=True) == H(v["coset_ifft"])
    for key, g2 in (("g1", False), ("g2", True)):
        G = group(curve, g2)
        v = gold["msm"][key]
        pts = [None if p is None else G.from_bytes(bytes.fromhex(p)) for p in v["points"]]
        assert G.to_bytes(G.msm(pts, H(v["scalars"]))).hex() == v["result"]


@pytest.mark.parametrize("curve", CURVES)
def test_cport_matches_golden_and_oracle(curve):
    """The C++ restatement (the CPU baseline that bench.py times) against the same fixtures and against the Python oracle."""
    gold = GOLD[NAMES[curve]]
    P = PARAMS[curve]
    for v in gold["ntt"]:
        c = H(v["input"])
        for key, inv, coset in (("fft", 0, 0), ("ifft", 1, 0), ("coset_fft", 0, 1), ("coset_ifft", 1, 1)):
            assert cport.unpack(cport.fft(curve, c, v["log_n"], bool(inv), bool(coset))) == H(v[key]), key
    rnd = random.Random(9)
    for log_n in (9, 11):
        c = [rnd.randrange(P.r) for _ in range(1 << log_n)]
        assert cport.unpack(cport.fft(curve, c, log_n)) == poly.fft(curve, c)
        assert cport.unpack(cport.fft(curve, c, log_n, True, True)) == poly.ifft(curve, c, coset=True)
    with pytest.raises(ValueError):
        cport.fft(curve, [1], P.two_adicity + 1)
    fl = cport.fq_limbs(curve)
    for key, grp in (("g1", 1), ("g2", 2)):
        G = group(curve, grp == 2)
        v = gold["msm"][key]
        pts = [None if p is None else G.from_bytes(bytes.fromhex(p)) for p in v["points"]]
        flat = np.zeros((len(pts), cport.affine_limbs(curve, grp)), dtype=np.uint64)
        for i, p in enumerate(pts):
            if p is None:
                continue
            coords = list(p) if grp == 1 else [p[0][0], p[0][1], p[1][0], p[1][1]]
            flat[i] = cport.pack(coords, fl * 8).reshape(-1)
        out, inf = cport.msm(curve, grp, flat, cport.pack(H(v["scalars"])))
        co = cport.unpack(out, fl * 8)
        got = None if inf else ((co[0], co[1]) if grp == 1 else ((co[0], co[1]), (co[2], co[3])))
        assert G.to_bytes(got).hex() == v["result"]
    with pytest.raises(ValueError, match="mismatch"):
        cport.msm(curve, 1, np.zeros((2, cport.affine_limbs(curve, 1)), np.uint64), np.zeros((1, 4), np.uint64))


@pytest.mark.parametrize("curve", CURVES)
def test_cport_msm_large_window_path_closed_form(curve):
    """n >= 32 takes ark's big-window branch; chain points (k0+i)G give sum s_i (k0+i) G exactly."""
    rnd = random.Random(21)
    r = PARAMS[curve].r
    for grp, n in ((1, 300), (2, 70)):
        G = group(curve, grp == 2)
        pts = cport.chain_points(curve, grp, 5, n)
        sc = [rnd.randrange(r) for _ in range(n)]
        sc[:4] = [0, 1, r - 1, r + 5]
        out, inf = cport.msm(curve, grp, pts, cport.pack([s % (1 << 256) for s in sc]))
        fl = cport.fq_limbs(curve)
        co = cport.unpack(out, fl * 8)
        got = (co[0], co[1]) if grp == 1 else ((co[0], co[1]), (co[2], co[3]))
        assert not inf and got == G.mul(G.gen, sum(s * (5 + i) for i, s in enumerate(sc)) % r)


@pytest.mark.parametrize("curve", CURVES)
def test_groth16_golden_oracle_and_cport(curve):
    """Fixed-seed proofs: oracle literal route == closed form == golden; the C++ prover reproduces H and the proof bytes."""
    gold = GOLD[NAMES[curve]]["groth16"]
    from tests.golden.make_golden import chain_triplets, readme_triplets
    r = PARAMS[curve].r
    G1, G2 = group(curve, False), group(curve, True)
    for v in gold:
        trip, n_rows, n_cols = (readme_triplets(r), 2, 4) if v["circuit"] == "readme" else (chain_triplets(4), 4, 6)
        wit, toxic = H(v["witness"]), tuple(H(v["toxic"]))
        rr, ss = int(v["r"], 16), int(v["s"], 16)
        st = og.Setup(curve, *trip, n_rows, n_cols, 2, toxic)
        A, B, C = og.prove_closed_form(st, wit, rr, ss)
        assert og.proof_bytes(curve, A, B, C).hex() == v["proof"]
        # C++ prover on the same explicit key
        inv_delta = pow(st.delta, -1, r)
        n = st.n
        pw = [pow(st.tau, i, r) for i in range(n)]

        def flat(G, pts, grp):
            fl = cport.fq_limbs(curve)
            arr = np.zeros((max(len(pts), 1), cport.affine_limbs(curve, grp)), dtype=np.uint64)
            for i, p in enumerate(pts):
                if p is not None:
                    coords = list(p) if grp == 1 else [p[0][0], p[0][1], p[1][0], p[1][1]]
                    arr[i] = cport.pack(coords, fl * 8).reshape(-1)
            return arr
        key = {
            "tau1": flat(G1, [G1.mul(G1.gen, x) for x in pw], 1), "tau2": flat(G2, [G2.mul(G2.gen, x) for x in pw], 2),
            "target1": flat(G1, [G1.mul(G1.gen, x * st.t % r * inv_delta % r) for x in pw], 1),
            "kdelta1": flat(G1, [G1.mul(G1.gen, k * inv_delta % r) for k in st.K[2:]], 1),
            "alpha1": flat(G1, [G1.mul(G1.gen, st.alpha)], 1)[0], "beta1": flat(G1, [G1.mul(G1.gen, st.beta)], 1)[0],
            "beta2": flat(G2, [G2.mul(G2.gen, st.beta)], 2)[0], "delta1": flat(G1, [G1.mul(G1.gen, st.delta)], 1)[0],
            "delta2": flat(G2, [G2.mul(G2.gen, st.delta)], 2)[0],
        }
        csrs = []
        for t in trip:
            rows = sorted(t)
            rp = np.zeros(n_rows + 1, dtype=np.uint64)
            for row, _, _ in rows:
                rp[row + 1] += 1
            csrs.append((np.cumsum(rp, dtype=np.uint64), np.array([x[1] for x in rows], dtype=np.uint32),
                         cport.pack([x[2] % r for x in rows])))
        log_n = n.bit_length() - 1
        oa, ob, oc, infs, h = cport.groth16_prove(curve, log_n, csrs, n_cols, 2, cport.pack(wit), key, rr, ss, want_h=True)
        assert poly.strip(cport.unpack(h)) == H(v["H"])
        fl = cport.fq_limbs(curve) * 8
        a_, b_, c_ = cport.unpack(oa, fl), cport.unpack(ob, fl), cport.unpack(oc, fl)
        got = og.proof_bytes(curve, None if infs[0] else (a_[0], a_[1]),
                             None if infs[1] else ((b_[0], b_[1]), (b_[2], b_[3])), None if infs[2] else (c_[0], c_[1]))
        assert got.hex() == v["proof"]
