"""GPU parity tests (run with -m gpu on a B200): every kernel family of libzkb200.so through the C ABI against the
CPU oracle on identical seeded inputs -- bit-exact."""
import ctypes
import os
import random

import numpy as np
import pytest

from oracle import poly
from oracle.curve import group
from oracle.fields import BN254, BLS12_381, PARAMS

pytestmark = pytest.mark.gpu
CURVES = [BN254, BLS12_381]
FIELDS = [PARAMS[BN254].r, PARAMS[BN254].q, PARAMS[BLS12_381].r, PARAMS[BLS12_381].q]


def fr_pack(nat, vals):
    return nat.ints_to_limbs(vals, 32)


def ntt_gpu(nat, curve, vals, log_n, inverse=False, coset=False):
    a = fr_pack(nat, vals) if len(vals) else np.zeros((1, 4), dtype=np.uint64)
    out = np.zeros((1 << log_n, 4), dtype=np.uint64)
    nat.check(nat.lib.zkb_ntt(curve, int(inverse), int(coset), log_n, nat.ptr(a), len(vals), nat.ptr(out)))
    return nat.limbs_to_ints(out)


@pytest.mark.parametrize("field", range(4))
def test_field_ops_device_matches_host_and_python(gpu, field):
    p = FIELDS[field]
    nl = (p.bit_length() + 31) // 32
    rnd = random.Random(field)
    n = 4096
    A = [rnd.randrange(p) for _ in range(n)]
    B = [rnd.randrange(p) for _ in range(n)]
    A[:6] = [0, 1, p - 1, p - 1, 0, 2]
    B[:6] = [0, p - 1, p - 1, 1, 5, (p + 1) // 2]
    a = np.frombuffer(b"".join(v.to_bytes(nl * 4, "little") for v in A), dtype=np.uint32).copy()
    b = np.frombuffer(b"".join(v.to_bytes(nl * 4, "little") for v in B), dtype=np.uint32).copy()
    ops = [lambda x, y: x * y % p, lambda x, y: (x + y) % p, lambda x, y: (x - y) % p, None, lambda x, y: (-x) % p,
           lambda x, y: (x * y + (x + y) * (x - y)) % p, lambda x, y: x * y % p]   # 5, 6: mont_dot2 (lazy Fp2 half product)
    for op, fn in enumerate(ops):
        nn = 64 if op == 3 else n
        out = np.zeros(nn * nl, dtype=np.uint32)
        gpu.check(gpu.lib.zkb_test_field_op_dev(field, op, nn, gpu.ptr(a), gpu.ptr(b), gpu.ptr(out)))
        got = [int.from_bytes(out[i * nl:(i + 1) * nl].tobytes(), "little") for i in range(nn)]
        if fn is None:
            exp = [pow(x, p - 2, p) for x in A[:nn]]
        else:
            exp = [fn(x, y) for x, y in zip(A, B)]
        assert got == exp, f"field {field} op {op}"


@pytest.mark.parametrize("curve", CURVES)
def test_ntt_exhaustive_small(gpu, curve):
    """All four transforms, N = 1 .. 2^12, every output index, against the oracle (which is itself pinned to the O(N^2)
    definition in tests/test_oracle_pins.py)."""
    r = PARAMS[curve].r
    rnd = random.Random(11 + curve)
    for log_n in range(0, 13):
        n = 1 << log_n
        c = [rnd.randrange(r) for _ in range(n)]
        if n >= 4:
            c[1], c[2] = 0, r - 1
        for coset in (False, True):
            f = ntt_gpu(gpu, curve, c, log_n, False, coset)
            assert f == poly.fft(curve, c, coset=coset), (log_n, coset)
            assert ntt_gpu(gpu, curve, f, log_n, True, coset) == c, (log_n, coset)
        assert ntt_gpu(gpu, curve, c, log_n, True, False) == poly.ifft(curve, c)
        assert ntt_gpu(gpu, curve, c, log_n, True, True) == poly.ifft(curve, c, coset=True)


@pytest.mark.parametrize("curve", CURVES)
def test_ntt_padding_truncation_and_reduction(gpu, curve):
    """Ragged inputs: shorter than the domain (zero padded), empty, longer (truncated), entries >= r (reduced)."""
    r = PARAMS[curve].r
    rnd = random.Random(5)
    for log_n, length in [(3, 5), (4, 1), (6, 33), (10, 1000), (11, 1025), (12, 3000), (3, 0)]:
        c = [rnd.randrange(r) for _ in range(length)]
        assert ntt_gpu(gpu, curve, c, log_n) == poly.fft(curve, c, size=1 << log_n)
    c = [rnd.randrange(r) for _ in range(20)]
    assert ntt_gpu(gpu, curve, c, 4) == poly.fft(curve, c, size=16)          # truncation
    big = [r, r + 1, 2 ** 256 - 1, 2 * r + 5, 7, r - 1, 0, 3]
    assert ntt_gpu(gpu, curve, big, 3) == poly.fft(curve, big)                # Fr::from reduction


@pytest.mark.parametrize("maxk", [3, 4, 5, 7])
def test_ntt_multipass_plans(gpu, maxk):
    """Force 2-, 3- and 4-pass decompositions at sizes the oracle checks exhaustively."""
    old = os.environ.get("ZKB_NTT_MAXK")
    os.environ["ZKB_NTT_MAXK"] = str(maxk)
    try:
        for curve in CURVES:
            r = PARAMS[curve].r
            rnd = random.Random(maxk)
            for log_n in (11, 12, 13, 14):
                if (log_n + maxk - 1) // maxk > 4:
                    continue
                n = 1 << log_n
                c = [rnd.randrange(r) for _ in range(n)]
                assert ntt_gpu(gpu, curve, c, log_n) == poly.fft(curve, c), (curve, log_n)
                assert ntt_gpu(gpu, curve, c, log_n, True, True) == poly.ifft(curve, c, coset=True), (curve, log_n)
    finally:
        if old is None:
            del os.environ["ZKB_NTT_MAXK"]
        else:
            os.environ["ZKB_NTT_MAXK"] = old


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("log_n", [16, 20, 22])
def test_ntt_large_sampled(gpu, curve, log_n):
    """Full-size transforms: sampled outputs against Horner evaluation of the definition, plus inverse round trip."""
    P = PARAMS[curve]
    n = 1 << log_n
    rng = np.random.Generator(np.random.PCG64(log_n))
    a = rng.integers(0, 2 ** 63, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 59) - 1)   # < 2^251 < r : already canonical
    out = np.zeros_like(a)
    gpu.check(gpu.lib.zkb_ntt(curve, 0, 0, log_n, gpu.ptr(a), n, gpu.ptr(out)))
    back = np.zeros_like(a)
    gpu.check(gpu.lib.zkb_ntt(curve, 1, 0, log_n, gpu.ptr(out), n, gpu.ptr(back)))
    assert np.array_equal(back, a)
    coeffs = gpu.limbs_to_ints(a)
    w = P.omega(log_n)
    rnd = random.Random(log_n)
    idx = [0, 1, n // 2, n - 1] + [rnd.randrange(n) for _ in range(4 if log_n > 20 else 8)]
    got = gpu.limbs_to_ints(out[idx])
    for i, g in zip(idx, got):
        assert g == poly.poly_eval(curve, coeffs, pow(w, i, P.r)), i


@pytest.mark.parametrize("curve", CURVES)
def test_vec_ops(gpu, curve):
    r = PARAMS[curve].r
    rnd = random.Random(3)
    n = 1000
    A = [rnd.randrange(r) for _ in range(n)]
    B = [rnd.randrange(r) for _ in range(700)]
    a, b = fr_pack(gpu, A), fr_pack(gpu, B)
    out = np.zeros((n, 4), dtype=np.uint64)
    gpu.check(gpu.lib.zkb_vec_op(curve, 0, n, gpu.ptr(a), n, gpu.ptr(b), len(B), gpu.ptr(out)))
    assert gpu.limbs_to_ints(out) == poly.mul_over_evaluation_domain(curve, n, A, B)
    B2 = [rnd.randrange(r) for _ in range(n)]
    b2 = fr_pack(gpu, B2)
    gpu.check(gpu.lib.zkb_vec_op(curve, 1, n, gpu.ptr(a), n, gpu.ptr(b2), n, gpu.ptr(out)))
    assert gpu.limbs_to_ints(out) == poly.add_over_evaluation_domain(curve, n, A, B2)
    gpu.check(gpu.lib.zkb_vec_op(curve, 2, n, gpu.ptr(a), n, gpu.ptr(b2), n, gpu.ptr(out)))
    assert gpu.limbs_to_ints(out) == [(x - y) % r for x, y in zip(A, B2)]


# ---------------------------------------------------------------------------------------------------- MSM helpers
def pts_pack(G, pts):
    nb8 = (G.P.fq_bytes + 7) // 8 * 8
    out = []
    for pt in pts:
        if pt is None:
            out.append(b"\0" * (nb8 * (4 if G.is_g2 else 2)))
        else:
            cs = (pt[0][0], pt[0][1], pt[1][0], pt[1][1]) if G.is_g2 else (pt[0], pt[1])
            out.append(b"".join(c.to_bytes(nb8, "little") for c in cs))
    return np.frombuffer(b"".join(out), dtype=np.uint64).copy()


def pt_unpack(G, arr, inf):
    if inf:
        return None
    nb8 = (G.P.fq_bytes + 7) // 8 * 8
    raw = arr.tobytes()
    cs = [int.from_bytes(raw[i * nb8:(i + 1) * nb8], "little") for i in range(4 if G.is_g2 else 2)]
    return ((cs[0], cs[1]), (cs[2], cs[3])) if G.is_g2 else (cs[0], cs[1])


def msm_gpu(nat, G, pts, scalars):
    p = pts_pack(G, pts) if pts else np.zeros(1, dtype=np.uint64)
    s = fr_pack(nat, scalars) if scalars else np.zeros((1, 4), dtype=np.uint64)
    out = np.zeros(nat.lib.zkb_affine_bytes(G.curve, 2 if G.is_g2 else 1) // 8, dtype=np.uint64)
    inf = ctypes.c_int(0)
    nat.check(nat.lib.zkb_msm(G.curve, 2 if G.is_g2 else 1, nat.ptr(p), len(pts), nat.ptr(s), len(scalars), nat.ptr(out),
                              ctypes.byref(inf)))
    return pt_unpack(G, out, inf.value)


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("g2", [False, True])
def test_msm_small_and_edge_cases(gpu, curve, g2):
    G = group(curve, g2)
    rnd = random.Random(21 + curve + 2 * g2)
    base_k = [rnd.randrange(1, G.r) for _ in range(24)]
    base = [G.mul(G.gen, k) for k in base_k]
    sizes = [1, 2, 3, 17, 64] if g2 else [1, 2, 3, 7, 17, 33, 64, 100]
    for n in sizes:
        pts = [base[i % len(base)] for i in range(n)]        # repeated points when n > 24
        sc = [rnd.randrange(G.r) for _ in range(n)]
        assert msm_gpu(gpu, G, pts, sc) == G.msm(pts, sc), n
    # edge mix: zero / one / r-1 scalars, identity points, P + P and P - P collisions, scalars >= r
    pts = [base[0], base[0], base[1], G.neg(base[1]), None, base[2], base[3], None, base[4]]
    sc = [5, 5, 9, 9, 12345, 0, 1, 0, G.r - 1]
    assert msm_gpu(gpu, G, pts, sc) == G.msm(pts, sc)
    sc2 = [G.r + 3, 2 ** 256 - 1, 7, G.r, 1, 2, 3, 4, 5]
    assert msm_gpu(gpu, G, pts, sc2) == G.msm(pts, sc2)
    # everything cancels -> identity
    assert msm_gpu(gpu, G, [base[0], G.neg(base[0])], [77, 77]) is None
    # empty
    assert msm_gpu(gpu, G, [], []) is None
    # mismatch -> ValueError("Number of points and scalars mismatch")
    with pytest.raises(ValueError, match="mismatch"):
        msm_gpu(gpu, G, [base[0]], [1, 2])


def dlog_msm_case(nat, G, n, scalar_kind, seed):
    """bases P_i = k_i * G built on the GPU by the fixed-base kernel; MSM must equal (sum s_i k_i) * G."""
    curve, grp = G.curve, 2 if G.is_g2 else 1
    rng = np.random.Generator(np.random.PCG64(seed))
    k = rng.integers(0, 2 ** 63, size=(n, 4), dtype=np.uint64)
    k[:, 3] &= np.uint64((1 << 59) - 1)
    if scalar_kind == "uniform":
        s = rng.integers(0, 2 ** 63, size=(n, 4), dtype=np.uint64)
        s[:, 3] &= np.uint64((1 << 59) - 1)
        s[:, 0] |= rng.integers(0, 2, size=n, dtype=np.uint64) << np.uint64(63)
    elif scalar_kind == "bits":      # witness-like: 0 / 1 / small
        s = np.zeros((n, 4), dtype=np.uint64)
        s[:, 0] = rng.integers(0, 3, size=n, dtype=np.uint64)
    elif scalar_kind == "same":      # one hot bucket per window
        s = np.tile(np.array([[0x123456789abcdef1, 0x0fedcba987654321, 0x1111111122222222, 0x0333333344444444]],
                             dtype=np.uint64), (n, 1))
    else:                            # powers of two, chain-circuit witness
        s = np.zeros((n, 4), dtype=np.uint64)
        for i in range(n):
            e = (i + 2) % 250
            s[i, e // 64] = np.uint64(1) << np.uint64(e % 64)
    ab = nat.lib.zkb_affine_bytes(curve, grp)
    d_k = nat.DeviceBuffer(n * 32).upload(k)
    d_s = nat.DeviceBuffer(n * 32).upload(s)
    d_gen = nat.DeviceBuffer(ab)
    gen = pts_pack(G, [G.gen])
    nat.check(nat.lib.zkb_points_upload(curve, grp, nat.ptr(gen), 1, d_gen.ptr))
    d_pts = nat.DeviceBuffer(n * ab)
    nat.check(nat.lib.zkb_batch_mul_dev(curve, grp, d_gen.ptr, 1, d_k.ptr, n, d_pts.ptr))
    out = np.zeros(ab // 8, dtype=np.uint64)
    inf = ctypes.c_int(0)
    nat.check(nat.lib.zkb_msm_dev(curve, grp, d_pts.ptr, d_s.ptr, n, nat.ptr(out), ctypes.byref(inf)))
    ks, ss = nat.limbs_to_ints(k), nat.limbs_to_ints(s)
    e = sum(a * b for a, b in zip(ks, ss)) % G.r
    assert pt_unpack(G, out, inf.value) == G.mul(G.gen, e), (n, scalar_kind)
    # spot-check the fixed-base kernel itself
    few = np.zeros(3 * ab // 8, dtype=np.uint64)
    nat.check(nat.lib.zkb_points_download(curve, grp, d_pts.ptr, 3, nat.ptr(few)))
    for i in range(3):
        assert pt_unpack(G, few[i * ab // 8:(i + 1) * ab // 8], 0) == G.mul(G.gen, ks[i])
    for b in (d_k, d_s, d_gen, d_pts):
        b.free()


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("kind", ["uniform", "bits", "same", "pow2"])
def test_msm_g1_discrete_log(gpu, curve, kind):
    for n in (1000, 1 << 14, 1 << 18):
        dlog_msm_case(gpu, group(curve, False), n, kind, seed=n)


def test_msm_g1_discrete_log_full_size(gpu):
    """2^20 points (BASELINE.json's size), uniform scalars: the result must equal (sum s_i k_i) * G."""
    dlog_msm_case(gpu, group(0, False), 1 << 20, "uniform", seed=2020)


@pytest.mark.parametrize("curve", CURVES)
def test_msm_g2_discrete_log(gpu, curve):
    # ("same": one hot bucket per window -- whole runs of the two-lanes-per-point kernel inside one bucket, then the folds;
    #  "pow2": the chain circuit's witness; 1000: ragged last run, lane pairs without a run)
    for n, kind in ((1 << 12, "uniform"), (1 << 14, "bits"), (1 << 16, "uniform"), (1 << 12, "same"), (1 << 13, "pow2"),
                    (1000, "uniform")):
        dlog_msm_case(gpu, group(curve, True), n, kind, seed=n + 1)


@pytest.mark.parametrize("curve", CURVES)
def test_msm_g2_repeated_and_opposite_points(gpu, curve):
    """The special cases of the group law inside ONE bucket of the G2 accumulation (msm_pair.cuh: selects over pair-uniform flags,
    the doubling walked by the whole warp): the same point many times (P + P right after the bucket's first point), P and -P
    alternating (the running sum returns to the identity over and over), identities in between -- against the oracle."""
    G = group(curve, True)
    rnd = random.Random(77 + curve)
    P1, P2 = G.mul(G.gen, rnd.randrange(1, G.r)), G.mul(G.gen, rnd.randrange(1, G.r))
    s = rnd.randrange(1, G.r)
    cases = [
        ([P1] * 40, [s] * 40),
        ([P1, G.neg(P1)] * 20 + [P2], [s] * 40 + [s]),
        ([P1, None, P1, None, G.neg(P1), P2, P2, P2], [s] * 8),
        ([P1, P1, G.neg(P1), G.neg(P1), P1], [3, 3, 3, 3, 3]),
    ]
    for pts, sc in cases:
        assert msm_gpu(gpu, G, pts, sc) == G.msm(pts, sc)


@pytest.mark.parametrize("curve", CURVES)
def test_groth16_quotient(gpu, curve):
    """H = (U V - W)/Z from A.w, B.w, C.w: exact coefficients vs the oracle's 2n-domain restatement of qap.py:42-71."""
    r = PARAMS[curve].r
    rnd = random.Random(17)
    for log_n in (1, 2, 3, 5, 8, 11):
        n = 1 << log_n
        a = [rnd.randrange(r) for _ in range(n)]
        b = [rnd.randrange(r) for _ in range(n)]
        c = [x * y % r for x, y in zip(a, b)]
        u, v, w, h = poly.evaluate_witness_evals(curve, a, b, c)
        bufs = [np.zeros((n, 4), dtype=np.uint64) for _ in range(4)]
        gpu.check(gpu.lib.zkb_groth16_h(curve, log_n, gpu.ptr(fr_pack(gpu, a)), gpu.ptr(fr_pack(gpu, b)),
                                        gpu.ptr(fr_pack(gpu, c)), *[gpu.ptr(x) for x in bufs]))
        got = [poly.strip(gpu.limbs_to_ints(x)) for x in bufs]
        assert got == [u, v, w, h], log_n
    # README circuit (x=3): H is the zero polynomial; chain circuit KAT of SURVEY.md section 8c
    for (a, b, c) in (([3, 9], [3, 3], [9, 27]), ([2, 4, 8, 16], [2, 2, 2, 1], [4, 8, 16, 16])):
        n = len(a)
        bufs = [np.zeros((n, 4), dtype=np.uint64) for _ in range(4)]
        gpu.check(gpu.lib.zkb_groth16_h(curve, n.bit_length() - 1, gpu.ptr(fr_pack(gpu, a)), gpu.ptr(fr_pack(gpu, b)),
                                        gpu.ptr(fr_pack(gpu, c)), *[gpu.ptr(x) for x in bufs]))
        assert [poly.strip(gpu.limbs_to_ints(x)) for x in bufs] == list(poly.evaluate_witness_evals(curve, a, b, c))
    # unsatisfied R1CS -> the reference's ValueError
    a, b, c = [3, 9], [3, 3], [9, 28]
    bufs = [np.zeros((2, 4), dtype=np.uint64) for _ in range(4)]
    with pytest.raises(ValueError, match="did not divided"):
        gpu.check(gpu.lib.zkb_groth16_h(curve, 1, gpu.ptr(fr_pack(gpu, a)), gpu.ptr(fr_pack(gpu, b)),
                                        gpu.ptr(fr_pack(gpu, c)), *[gpu.ptr(x) for x in bufs]))


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("world", [2, 5, 40])
def test_msm_window_shards_add_up(gpu, curve, world):
    """zkb_msm_dev_windows: the partial sums over disjoint scalar-window ranges add up to the full MSM (also when there are
    more ranks than windows: the surplus ranks return the identity)."""
    nat = gpu
    G = group(curve, False)
    n = 3000
    rng = np.random.Generator(np.random.PCG64(77 + world))
    k = rng.integers(0, 2 ** 63, size=(n, 4), dtype=np.uint64)
    k[:, 3] &= np.uint64((1 << 59) - 1)
    s = rng.integers(0, 2 ** 64, size=(n, 4), dtype=np.uint64)
    s[:, 3] &= np.uint64((1 << 60) - 1)   # the _dev entry takes canonical scalars: keep them below r
    ab = nat.lib.zkb_affine_bytes(curve, 1)
    d_k = nat.DeviceBuffer(n * 32).upload(k)
    d_s = nat.DeviceBuffer(n * 32).upload(s)
    d_gen = nat.DeviceBuffer(ab)
    gen = pts_pack(G, [G.gen])
    nat.check(nat.lib.zkb_points_upload(curve, 1, nat.ptr(gen), 1, d_gen.ptr))
    d_pts = nat.DeviceBuffer(n * ab)
    nat.check(nat.lib.zkb_batch_mul_dev(curve, 1, d_gen.ptr, 1, d_k.ptr, n, d_pts.ptr))
    acc = None
    for rank in range(world):
        out = np.zeros(ab // 8, dtype=np.uint64)
        inf = ctypes.c_int(0)
        nat.check(nat.lib.zkb_msm_dev_windows(curve, 1, d_pts.ptr, d_s.ptr, n, rank, world, nat.ptr(out), ctypes.byref(inf)))
        acc = G.add(acc, pt_unpack(G, out, inf.value))
    e = sum(a * b for a, b in zip(nat.limbs_to_ints(k), nat.limbs_to_ints(s))) % G.r
    assert acc == G.mul(G.gen, e)
    for b in (d_k, d_s, d_gen, d_pts):
        b.free()


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("g2", [False, True])
def test_msm_fixed_base_table(gpu, curve, g2):
    """zkb_msm_table_*: table[w*n+i] = 2^(c w) P_i, one bucket set for all windows.  Must equal the discrete-log closed form for
    the full vector, a prefix (n_scalars < n), window shards, skewed scalars (all equal -> one hot bucket; 0/1/2), and report a
    mismatch beyond the table."""
    nat = gpu
    G = group(curve, g2)
    grp = 2 if g2 else 1
    n = 5000 if g2 else 20000
    rng = np.random.Generator(np.random.PCG64(4242 + curve))
    k = rng.integers(0, 2 ** 63, size=(n, 4), dtype=np.uint64)
    k[:, 3] &= np.uint64((1 << 59) - 1)
    ab = nat.lib.zkb_affine_bytes(curve, grp)
    d_k = nat.DeviceBuffer(n * 32).upload(k)
    d_gen = nat.DeviceBuffer(ab)
    gen = pts_pack(G, [G.gen])
    nat.check(nat.lib.zkb_points_upload(curve, grp, nat.ptr(gen), 1, d_gen.ptr))
    d_pts = nat.DeviceBuffer(n * ab)
    nat.check(nat.lib.zkb_batch_mul_dev(curve, grp, d_gen.ptr, 1, d_k.ptr, n, d_pts.ptr))
    tab = ctypes.c_void_p()
    nat.check(nat.lib.zkb_msm_table_create(curve, grp, d_pts.ptr, n, 0, 1, ctypes.byref(tab)))
    cbits, W, nbytes = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_size_t()
    nat.check(nat.lib.zkb_msm_table_info(tab, ctypes.byref(cbits), ctypes.byref(W), ctypes.byref(nbytes)))
    assert nbytes.value == W.value * n * ab and 3 <= cbits.value <= 22
    ks = nat.limbs_to_ints(k)

    def run(s, count, rank=0, world=1):
        d_s = nat.DeviceBuffer(max(count, 1) * 32).upload(s[:max(count, 1)])
        out = np.zeros(ab // 8, dtype=np.uint64)
        inf = ctypes.c_int(0)
        nat.check(nat.lib.zkb_msm_table_dev(tab, d_s.ptr, count, rank, world, nat.ptr(out), ctypes.byref(inf)))
        d_s.free()
        return pt_unpack(G, out, inf.value)

    def want(s, count):
        ss = nat.limbs_to_ints(s[:count]) if count else []
        return G.mul(G.gen, sum(a * b for a, b in zip(ks, ss)) % G.r)

    uni = rng.integers(0, 2 ** 64, size=(n, 4), dtype=np.uint64)
    uni[:, 3] &= np.uint64((1 << 60) - 1)
    assert run(uni, n) == want(uni, n)
    assert run(uni, n - 777) == want(uni, n - 777)          # prefix of the table
    assert run(uni, 1) == want(uni, 1)
    assert run(uni, 0) is None
    acc = None
    for rank in range(3):                                    # window shards add up
        acc = G.add(acc, run(uni, n, rank, 3))
    assert acc == want(uni, n)
    same = np.tile(np.array([[0x123456789abcdef1, 0x0fedcba987654321, 0x1111111122222222, 0x0333333344444444]],
                            dtype=np.uint64), (n, 1))
    assert run(same, n) == want(same, n)
    bits = np.zeros((n, 4), dtype=np.uint64)
    bits[:, 0] = rng.integers(0, 3, size=n, dtype=np.uint64)
    assert run(bits, n) == want(bits, n)
    with pytest.raises(ValueError, match="mismatch"):
        d_s = nat.DeviceBuffer((n + 1) * 32)
        nat.check(nat.lib.zkb_msm_table_dev(tab, d_s.ptr, n + 1, 0, 1, nat.ptr(np.zeros(ab // 8, dtype=np.uint64)),
                                            ctypes.byref(ctypes.c_int())))
    nat.lib.zkb_msm_table_free(tab)
    for b in (d_k, d_gen, d_pts):
        b.free()
