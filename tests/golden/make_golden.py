"""Regenerates tests/golden/derived_kats.json from the pure-Python oracle.

The reference (Merricx/zksnake) cannot be built or imported in this environment (its arithmetic lives in Rust crates that are
not vendored; no Rust toolchain), and its own tests hold no numeric golden vectors for this path.  The
vectors written here are therefore DERIVED known-answer vectors -- computed from the published definitions by oracle/*.py -- and
serve as (a) a regression pin on the oracle, (b) fixed inputs/outputs the C++ restatement (oracle/cport) and the CUDA path are
both checked against.  Independent pins of the oracle itself (published constants, standard generator encodings, the O(N^2)
definition, closed forms) live in tests/test_oracle_pins.py.

    python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import groth16 as og  # noqa: E402
from oracle import poly  # noqa: E402
from oracle.curve import group  # noqa: E402
from oracle.fields import PARAMS  # noqa: E402

NAMES = {0: "BN254", 1: "BLS12_381"}


def readme_triplets(p):
    a = [(0, 2, 1), (1, 3, 1)]
    b = [(0, 2, 1), (1, 2, 1)]
    c = [(0, 3, 1), (1, 1, 1), (1, 2, p - 1), (1, 0, p - 5)]
    return a, b, c


def chain_triplets(N):
    a, b, c = [(0, 2, 1)], [(0, 2, 1)], [(0, 3, 1)]
    for i in range(1, N - 1):
        a.append((i, 3 + i - 1, 1)); b.append((i, 2, 1)); c.append((i, 3 + i, 1))
    a.append((N - 1, 3 + N - 2, 1)); b.append((N - 1, 0, 1)); c.append((N - 1, 1, 1))
    return a, b, c


def chain_witness(N, p, inp=2):
    v, cur = [], inp
    for _ in range(N - 1):
        cur = cur * inp % p
        v.append(cur)
    return [1, v[-1], inp] + v


def main():
    out = {"_comment": "derived KATs (oracle/*.py), NOT reference output; see make_golden.py", "curves": {}}
    for cid in (0, 1):
        P = PARAMS[cid]
        r = P.r
        rnd = random.Random(1000 + cid)
        cur = {}
        # ---- NTT vectors
        ntt = []
        for log_n, length in ((0, 1), (3, 8), (3, 5), (5, 32), (6, 70)):
            c = [rnd.randrange(r) for _ in range(length)]
            n = 1 << log_n
            ntt.append({
                "log_n": log_n, "input": [hex(x) for x in c],
                "fft": [hex(x) for x in poly.fft(cid, c, n)],
                "ifft": [hex(x) for x in poly.ifft(cid, c, n)],
                "coset_fft": [hex(x) for x in poly.fft(cid, c, n, coset=True)],
                "coset_ifft": [hex(x) for x in poly.ifft(cid, c, n, coset=True)],
            })
        cur["ntt"] = ntt
        cur["omega"] = {str(k): hex(P.omega(k)) for k in (1, 2, 16, 20)}
        # ---- MSM vectors: P_i = k_i * G
        msm = {}
        for g2 in (False, True):
            G = group(cid, g2)
            ks = [rnd.randrange(1, r) for _ in range(6)] + [1, 1]
            pts = [G.mul(G.gen, k) for k in ks]
            pts[3] = None                                   # identity among the bases
            sc = [rnd.randrange(r) for _ in range(8)]
            sc[0], sc[1], sc[2] = 0, 1, r - 1               # edge scalars
            res = G.msm(pts, sc)
            msm["g2" if g2 else "g1"] = {
                "points": [None if p is None else G.to_bytes(p).hex() for p in pts],
                "scalars": [hex(s) for s in sc], "result": G.to_bytes(res).hex(),
            }
        cur["msm"] = msm
        # ---- Groth16: README circuit and the 4-constraint chain, fixed toxic waste and prover randomness
        proofs = []
        for name, trip, n_rows, n_cols, n_pub, wit in (
            ("readme", readme_triplets(r), 2, 4, 2, [1, 35, 3, 9]),
            ("chain4", chain_triplets(4), 4, 6, 2, chain_witness(4, r)),
        ):
            tr = random.Random(1)
            toxic = tuple(tr.randint(1, r - 1) for _ in range(5))
            rs = random.Random(2)
            rr, ss = rs.randint(1, r - 1), rs.randint(1, r - 1)
            st = og.Setup(cid, *trip, n_rows, n_cols, n_pub, toxic)
            A, B, C, U, V, W, H = og.prove_literal(st, wit, rr, ss)
            assert (A, B, C) == og.prove_closed_form(st, wit, rr, ss)
            proofs.append({"circuit": name, "toxic": [hex(x) for x in toxic], "r": hex(rr), "s": hex(ss),
                           "witness": [hex(x) for x in wit], "U": [hex(x) for x in U], "V": [hex(x) for x in V],
                           "W": [hex(x) for x in W], "H": [hex(x) for x in H], "proof": og.proof_bytes(cid, A, B, C).hex()})
        cur["groth16"] = proofs
        out["curves"][NAMES[cid]] = cur
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "derived_kats.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
