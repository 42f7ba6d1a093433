"""Key-file round trip at benchmark size: Groth16 proving key of a 2^log_n chain circuit ->
ProvingKey.to_bytes() -> ProvingKey.from_bytes() -> prove with the re-read key; prints one JSON line.
The per-point host decoder (the mirror of the reference's from_hex loop, ecc.py:128-142) is timed on a small sample beside it.

    python tools/keyfile_bench.py [--log-n 20] [--curve BN254]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--curve", default="BN254")
    args = ap.parse_args()
    from zksnake_b200 import groth16 as gm
    from zksnake_b200 import r1cs as rm
    n = 1 << args.log_n
    circuit, pub, priv = rm.chain_circuit(n, args.curve)
    g = gm.Groth16(circuit, args.curve)
    t0 = time.perf_counter()
    g.setup()
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    blob = g.proving_key.to_bytes()
    t_ser = time.perf_counter() - t0
    t0 = time.perf_counter()
    pk2 = gm.ProvingKey.from_bytes(blob, crv=args.curve)
    t_de = time.perf_counter() - t0
    assert pk2.to_bytes() == blob
    seq = iter([5, 7, 5, 7])
    old = gm.get_random_int
    gm.get_random_int = lambda n_max: next(seq)
    try:
        p1 = g.prove(pub, priv).to_bytes()
        g2 = gm.Groth16(rm.chain_circuit(n, args.curve)[0], args.curve)
        g2.proving_key, g2.verifying_key = pk2, g.verifying_key
        p2 = g2.prove(pub, priv).to_bytes()
    finally:
        gm.get_random_int = old
    assert p1 == p2
    # host per-point decoder on a sample
    ec = g.ec
    size = len(pk2.tau_1.to_bytes()) // len(pk2.tau_1)
    off = 7 * size + 8
    sample = 64
    t0 = time.perf_counter()
    for i in range(sample):
        ec.PointG1.from_bytes(blob[off + i * size:off + (i + 1) * size])
    t_host_g1 = (time.perf_counter() - t0) / sample
    n_points = len(pk2.tau_1) + len(pk2.tau_2) + len(pk2.target_1) + len(pk2.kdelta_1)
    print(json.dumps({"tool": "keyfile_bench", "curve": args.curve, "log_n": args.log_n, "key_bytes": len(blob), "points": n_points,
                      "setup_s": t_setup, "to_bytes_s": t_ser, "from_bytes_s": t_de,
                      "host_per_point_g1_ms": 1e3 * t_host_g1,
                      "host_loop_estimate_s": t_host_g1 * (n_points + len(pk2.tau_2) * 3),
                      "same_proof_from_reloaded_key": True}))


if __name__ == "__main__":
    main()
