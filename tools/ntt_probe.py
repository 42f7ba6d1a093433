"""NTT kernel probe (GPU): the register-resident passes (csrc/ntt_warp.cuh, 4 elements per lane) against the shared-memory passes
(csrc/ntt.cuh) -- bit-exact comparison on random inputs for every transform flavour, then device timings (CUDA events, median
of `reps`), and the same for the batched Groth16 quotient pipeline.

  python tools/ntt_probe.py [min_log=11] [max_log=24] [reps=20]        -> one JSON line per size on stdout"""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zksnake_b200 import _native as nat  # noqa: E402

R = {0: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
     1: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001}


def rand_fr(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 59) - 1)        # < r on both curves
    return a


def mode(warp, el=0):
    os.environ["ZKB_NTT_WARP"] = "1" if warp else "0"
    os.environ["ZKB_NTT_WARP_SINGLE"] = "1" if warp else "0"
    os.environ["ZKB_NTT_EL"] = str(el)


def run_ntt(curve, log_n, d_in, in_len, d_out, inverse, coset, reps):
    n = 1 << log_n
    nat.check(nat.lib.zkb_ntt_dev(curve, inverse, coset, log_n, d_in.ptr, in_len, d_out.ptr))     # warm-up (tables)
    nat.check(nat.lib.zkb_sync())
    times = []
    for _ in range(reps):
        with nat.Timer() as t:
            nat.check(nat.lib.zkb_ntt_dev(curve, inverse, coset, log_n, d_in.ptr, in_len, d_out.ptr))
        times.append(t.ms)
    return float(np.median(times)), d_out.download(count=n * 32)


def main():
    lo = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    hi = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    nat.ensure_init()
    ok = True
    for curve in (0, 1):
        for log_n in range(lo, hi + 1):
            if curve == 1 and log_n not in (lo, 16, 20, 22):
                continue
            n = 1 << log_n
            x = rand_fr(n, log_n * 7 + curve)
            d_in = nat.DeviceBuffer(n * 32).upload(x)
            d_out = nat.DeviceBuffer(n * 32)
            row = {"curve": curve, "log_n": log_n}
            for inverse, coset, in_len in ((0, 0, n), (1, 0, n), (0, 1, n), (1, 1, n), (0, 2, n), (1, 2, n), (0, 0, n - n // 3), (0, 2, 5)):
                if log_n > 20 and (coset == 1 or in_len != n):
                    continue
                mode(False)
                t_old, ref = run_ntt(curve, log_n, d_in, in_len, d_out, inverse, coset, reps if (coset, in_len) == (0, n) else 1)
                res = {}
                for el in (2,):
                    mode(True, el)
                    t_new, got = run_ntt(curve, log_n, d_in, in_len, d_out, inverse, coset, reps if (coset, in_len) == (0, n) else 1)
                    same = bool((got == ref).all())
                    ok &= same
                    res[f"el{el}"] = (round(t_new, 4), same)
                if (coset, in_len) == (0, n):
                    row["inv" if inverse else "fwd"] = {"smem_ms": round(t_old, 4), **{k: v[0] for k, v in res.items()},
                                                        "exact": all(v[1] for v in res.values())}
                else:
                    row.setdefault("variants_exact", True)
                    row["variants_exact"] &= all(v[1] for v in res.values())
            d_in.free()
            d_out.free()
            print(json.dumps(row), flush=True)
        # Groth16 quotient pipeline (3 + 3 + 1 transforms), batched vs one by one
        for log_n in (12, 16, 20):
            n = 1 << log_n
            a, b = rand_fr(n, 100 + log_n), rand_fr(n, 200 + log_n)
            work = nat.DeviceBuffer(7 * n * 32)
            nat.check(nat.lib.zkb_h2d(work.at(0), nat.ptr(a), n * 32))
            nat.check(nat.lib.zkb_h2d(work.at(n * 32), nat.ptr(b), n * 32))
            nat.check(nat.lib.zkb_vec_op_dev(curve, 0, n, work.at(0), n, work.at(n * 32), n, work.at(2 * n * 32)))     # c = a * b
            outs = {}
            for name, warp, el in (("smem", False, 0), ("el2", True, 2)):
                mode(warp, el)
                args = [work.at(k * n * 32) for k in range(7)]
                nat.check(nat.lib.zkb_groth16_h_dev(curve, log_n, *args, 1))
                nat.check(nat.lib.zkb_sync())
                ts = []
                for _ in range(max(reps // 2, 3)):
                    with nat.Timer() as t:
                        nat.check(nat.lib.zkb_groth16_h_dev(curve, log_n, *args, 1))
                    ts.append(t.ms)
                res = np.zeros((4 * n, 4), dtype=np.uint64)
                nat.check(nat.lib.zkb_d2h(nat.ptr(res), work.at(3 * n * 32), 4 * n * 32))
                outs[name] = (float(np.median(ts)), res[:2 * n].copy(), res[3 * n:].copy())
            same = all((outs[k][1] == outs["smem"][1]).all() and (outs[k][2] == outs["smem"][2]).all() for k in ("el2",))
            ok &= bool(same)
            print(json.dumps({"curve": curve, "groth16_h_log_n": log_n, "exact": bool(same),
                              **{k + "_ms": round(v[0], 4) for k, v in outs.items()}}), flush=True)
            work.free()
    print(json.dumps({"all_exact": ok}), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
