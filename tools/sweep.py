"""Raw NTT / MSM sweep (BASELINE.json configs[4], SURVEY.md section 8d "Config 5"): device-resident inputs, CUDA-event timing on
the library stream, warm-up 2, median of 5.  Inputs: uniform Fr from numpy PCG64(seed = log_n); MSM bases P_i = k_i * G built on
the GPU by the fixed-base kernel.  Prints one JSON object per line and writes them to gpurun_out/sweep_<tag>.jsonl.

  python tools/sweep.py [tag] [max_log_g1=26] [max_log_ntt=26]
Rooflines: NTT against the HBM copy peak of MEASURED_PEAKS.json with 64 B/element of algorithmic traffic (the transform is
integer-bound); MSM against the IMAD.WIDE issue rate measured here, with the canonical 16 windows x 10
products x (2L^2+L) multiply-adds per point (G2: x3)."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from zksnake_b200 import _native as nat  # noqa: E402
import perf_probe as pp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def median_ms(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        with nat.Timer() as t:
            fn()
        ts.append(t.ms)
    return float(np.median(ts))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    max_g1 = int(sys.argv[2]) if len(sys.argv) > 2 else 26
    max_ntt = int(sys.argv[3]) if len(sys.argv) > 3 else 26
    nat.ensure_init()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    wide = ctypes.c_double()
    nat.check(nat.lib.zkb_imad_peak(1, ctypes.byref(wide)))
    out = open(os.path.join(ROOT, "gpurun_out", f"sweep_{tag}.jsonl"), "w")

    def emit(rec):
        line = json.dumps(rec)
        print(line, flush=True)
        out.write(line + "\n")
        out.flush()

    emit({"kind": "peaks", "hbm_gbs": hbm, "imad_wide_tops": wide.value / 1e12})
    for curve, cname in ((0, "BN254"), (1, "BLS12_381")):
        for log_n in range(16, max_ntt + 1, 2):
            n = 1 << log_n
            d_in = nat.DeviceBuffer(n * 32).upload(pp.rand_fr(n, log_n, curve))
            d_out = nat.DeviceBuffer(n * 32)
            for name, inv, coset in (("ntt", 0, 0), ("intt", 1, 0), ("coset_ntt", 0, 1)):
                ms = median_ms(lambda: nat.check(nat.lib.zkb_ntt_dev(curve, inv, coset, log_n, d_in.ptr, n, d_out.ptr)))
                emit({"kind": name, "curve": cname, "log_n": log_n, "ms": ms, "gelem_s": n / ms / 1e6,
                      "alg_gb_s": 64 * n / ms / 1e6, "hbm_frac": 64 * n / ms / 1e6 / hbm})
            d_in.free(); d_out.free()
    for curve, cname, grp, top in ((0, "BN254", 1, max_g1), (1, "BLS12_381", 1, min(max_g1, 24)), (0, "BN254", 2, min(max_g1, 22)),
                                   (1, "BLS12_381", 2, min(max_g1, 22))):
        L = 8 if curve == 0 else 12
        ops = 16 * 10 * (2 * L * L + L) * (3 if grp == 2 else 1)
        for log_n in range(16, top + 1, 2):
            n = 1 << log_n
            ab = nat.lib.zkb_affine_bytes(curve, grp)
            d_pts = pp.make_points(curve, grp, n, 1)
            res = np.zeros(ab // 8, dtype=np.uint64)
            inf = ctypes.c_int()
            for sname in ("uniform", "pow2_chain", "edge_mix") if (grp == 1 and log_n <= 22) else ("uniform",):
                s = pp.rand_fr(n, 1000 + log_n, curve)
                if sname == "pow2_chain":      # chain-circuit witness: 2^k
                    s = np.zeros((n, 4), dtype=np.uint64)
                    e = (np.arange(n) + 2) % 250
                    s[np.arange(n), e // 64] = np.uint64(1) << (e % 64).astype(np.uint64)
                elif sname == "edge_mix":      # 10 % zeros and ones
                    idx = np.arange(n)
                    s[idx % 20 == 0] = 0
                    s[idx % 20 == 1] = 0
                    s[idx % 20 == 1, 0] = 1
                d_s = nat.DeviceBuffer(n * 32).upload(s)
                ms = median_ms(lambda: nat.check(nat.lib.zkb_msm_dev(curve, grp, d_pts.ptr, d_s.ptr, n, nat.ptr(res), ctypes.byref(inf))),
                               warm=1, reps=3)
                emit({"kind": f"msm_g{grp}", "curve": cname, "log_n": log_n, "scalars": sname, "ms": ms, "mpts_s": n / ms / 1e3,
                      "imad_wide_frac": n * ops / (ms * 1e-3) / wide.value})
                d_s.free()
            d_pts.free()
    # (the CPU restatement beside these numbers is timed by tests/cpu_sweep.py: only tests/ and bench.py's CPU legs run oracle/)
    out.close()


if __name__ == "__main__":
    main()
