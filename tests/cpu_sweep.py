"""CPU side of the raw NTT / MSM sweep (SURVEY.md section 8d "Config 5": the CPU restatement timed beside the GPU numbers of
tools/sweep.py): oracle/cport -- ark-style radix-2 FFT and signed-digit Pippenger -- at bounded sizes, on 1 thread (what the
shipped reference wheel does) and on all host threads.  Lives under tests/ because only tests/, smoke() and bench.py's CPU legs
may execute oracle/.  Not collected by pytest (no test_ prefix); run it by hand:

  python tests/cpu_sweep.py [tag]        # appends JSON lines to gpurun_out/sweep_<tag>.jsonl and prints them"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TOP_LIMB = {0: 0x30644e72e131a029, 1: 0x73eda753299d7d48}


def rand_fr(n, seed, curve=0):
    """uniform 256-bit values below (top limb of r) * 2^192 (the generator tools/perf_probe.py uses)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 2 ** 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] %= np.uint64(TOP_LIMB[curve])
    return a


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "cpu"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", f"sweep_{tag}.jsonl"), "a")

    def emit(row):
        line = json.dumps(row)
        print(line, flush=True)
        out.write(line + "\n")

    from oracle import cport
    cport.build()
    try:
        nthreads = len(os.sched_getaffinity(0))
    except AttributeError:
        nthreads = os.cpu_count() or 1
    for curve, cname in ((0, "BN254"), (1, "BLS12_381")):
        for log_n in (16, 20):
            a = rand_fr(1 << log_n, log_n, curve)
            for th in (1, nthreads):
                cport.fft(curve, a, log_n, nthreads=th)
                t0 = time.perf_counter()
                cport.fft(curve, a, log_n, nthreads=th)
                ms = (time.perf_counter() - t0) * 1e3
                emit({"kind": "cpu_ntt", "curve": cname, "log_n": log_n, "threads": th, "ms": ms,
                      "gelem_s": (1 << log_n) / ms / 1e6, "impl": "oracle/cport (C++ restatement)"})
        for log_n in (16, 18):
            n = 1 << log_n
            pts = cport.chain_points(curve, 1, 7, n, nthreads)
            sc = rand_fr(n, 1000 + log_n, curve)
            for th in (1, nthreads):
                t0 = time.perf_counter()
                cport.msm(curve, 1, pts, sc, nthreads=th)
                ms = (time.perf_counter() - t0) * 1e3
                emit({"kind": "cpu_msm_g1", "curve": cname, "log_n": log_n, "threads": th, "ms": ms, "mpts_s": n / ms / 1e3,
                      "impl": "oracle/cport (C++ restatement)"})
    out.close()


if __name__ == "__main__":
    main()
