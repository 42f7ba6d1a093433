"""bench.py -- the measurement contract.

Workload (BASELINE.json `metric`: "Groth16 prove ms @2^20 BN254 at 1/2/4/8 GPU; MSM Mpts/s; NTT GB/s"):
one Groth16 proof of the reference's benchmark circuit (benchmarks/benchmark_groth16.py:11-24, the multiplication chain) with
2^20 constraints on BN254.  A "step" is one prove: witness -> A.w,B.w,C.w (SpMV) -> quotient H (7 NTTs) -> 5 MSMs -> proof.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 20] [--curve BN254]
  python bench.py --impl reference ...      # the CPU arm: the C++ restatement of the reference's arkworks path (oracle/cport)

Own arm, N ranks (torchrun): every rank holds 1/N of each proving-key vector, computes the witness polynomials redundantly
(~1 ms of NTT work) and its slice of the five MSMs; one NCCL all-gather of the partial sums (< 1 KiB per rank) finishes the
proof.  One proof is split over N GPUs, so `scaling` is "strong" and `value` is the latency of that one proof.

JSON keys beyond the base contract: `roofline` (dominant kernel: MSM bucket accumulation, integer-pipe bound, SURVEY.md section
8d), `roofline_ntt` (HBM bound), `cpu_baseline`, `breakdown_ms`, `msm_mpts_s`, `ntt_gelem_s`.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CURVE_IDS = {"BN254": 0, "BLS12_381": 1}
METRIC = "groth16_prove_ms"
UNIT = "ms"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="zkb200", choices=["zkb200", "reference"])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--curve", default="BN254", choices=sorted(CURVE_IDS))
    ap.add_argument("--workload", default="groth16", choices=["groth16", "plonk"],
                    help="groth16: BASELINE.json's headline (Groth16 prove of the chain circuit); plonk: configs[3], PlonK prove of "
                         "the same chain as gates through DevicePlonk.prove_packed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the closed-form parity check of the untimed region")
    ap.add_argument("--no-list-api", action="store_true", help="skip the Groth16.prove(list, list) end-to-end measurement")
    ap.add_argument("--shard-mode", default="windows", choices=["windows", "points"],
                    help="multi-GPU split of every MSM: scalar windows (whole key per GPU) or point ranges (1/N of the key per GPU)")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0, help="CPU seconds (wall) the cpu_baseline sample may take")
    return ap.parse_args()


def workload_config(args):
    if args.workload == "plonk":
        return {
            "workload": f"plonk-prove chain circuit (benchmarks/benchmark_plonk.py) 2^{args.log_n} gates {args.curve}",
            "log_n": args.log_n,
            "curve": args.curve,
            "msm": "9 x G1 of 2^%d + 6 points over the SRS" % args.log_n,
            "ntt": "reference: 5 x n, 15 x 4n, 6 x 8n; this prover: 5 x n + 6 x 4n (quotient on one coset)",
            "l2": "working set (SRS table + 14 quotient-domain vectors, > 2 GiB) exceeds L2; no flush",
        }
    return {
        "workload": f"groth16-prove chain circuit (benchmarks/benchmark_groth16.py) 2^{args.log_n} constraints {args.curve}",
        "log_n": args.log_n,
        "curve": args.curve,
        "msm": "4 x G1 + 1 x G2 of 2^%d points" % args.log_n,
        "ntt": "7 x 2^%d (3 inverse, 3 coset forward, 1 coset inverse)" % args.log_n,
        "l2": "working set 450 MiB (key 320 + scalars 130; 4.6 GiB with the fixed-base tables) exceeds L2; no flush",
    }


# ------------------------------------------------------------------------------------------------------------------------
# synthetic circuit in packed form (no Python big-int lists at 2^20: numpy all the way)
# ------------------------------------------------------------------------------------------------------------------------
def chain_csr(n_constraints):
    """CSR (row_ptr u64, col u32, val (nnz,4) u64) of A, B, C for the chain circuit -- same layout as
    zksnake_b200.r1cs.chain_circuit: columns [1, out, inp, v0..v_{N-2}]."""
    N = n_constraints
    row_ptr = np.arange(N + 1, dtype=np.uint64)
    one = np.zeros((N, 4), dtype=np.uint64)
    one[:, 0] = 1
    a_col = np.empty(N, dtype=np.uint32)
    a_col[0] = 2
    a_col[1:] = 3 + np.arange(N - 1, dtype=np.uint32)
    b_col = np.full(N, 2, dtype=np.uint32)
    b_col[N - 1] = 0
    c_col = np.empty(N, dtype=np.uint32)
    c_col[:N - 1] = 3 + np.arange(N - 1, dtype=np.uint32)
    c_col[N - 1] = 1
    return [(row_ptr, a_col, one.copy()), (row_ptr, b_col, one.copy()), (row_ptr, c_col, one.copy())]


def chain_witness(n_constraints, r, inp=2):
    """[1, out, inp, v0..v_{N-2}] with v_i = inp^(i+2) mod r, as an (N+2, 4) uint64 array."""
    N = n_constraints
    vals = bytearray()
    cur = inp % r
    vs = []
    for _ in range(N - 1):
        cur = cur * inp % r
        vs.append(cur)
    w = [1, vs[-1], inp % r] + vs
    for x in w:
        vals += x.to_bytes(32, "little")
    return np.frombuffer(bytes(vals), dtype=np.uint64).reshape(N + 2, 4).copy()


# ------------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.15] or self.rows
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------
# CPU arm (C++ restatement of the reference's arkworks path) -- also the cpu_baseline of the own arm
# ------------------------------------------------------------------------------------------------------------------------
class CpuProver:
    """Groth16.prove of the chain circuit on the host cores through oracle/cport (kind "port").  The proving-key vectors are
    synthetic ((k0+i)*G: the MSM cost does not depend on which points they are); everything else is the real pipeline."""

    def __init__(self, curve, log_n):
        from oracle import cport
        from oracle.fields import PARAMS
        self.cport = cport
        cport.build()
        self.curve, self.log_n = curve, log_n
        self.n = n = 1 << log_n
        self.r = PARAMS[curve].r
        self.csr = chain_csr(n)
        self.witness = chain_witness(n, self.r)
        self.m = n + 2
        # every host core this process may use, passed explicitly: torchrun exports OMP_NUM_THREADS=1, which would otherwise
        # turn the "all host threads" arm into a single-threaded one
        try:
            self.threads = len(os.sched_getaffinity(0))
        except AttributeError:
            self.threads = os.cpu_count() or 1
        nt = self.threads
        g1 = lambda k0, cnt: cport.chain_points(curve, 1, k0, cnt, nt)  # noqa: E731
        g2 = lambda k0, cnt: cport.chain_points(curve, 2, k0, cnt, nt)  # noqa: E731
        self.key = {
            "tau1": g1(1, n), "tau2": g2(1, n), "target1": g1(7, n), "kdelta1": g1(11, self.m - 2),
            "alpha1": g1(3, 1)[0], "beta1": g1(5, 1)[0], "beta2": g2(5, 1)[0],
            "delta1": g1(9, 1)[0], "delta2": g2(9, 1)[0],
        }

    def prove(self):
        t0 = time.perf_counter()
        out = self.cport.groth16_prove(self.curve, self.log_n, self.csr, self.m, 2, self.witness, self.key, 12345, 67890,
                                       nthreads=self.threads)
        return (time.perf_counter() - t0) * 1e3, out


def cpu_sample_log_n(curve, target_log_n, budget_s):
    """Largest log_n <= target whose single prove is predicted to fit the budget (calibrated on a 2^12 prove; prove time is
    ~linear in n at these sizes)."""
    probe = CpuProver(curve, 12)
    probe.prove()
    ms, _ = probe.prove()
    log_n = 12
    while log_n < target_log_n and ms * (1 << (log_n + 1 - 12)) / 1e3 <= budget_s:
        log_n += 1
    return log_n


def run_cpu_arm(args, steps, warmup, budget_total_s):
    """Times the CPU prover.  Each step is one prove at the sample size -- the full workload size whenever one prove fits the
    per-step budget (2^20 BN254 on 16 cores: ~8 s); otherwise the largest power of two that fits, scaled linearly and labelled
    (Pippenger's per-point cost falls slowly with n, so scaling from a smaller n slightly OVERSTATES the CPU time)."""
    if args.workload == "plonk":
        return run_cpu_arm_plonk(args, steps, warmup, budget_total_s)
    curve = CURVE_IDS[args.curve]
    per_step = budget_total_s / max(1, steps + warmup)
    s_log = cpu_sample_log_n(curve, args.log_n, per_step)
    prover = CpuProver(curve, s_log)
    for _ in range(warmup):
        prover.prove()
    times = [prover.prove()[0] for _ in range(steps)]
    scale = 1 << (args.log_n - s_log)
    ms = float(np.mean(times)) * scale
    sample = (f"full Groth16.prove (SpMV + QAP quotient on the 2n domain + 5 ark-style Pippenger MSMs + assembly) at "
              f"2^{s_log} constraints, {steps} runs, mean"
              + ("" if scale == 1 else f", scaled x{scale} linearly to 2^{args.log_n} (slightly overstates the CPU time)")
              + "; synthetic key points (k0+i)*G; OpenMP over MSM windows / FFT butterflies (the shipped reference wheel runs "
                "these single-threaded)")
    return ms, {"value": ms, "unit": UNIT, "cores": prover.threads, "kind": "port", "sample": sample,
                "sample_log_n": s_log, "same_size_as_workload": scale == 1,
                "sample_ms_unscaled": float(np.mean(times)), "host_cpus": os.cpu_count()}


def run_cpu_arm_plonk(args, steps, warmup, budget_total_s):
    """PlonK prove on the host cores, composed from its measured heavy parts: the reference's Plonk.prove
    (python/zksnake/plonk/protocol.py:157-484) makes 9 multiexps over the SRS (n + 6 points) and 26 transforms (5 of size n, 15
    of size 4n, 6 of size 8n -- SURVEY.md section 3.3); one of each kind is timed per step through oracle/cport (ark's
    algorithms) and multiplied by its count.  The Python list glue between them is NOT included, so this UNDERSTATES the
    reference's CPU time."""
    from oracle import cport
    from oracle.fields import PARAMS
    cport.build()
    curve = CURVE_IDS[args.curve]
    r = PARAMS[curve].r
    try:
        nt = len(os.sched_getaffinity(0))
    except AttributeError:
        nt = os.cpu_count() or 1
    per_step = budget_total_s / max(1, steps + warmup)
    # calibrate at 2^14, then the largest size whose composed step fits the budget
    def parts(log_n):
        n = 1 << log_n
        rng = np.random.default_rng(log_n)
        def rand_fr(count):
            a = rng.integers(0, 1 << 60, size=(count, 4), dtype=np.uint64)      # < r on both curves
            return a
        pts = cport.chain_points(curve, 1, 5, n, nt)
        sc = rand_fr(n)
        vec = {k: rand_fr(n << k) for k in (0, 2, 3)}
        return pts, sc, vec

    def one(log_n, data):
        pts, sc, vec = data
        t0 = time.perf_counter()
        cport.msm(curve, 1, pts, sc, nt)
        t_msm = time.perf_counter() - t0
        t_fft = {}
        for k, v in vec.items():
            t0 = time.perf_counter()
            cport.fft(curve, v, log_n + k, nthreads=nt)
            t_fft[k] = time.perf_counter() - t0
        return (9 * t_msm + 5 * t_fft[0] + 15 * t_fft[2] + 6 * t_fft[3]) * 1e3, t_msm * 1e3, {k: v * 1e3 for k, v in t_fft.items()}

    base_log = min(14, args.log_n)
    s_log = base_log
    ms_base, _, _ = one(base_log, parts(base_log))
    while s_log < args.log_n and ms_base * (1 << (s_log + 1 - base_log)) / 1e3 <= per_step:
        s_log += 1
    data = parts(s_log)
    for _ in range(warmup):
        one(s_log, data)
    runs = [one(s_log, data) for _ in range(steps)]
    scale = 1 << (args.log_n - s_log)
    ms = float(np.mean([x[0] for x in runs])) * scale
    sample = (f"composed from measured parts at 2^{s_log} gates: 9 x (one ark-style Pippenger G1 MSM) + 5 x FFT(n) + 15 x FFT(4n) + "
              f"6 x FFT(8n), {steps} runs, mean" + ("" if scale == 1 else f", scaled x{scale} linearly to 2^{args.log_n}")
              + "; the reference's Python list glue, batch inversion and divisions are NOT included (understates the CPU time)")
    return ms, {"value": ms, "unit": UNIT, "cores": nt, "kind": "port", "sample": sample, "sample_log_n": s_log,
                "same_size_as_workload": scale == 1, "msm_ms": float(np.mean([x[1] for x in runs])),
                "fft_ms": {f"{1 << k}n": float(np.mean([x[2][k] for x in runs])) for k in (0, 2, 3)}, "host_cpus": os.cpu_count()}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = args.steps, args.warmup
    # every step runs the FULL workload size when it fits ~25 s (2^20 BN254 Groth16 on 16 cores: ~8 s per prove, so K = 20, W = 5
    # is ~3.5 minutes); the budget only bites on small hosts / the larger curve
    ms, base = run_cpu_arm(args, steps, warmup, budget_total_s=25.0 * max(1, steps + warmup))
    line = {
        "impl": "reference", "metric": metric_name(args), "value": ms, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64 limbs (int)",
        "data": "synthetic", "config": workload_config(args), "cpu_baseline": base,
        "e2e": {"value": ms, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def metric_name(args):
    return "plonk_prove_ms" if args.workload == "plonk" else METRIC


# ------------------------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------------------------
def load_traffic():
    """Per-launch DRAM bytes of the dominant kernels from the committed `ncu --set full` capture (profiles/traffic.json)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def main_own(args):
    import random

    from zksnake_b200 import _native as nat
    from zksnake_b200 import groth16 as zg
    from zksnake_b200.r1cs import chain_circuit

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not nat.gpu_available():
        raise SystemExit("bench.py: no CUDA device visible -- zksnake_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    td = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(local_rank)
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nat.ensure_init(local_rank)
    curve = CURVE_IDS[args.curve]
    n = 1 << args.log_n

    def barrier():
        if td is not None:
            td.barrier()
            torch.cuda.synchronize()
        nat.check(nat.lib.zkb_sync())

    # ---- setup (untimed): circuit, device-resident proving key slice, device-resident R1CS ----
    t_setup = time.perf_counter()
    r1cs, pub, priv = chain_circuit(n, args.curve)
    rnd = random.Random(1)
    toxic = [rnd.randint(1, r1cs.p - 1) for _ in range(5)]
    rs = random.Random(2)
    r_rand, s_rand = rs.randint(1, r1cs.p - 1), rs.randint(1, r1cs.p - 1)
    # the randomness hook the reference's harnesses patch (protocol.py:11), seeded identically on every rank: setup() draws the
    # toxic waste from it (rank 0's draws are what every rank uses), so the key and the proof bytes are the same at every N
    hook_values = list(toxic)
    zg.get_random_int = lambda n_max: hook_values.pop(0)
    prover = zg.Groth16(r1cs, args.curve, shard=(rank, world), shard_mode=args.shard_mode)
    prover.setup()
    m = n + 2
    w_host_np = nat.ints_to_limbs(pub + priv)
    # pinned host copy of the witness (the e2e path copies from here every step) and a device-resident copy (the `value` path)
    pinned = ctypes.c_void_p()
    nat.check(nat.lib.zkb_host_alloc(m * 32, ctypes.byref(pinned)))
    w_pinned = np.ctypeslib.as_array(ctypes.cast(pinned, ctypes.POINTER(ctypes.c_uint64)), shape=(m, 4))
    w_pinned[:] = w_host_np
    w_dev = nat.DeviceBuffer(m * 32).upload(w_host_np)
    setup_s = time.perf_counter() - t_setup

    # ---- first proof (builds the NTT tables, grows the scratch arena); the parity check below uses its bytes ----
    proof = prover.prove_packed(w_dev, r_rand, s_rand)
    proof_hex = proof.to_bytes().hex()

    # ---- parity (untimed): the proof against the closed-form exponents of the known toxic waste, computed
    # by the CPU oracle with Python ints -- an independent route (no NTT, no MSM).  Rank 0 checks; every rank holds the same bytes.
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import groth16 as og
        st = og.Setup(curve, r1cs.A.triplets, r1cs.B.triplets, r1cs.C.triplets, n, n + 2, 2, tuple(toxic))
        pa, pb, pc = og.prove_closed_form(st, pub + priv, r_rand, s_rand)
        parity = og.proof_bytes(curve, pa, pb, pc).hex() == proof_hex
        del st

    # ---- warm-up (after the oracle's seconds of CPU work on rank 0: the timed region must start on a warm, clocked-up GPU) ----
    barrier()
    for _ in range(max(args.warmup, 1)):
        proof = prover.prove_packed(w_dev, r_rand, s_rand)
        proof_e2e = prover.prove_packed(w_pinned, r_rand, s_rand)
    assert proof.to_bytes().hex() == proof_hex and proof_e2e.to_bytes().hex() == proof_hex

    # ---- timed: device-resident ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    nat.check(nat.lib.zkb_prof_enable(1))
    barrier()
    launches0 = nat.lib.zkb_launch_count()
    prover.phase_ms.clear()
    t0 = time.perf_counter()
    with nat.Timer() as tm:
        for _ in range(args.steps):
            prover.prove_packed(w_dev, r_rand, s_rand)
    barrier()
    t1 = time.perf_counter()
    host_phases = {k: round(v / max(prover.phase_ms.get("proofs", 1), 1), 3) for k, v in prover.phase_ms.items() if k != "proofs"}
    launches = nat.lib.zkb_launch_count() - launches0
    dev_ms = tm.ms / args.steps
    wall_ms = (t1 - t0) * 1e3 / args.steps
    prof = nat.prof_read()
    nat.check(nat.lib.zkb_prof_enable(0))

    # ---- timed: end to end from pinned host memory through the public API ----
    barrier()
    h2d0, d2h0 = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    nat.lib.zkb_transfer_count(ctypes.byref(h2d0), ctypes.byref(d2h0))
    from zksnake_b200 import dist as zdist
    dist_t0 = dict(zdist.TRANSFER)
    t2 = time.perf_counter()
    for _ in range(args.steps):
        p2 = prover.prove_packed(w_pinned, r_rand, s_rand)
        _ = p2.to_bytes()
    barrier()
    t3 = time.perf_counter()
    e2e_ms = (t3 - t2) * 1e3 / args.steps
    h2d1, d2h1 = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    nat.lib.zkb_transfer_count(ctypes.byref(h2d1), ctypes.byref(d2h1))
    h2d_step = (h2d1.value - h2d0.value) // args.steps   # counted by the library around every cudaMemcpy it issues
    d2h_step = (d2h1.value - d2h0.value) // args.steps
    h2d_step += (zdist.TRANSFER["h2d"] - dist_t0["h2d"]) // args.steps   # ... plus what zksnake_b200.dist moved through torch
    d2h_step += (zdist.TRANSFER["d2h"] - dist_t0["d2h"]) // args.steps
    clocks = sampler.stop(t0, t3) if rank == 0 else None

    # ---- the reference-signature call: Groth16.prove(list[int], list[int]) (protocol.py:115-131), wall clock, every step
    # marshalling 2^20 Python ints (csrc/pymarshal.cpp) + H2D + prove + three points back
    list_api_ms = None
    if world == 1 and not args.no_list_api:
        hook_values[:] = [r_rand, s_rand] * (args.steps + 2)
        p3 = prover.prove(pub, priv)
        assert p3.to_bytes().hex() == proof_hex
        nat.check(nat.lib.zkb_sync())
        t_l0 = time.perf_counter()
        for _ in range(args.steps):
            _ = prover.prove(pub, priv).to_bytes()
        list_api_ms = (time.perf_counter() - t_l0) * 1e3 / args.steps

    if td is not None:
        t = torch.tensor([dev_ms, wall_ms, e2e_ms], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        dev_ms, wall_ms, e2e_ms = t.tolist()
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        td.all_reduce(lt)
        launches = int(lt.item())
        # where each rank's host spent the sharded proof: its own work up to the partial sums | waiting for the slowest rank | assembly
        ph = torch.tensor([host_phases.get(k, 0.0) for k in ("partial", "exchange", "assemble")], dtype=torch.float64, device="cuda")
        allph = torch.empty(world * 3, dtype=torch.float64, device="cuda")
        td.all_gather_into_tensor(allph, ph)
        allph = allph.cpu().reshape(world, 3).tolist()
        host_phases = {"partial": [round(r[0], 3) for r in allph], "exchange": [round(r[1], 3) for r in allph],
                       "assemble": [round(r[2], 3) for r in allph], "note": "per rank, mean ms per proof, device-resident leg"}
        if zdist.SPREAD_TRACE is not None:      # ZKB_SPREAD_TRACE=1: device timestamps of the exchange steps, per rank
            tk = ("chains", "uv", "recv", "quot", "h", "end")
            cnt = max(zdist.SPREAD_TRACE.get("proofs", 0), 1)
            tv = torch.tensor([zdist.SPREAD_TRACE.get(k, 0.0) / cnt for k in tk], dtype=torch.float64, device="cuda")
            allt = torch.empty(world * len(tk), dtype=torch.float64, device="cuda")
            td.all_gather_into_tensor(allt, tv)
            allt = allt.cpu().reshape(world, len(tk)).tolist()
            host_phases["device_trace_ms"] = {k: [round(r[i], 3) for r in allt] for i, k in enumerate(tk)}
            pk_ = sorted(prof)
            pv = torch.tensor([prof[k][0] / args.steps for k in pk_], dtype=torch.float64, device="cuda")
            allp = torch.empty(world * len(pk_), dtype=torch.float64, device="cuda")
            td.all_gather_into_tensor(allp, pv)
            allp = allp.cpu().reshape(world, len(pk_)).tolist()
            host_phases["breakdown_ms_per_rank"] = {k: [round(r[i], 3) for r in allp] for i, k in enumerate(pk_)}

    if rank != 0:
        if td is not None:
            td.destroy_process_group()
        return 0

    # ---- roofline (rank 0's kernels) ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"
    imad = ctypes.c_double()
    nat.check(nat.lib.zkb_imad_peak(0, ctypes.byref(imad)))
    imad_wide = ctypes.c_double()
    nat.check(nat.lib.zkb_imad_peak(1, ctypes.byref(imad_wide)))
    traffic = load_traffic()
    lo, hi = prover._slice
    pts_per_launch = (hi - lo) / (world if args.shard_mode == "windows" else 1)   # point-equivalents of this rank's share
    # canonical algorithmic work of a G1 MSM = W*10 Fq products per point with c = 16 (W = 16), one
    # product = 2L^2+L 32x32->64 multiply-adds (L = 8 limbs BN254, 12 BLS12-381)  => 21760 / 48000 per point.  In this
    # library every one of them is one IMAD.WIDE.U32, whose issue rate (32 lanes/clk/SM, HALF the IMAD.LO rate; measured by
    # zkb_imad_peak(1) with loop-variant operands, tools/ffbench.cu) is therefore the roofline denominator.
    L = 8 if curve == 0 else 12
    ops_per_pt = 16 * 10 * (2 * L * L + L)
    c_c, c_w = ctypes.c_uint32(), ctypes.c_uint32()
    nat.check(nat.lib.zkb_groth16_pk_msm_info(prover._pk_handle, 0, ctypes.byref(c_c), ctypes.byref(c_w)))
    win_c, win_w = int(c_c.value), int(c_w.value)
    acc_ms, acc_cnt = prof["msm_accum_g1"]
    roofline = None
    peak_t = imad_wide.value / 1e12
    mul_ops = 2 * L * L + L                      # 32x32->64 multiply-adds of one Fq Montgomery product
    if acc_cnt:
        per_launch_ms = acc_ms / acc_cnt
        # `achieved` counts what the kernel EXECUTES: W (this key's window count, not the canonical 16) XYZZ mixed additions of
        # 10 Fq products per point ("W = the implementation's window, report it")
        # (per proof there are four G1 MSMs of win_w / world windows each; with the chains spread over the ranks this rank's [K w]
        # MSM covers prover._kw_windows[1] windows instead -- possibly none, then there are three launches: count the work per proof)
        g1_msms = 4.0
        if getattr(prover, "_kw_windows", None) is not None and args.shard_mode == "windows":
            g1_msms = 3.0 + prover._kw_windows[1] / (win_w / world)
        step_s = acc_ms / args.steps * 1e-3
        executed = g1_msms * pts_per_launch * win_w * 10 * mul_ops / step_s / 1e12
        canonical = g1_msms * pts_per_launch * ops_per_pt / step_s / 1e12
        roofline = {"kernel": "msm_accumulate_kernel<G1>", "bound": "int32-pipe", "achieved": executed,
                    "peak": peak_t, "unit": "T 32x32->64 multiply-add lane-ops/s (IMAD.WIDE)",
                    "frac": executed / peak_t,
                    "peak_source": "zkb_imad_peak(1) microbenchmark run inside this bench (mad.wide.u32, loop-variant "
                                   "multiplicand, 8 chains/thread)",
                    "imad_lo_peak": imad.value / 1e12,
                    "window_bits": win_c, "windows": win_w, "executed_ops_per_point": win_w * 10 * mul_ops,
                    "points_per_launch": pts_per_launch,
                    # the same launch against the CANONICAL count of SURVEY 8d (c = 16, W = 16): work the table path avoids counts
                    "canonical_ops_per_point": ops_per_pt, "achieved_canonical": canonical, "frac_canonical": canonical / peak_t,
                    "launch_ms": per_launch_ms, "launches": acc_cnt,
                    "traffic": (traffic.get("msm_accumulate_g1") or {}).get("dram_bytes"),
                    "traffic_detail": traffic.get("msm_accumulate_g1"), "share_of_step": acc_ms / args.steps / dev_ms}
        msm_all_ms = sum(prof[k][0] for k in ("msm_sort", "msm_accum_g1", "msm_accum_g2", "msm_reduce")) / args.steps
        # executed additions of ALL FIVE MSMs (the G2 mixed addition is 8 Fq2 products + 2 Fq2 squarings = 28 Fq products)
        # over all MSM kernels of the proof, sorts and reductions included
        ops_step = pts_per_launch * win_w * mul_ops * (g1_msms * 10 + 28)   # (28: the one-thread Karatsuba count, kept as the work unit)
        roofline["whole_msm_frac"] = ops_step / (msm_all_ms * 1e-3) / imad_wide.value if msm_all_ms else None
    roofline_g2 = None
    g2_ms, g2_cnt = prof["msm_accum_g2"]
    if g2_cnt:
        c2_c, c2_w = ctypes.c_uint32(), ctypes.c_uint32()
        nat.check(nat.lib.zkb_groth16_pk_msm_info(prover._pk_handle, 1, ctypes.byref(c2_c), ctypes.byref(c2_w)))
        per2 = g2_ms / g2_cnt
        lanes, ctas = ctypes.c_int(), ctypes.c_int()
        nat.check(nat.lib.zkb_msm_kernel_info(curve, 2, ctypes.byref(lanes), ctypes.byref(ctas)))
        # one thread per point: 8 Fq2 products + 2 Fq2 squarings = 28 reduced Fq products (Karatsuba); two lanes per point
        # (msm_pair.cuh): per lane 8 lazy dot products of 3L^2+L multiply-adds + 2 plain products
        add_ops = 28 * mul_ops if lanes.value == 1 else 2 * (8 * (3 * L * L + L) + 2 * mul_ops)
        ex2 = pts_per_launch * int(c2_w.value) * add_ops / (per2 * 1e-3) / 1e12
        roofline_g2 = {"kernel": "msm_accumulate_kernel<G2>" if lanes.value == 1 else "msm_accumulate_pair_kernel<G2>",
                       "bound": "int32-pipe", "achieved": ex2, "peak": peak_t,
                       "unit": "T 32x32->64 multiply-add lane-ops/s (IMAD.WIDE)", "frac": ex2 / peak_t,
                       "window_bits": int(c2_c.value), "windows": int(c2_w.value), "lanes_per_point": lanes.value,
                       "ctas_per_sm": ctas.value, "executed_ops_per_addition": add_ops,
                       "executed_ops_per_point": int(c2_w.value) * add_ops, "launch_ms": per2, "launches": g2_cnt,
                       "traffic": (traffic.get("msm_accumulate_pair_kernel") or {}).get("dram_bytes") if lanes.value == 2 else None,
                       "traffic_detail": traffic.get("msm_accumulate_pair_kernel") if lanes.value == 2 else None,
                       "share_of_step": g2_ms / args.steps / dev_ms}
    ntt_ms, ntt_cnt = prof["ntt"]
    roofline_ntt = None
    if ntt_cnt:
        # The timed brackets are per ENQUEUE CALL: one whole transform (both passes) or a batch of three (the quotient's three
        # interpolations, then its three coset evaluations, are one batched launch per pass).  Transforms rank 0 runs per proof: all
        # 7 on one GPU; with the chains spread over the ranks, its own chains x 2 (rank 0 never forms H, zksnake_b200/dist.py).
        if world > 1 and prover._spread_chains():
            transforms = 2 * bin(zdist.chain_mask(rank, world)).count("1")
        else:
            transforms = 7
        per = ntt_ms / args.steps / max(transforms, 1)
        ach = 64.0 * n / (per * 1e-3) / 1e9
        roofline_ntt = {"kernel": "ntt_warp_pass_kernel / ntt_pass_kernel (all passes of one 2^%d transform)" % args.log_n, "bound": "hbm",
                        "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "peak_source": hbm_src,
                        "algorithmic_bytes": 64 * n, "launch_ms": per, "launches": ntt_cnt, "transforms_per_step": transforms,
                        "note": "launch_ms = the transforms' device time per proof / transforms_per_step; on one GPU the digit sorts of "
                                "the witness, U and V MSMs run beside the transforms on a side stream and lengthen them",
                        "traffic": (traffic.get("ntt_pass") or {}).get("dram_bytes"),
                        "traffic_detail": dict(traffic.get("ntt_pass") or {}, note="one PASS of the transform (a 2^20 transform is 2 passes)"),
                        "share_of_step": ntt_ms / args.steps / dev_ms}
    breakdown = {k: v[0] / args.steps for k, v in prof.items()}
    msm_ms_total = breakdown["msm_sort"] + breakdown["msm_accum_g1"] + breakdown["msm_accum_g2"] + breakdown["msm_reduce"]

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            _, cpu_base = run_cpu_arm(args, steps=1, warmup=0, budget_total_s=args.cpu_budget_s)
        except Exception as exc:  # the CPU checker is not the product: report, do not fail the bench
            cpu_base = {"unavailable": repr(exc)}

    g1b = nat.lib.zkb_affine_bytes(curve, 1)
    g2b = nat.lib.zkb_affine_bytes(curve, 2)
    line = {
        "metric": metric_name(args), "value": dev_ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall_ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32 limbs (int32 pipe)", "data": "synthetic",
        "config": workload_config(args),
        "parallelism": f"msm-{args.shard_mode}-shard{world}" + ("+quotient-chains-spread" if world > 1 and prover._spread_chains() else ""),
        "timing": "CUDA events on the library stream around the K steps (value); wall clock between barriers (ms_per_step, e2e)",
        "clocks": clocks,
        "e2e": {"value": e2e_ms, "unit": UNIT, "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                "api": "zksnake_b200.groth16.Groth16.prove_packed(pinned witness) -> Proof.to_bytes()"},
        "gpu_launches": int(launches),
        "roofline": roofline, "roofline_g2": roofline_g2, "roofline_ntt": roofline_ntt, "cpu_baseline": cpu_base,
        "parity": parity,
        "parity_check": None if parity is None else "proof bytes == closed-form exponents of the seeded toxic waste (oracle.groth16."
                                                    "prove_closed_form, Python ints, untimed region); identical proof_sha expected at every N",
        "e2e_list_api": None if list_api_ms is None else {
            "value": list_api_ms, "unit": UNIT,
            "api": "zksnake_b200.groth16.Groth16.prove(public: list[int], private: list[int]) -> Proof.to_bytes() (the reference's "
                   "signature, protocol.py:115-131): 2^%d Python ints marshalled by csrc/pymarshal.cpp every step" % args.log_n},
        "breakdown_ms": breakdown,
        "host_phases_ms": host_phases if world > 1 else None,
        "msm_mpts_s": (4 + 1) * pts_per_launch / (msm_ms_total * 1e-3) / 1e6 if msm_ms_total else None,
        "ntt_gelem_s": n / (roofline_ntt["launch_ms"] * 1e-3) / 1e9 if roofline_ntt else None,
        "proof_sha": __import__("hashlib").sha256(bytes.fromhex(proof_hex)).hexdigest()[:16],
        "setup_s": setup_s,
    }
    print(json.dumps(line), flush=True)
    if td is not None:
        td.destroy_process_group()
    if parity is False:
        print("bench.py: PARITY FAILURE -- the proof bytes differ from the closed form", file=sys.stderr)
        return 3
    return 0


# ------------------------------------------------------------------------------------------------------------------------
# own arm, PlonK (BASELINE.json configs[3])
# ------------------------------------------------------------------------------------------------------------------------
def main_plonk(args):
    """One step = one PlonK proof of the chain circuit as 2^log_n gates through DevicePlonk.prove_packed: the three wire columns
    (host memory) -> five rounds (11 transforms, 9 commitment MSMs over the SRS table, the prover glue as kernels) -> proof
    bytes.  The protocol's Fiat-Shamir challenges need each round's commitments on the host, so the proof has five host
    synchronisation points by construction; `value` (CUDA events around the K proofs) therefore includes them."""
    import hashlib
    import random

    from zksnake_b200 import _native as nat
    from zksnake_b200 import dist as zdist
    from zksnake_b200 import plonk as pm
    from zksnake_b200.plonk_device import DevicePlonk
    from zksnake_b200.plonkish import chain_gates

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not nat.gpu_available():
        raise SystemExit("bench.py: no CUDA device visible -- zksnake_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    td = torch = None
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(local_rank)
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nat.ensure_init(local_rank)
    curve = CURVE_IDS[args.curve]
    n = 1 << args.log_n

    def barrier():
        if td is not None:
            td.barrier()
            torch.cuda.synchronize()
        nat.check(nat.lib.zkb_sync())

    t_setup = time.perf_counter()
    cs, pub, priv = chain_gates(n, args.curve)
    p = cs.p
    rnd = random.Random(3)                              # blinding scalars from random.Random(3)
    tau = random.Random(1).randint(1, p - 1)
    blinders = [rnd.randint(1, p - 1) for _ in range(11)]
    hook = []
    pm.get_random_int = lambda n_max: hook.pop(0)       # seeded identically on every rank
    plonk = DevicePlonk(cs, args.curve, shard=(rank, world))
    hook[:] = [tau]
    plonk.setup()
    cols = []
    for k in range(3):                                  # wire columns in pinned host memory
        src = nat.ints_to_limbs(priv[k::3])
        pinned = ctypes.c_void_p()
        nat.check(nat.lib.zkb_host_alloc(src.nbytes, ctypes.byref(pinned)))
        arr = np.ctypeslib.as_array(ctypes.cast(pinned, ctypes.POINTER(ctypes.c_uint64)), shape=src.shape)
        arr[:] = src
        cols.append(arr)
    setup_s = time.perf_counter() - t_setup

    def prove():
        hook[:] = list(blinders)
        return plonk.prove_packed(pub, cols)

    plonk.keep_polys = True
    proof = prove()                      # first proof: builds the NTT tables, grows the scratch arena; parity is checked on it
    blob = proof.to_bytes()
    # ---- parity (untimed): verify() through the host pairing, and the first-round commitment [A(tau)]G1 against the closed form
    # with A downloaded from the device and Horner-evaluated in Python ints (rank 0)
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle.curve import group
        G1 = group(curve, False)
        coeffs = plonk.last_polys["a"].to_ints()
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * tau + c) % p
        want = G1.mul(G1.gen, acc)
        parity = bool(plonk.verify(proof, pub)) and (proof.tau_a.x, proof.tau_a.y) == want
    plonk.keep_polys = False
    plonk.last_polys = {}
    barrier()
    for _ in range(max(args.warmup, 1)):     # warm-up AFTER the parity check's CPU work: the timed region starts on a warm GPU
        assert prove().to_bytes() == blob

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    nat.check(nat.lib.zkb_prof_enable(1))
    barrier()
    launches0 = nat.lib.zkb_launch_count()
    h2d0, d2h0 = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    nat.lib.zkb_transfer_count(ctypes.byref(h2d0), ctypes.byref(d2h0))
    dist_t0 = dict(zdist.TRANSFER)
    rounds = []
    t0 = time.perf_counter()
    with nat.Timer() as tm:
        for _ in range(args.steps):
            ts = time.perf_counter()
            pr = prove()
            _ = pr.to_bytes()
            rounds.append(dict(plonk.timings, step_wall=(time.perf_counter() - ts) * 1e3))
    barrier()
    t1 = time.perf_counter()
    launches = nat.lib.zkb_launch_count() - launches0
    h2d1, d2h1 = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    nat.lib.zkb_transfer_count(ctypes.byref(h2d1), ctypes.byref(d2h1))
    h2d_step = (h2d1.value - h2d0.value + zdist.TRANSFER["h2d"] - dist_t0["h2d"]) // args.steps
    d2h_step = (d2h1.value - d2h0.value + zdist.TRANSFER["d2h"] - dist_t0["d2h"]) // args.steps
    dev_ms = tm.ms / args.steps
    wall_ms = (t1 - t0) * 1e3 / args.steps
    prof = nat.prof_read()
    nat.check(nat.lib.zkb_prof_enable(0))
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    if td is not None:
        t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        dev_ms, wall_ms = t.tolist()
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        td.all_reduce(lt)
        launches = int(lt.item())
        # where each rank's host spent the sharded proof: its own work up to the partial sums | waiting for the slowest rank | assembly
        ph = torch.tensor([host_phases.get(k, 0.0) for k in ("partial", "exchange", "assemble")], dtype=torch.float64, device="cuda")
        allph = torch.empty(world * 3, dtype=torch.float64, device="cuda")
        td.all_gather_into_tensor(allph, ph)
        allph = allph.cpu().reshape(world, 3).tolist()
        host_phases = {"partial": [round(r[0], 3) for r in allph], "exchange": [round(r[1], 3) for r in allph],
                       "assemble": [round(r[2], 3) for r in allph], "note": "per rank, mean ms per proof, device-resident leg"}
        if zdist.SPREAD_TRACE is not None:      # ZKB_SPREAD_TRACE=1: device timestamps of the exchange steps, per rank
            tk = ("chains", "uv", "recv", "quot", "h", "end")
            cnt = max(zdist.SPREAD_TRACE.get("proofs", 0), 1)
            tv = torch.tensor([zdist.SPREAD_TRACE.get(k, 0.0) / cnt for k in tk], dtype=torch.float64, device="cuda")
            allt = torch.empty(world * len(tk), dtype=torch.float64, device="cuda")
            td.all_gather_into_tensor(allt, tv)
            allt = allt.cpu().reshape(world, len(tk)).tolist()
            host_phases["device_trace_ms"] = {k: [round(r[i], 3) for r in allt] for i, k in enumerate(tk)}
            pk_ = sorted(prof)
            pv = torch.tensor([prof[k][0] / args.steps for k in pk_], dtype=torch.float64, device="cuda")
            allp = torch.empty(world * len(pk_), dtype=torch.float64, device="cuda")
            td.all_gather_into_tensor(allp, pv)
            allp = allp.cpu().reshape(world, len(pk_)).tolist()
            host_phases["breakdown_ms_per_rank"] = {k: [round(r[i], 3) for r in allp] for i, k in enumerate(pk_)}
    if rank != 0:
        if td is not None:
            td.destroy_process_group()
        return 0

    imad_wide = ctypes.c_double()
    nat.check(nat.lib.zkb_imad_peak(1, ctypes.byref(imad_wide)))
    peak_t = imad_wide.value / 1e12
    L = 8 if curve == 0 else 12
    mul_ops = 2 * L * L + L
    c_c, c_w, c_b = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_size_t()
    nat.check(nat.lib.zkb_msm_table_info(plonk.table, ctypes.byref(c_c), ctypes.byref(c_w), ctypes.byref(c_b)))
    acc_ms, acc_cnt = prof["msm_accum_g1"]
    roofline = None
    if acc_cnt:
        per = acc_ms / acc_cnt
        pts = (n + 6) / world                                   # average launch: 7 of the 9 MSMs are n + 5 / n + 6 points
        ex = pts * int(c_w.value) * 10 * mul_ops / (per * 1e-3) / 1e12
        roofline = {"kernel": "msm_accumulate_kernel<G1>", "bound": "int32-pipe", "achieved": ex, "peak": peak_t,
                    "unit": "T 32x32->64 multiply-add lane-ops/s (IMAD.WIDE)", "frac": ex / peak_t,
                    "note": "upper estimate: every launch counted as n + 6 points (T_hi has n + 6, the wire / Z / T commitments n + 2 "
                            "... n + 3, the two opening proofs n + 5)",
                    "window_bits": int(c_c.value), "windows": int(c_w.value), "launch_ms": per, "launches": acc_cnt,
                    "share_of_step": acc_ms / args.steps / dev_ms, "traffic": None}
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            _, cpu_base = run_cpu_arm(args, steps=1, warmup=0, budget_total_s=args.cpu_budget_s)
        except Exception as exc:
            cpu_base = {"unavailable": repr(exc)}
    line = {
        "metric": metric_name(args), "value": dev_ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall_ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32 limbs (int32 pipe)", "data": "synthetic", "config": workload_config(args),
        "parallelism": f"msm-windows-shard{world}",
        "timing": "CUDA events on the library stream around the K proofs (value); wall clock between barriers (ms_per_step = e2e: "
                  "the wire columns start in pinned host memory every step)",
        "clocks": clocks,
        "e2e": {"value": wall_ms, "unit": UNIT, "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                "api": "zksnake_b200.plonk_device.DevicePlonk.prove_packed(public dict, 3 pinned wire columns) -> Proof.to_bytes()"},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_base,
        "rounds_ms": {k: float(np.median([r[k] for r in rounds])) for k in rounds[0]},
        "breakdown_ms": {k: v[0] / args.steps for k, v in prof.items()},
        "parity": parity,
        "parity_check": None if parity is None else "verify() through the host pairing and [A(tau)]G1 == tau_a with A downloaded and "
                                                    "Horner-evaluated in Python ints (untimed region)",
        "proof_sha": hashlib.sha256(blob).hexdigest()[:16], "setup_s": setup_s,
    }
    print(json.dumps(line), flush=True)
    if td is not None:
        td.destroy_process_group()
    if parity is False:
        print("bench.py: PARITY FAILURE (plonk)", file=sys.stderr)
        return 3
    return 0


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        sys.exit(main_reference(a))
    sys.exit(main_plonk(a) if a.workload == "plonk" else main_own(a))
