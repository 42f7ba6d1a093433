"""PlonK prove timing (BASELINE.json configs[3]: 2^20 gates, BN254; the chain circuit of benchmarks/benchmark_plonk.py:12-25
synthesised as gates) with the device-resident prover.  Prints one JSON line: total prove ms (wall clock around prove_packed,
host-resident witness columns in, proof bytes out), per-round ms, and the device time per kernel family.

  python tools/plonk_bench.py [log_n=20] [curve=BN254] [steps=3]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zksnake_b200 import _native as nat  # noqa: E402
from zksnake_b200 import plonk as pm  # noqa: E402
from zksnake_b200.plonk_device import DevicePlonk  # noqa: E402
from zksnake_b200.plonkish import chain_gates  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    curve = sys.argv[2] if len(sys.argv) > 2 else "BN254"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    n = 1 << log_n
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    td = None
    if world > 1:   # torchrun: one process per GPU; the commitment MSMs are window-sharded, everything else is replicated
        import torch
        import torch.distributed as td
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        td.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    nat.ensure_init(int(os.environ.get("LOCAL_RANK", "0")))
    t0 = time.perf_counter()
    cs, pub, priv = chain_gates(n, curve)
    t_circuit = time.perf_counter() - t0
    import random
    rnd = random.Random(3)
    pm.get_random_int = lambda n_max: rnd.randint(1, n_max)
    plonk = DevicePlonk(cs, curve, shard=(rank, world))
    t0 = time.perf_counter()
    plonk.setup()
    t_setup = time.perf_counter() - t0
    # wire columns in pinned host memory (the e2e path copies them to the device inside the timed region)
    import ctypes
    cols = []
    for k in range(3):
        src = nat.ints_to_limbs(priv[k::3])
        pinned = ctypes.c_void_p()
        nat.check(nat.lib.zkb_host_alloc(src.nbytes, ctypes.byref(pinned)))
        arr = np.ctypeslib.as_array(ctypes.cast(pinned, ctypes.POINTER(ctypes.c_uint64)), shape=src.shape)
        arr[:] = src
        cols.append(arr)
    proof = plonk.prove_packed(pub, cols)          # warm-up (builds NTT tables, grows the scratch arena)
    ok = plonk.verify(proof, pub)
    times, rounds = [], []
    nat.check(nat.lib.zkb_prof_enable(1))
    l0 = nat.lib.zkb_launch_count()
    for _ in range(steps):
        if td is not None:
            td.barrier()
        t0 = time.perf_counter()
        proof = plonk.prove_packed(pub, cols)
        blob = proof.to_bytes()
        times.append((time.perf_counter() - t0) * 1e3)
        rounds.append(dict(plonk.timings))
    launches = (nat.lib.zkb_launch_count() - l0) // steps
    prof = nat.prof_read()
    nat.check(nat.lib.zkb_prof_enable(0))
    line = {
        "metric": "plonk_prove_ms", "value": float(np.median(times)), "unit": "ms", "n_gpus": 1, "steps": steps,
        "config": {"workload": f"plonk-prove chain circuit (benchmarks/benchmark_plonk.py) 2^{log_n} gates {curve}",
                   "msm": "9 x G1 over the SRS table", "ntt": "5 x n, 6 x 4n (coset quotient)"},
        "rounds_ms": {k: float(np.median([r[k] for r in rounds])) for k in rounds[0]},
        "device_ms_by_family": {k: v[0] / steps for k, v in prof.items() if v[1]},
        "family_launch_groups": {k: v[1] // steps for k, v in prof.items() if v[1]},
        "gpu_launches": int(launches), "verify": bool(ok), "proof_bytes": len(blob),
        "setup_s": t_setup, "circuit_s": t_circuit,
    }
    line["n_gpus"] = world
    line["proof_sha"] = __import__("hashlib").sha256(blob).hexdigest()[:16]
    if td is not None:
        import torch
        t = torch.tensor([line["value"]], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        line["value"] = float(t.item())
        td.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
