"""Raw MSM at 1 / 2 / 4 / 8 GPUs (BASELINE.json configs[4]): every rank holds the points and the scalars and runs the scalar
windows [W*rank/world, W*(rank+1)/world) of the MSM (zkb_msm_dev_windows: digit sort, accumulation, bucket reduction and the
host Horner tail all shrink by 1/world); the partial points are all-gathered (< 200 B per rank) and added with the exact group
law on the host.  Timing: barrier, wall clock around [windows MSM + all-gather + add] on every rank (the MSM call returns only
after its device work has completed), MAX over ranks, median of `steps`.  One JSON line per (curve, group, log_n) on rank 0.

  python tools/msm_scale.py                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/msm_scale.py"""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zksnake_b200 import _native as nat  # noqa: E402
from zksnake_b200 import dist  # noqa: E402
from zksnake_b200.ecc import EllipticCurve  # noqa: E402
from zksnake_b200.frvec import FrVec  # noqa: E402

CASES = [("BN254", 1, 20), ("BN254", 1, 22), ("BN254", 1, 24), ("BN254", 2, 20), ("BLS12_381", 1, 20), ("BLS12_381", 1, 22),
         ("BLS12_381", 2, 20)]


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    td = None
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        td.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    nat.ensure_init(int(os.environ.get("LOCAL_RANK", "0")))
    for curve, grp, log_n in CASES:
        E = EllipticCurve(curve)
        cid = E.curve.CURVE_ID
        n = 1 << log_n
        r = E.order
        # bases k_i * G with k_i = 7 * 3^i (fixed-base kernel); scalars: a fixed geometric sequence spread over the whole field
        gen = E.curve.upload_points([E.G1() if grp == 1 else E.G2()], grp)
        ks = FrVec.powers(cid, n, 3, 7)
        pts = E.curve.PointVector(cid, grp, n)
        nat.check(nat.lib.zkb_batch_mul_dev(cid, grp, gen.ptr, 1, ks.ptr, n, pts.ptr))
        scal = FrVec.powers(cid, n, 0x1234567890ABCDEF1234567890ABCDEF1234567 % r, 0xFEDCBA9876543210FEDCBA987654321 % r)
        nat.check(nat.lib.zkb_sync())
        limbs = nat.lib.zkb_affine_bytes(cid, grp) // 8

        def once():
            xy = np.zeros(limbs, dtype=np.uint64)
            inf = ctypes.c_int()
            nat.check(nat.lib.zkb_msm_dev_windows(cid, grp, pts.ptr, scal.ptr, n, rank, world, nat.ptr(xy), ctypes.byref(inf)))
            if world == 1:
                return xy, inf.value
            row = np.zeros((1, limbs + 1), dtype=np.uint64)
            row[0, :limbs] = xy
            row[0, limbs] = inf.value
            parts = dist.all_gather_array(row)          # (world, 1, limbs + 1)
            p = np.ascontiguousarray(parts[:, 0, :limbs])
            infs = np.ascontiguousarray(parts[:, 0, limbs].astype(np.int32))
            sc = np.zeros((world, 4), dtype=np.uint64)
            has = np.zeros(world, dtype=np.int32)
            acc = np.zeros(limbs, dtype=np.uint64)
            ainf = ctypes.c_int()
            nat.check(nat.lib.zkb_point_lincomb(cid, grp, world, nat.ptr(p), nat.ptr(infs), nat.ptr(sc), nat.ptr(has), nat.ptr(acc),
                                                ctypes.byref(ainf)))
            return acc, ainf.value

        for _ in range(2):
            res = once()
        times = []
        for _ in range(steps):
            if td is not None:
                td.barrier()
            t0 = time.perf_counter()
            res = once()
            dt = (time.perf_counter() - t0) * 1e3
            if td is not None:
                import torch
                t = torch.tensor([dt], device="cuda")
                td.all_reduce(t, op=td.ReduceOp.MAX)
                dt = float(t.item())
            times.append(dt)
        if rank == 0:
            ms = float(np.median(times))
            digest = __import__("hashlib").sha256(res[0].tobytes() + bytes([res[1]])).hexdigest()[:16]
            print(json.dumps({"tool": "msm_scale", "curve": curve, "group": grp, "log_n": log_n, "n_gpus": world, "ms": ms,
                              "mpts_s": n / ms / 1e3, "result_sha": digest, "steps": steps}), flush=True)
        del pts, ks, scal
    if td is not None:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
