"""Multi-GPU split of the proving path: one process per GPU, contiguous slices of every MSM's point and
scalar range per rank, and ONE small collective per proof that exchanges the per-rank partial sums.

The reference has no multi-device path; its MSM (/root/reference/src/bn254/curve.rs:356-392) is a sum of independent terms, so
the slice sums add up to the same group element and the proof bytes do not depend on the world size.

`torch.distributed` is only the plumbing (NCCL on GPUs, gloo in the CPU tests): the payload is 5 affine points + 5 flags per
rank (< 1 KiB).
"""
import ctypes

import numpy as np

from . import _native as nat

MSM_SLOTS = 5            # A, B1, B2 (G2), HZ, KW  -- protocol.py:133-155
SLOT_LIMBS = 24          # uint64 per slot in the C ABI (room for a BLS12-381 G2 affine point)
SLOT_GROUP = (1, 1, 2, 1, 1)


def shard_range(total, rank, world):
    """Contiguous slice [lo, hi) of `total` items owned by `rank` (sizes differ by at most one item)."""
    lo = total * rank // world
    hi = total * (rank + 1) // world
    return lo, hi


def world():
    """(rank, world_size) of the current torch.distributed job, (0, 1) when not initialised."""
    try:
        import torch.distributed as td
    except ImportError:
        return 0, 1
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


# host<->device bytes moved by this module through torch (the library counts its own copies: zkb_transfer_count)
TRANSFER = {"h2d": 0, "d2h": 0}


def nccl_ready():
    try:
        import torch.distributed as td
        return td.is_available() and td.is_initialized() and td.get_backend() == "nccl"
    except Exception:
        return False


def upload_sharded(arr):
    """A host array that every rank holds (the witness) -> a device copy on every rank, moving only 1/world of it over each
    rank's PCIe link: every rank uploads its own slice and one NCCL all-gather over NVLink completes the vector.  With N
    ranks uploading the whole 32 MiB witness at once the host side is the bottleneck (measured: +1.0 ms at 4 GPUs against
    +0.4 ms at 1).  Returns a torch uint8 CUDA tensor (keep it alive while the library reads `data_ptr()`)."""
    import torch
    import torch.distributed as td
    rank, ws = world()
    flat = np.ascontiguousarray(arr).reshape(-1).view(np.uint8)
    nbytes = flat.size
    per = -(-nbytes // (ws * 256)) * 256                      # slice size, 256-byte aligned
    full = torch.empty(per * ws, dtype=torch.uint8, device="cuda")
    mine = torch.empty(per, dtype=torch.uint8, device="cuda")
    lo = min(rank * per, nbytes)
    hi = min(lo + per, nbytes)
    if hi > lo:
        mine[:hi - lo].copy_(torch.from_numpy(flat[lo:hi]), non_blocking=True)
        TRANSFER["h2d"] += hi - lo
    td.all_gather_into_tensor(full, mine)
    torch.cuda.current_stream().synchronize()                 # the library reads it on its own stream
    return full


def all_gather_partials(msm_xy, msm_inf, device=None):
    """Exchange the per-rank partial MSM results.  msm_xy: (5, 24) uint64, msm_inf: (5,) int32.
    Returns (world, 5, 24) uint64 and (world, 5) int32 arrays, identical on every rank."""
    import torch
    import torch.distributed as td
    rank, ws = world()
    if ws == 1:
        return msm_xy[None].copy(), msm_inf[None].copy()
    payload = np.zeros(MSM_SLOTS * SLOT_LIMBS + MSM_SLOTS, dtype=np.int64)
    payload[:MSM_SLOTS * SLOT_LIMBS] = msm_xy.reshape(-1).view(np.int64)
    payload[MSM_SLOTS * SLOT_LIMBS:] = msm_inf
    t = torch.from_numpy(payload)
    on_gpu = td.get_backend() == "nccl"
    if on_gpu:
        t = t.cuda(device) if device is not None else t.cuda()
        TRANSFER["h2d"] += payload.nbytes
        TRANSFER["d2h"] += ws * payload.nbytes
    out = torch.empty(ws * t.numel(), dtype=torch.int64, device=t.device)
    td.all_gather_into_tensor(out, t)
    arr = out.cpu().numpy().reshape(ws, -1)
    xy = arr[:, :MSM_SLOTS * SLOT_LIMBS].copy().view(np.uint64).reshape(ws, MSM_SLOTS, SLOT_LIMBS)
    inf = arr[:, MSM_SLOTS * SLOT_LIMBS:].astype(np.int32)
    return xy, inf


def add_partials(curve, all_xy, all_inf):
    """Sum the ranks' partial points slot by slot on the host (zkb_point_lincomb: exact group law, no GPU needed).
    Returns (5, 24) uint64 and (5,) int32."""
    ws = all_xy.shape[0]
    out_xy = np.zeros((MSM_SLOTS, SLOT_LIMBS), dtype=np.uint64)
    out_inf = np.zeros(MSM_SLOTS, dtype=np.int32)
    for slot, grp in enumerate(SLOT_GROUP):
        limbs = nat.lib.zkb_affine_bytes(curve, grp) // 8
        pts = np.ascontiguousarray(all_xy[:, slot, :limbs])
        infs = np.ascontiguousarray(all_inf[:, slot].astype(np.int32))
        scal = np.zeros((ws, 4), dtype=np.uint64)
        has = np.zeros(ws, dtype=np.int32)
        res = np.zeros(limbs, dtype=np.uint64)
        inf = ctypes.c_int()
        nat.check(nat.lib.zkb_point_lincomb(curve, grp, ws, nat.ptr(pts), nat.ptr(infs), nat.ptr(scal), nat.ptr(has),
                                            nat.ptr(res), ctypes.byref(inf)))
        out_xy[slot, :limbs] = res
        out_inf[slot] = inf.value
    return out_xy, out_inf


def all_gather_array(arr, device=None):
    """all-gather of one small numpy array per rank (same shape and dtype everywhere) -> array of shape (world,) + arr.shape.
    NCCL on GPUs (the payload is staged through a CUDA tensor), gloo on CPU; identity when not distributed."""
    import torch
    import torch.distributed as td
    rank, ws = world()
    if ws == 1:
        return arr[None].copy()
    flat = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    t = torch.from_numpy(flat.copy())
    if td.get_backend() == "nccl":
        t = t.cuda(device) if device is not None else t.cuda()
    out = torch.empty(ws * t.numel(), dtype=torch.uint8, device=t.device)
    td.all_gather_into_tensor(out, t)
    return out.cpu().numpy().view(arr.dtype).reshape((ws,) + arr.shape)


def require_world(rank, world_size, emulate=False):
    """An explicit shard=(rank, world) must BE the torch.distributed world: the cooperative provers draw shared randomness and
    exchange partial sums with collectives, and without a process group of that size every rank would silently produce a proof
    from its own slice and its own randomness.  `emulate=True` (tests that drive zkb_groth16_partial / assemble rank by rank in one
    process) skips the check; the collectives below then refuse to run."""
    if world_size == 1 or emulate:
        return
    if world() != (rank, world_size):
        raise RuntimeError(f"shard=({rank}, {world_size}) but the torch.distributed world is {world()}: initialise a process "
                           "group of that size (torchrun) before constructing a sharded prover")


def shared_draws(draw, n_max, count, world_size=None):
    """`count` values of the randomness hook `draw(n_max)` (uniform in [1, n_max]) that every rank agrees on: rank 0 draws
    them, one small all-gather hands them to everybody.  Toxic waste, prover randomness and blinding scalars must be identical
    on every rank, otherwise the partial sums belong to different keys / proofs.  Because the values ARE rank 0's draws, a
    seeded hook gives the same key and the same proof bytes at every world size (bench.py's `proof_sha`)."""
    rank, ws = world()
    if world_size is not None and world_size > 1 and ws != world_size:
        raise RuntimeError(f"shared randomness for {world_size} ranks needs a torch.distributed world of that size (have {ws})")
    mine = [int(draw(n_max)) for _ in range(count)] if (rank == 0 or ws == 1) else [0] * count
    if ws == 1:
        return mine
    nbytes = (max(int(n_max).bit_length(), 1) + 7) // 8
    raw = np.frombuffer(b"".join(v.to_bytes(nbytes, "little") for v in mine), dtype=np.uint8).copy()
    got = bytes(all_gather_array(raw)[0])
    return [int.from_bytes(got[i * nbytes:(i + 1) * nbytes], "little") for i in range(count)]


def shared_random(draw):
    """Function form of shared_draws for call sites that draw one value at a time (one collective per draw)."""
    return lambda n_max: shared_draws(draw, n_max, 1)[0]
