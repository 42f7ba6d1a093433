"""Quick device-side timing probe (not the bench contract): NTT and MSM kernels with device-resident data."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zksnake_b200 import _native as nat  # noqa: E402


TOP_LIMB = {0: 0x30644e72e131a029, 1: 0x73eda753299d7d48}


def rand_fr(n, seed, curve=0):
    """uniform 256-bit values below (top limb of r) * 2^192: the digit statistics of uniform scalars mod r"""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 2 ** 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] %= np.uint64(TOP_LIMB[curve])
    return a


def time_ntt(curve, log_n, reps=10):
    n = 1 << log_n
    d_in = nat.DeviceBuffer(n * 32).upload(rand_fr(n, log_n))
    d_out = nat.DeviceBuffer(n * 32)
    for _ in range(3):
        nat.check(nat.lib.zkb_ntt_dev(curve, 0, 0, log_n, d_in.ptr, n, d_out.ptr))
    with nat.Timer() as t:
        for _ in range(reps):
            nat.check(nat.lib.zkb_ntt_dev(curve, 0, 0, log_n, d_in.ptr, n, d_out.ptr))
    ms = t.ms / reps
    print(f"ntt curve={curve} 2^{log_n}: {ms*1e3:9.1f} us  {64*n/ms/1e6:8.1f} GB/s(alg)  {n/ms/1e6:8.2f} Gelem/s", flush=True)
    d_in.free(); d_out.free()
    return ms


def make_points(curve, grp, n, seed):
    ab = nat.lib.zkb_affine_bytes(curve, grp)
    gens = {
        (0, 1): [1, 2],
        (1, 1): [0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
                 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1],
        (0, 2): [10857046999023057135944570762232829481370756359578518086990519993285655852781,
                 11559732032986387107991004021392285783925812861821192530917403151452391805634,
                 8495653923123431417604973247489272438418190587263600148770280649306958101930,
                 4082367875863433681332203403145435568316851327593401208105741076214120093531],
        (1, 2): [0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
                 0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E,
                 0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
                 0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE],
    }[(curve, grp)]
    nb = 32 if curve == 0 else 48
    gen = np.frombuffer(b"".join(c.to_bytes(nb, "little") for c in gens), dtype=np.uint64).copy()
    d_gen = nat.DeviceBuffer(ab)
    nat.check(nat.lib.zkb_points_upload(curve, grp, nat.ptr(gen), 1, d_gen.ptr))
    d_k = nat.DeviceBuffer(n * 32).upload(rand_fr(n, seed, curve))
    d_pts = nat.DeviceBuffer(n * ab)
    t0 = time.time()
    nat.check(nat.lib.zkb_batch_mul_dev(curve, grp, d_gen.ptr, 1, d_k.ptr, n, d_pts.ptr))
    nat.check(nat.lib.zkb_sync())
    print(f"  batch_mul curve={curve} g{grp} n={n}: {time.time()-t0:.3f} s", flush=True)
    d_k.free(); d_gen.free()
    return d_pts


def time_msm(curve, grp, log_n, reps=3, tunings=((0, 0, 0),), scalars=None, label=""):
    n = 1 << log_n
    ab = nat.lib.zkb_affine_bytes(curve, grp)
    d_pts = make_points(curve, grp, n, 1)
    d_s = nat.DeviceBuffer(n * 32).upload(scalars if scalars is not None else rand_fr(n, 2, curve))
    out = np.zeros(ab // 8, dtype=np.uint64)
    inf = ctypes.c_int()
    for tun in tunings:
        nat.lib.zkb_msm_set_tuning(*tun)
        nat.check(nat.lib.zkb_msm_dev(curve, grp, d_pts.ptr, d_s.ptr, n, nat.ptr(out), ctypes.byref(inf)))
        t0 = time.perf_counter()
        with nat.Timer() as t:
            for _ in range(reps):
                nat.check(nat.lib.zkb_msm_dev(curve, grp, d_pts.ptr, d_s.ptr, n, nat.ptr(out), ctypes.byref(inf)))
        wall = (time.perf_counter() - t0) / reps * 1e3
        ms = t.ms / reps
        nat.check(nat.lib.zkb_prof_enable(1))
        nat.check(nat.lib.zkb_msm_dev(curve, grp, d_pts.ptr, d_s.ptr, n, nat.ptr(out), ctypes.byref(inf)))
        parts = []
        for tag, name in ((1, "sort"), (2, "accG1"), (3, "accG2"), (4, "reduce")):
            tms = ctypes.c_float()
            cnt = ctypes.c_ulonglong()
            nat.check(nat.lib.zkb_prof_read(tag, ctypes.byref(tms), ctypes.byref(cnt)))
            if cnt.value:
                parts.append(f"{name}={tms.value:.3f}")
        nat.check(nat.lib.zkb_prof_enable(0))
        print(f"msm{label} curve={curve} g{grp} 2^{log_n} tuning={tun}: {ms:8.3f} ms (wall {wall:8.3f})  {n/ms/1e3:8.2f} Mpts/s  [{' '.join(parts)}]", flush=True)
    nat.lib.zkb_msm_set_tuning(0, 0, 0)
    d_pts.free(); d_s.free()


if __name__ == "__main__":
    nat.ensure_init()
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("ntt", "all"):
        for maxk in (11, 10, 8):
            os.environ["ZKB_NTT_MAXK"] = str(maxk)
            print("ZKB_NTT_MAXK", maxk)
            for ln in (16, 20, 21, 22, 23, 24):
                time_ntt(0, ln)
        os.environ["ZKB_NTT_MAXK"] = "10"
        time_ntt(1, 20)
    if what == "table":
        for curve, grp, ln in ((0, 1, 20), (0, 2, 20), (1, 1, 20)):
            n = 1 << ln
            ab = nat.lib.zkb_affine_bytes(curve, grp)
            d_pts = make_points(curve, grp, n, 1)
            d_s = nat.DeviceBuffer(n * 32).upload(rand_fr(n, 2, curve))
            out = np.zeros(ab // 8, dtype=np.uint64)
            inf = ctypes.c_int()
            for cb in (0, 16, 17, 18, 19, 20):
                tab = ctypes.c_void_p()
                t0 = time.time()
                nat.check(nat.lib.zkb_msm_table_create(curve, grp, d_pts.ptr, n, cb, 1, ctypes.byref(tab)))
                nat.check(nat.lib.zkb_sync())
                tb = time.time() - t0
                c_, W_, by_ = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_size_t()
                nat.lib.zkb_msm_table_info(tab, ctypes.byref(c_), ctypes.byref(W_), ctypes.byref(by_))
                nat.check(nat.lib.zkb_msm_table_dev(tab, d_s.ptr, n, 0, 1, nat.ptr(out), ctypes.byref(inf)))
                with nat.Timer() as t:
                    for _ in range(3):
                        nat.check(nat.lib.zkb_msm_table_dev(tab, d_s.ptr, n, 0, 1, nat.ptr(out), ctypes.byref(inf)))
                nat.check(nat.lib.zkb_prof_enable(1))
                nat.check(nat.lib.zkb_msm_table_dev(tab, d_s.ptr, n, 0, 1, nat.ptr(out), ctypes.byref(inf)))
                pr = nat.prof_read()
                nat.check(nat.lib.zkb_prof_enable(0))
                print(f"table msm curve={curve} g{grp} 2^{ln} c={c_.value} W={W_.value} table={by_.value/2**20:.0f} MiB build={tb:.2f}s: "
                      f"{t.ms/3:8.3f} ms {n/(t.ms/3)/1e3:8.1f} Mpts/s  " + " ".join(f"{k}={v[0]:.3f}" for k, v in pr.items() if v[1]), flush=True)
                nat.lib.zkb_msm_table_free(tab)
            d_pts.free(); d_s.free()
    if what == "small":
        # latency-bound sizes: window / run-length choices for plain and table MSMs at 2^14 .. 2^18
        for ln in (14, 16, 18):
            time_msm(0, 1, ln, tunings=[(0, 0, 0)] + [(c, seg, 3) for c in (ln - 6, ln - 5, ln - 4, ln - 3) for seg in (8, 16, 32)])
        for curve, grp, ln in ((0, 1, 16), (0, 1, 18), (0, 2, 16)):
            n = 1 << ln
            ab = nat.lib.zkb_affine_bytes(curve, grp)
            d_pts = make_points(curve, grp, n, 1)
            d_s = nat.DeviceBuffer(n * 32).upload(rand_fr(n, 2, curve))
            out = np.zeros(ab // 8, dtype=np.uint64)
            inf = ctypes.c_int()
            for cb in (0, ln - 5, ln - 4, ln - 3, ln - 2, ln - 1):
                tab = ctypes.c_void_p()
                nat.check(nat.lib.zkb_msm_table_create(curve, grp, d_pts.ptr, n, cb, 1, ctypes.byref(tab)))
                c_, W_, by_ = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_size_t()
                nat.lib.zkb_msm_table_info(tab, ctypes.byref(c_), ctypes.byref(W_), ctypes.byref(by_))
                for seg in (0, 8, 16, 32, 64):
                    nat.lib.zkb_msm_set_tuning(0, seg, 0)
                    nat.check(nat.lib.zkb_msm_table_dev(tab, d_s.ptr, n, 0, 1, nat.ptr(out), ctypes.byref(inf)))
                    with nat.Timer() as t:
                        for _ in range(5):
                            nat.check(nat.lib.zkb_msm_table_dev(tab, d_s.ptr, n, 0, 1, nat.ptr(out), ctypes.byref(inf)))
                    nat.check(nat.lib.zkb_prof_enable(1))
                    nat.check(nat.lib.zkb_msm_table_dev(tab, d_s.ptr, n, 0, 1, nat.ptr(out), ctypes.byref(inf)))
                    pr = nat.prof_read()
                    nat.check(nat.lib.zkb_prof_enable(0))
                    print(f"table msm curve={curve} g{grp} 2^{ln} c={c_.value}{'(auto)' if cb == 0 else ''} W={W_.value} seg={seg}: "
                          f"{t.ms/5:8.3f} ms  " + " ".join(f"{k}={v[0]:.3f}" for k, v in pr.items() if v[1]), flush=True)
                nat.lib.zkb_msm_set_tuning(0, 0, 0)
                nat.lib.zkb_msm_table_free(tab)
            d_pts.free(); d_s.free()
    if what == "msmx":
        time_msm(1, 1, 20)
        time_msm(1, 1, 22, tunings=((0, 0, 0), (15, 32, 3), (16, 32, 3), (17, 32, 3)))
        time_msm(1, 2, 20)
        time_msm(1, 2, 22, reps=2)
        time_msm(0, 1, 24, reps=2, tunings=((0, 0, 0), (18, 32, 3), (19, 32, 3)))
        time_msm(0, 1, 26, reps=2, tunings=((0, 0, 0), (20, 32, 3), (19, 32, 3)))
    if what == "msm1":
        time_msm(0, 1, 20, reps=1)
    if what in ("msm", "all"):
        time_msm(0, 1, 20, tunings=((0, 0, 0), (16, 32, 2), (15, 32, 3), (17, 32, 3), (18, 32, 3), (16, 64, 3), (16, 24, 3)))
        n = 1 << 20
        ones = np.zeros((n, 4), dtype=np.uint64); ones[:, 0] = 1
        time_msm(0, 1, 20, scalars=ones, label="[all ones]")
        bits = np.zeros((n, 4), dtype=np.uint64); bits[:, 0] = np.arange(n) % 3
        time_msm(0, 1, 20, scalars=bits, label="[0/1/2]")
        time_msm(0, 1, 16)
        time_msm(0, 1, 22)
        time_msm(0, 1, 24, reps=2)
        time_msm(0, 2, 18)
        time_msm(1, 1, 20)
        time_msm(1, 2, 18)
