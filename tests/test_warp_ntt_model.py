"""CPU model check of the register-resident NTT pass (csrc/ntt_warp.cuh): the lane-bit / register-slot swap schedule, the twiddle
indices and the output-row formula, restated in plain Python over a small prime field (tools/warp_ntt_model.py), must compute the
DFT of a column for every (elements per lane, column size) the kernel template accepts.  The kernel itself is compared bit-exactly
with the shared-memory kernel on the GPU (tools/ntt_probe.py, tests/test_gpu_parity.py through the Groth16 quotient)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lane_slot_bookkeeping_computes_the_dft():
    spec = importlib.util.spec_from_file_location("warp_ntt_model", os.path.join(ROOT, "tools", "warp_ntt_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for el in (1, 2, 3):
        for k in range(el + 1, el + 6):
            if k <= 8:
                mod.check(el, k, seed=31 * k + el)


def test_final_permutation_is_a_bijection():
    spec = importlib.util.spec_from_file_location("warp_ntt_model", os.path.join(ROOT, "tools", "warp_ntt_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for el, k in ((2, 7), (2, 6), (2, 3), (3, 8)):
        rows = sorted(mod.out_row(el, k, lane, slot) for lane in range(1 << (k - el)) for slot in range(1 << el))
        assert rows == list(range(1 << k))
