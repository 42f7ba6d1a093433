"""CPU oracle for the zksnake proving hot path -- TEST INFRASTRUCTURE ONLY.

Pure-Python big-int restatement of what the reference computes through arkworks
(ark-poly 0.4.2 / ark-ec 0.4.2 / ark-ff 0.4.2 / ark-bn254 0.4.0 / ark-bls12-381 0.4.0 /
ark-serialize 0.4.2, pinned in /root/reference/Cargo.lock:18-140; none of them vendored).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (zksnake_b200/) never does.

PARITY STATUS: "parity unpinned" at the numeric level -- the reference ships no golden vector
for NTT output, MSM results, point encodings or proof bytes and cannot be
built here (no Rust toolchain).  What *is* pinned (tests/test_oracle_pins.py):
  * the polynomial KATs of /root/reference/tests/test_algebra.py:6-26,
  * the published constants (roots of unity, Montgomery constants, curve generators and the
    standard Zcash/IETF compressed encodings of the BLS12-381 generators),
  * algebraic identities (NTT by definition, H*Z == U*V-W, MSM == discrete-log closed form,
    Groth16 verification equation through an independent pairing implementation),
  * ONE outside anchor the reference's own tests do hold: the five Poseidon hashes over BN254's
    Fr of /root/reference/tests/test_gadgets.py:19-50 (Hades reference code, circomlib) --
    oracle/poseidon.py, tests/test_poseidon_kat.py.  They pin Fr addition and multiplication
    (ints, the host build of csrc/ff.cuh, the device) to implementations independent of this
    repository; NTT, MSM and encodings stay unpinned in the contract's sense.  The reference's
    fixture tests/stub/test_poseidon.r1cs (circom-compiled Poseidon(3)) carries the same anchor to
    the .r1cs reader and the row semantics (oracle/circom.py).
"""
