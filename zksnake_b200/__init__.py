"""zksnake_b200 -- B200 (sm_100a) back end for zksnake's proving hot path (Fr NTT + G1/G2 MSM -> Groth16 / KZG).

Layout: csrc/ (CUDA kernels + the C ABI of include/zkb200.h), _native.py (ctypes binding), _algebra/ (mirror of the
reference's `zksnake._algebra` extension-module surface), groth16.py (Groth16 setup / prove / verify on top of it).
"""
__all__ = ["_native"]
