"""Drop-in for the reference's pyo3 extension module `zksnake._algebra` (/root/reference/src/lib.rs:178-185): submodules
ec_bn254, ec_bls12_381, polynomial_bn254, polynomial_bls12_381 (the proving hot path, over libzkb200.so on the GPU) and
`circuit` (a host-side stand-in for the symbolic front end, so that the reference's Python layer imports and runs unmodified:
`zksnake_b200.dropin.install()`).  MultilinearPolynomial (sumcheck / GKR) is out of scope."""
from . import circuit, ec_bls12_381, ec_bn254, polynomial_bls12_381, polynomial_bn254  # noqa: F401
