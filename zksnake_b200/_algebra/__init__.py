"""Drop-in for the reference's pyo3 extension module `zksnake._algebra` (/root/reference/src/lib.rs:178-185), restricted to
the proving hot path: submodules ec_bn254, ec_bls12_381, polynomial_bn254, polynomial_bls12_381.  The `circuit` submodule
(symbolic front end) and MultilinearPolynomial are out of scope."""
from . import ec_bls12_381, ec_bn254, polynomial_bls12_381, polynomial_bn254  # noqa: F401
