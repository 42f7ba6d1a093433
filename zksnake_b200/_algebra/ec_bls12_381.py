"""Mirror of the reference's `zksnake._algebra.ec_bls12_381` submodule (/root/reference/src/lib.rs) over libzkb200.so."""
from ._ec import build as _build

globals().update(_build(1))
