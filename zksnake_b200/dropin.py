"""Plug `zksnake_b200._algebra` in where the reference loads its pyo3 extension module `zksnake._algebra`
(/root/reference/src/lib.rs:178-185; imported at python/zksnake/polynomial.py:4-7, ecc.py:3, arithmetization/__init__.py:4,
arithmetization/r1cs.py:4, plonkish.py:4, parser.py:6).  After `install()` the reference's unmodified Python layer --
`zksnake.groth16.Groth16`, `zksnake.plonk.Plonk`, `zksnake.commitment.polynomial.KZG`, ... -- runs every field / curve
operation through libzkb200.so on the GPU.

    import zksnake_b200.dropin; zksnake_b200.dropin.install()      # before the first `import zksnake...`
    from zksnake.groth16 import Groth16

It is also a pytest plugin: `python -m pytest -p zksnake_b200.dropin <reference>/tests/test_groth16.py` runs the reference's own
test files over the mirror (tests/test_gpu_dropin.py does exactly that).  The reference package itself is found on sys.path:
`reference_python_dir()` names where this repository looks for it.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODULE = "zksnake._algebra"


def reference_python_dir():
    """Directory holding the reference's pure-Python package `zksnake/`: $ZKSNAKE_REF, else /root/reference/python (the build
    container), else baseline/_ref (git-ignored copy made by __graft_entry__.build(), which travels to the GPU box)."""
    for cand in (os.environ.get("ZKSNAKE_REF"), "/root/reference/python", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "zksnake", "__init__.py")):
            return cand
    return None


def install(add_reference_path=True):
    """Register the mirror as `zksnake._algebra` (idempotent).  Returns the directory the reference package is taken from, or
    None when it is already importable / not present."""
    from . import _algebra
    where = None
    if add_reference_path:
        where = reference_python_dir()
        if where and where not in sys.path:
            sys.path.insert(0, where)
    sys.modules[MODULE] = _algebra
    for sub in ("ec_bn254", "ec_bls12_381", "polynomial_bn254", "polynomial_bls12_381", "circuit"):
        sys.modules[f"{MODULE}.{sub}"] = getattr(_algebra, sub)
    pkg = sys.modules.get("zksnake")
    if pkg is not None:
        setattr(pkg, "_algebra", _algebra)
    return where


def installed():
    from . import _algebra
    return sys.modules.get(MODULE) is _algebra


# pytest plugin behaviour: `-p zksnake_b200.dropin` imports this module before collection
if os.environ.get("ZKB_DROPIN_AUTOINSTALL", "1") != "0" and any(a.endswith("zksnake_b200.dropin") for a in sys.argv):
    install()
