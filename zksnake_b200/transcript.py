"""Fiat-Shamir transcript of the PlonK / KZG provers.

Byte conventions are the reference's (python/zksnake/transcript.py:29-71), because every challenge -- hence every proof byte --
depends on them: blake2b; an int is absorbed big-endian, left-padded to `bit_length()` BYTES (the reference passes the bit length
where a byte length is meant; reproduced, not fixed); a curve point is absorbed in its compressed encoding; str as UTF-8; a list
element by element; after a challenge is squeezed the hash restarts from that digest.  The implementation is a small encoder
(`absorbed_bytes`) in front of hashlib, shared by `append`.
"""
import hashlib

from ._algebra import ec_bls12_381, ec_bn254
from ._algebra._poly import _R

_POINT_TYPES = (ec_bn254.PointG1, ec_bn254.PointG2, ec_bls12_381.PointG1, ec_bls12_381.PointG2)


def absorbed_bytes(item):
    """The byte string one transcript item contributes."""
    if isinstance(item, (bytes, bytearray)):
        return bytes(item)
    if isinstance(item, str):
        return item.encode()
    if isinstance(item, int):
        return item.to_bytes(item.bit_length(), "big")      # sic: bit length used as the byte count
    if isinstance(item, _POINT_TYPES):
        return bytes(item.to_bytes())
    if isinstance(item, list) and item and isinstance(item[0], (int,) + _POINT_TYPES):
        return b"".join(absorbed_bytes(x) for x in item)
    raise TypeError(f"Type of {type(item)} is not supported as transcript")


class FiatShamirTranscript:
    def __init__(self, label=b"", field=None, alg="blake2b"):
        self._alg, self._label = alg, label
        self.field = field or _R[0]
        self._state = hashlib.new(alg, label)

    def reset(self):
        self._state = hashlib.new(self._alg, self._label)

    def append(self, data):
        self._state.update(absorbed_bytes(data))

    def get_challenge(self):
        out = self._state.digest()
        self._state = hashlib.new(self._alg, out)            # ratchet: the next challenge depends on this one
        return out

    def get_challenge_scalar(self):
        return int.from_bytes(self.get_challenge(), "big") % self.field
