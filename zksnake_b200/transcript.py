"""Fiat-Shamir transcript with the reference's exact byte conventions (python/zksnake/transcript.py:29-71): blake2b, ints
appended big-endian with `bit_length()` BYTES (sic -- the reference passes the bit length as the byte length, so an int is
left-padded with zeros to that many bytes; reproduced because every challenge depends on it), points appended in their
compressed encoding, and the hasher re-seeded with the digest after every challenge."""
import hashlib

from .ecc import ispointG1, ispointG2


class FiatShamirTranscript:
    def __init__(self, label=b"", field=None, alg="blake2b"):
        from .polynomial import BN254_SCALAR_FIELD
        self.alg, self.label = alg, label
        self.hasher = hashlib.new(alg, label)
        self.field = field or BN254_SCALAR_FIELD

    def reset(self):
        self.hasher = hashlib.new(self.alg, self.label)

    @staticmethod
    def _int_bytes(v):
        return int.to_bytes(v, v.bit_length(), "big")

    def append(self, data):
        if isinstance(data, bytes):
            self.hasher.update(data)
        elif isinstance(data, str):
            self.hasher.update(data.encode())
        elif isinstance(data, int):
            self.hasher.update(self._int_bytes(data))
        elif data and isinstance(data, list) and isinstance(data[0], int):
            for d in data:
                self.hasher.update(self._int_bytes(d))
        elif ispointG1(data) or ispointG2(data):
            self.hasher.update(bytes(data.to_bytes()))
        elif data and isinstance(data, list) and (ispointG1(data[0]) or ispointG2(data[0])):
            for d in data:
                self.hasher.update(bytes(d.to_bytes()))
        else:
            raise TypeError(f"Type of {type(data)} is not supported as transcript")

    def get_challenge(self):
        digest = self.hasher.digest()
        self.hasher = hashlib.new(self.alg, digest)
        return digest

    def get_challenge_scalar(self):
        return int.from_bytes(self.get_challenge(), "big") % self.field
