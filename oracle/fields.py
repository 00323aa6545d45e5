"""Pasta-cycle field constants and byte conventions (oracle; test infrastructure only).

Fq  = Pallas scalar field = Vesta base field.  Pinned in-tree:
      /root/reference/src/backend/r1cs_helper.rs:37-38 (CirC custom modulus).
Fp  = Pallas base field = Vesta scalar field.  Public Pasta parameter
      (fil_pasta_curves 0.5.2, not vendored under /root/reference).

Element bytes are 32-byte little-endian canonical integers:
      r1cs_helper.rs:488, commitment.rs:528 (`Integer::from_digits(.., Order::Lsf)`).
"""

FQ = 28948022309329048855892746252171976963363056481941647379679742748393362948097
FP = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001

assert FQ == 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
assert FP == 28948022309329048855892746252171976963363056481941560715954676764349967630337

# BLS12-381 scalar field: only used to replay neptune's upstream known-answer
# vectors (which are published for that field) through the generic algorithm.
BLS_FR = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def to_le32(x: int) -> bytes:
    """Canonical 32-byte little-endian encoding (ff::PrimeField::to_repr for Pasta)."""
    return int(x).to_bytes(32, "little")


def from_le32(b: bytes) -> int:
    return int.from_bytes(bytes(b), "little")


def pack(xs, mod=None) -> bytes:
    """Concatenate canonical encodings; values are reduced with rem_floor semantics."""
    if mod is None:
        return b"".join(to_le32(x) for x in xs)
    return b"".join(to_le32(x % mod) for x in xs)


def unpack(b: bytes):
    b = bytes(b)
    assert len(b) % 32 == 0
    return [from_le32(b[i:i + 32]) for i in range(0, len(b), 32)]
