"""reef_b200 -- B200-native prover hot path for eniac/Reef.

The product is `libreef_b200.so` (hand-written sm_100a CUDA behind the C ABI declared in
`include/reef_b200.h`).  This package is only the ctypes binding used by the tests and the
benchmark; a Rust host binds the same symbols with `extern "C"`.

There is NO CPU fallback: importing works everywhere (so the ABI can be inspected), but every
compute call raises `ReefError` unless the library was built and an sm_100 GPU is present.
"""
from ._lib import ReefError, lib, lib_path  # noqa: F401
from .backend import (  # noqa: F401
    TAG_NL, TAG_NLDOC, TAG_NLHYBRID, Bases, Context, MerkleCommitment, NlookupResult, ShardedNlookup, Sponge, Sumcheck, Table, Witness,
    combined_q, doc_transform, io_pattern_tag, logmn,
)
