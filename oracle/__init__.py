"""CPU oracle for the Reef hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product path (reef_b200/) never
does: it fails loudly when libreef_b200.so is missing.

PARITY STATUS
  * MLE sweeps, nlookup driver, doc encoding, Merkle layout: restated from
    in-tree Rust (file:line cited per function) and pinned by the reference's
    own KATs (mle_partial, mle_linear_basic, make_mt) -- see tests/.
  * Poseidon (neptune 8.1.0), Pasta curves (fil_pasta_curves 0.5.2): the
    algorithm lives in crates that are NOT under /root/reference and no Rust
    toolchain exists here, so digests are "parity unpinned" against the
    reference binary; they are pinned instead against neptune's published
    algorithm (Grain-LFSR constants, Cauchy MDS, SAFE sponge tag) and against
    two upstream known-answer values reproduced in tests/test_oracle_kats.py.
"""
