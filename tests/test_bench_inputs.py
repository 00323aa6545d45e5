"""CPU tier: bench.py's input generators (kept free of oracle/ imports on the GPU arm) agree with the
oracle, and the weak-scaling document lengths stay clear of the reference's f32 `logmn` quirk."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle.curves import PALLAS, VESTA  # noqa: E402
from oracle.nlookup import ASCII_AB, doc_transform, logmn  # noqa: E402


def test_curve_multiples_match_the_oracle():
    assert bench.curve_multiples(bench.FP, 40) == PALLAS.multiples(40)
    assert bench.curve_multiples(bench.FQ, 40) == VESTA.multiples(40)
    assert all(PALLAS.on_curve(P) for P in bench.curve_multiples(bench.FP, 40))


def test_workload_document_is_the_reference_encoding():
    assert bench.ASCII_AB == ASCII_AB
    w = bench.make_workload("cfg2")
    exp = np.asarray(doc_transform(ASCII_AB, "a" * 65535 + "b"), dtype=np.uint32)
    assert w["doc_len"] == 1 << 16 and (w["udoc"] == exp).all()
    assert len(w["bases_pri"]) == 64 * w["n_pri"] and len(w["bases_sec"]) == 64 * w["n_sec"]


def test_weak_scaling_lengths_avoid_the_f32_logmn_quirk():
    # framework.rs:1007 underflows when 2^logmn(len + 2) < len + 2, which the f32 logmn (costs.rs:10-15)
    # causes at exactly 2^22 and 2^23 characters
    base = bench.WORKLOADS["target"]["doc_len"]
    for world in (1, 2, 4, 8):
        n = base * world
        if (1 << logmn(n + 2)) < n + 2:
            n += 64
        assert (1 << logmn(n + 2)) >= n + 2
        assert (1 << logmn(n + 2)) == 2 * base * world          # 2^21 table entries per GPU at every size
