"""Spartan-side sum-check rounds, R1CS mat-vec and IPA generator folding (oracle; test
infrastructure only -- never imported by reef_b200/).

PARITY UNPINNED.  These run inside nova-snark (`CompressedSNARK::prove`,
/root/reference/src/backend/framework.rs:695-698 with S = spartan::RelaxedR1CSSNARK<G,
ipa_pc::EvaluationEngine<G>>, framework.rs:5-8; `cap_prove`, commitment.rs:261-268;
`hyrax_gen.prove_eval`, commitment.rs:371-393).  nova-snark is a git dependency on
github.com/sga001/Nova with NO pinned revision (/root/reference/Cargo.toml:12, Cargo.lock is
git-ignored) and its source is not under /root/reference; the reference's own tests only check
prove -> verify round trips (framework.rs:1071-1172).  What is restated is therefore the
published upstream algorithm (microsoft/Nova `spartan/sumcheck.rs`, `spartan/polynomial.rs`,
`provider/ipa_pc.rs`), anchored on the in-tree call sites above:

  * `MultilinearPolynomial::bound_poly_var_top(r)`: Z[i] <- Z[i] + r (Z[i + n/2] - Z[i]) -- the
    same MSB-first fold as Reef's own `linear_mle_product` (r1cs_helper.rs:490-503).
  * `SumcheckProof::prove_quad`: per round, with `len = n/2`,
        eval_0 = sum_i A[i] B[i],   eval_2 = sum_i (2 A[len+i] - A[i]) (2 B[len+i] - B[i]),
    message polynomial from evaluations [eval_0, claim - eval_0, eval_2].
  * `prove_cubic_with_additive_term`, comb(a, b, c, d) = a (b c - d): evaluations at 0, 2, 3 with the
    bound points 2 hi - lo and 3 hi - 2 lo; message from [eval_0, claim - eval_0, eval_2, eval_3].
  * IPA round: G' = G_L * r^-1 + G_R * r (`ck.fold`).
"""
from __future__ import annotations

from .curves import INF


def bound_top(Z, r, p):
    n = len(Z) // 2
    return [(Z[i] + r * (Z[i + n] - Z[i])) % p for i in range(n)]


def round_quad(A, B, p):
    n = len(A) // 2
    e0 = sum(A[i] * B[i] for i in range(n)) % p
    e2 = sum((2 * A[n + i] - A[i]) * (2 * B[n + i] - B[i]) for i in range(n)) % p
    return e0, e2


def comb_cubic(a, b, c, d):
    return a * (b * c - d)


def round_cubic(A, B, C, D, p):
    n = len(A) // 2
    e0 = e2 = e3 = 0
    for i in range(n):
        lo = (A[i], B[i], C[i], D[i])
        hi = (A[n + i], B[n + i], C[n + i], D[n + i])
        e0 += comb_cubic(*lo)
        e2 += comb_cubic(*[2 * h - l for l, h in zip(lo, hi)])
        e3 += comb_cubic(*[3 * h - 2 * l for l, h in zip(lo, hi)])
    return e0 % p, e2 % p, e3 % p


def interpolate_eval(evals, r, p):
    """Value at r of the polynomial of degree len(evals)-1 through (k, evals[k]), k = 0.."""
    total = 0
    for k, yk in enumerate(evals):
        num, den = 1, 1
        for j in range(len(evals)):
            if j != k:
                num = num * (r - j) % p
                den = den * (k - j) % p
        total += yk * num * pow(den, -1, p)
    return total % p


def prove(tables, challenges, p):
    """Runs a whole sum-check (kind = len(tables) in {2, 4}) with the given challenges.
    Returns (claim, [evaluations per round incl. eval_1], final bound values)."""
    kind = len(tables)
    tabs = [list(t) for t in tables]
    n = len(tabs[0])
    if kind == 2:
        claim = sum(a * b for a, b in zip(*tabs)) % p
    else:
        claim = sum(comb_cubic(*x) for x in zip(*tabs)) % p
    claim0 = claim
    rounds = []
    for r in challenges:
        ev = round_quad(*tabs, p) if kind == 2 else round_cubic(*tabs, p)
        full = [ev[0], (claim - ev[0]) % p] + list(ev[1:])
        rounds.append(full)
        claim = interpolate_eval(full, r, p)
        tabs = [bound_top(t, r, p) for t in tabs]
    assert all(len(t) == 1 for t in tabs) or len(challenges) < n.bit_length() - 1
    return claim0, rounds, [t[0] for t in tabs], claim


def spmv(row_ptr, col_idx, vals, z, p):
    out = []
    for r in range(len(row_ptr) - 1):
        out.append(sum(vals[k] * z[col_idx[k]] for k in range(row_ptr[r], row_ptr[r + 1])) % p)
    return out


def ipa_fold_bases(curve, G, s_lo, s_hi):
    n = len(G) // 2
    return [curve.add(curve.mul(s_lo, G[i]), curve.mul(s_hi, G[i + n])) for i in range(n)]
