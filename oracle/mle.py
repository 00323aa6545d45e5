"""nlookup sum-check MLE kernels, restated line by line (oracle; test infrastructure only).

Follows /root/reference/src/backend/r1cs_helper.rs:441-641.  All values are Python ints;
`% FQ` is GMP `rem_floor(modulus)` (non-negative representative).
"""
from __future__ import annotations

from .fields import FQ


def gen_eq_table(rs, qs, last_q):
    """r1cs_helper.rs:508-544.  len(rs) == len(qs)+1; bit j of index i pairs with last_q[j]."""
    ell = len(last_q)
    n = 1 << ell
    assert len(rs) == len(qs) + 1
    eq_t = [0] * n
    for i, q in enumerate(qs):
        eq_t[q] += rs[i]
    rm = rs[len(qs)]
    for i in range(n):
        term = rm
        for j in range(ell - 1, -1, -1):
            xi = (i >> j) & 1
            term *= xi * last_q[j] + (1 - xi) * (1 - last_q[j])
        eq_t[i] = (eq_t[i] + term) % FQ
    return eq_t


def gen_eq_table_fast(rs, qs, last_q):
    """Same values as gen_eq_table in O(N) (doubling construction); used for large parity sizes."""
    ell = len(last_q)
    tab = [rs[len(qs)] % FQ]
    # bit j of the index pairs with last_q[j]; build from the top bit down so that the
    # lowest bit varies fastest.
    for j in range(ell - 1, -1, -1):
        qj = last_q[j] % FQ
        one_m = (1 - qj) % FQ
        nxt = [0] * (2 * len(tab))
        for k, v in enumerate(tab):
            nxt[2 * k] = v * one_m % FQ
            nxt[2 * k + 1] = v * qj % FQ
        tab = nxt
    # tab index so far has bit (ell-1) as the MOST significant factor applied first:
    # after processing j = ell-1 .. 0, position bit for j is at weight 2^j already.
    for i, q in enumerate(qs):
        tab[q] = (tab[q] + rs[i]) % FQ
    return tab


def linear_mle_product_coeffs(table_t, table_eq, ell, i):
    """First loop of linear_mle_product (r1cs_helper.rs:449-477): returns (xsq, x, con)."""
    pw = 1 << (ell - i)
    assert len(table_t) == 1 << ell and len(table_eq) == 1 << ell
    xsq = x = con = 0
    for b in range(pw):
        t0, t1 = table_t[b], table_t[b + pw]
        e0, e1 = table_eq[b], table_eq[b + pw]
        ts, es = t1 - t0, e1 - e0
        xsq += ts * es
        x += es * t0
        x += ts * e0
        con += t0 * e0
    return xsq % FQ, x % FQ, con % FQ


def linear_mle_fold(table_t, table_eq, ell, i, r_i):
    """Second loop of linear_mle_product (r1cs_helper.rs:490-503), in place."""
    pw = 1 << (ell - i)
    for b in range(pw):
        table_t[b] = (table_t[b] * (1 - r_i) + table_t[b + pw] * r_i) % FQ
        table_eq[b] = (table_eq[b] * (1 - r_i) + table_eq[b + pw] * r_i) % FQ


def linear_mle_product(table_t, table_eq, ell, i, sponge):
    """r1cs_helper.rs:441-506: coefficients, absorb [con, x, xsq], squeeze r_i, fold."""
    xsq, x, con = linear_mle_product_coeffs(table_t, table_eq, ell, i)
    sponge.absorb([con, x, xsq])
    r_i = sponge.squeeze(1)[0]
    linear_mle_fold(table_t, table_eq, ell, i, r_i)
    return r_i, xsq, x, con


def prover_mle_partial_eval(prods, x, es, for_t, last_q=None):
    """r1cs_helper.rs:551-634.  x[k] == -1 marks the hole; returns (hole_coeff, const)."""
    m = len(x)
    if for_t:
        assert (1 << (m - 1)) <= len(prods) <= (1 << m)
        assert len(es) == len(prods)
    elif last_q is not None:
        assert len(es) + 1 == len(prods)
    hole = minus = 0
    for i in range(len(es) + 1):
        if i < len(es):
            prod = prods[i]
            next_hole = 0
            for j in range(m - 1, -1, -1):
                ej = (es[i] >> j) & 1
                xv = x[m - j - 1]
                if xv == -1:
                    next_hole = ej
                else:
                    prod *= xv if ej == 1 else 1 - xv
                    prod %= FQ          # value-preserving: final results are reduced mod FQ
            if next_hole == 1:
                hole += prod
            else:
                minus += prod
        elif last_q is not None:
            prod = prods[i]
            nh, nm = 1, 1
            for j in range(m):
                ej = last_q[j]
                if x[j] == -1:
                    nh, nm = ej, 1 - ej
                else:
                    prod *= ej * x[j] + (1 - ej) * (1 - x[j])
                    prod %= FQ
            hole += prod * nh
            minus += prod * nm
    hole -= minus
    return hole % FQ, minus % FQ


def verifier_mle_eval(table, q):
    """r1cs_helper.rs:637-641."""
    return prover_mle_partial_eval(table, q, list(range(len(table))), True, None)[1]


def mle_eval_fast(table, x):
    """T~(x), x[0] <-> top index bit; O(N) folding (equals verifier_mle_eval on full tables)."""
    tab = [v % FQ for v in table]
    n = 1 << len(x)
    tab += [0] * (n - len(tab))
    for r in x:
        h = len(tab) // 2
        tab = [(tab[b] + r * (tab[b + h] - tab[b])) % FQ for b in range(h)]
    return tab[0]
