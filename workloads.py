"""Synthetic documents of BASELINE.json's configs, as document-table codes.

Pure numpy, no oracle/ and no reef_b200 imports: bench.py's GPU arm, its reference arm and the
tests all generate their inputs here, so the two arms of a measurement see the same bytes.
`encode()` restates framework.rs:978-1011 `doc_transform` for the three alphabets of config.rs
(`ascii` :232-233, `utf8` :255-256, `dna` :269) on code-point arrays; tests/test_bench_inputs.py
pins it against oracle.nlookup.doc_transform and against the library's reef_doc_transform.
"""
from __future__ import annotations

import math
import os

import numpy as np

FQ = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
FP = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001

N_UTF8 = 0x110000 - 0x800          # config.rs:255-256: every code point that is a `char` (no surrogates)


def logmn(mn: int) -> int:
    """costs.rs:10-15: (mn as f32).log2().ceil(), logmn(1) == 1."""
    if mn == 1:
        return 1
    return int(math.ceil(float(np.log2(np.float32(mn)))))


def alphabet_size(ab: str) -> int:
    return {"ascii": 128, "utf8": N_UTF8, "dna": 4}[ab]


def encode(ab: str, cps: np.ndarray) -> np.ndarray:
    """doc_transform on an array of Unicode code points -> padded u32 document-table codes.
    index = position in the alphabet; EPSILON = |ab| + 1; EOF (char 26) = |ab| + 2 (for ascii / utf8
    that assignment OVERWRITES char 26's own index, framework.rs:985-986)."""
    cps = np.asarray(cps, dtype=np.int64)
    n_ab = alphabet_size(ab)
    if ab == "dna":
        lut = np.full(128, -1, dtype=np.int64)
        for i, c in enumerate("ACGT"):
            lut[ord(c)] = i
        codes = lut[cps]
        if (codes < 0).any():
            raise ValueError("Character in document that's not in alphabet")
    else:
        if ((cps >= 0xD800) & (cps < 0xE000)).any() or (cps >= (128 if ab == "ascii" else 0x110000)).any():
            raise ValueError("Character in document that's not in alphabet")
        codes = np.where(cps >= 0xE000, cps - 0x800, cps)
        codes = np.where(cps == 26, n_ab + 2, codes)
    total = len(cps) + 2
    lg = logmn(total)
    if (1 << lg) < total:
        raise OverflowError("attempt to subtract with overflow (f32 logmn, framework.rs:1007)")
    out = np.zeros(1 << lg, dtype=np.uint32)
    out[:len(cps)] = codes
    out[len(cps)] = n_ab + 2           # EOF
    out[len(cps) + 1] = n_ab + 1       # EPSILON
    return out


def _put(cps, pos, s):
    cps[pos:pos + len(s)] = [ord(c) for c in s]


def document(cfg: str, doc_len: int | None = None, seed_shift: int = 0):
    """(alphabet, code points) of BASELINE config `cfg`, optionally at another length."""
    if cfg in ("cfg2", "target"):
        n = doc_len or (1 << 16 if cfg == "cfg2" else 1 << 20)
        cps = np.full(n, ord("a"), dtype=np.int64)
        cps[-1] = ord("b")
        return "ascii", cps
    if cfg == "cfg3":          # dna, '(A|C|G|T){4}TATA.*', one match
        n = doc_len or 1 << 20
        cps = np.asarray([ord(c) for c in "ACG"], dtype=np.int64)[np.random.default_rng(20 + seed_shift).integers(0, 3, size=n)]
        _put(cps, min(4100, n - 8), "TATA")
        return "dna", cps
    if cfg == "cfg4":          # ascii printable without h, w; 'hello.*world', one match
        n = doc_len or 1 << 20
        pool = np.asarray([c for c in range(0x20, 0x7F) if chr(c) not in "hw"], dtype=np.int64)
        cps = pool[np.random.default_rng(21 + seed_shift).integers(0, len(pool), size=n)]
        _put(cps, min(1000, n // 4), "hello")
        _put(cps, min(900000, n - 16), "world")
        return "ascii", cps
    if cfg == "cfg5":          # utf8 0x20..0x2FFF without f, b; '.*(foo|bar|baz).*', one match
        n = doc_len or 1 << 22
        pool = np.asarray([c for c in range(0x20, 0x3000) if chr(c) not in "fb"], dtype=np.int64)
        cps = pool[np.random.default_rng(22 + seed_shift).integers(0, len(pool), size=n)]
        _put(cps, min(123456, n // 2), "foo")
        return "utf8", cps
    raise KeyError(cfg)


def safe_len(n: int) -> int:
    """The reference's f32 logmn mis-rounds 2^22 + 2 and 2^23 + 2 (doc_transform would panic at exactly
    2^22 / 2^23 characters): 64 more characters put the length where logmn is exact."""
    return n + 64 if (1 << logmn(n + 2)) < n + 2 else n


def hyrax_dims(ell: int):
    """nova-snark `EqPolynomial::compute_factored_lens(ell)` = (ell / 2, ell - ell / 2) as used at
    commitment.rs:173-174: rows = 2^left, cols = 2^right."""
    left = ell // 2
    return 1 << left, 1 << (ell - left)


# ---------------------------------------------------------------------------------------------
# generators k*G for synthetic commitment keys: distinct, cheap, checkable
# ---------------------------------------------------------------------------------------------
def curve_multiples(p: int, n: int):
    """[G, 2G, ..., nG] on y^2 = x^3 + 5 over F_p with G = (-1, 2)."""
    gx, gy = p - 1, 2
    pts, (x, y) = [], (gx, gy)
    for _ in range(n):
        pts.append((x, y))
        if x == gx and y == gy:
            lam = 3 * x * x * pow(2 * y, -1, p) % p
        else:
            lam = (y - gy) * pow(x - gx, -1, p) % p
        x3 = (lam * lam - x - gx) % p
        x, y = x3, (lam * (x - x3) - y) % p
    return pts


_CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "build", "gens")


_CACHE_MAX = 1 << 18


def _expand(curve: str, n: int) -> bytes:
    """More than 2^18 distinct curve points without more big-int point additions: the images of the
    first 2^18 multiples under the curve's endomorphism (x, y) -> (zeta x, y), zeta^3 = 1, and under
    negation -- P, phi(P), phi^2(P), -P are distinct points of the same prime-order group."""
    p = FP if curve == "pallas" else FQ
    assert n <= 4 * _CACHE_MAX
    base = generators(curve, _CACHE_MAX)
    zeta = next(z for z in (pow(g, (p - 1) // 3, p) for g in range(2, 50)) if z != 1)
    xs = [int.from_bytes(base[i * 64:i * 64 + 32], "little") for i in range(_CACHE_MAX)]
    ys = [base[i * 64 + 32:i * 64 + 64] for i in range(_CACHE_MAX)]
    out = [base]
    for k in (1, 2):
        z = pow(zeta, k, p)
        out.append(b"".join((x * z % p).to_bytes(32, "little") + y for x, y in zip(xs, ys)))
    out.append(b"".join(x.to_bytes(32, "little") + (p - int.from_bytes(y, "little")).to_bytes(32, "little") for x, y in zip(xs, ys)))
    return b"".join(out)[:64 * n]


def generators(curve: str, n: int) -> bytes:
    """n affine generators (64 B each, x || y LE) of `curve`, cached on disk (Python big-int k*G is
    ~10 us per point: 30 s per bench run in round 1)."""
    p = FP if curve == "pallas" else FQ
    if n > _CACHE_MAX:
        return _expand(curve, n)
    path = os.path.join(_CACHE_DIR, f"{curve}_{n}.bin")
    if os.path.exists(path) and os.path.getsize(path) == 64 * n:
        return open(path, "rb").read()
    # reuse a longer cached prefix if there is one
    if os.path.isdir(_CACHE_DIR):
        for f in os.listdir(_CACHE_DIR):
            if f.startswith(curve + "_") and f.endswith(".bin"):
                try:
                    m = int(f[len(curve) + 1:-4])
                except ValueError:
                    continue
                fp = os.path.join(_CACHE_DIR, f)
                if m >= n and os.path.getsize(fp) == 64 * m:
                    with open(fp, "rb") as fh:
                        return fh.read(64 * n)
    raw = b"".join(int(P[0]).to_bytes(32, "little") + int(P[1]).to_bytes(32, "little") for P in curve_multiples(p, n))
    try:
        os.makedirs(_CACHE_DIR, exist_ok=True)
        tmp = path + f".{os.getpid()}.tmp"
        with open(tmp, "wb") as fh:
            fh.write(raw)
        os.replace(tmp, path)
    except OSError:
        pass
    return raw
