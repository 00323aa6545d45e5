"""Poseidon permutation + SAFE sponge as used by Reef (oracle; test infrastructure only).

Reference call sites (the hash itself lives in the un-vendored crate `neptune 8.1.0`,
Cargo.toml:19 -- "parity unpinned" against the reference binary, see oracle/__init__.py):

  * constants   `Sponge::<Fq, U4>::api_constants(Strength::Standard)`
                /root/reference/src/backend/framework.rs:71-73   (arity 4 => width t = 5)
  * one-shot    `Sponge::new_with_constants(pc, Mode::Simplex)`; start(IOPattern[Absorb(n),Squeeze(1)]);
                absorb; squeeze; finish
                /root/reference/src/backend/merkle_tree.rs:80-114, commitment.rs:495-510
  * transcript  /root/reference/src/backend/r1cs.rs:2260-2311, r1cs_helper.rs:479-488

Published algorithm restated here (neptune 8.1.0):
  round numbers   security inequalities of the Poseidon paper with +2 full rounds / +7.5% partial
                  rounds margin                         => (R_F, R_P) = (8, 56) at t = 5
  round constants Grain LFSR, init = [field=1 (2b), sbox=1 (4b), n=255 (12b), t (12b), R_F (10b),
                  R_P (10b), 30 ones], 160 warm-up clocks, pairs-sampling, rejection >= p
  MDS             Cauchy  M[i][j] = 1 / (x_i + y_j),  x_i = i, y_j = t + j
  round           add round constants -> S-box x^5 (all lanes in full rounds, lane 0 in partial
                  rounds) -> state * M
  sponge (SAFE)   state[0] = capacity = tag(IOPattern, domain separator 0); rate = state[1..5];
                  absorb ADDS into the rate, permuting when the rate is full; the first squeeze
                  after an absorb always permutes; squeeze reads state[1 + pos].
  tag             u128 polynomial hash in wrapping arithmetic, base 2^128 - 159, of the
                  run-length-merged op words (Absorb(n) -> n + 2^31, Squeeze(n) -> n), then the
                  domain separator.
"""
from __future__ import annotations

import math
from functools import lru_cache

from .fields import FQ

MASK128 = (1 << 128) - 1
HASHER_BASE = (1 << 128) - 159


# --------------------------------------------------------------------------- round numbers
def _secure(t: int, rf: int, rp: int, n: int = 255, m: int = 128) -> bool:
    rf_stat = 6.0 if m <= (n - 3.0) * (t + 1.0) else 10.0
    rf_interp = 0.43 * m + math.log2(t) - rp
    rf_grob_1 = 0.21 * n - rp
    rf_grob_2 = (0.14 * n - 1.0 - rp) / (t - 1.0)
    rf_max = max(math.ceil(x) for x in (rf_stat, rf_interp, rf_grob_1, rf_grob_2))
    return rf >= rf_max


@lru_cache(None)
def round_numbers(t: int, n: int = 255) -> tuple[int, int]:
    """(R_F, R_P) with neptune's `Strength::Standard` security margin."""
    best = None
    for rf_t in range(2, 40, 2):
        for rp_t in range(4, 200):
            if _secure(t, rf_t, rp_t, n):
                rf = rf_t + 2
                rp = math.ceil(1.075 * rp_t)
                cost = t * rf + rp
                if best is None or cost < best[0] or (cost == best[0] and rf < best[1]):
                    best = (cost, rf, rp)
    return best[1], best[2]


# --------------------------------------------------------------------------- Grain LFSR
class _Grain:
    def __init__(self, field: int, sbox: int, n: int, t: int, rf: int, rp: int):
        bits = []
        for width, val in ((2, field), (4, sbox), (12, n), (12, t), (10, rf), (10, rp), (30, (1 << 30) - 1)):
            bits += [(val >> (width - 1 - i)) & 1 for i in range(width)]
        assert len(bits) == 80
        self.s = bits
        for _ in range(160):
            self._clock()

    def _clock(self) -> int:
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def bit(self) -> int:
        b = self._clock()
        while b == 0:
            self._clock()
            b = self._clock()
        return self._clock()

    def bits(self, k: int) -> int:
        x = 0
        for _ in range(k):
            x = (x << 1) | self.bit()
        return x


@lru_cache(None)
def constants(p: int = FQ, t: int = 5, nbits: int = 255, sbox_flag: int = 1):
    """(R_F, R_P, round_constants[(R_F+R_P)*t], mds[t][t])."""
    rf, rp = round_numbers(t, nbits)
    g = _Grain(1, sbox_flag, nbits, t, rf, rp)
    rc = []
    while len(rc) < (rf + rp) * t:
        x = g.bits(nbits)
        if x < p:
            rc.append(x)
    mds = [[pow(i + t + j, -1, p) for j in range(t)] for i in range(t)]
    return rf, rp, tuple(rc), tuple(tuple(r) for r in mds)


# --------------------------------------------------------------------------- permutation
def permute(state, p: int = FQ, t: int = 5, nbits: int = 255):
    """One Poseidon permutation of a width-t state (list of ints), textbook form."""
    rf, rp, rc, mds = constants(p, t, nbits)
    s = list(state)
    assert len(s) == t
    k = 0
    for r in range(rf + rp):
        full = r < rf // 2 or r >= rf // 2 + rp
        s = [(x + rc[k + i]) % p for i, x in enumerate(s)]
        k += t
        if full:
            s = [pow(x, 5, p) for x in s]
        else:
            s[0] = pow(s[0], 5, p)
        s = [sum(s[i] * mds[i][j] for i in range(t)) % p for j in range(t)]
    return s


# --------------------------------------------------------------------------- SAFE sponge
ABSORB, SQUEEZE = "A", "S"


def io_pattern_tag(pattern, domain_separator: int = 0) -> int:
    """u128 tag of an IOPattern [(ABSORB|SQUEEZE, n), ...] (neptune sponge::api)."""
    x_i, state = 1, 0
    cur_kind, cur_n = ABSORB, 0

    def update(a):
        nonlocal x_i, state
        x_i = (x_i * HASHER_BASE) & MASK128
        state = (state + x_i * a) & MASK128

    def finish_op():
        if cur_n == 0:
            return
        assert cur_n >> 31 == 0
        update(cur_n + (1 << 31) if cur_kind == ABSORB else cur_n)

    for kind, n in pattern:
        if kind == cur_kind:
            cur_n += n
        else:
            finish_op()
            cur_kind, cur_n = kind, n
    finish_op()
    update(domain_separator)
    return state


class Sponge:
    """`Sponge<Fq, U4>` in `Mode::Simplex`, driven through the SpongeAPI (start/absorb/squeeze/finish)."""

    def __init__(self, p: int = FQ, t: int = 5):
        self.p, self.t, self.rate = p, t, t - 1
        self.state = [0] * t
        self.pattern = None
        self.io = 0
        self.apos = self.spos = 0
        self.n_perm = 0

    def start(self, pattern, domain_separator: int = 0):
        self.pattern = list(pattern)
        self.state = [io_pattern_tag(pattern, domain_separator) % self.p] + [0] * self.rate
        self.io = 0
        self.apos = self.spos = 0

    def _permute(self):
        self.state = permute(self.state, self.p, self.t)
        self.n_perm += 1

    def absorb(self, elems):
        for e in elems:
            if self.apos == self.rate:
                self._permute()
                self.apos = 0
            self.state[1 + self.apos] = (self.state[1 + self.apos] + e) % self.p
            self.apos += 1
        assert self.pattern[self.io] == (ABSORB, len(elems)), "IOPattern mismatch"
        self.io += 1
        self.spos = self.rate

    def squeeze(self, n: int):
        out = []
        for _ in range(n):
            if self.spos == self.rate:
                self._permute()
                self.spos = 0
                self.apos = 0
            out.append(self.state[1 + self.spos])
            self.spos += 1
        assert self.pattern[self.io] == (SQUEEZE, n), "IOPattern mismatch"
        self.io += 1
        return out

    def finish(self):
        self.state = [0] * self.t
        if self.io != len(self.pattern):
            raise ValueError("ParameterUsageMismatch")


def hash_once(elems, p: int = FQ) -> int:
    """IOPattern [Absorb(n), Squeeze(1)] one-shot: merkle_tree.rs:80-114, commitment.rs:495-510."""
    sp = Sponge(p)
    sp.start([(ABSORB, len(elems)), (SQUEEZE, 1)])
    sp.absorb(elems)
    out = sp.squeeze(1)[0]
    sp.finish()
    return out


def calc_d(v: int, salt: int) -> int:
    """commitment.rs:495-510."""
    return hash_once([v, salt])


# --------------------------------------------------------------------------- PoseidonRO (nova-snark)
RO_ARITY = 24          # nova-snark `PoseidonConstantsCircuit` = PoseidonConstants<Scalar, U24>


def poseidon_ro(elems, base_p: int, scalar_p: int, num_bits: int = 256) -> int:
    """nova-snark `PoseidonRO<Base, Scalar>`: new(constants, len(elems)); absorb each; squeeze(num_bits).
    Reference call site: commitment.rs:190-198 (doc_commit_hash over the decompressed Hyrax row
    commitments; Base = Pallas base field Fp, Scalar = Fq, num_bits = 256).
    nova-snark is an un-pinned git dependency that is not under /root/reference (Cargo.toml:12): this
    restates the published upstream provider/poseidon.rs -- one neptune sponge of arity 24 in Simplex
    mode with IOPattern [Absorb(n), Squeeze(1)], the digest's low `num_bits` bits re-assembled in the
    Scalar field.  PARITY UNPINNED."""
    sp = Sponge(base_p, RO_ARITY + 1)
    sp.start([(ABSORB, len(elems)), (SQUEEZE, 1)])
    sp.absorb([int(e) % base_p for e in elems])
    h = sp.squeeze(1)[0]
    sp.finish()
    return (h & ((1 << num_bits) - 1)) % scalar_p


def point_coordinates(P):
    """`to_coordinates` of a nova-snark group element: (x, y, is_infinity); the identity is (0, 0, true)."""
    return (0, 0, 1) if P is None else (int(P[0]), int(P[1]), 0)


def doc_commit_hash(row_commitments, base_p: int, scalar_p: int) -> int:
    """commitment.rs:190-198: RO over (x, y, is_infinity) of every row commitment, squeeze(256)."""
    elems = []
    for P in row_commitments:
        elems += list(point_coordinates(P))
    return poseidon_ro(elems, base_p, scalar_p, 256)
