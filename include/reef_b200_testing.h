/* reef_b200_testing.h -- TEST HOOKS exported by libreef_b200.so.  Not part of the product ABI.
 *
 * The field arithmetic and the single-thread Poseidon permutation are written once as
 * __host__ __device__ code (reef_b200/csrc/fp.cuh, poseidon.cuh).  These hooks evaluate the
 * HOST instantiation of that shared code so that its structure (column bookkeeping, Montgomery
 * reduction, lazy accumulation, optimised round schedule) is unit-tested in the CPU-only test
 * tier.  No product entry point in reef_b200.h ever calls them, and they never touch a GPU.
 */
#ifndef REEF_B200_TESTING_H
#define REEF_B200_TESTING_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* field: 0 = Fq (Pallas scalar), 1 = Fp (Pallas base).  Operands/outputs canonical 32-byte LE.
 * op: 0 a*b, 1 a+b, 2 a-b, 3 1/a, 4 lazy(a*b + a*a + b*b), 5 lazy(40000 * a*b), 6 a[limb0]*b,
 *     7 a*a, 8 a*b in 29-bit limbs (fp29.cuh), 9 (a+b)(2a+b) + (a+b)^2 in 29-bit limbs */
int reef_hosttest_field_op(int field, int op, const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
/* plain 256x256 -> 512-bit integer product */
int reef_hosttest_mul_wide(const uint8_t a[32], const uint8_t b[32], uint8_t out[64]);
/* one Poseidon permutation of a width-5 state (canonical in/out) */
int reef_hosttest_poseidon_permute(const uint8_t in[160], uint8_t out[160]);

/* PoseidonRO (width 25; poseidon_ro.cu) on the HOST through the shared field code: the sponge over `n` canonical elements
 * of field (0 = Fq, 1 = Fp), out = state[1] canonical (no truncation); and the derived constants, canonical:
 * rc = 67 * 25 * 32 bytes, mds = 25 * 25 * 32 bytes */
int reef_hosttest_poseidon_ro(int field, const uint8_t* elems, uint64_t n, uint8_t out[32]);
int reef_hosttest_poseidon_ro_constants(int field, uint8_t* rc, uint8_t* mds);
/* 1 when the optimised width-25 schedule derived on the host (sparse partial rounds, rescaled lane 0: what the GPU
 * kernel runs) reproduced the textbook permutation on the host test vectors; 0 makes the library use the textbook kernel */
int reef_hosttest_poseidon_ro_fast_ok(int field);

/* curve formulas (ec.cuh), host instantiation.  curve: 0 Pallas, 1 Vesta.  Points affine 64 B.
 * op: 0 P+Q via XYZZ full add, 1 P+Q via mixed add, 2 2P, 3 P-Q via mixed add (neg), 4 k*P (k = first 8 bytes of q) */
int reef_hosttest_ec_op(int curve, int op, const uint8_t p[64], const uint8_t q[64], uint8_t out[64]);

/* GPU test hook for the lane-parallel transcript permutation (poseidon_lp.cuh): `n_perms` successive
 * permutations of one width-5 state (canonical in / out) by one 256-thread CTA; cycles_per_perm
 * (may be NULL, else 9 entries) receives the SM cycles per permutation and, for the last one, the
 * cycles of its four phases and of the two waits inside the chain.  ctx: a reef_ctx*. */
int reef_gputest_poseidon_permute_lp(void* ctx, const uint8_t in[160], uint32_t n_perms, uint8_t out[160], uint64_t* cycles_per_perm);

#ifdef __cplusplus
}
#endif
#endif
