// Latency of the lane-parallel multiplication (reef_b200/csrc/lpmul.cuh) in a dependent chain,
// next to the single-thread sqr29 / mul29 it replaces on the Fiat-Shamir path; results checked
// against the host instantiation of fp.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo -o build/bench_lp tools/bench_lp.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../reef_b200/csrc/fp.cuh"
#include "../reef_b200/csrc/fp29.cuh"
#include "../reef_b200/csrc/lpmul.cuh"
using namespace reef;
typedef Fe<FqCfg> Fq;

// w <- w^2, iters times, one warp
__global__ void __launch_bounds__(32) k_lp_sqr_chain(const u32* in /*9*/, u32* out /*9*/, int iters, long long* cycles) {
  __shared__ __align__(16) u32 pad[LP_PAD];
  __shared__ __align__(16) u32 hbuf[12];
  const int lane = threadIdx.x;
  const LpLane c = lp_lane_consts<0>(lane);
  lp_pad_clear(pad, lane);
  __syncwarp();
  u32 limb = lane < 10 ? in[lane] : 0;
  lp_store(pad, lane, limb);
  __syncwarp();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    limb = lp_mul<false>(pad, pad, c, hbuf, lane);
    __syncwarp();
    lp_store(pad, lane, limb);
    __syncwarp();
  }
  const long long t1 = clock64();
  if (lane < 10) out[lane] = limb;
  if (lane == 0) cycles[0] = (t1 - t0) / iters;
}

// the S-box chain of a partial round: w -> w^2 -> w^4 -> w^5 (+ addend), iters times
__global__ void __launch_bounds__(32) k_lp_quintic_chain(const u32* in, u32* out, int iters, long long* cycles) {
  __shared__ __align__(16) u32 p1[LP_PAD], p2[LP_PAD], p4[LP_PAD];
  __shared__ __align__(16) u32 hbuf[12];
  const int lane = threadIdx.x;
  const LpLane c = lp_lane_consts<0>(lane);
  lp_pad_clear(p1, lane); lp_pad_clear(p2, lane); lp_pad_clear(p4, lane);
  __syncwarp();
  u32 limb = lane < 10 ? in[lane] : 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    lp_store(p1, lane, limb);
    __syncwarp();
    const u32 m2 = lp_mul<false>(p1, p1, c, hbuf, lane);
    lp_store(p2, lane, m2);
    __syncwarp();
    const u32 m4 = lp_mul<false>(p2, p2, c, hbuf, lane);
    lp_store(p4, lane, m4);
    __syncwarp();
    limb = lp_mul<false>(p4, p1, c, hbuf, lane, lane < 9 ? 3u + lane : 0u);
  }
  const long long t1 = clock64();
  if (lane < 10) out[lane] = limb;
  if (lane == 0) cycles[0] = (t1 - t0) / iters;
}

// the same chain with fold A fed through shared-memory pieces (lp_fold_sm)
__global__ void __launch_bounds__(32) k_lp_quintic_chain_sm(const u32* in, u32* out, int iters, long long* cycles) {
  __shared__ __align__(16) u32 p1[LP_PAD], p2[LP_PAD], p4[LP_PAD];
  __shared__ __align__(16) u32 pieces[72];
  const int lane = threadIdx.x;
  const LpLane c = lp_lane_consts<0>(lane);
  lp_pad_clear(p1, lane); lp_pad_clear(p2, lane); lp_pad_clear(p4, lane);
  for (int i = lane; i < 72; i += 32) pieces[i] = 0;
  __syncwarp();
  u32 limb = lane < 10 ? in[lane] : 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    lp_store(p1, lane, limb);
    __syncwarp();
    const u32 m2 = lp_fold_sm<false>(lp_cols(p1, p1, lane), c, pieces, lane, 0);
    lp_store(p2, lane, m2);
    __syncwarp();
    const u32 m4 = lp_fold_sm<false>(lp_cols(p2, p2, lane), c, pieces, lane, 0);
    lp_store(p4, lane, m4);
    __syncwarp();
    limb = lp_fold_sm<false>(lp_cols(p4, p1, lane), c, pieces, lane, lane < 9 ? 3u + lane : 0u);
  }
  const long long t1 = clock64();
  if (lane < 10) out[lane] = limb;
  if (lane == 0) cycles[0] = (t1 - t0) / iters;
}

// interference experiment: warp 0 runs the quintic chain while warps 1..7 of the same CTA
//   mode 0: exit at once   mode 1: spin on a shared flag (tight loop)   mode 2: spin with nanosleep(40)
//   mode 3: run one-thread sqr29 chains (integer pipe load)   mode 4: only warp 4 (same scheduler as warp 0?) spins
//   modes 5..9: sqr29 chains on warp 4 only / warps 1-3 / warps 5-7 / warps 2,3,6,7 / warp 1 only
__global__ void __launch_bounds__(256) k_lp_quintic_interf(const u32* in, u32* out, int iters, long long* cycles, int mode) {
  __shared__ __align__(16) u32 p1[LP_PAD], p2[LP_PAD], p4[LP_PAD];
  __shared__ __align__(16) u32 hbuf[12];
  __shared__ __align__(16) u32 q1[LP_PAD], q2[LP_PAD], q4[LP_PAD];
  __shared__ __align__(16) u32 hbuf2[12];
  __shared__ volatile u32 flag;
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) flag = 0;
  if (warp == 0) { lp_pad_clear(p1, lane); lp_pad_clear(p2, lane); lp_pad_clear(p4, lane); }
  if (warp == 1) { lp_pad_clear(q1, lane); lp_pad_clear(q2, lane); lp_pad_clear(q4, lane); }
  __syncthreads();
  if ((mode == 10 && warp == 4) || (mode == 11 && warp == 1)) {
    // a second lane-parallel chain: on warp 0's scheduler (mode 10) or on another one (mode 11)
    const LpLane c = lp_lane_consts<0>(lane);
    u32 limb = lane < 10 ? in[lane] + 1 : 0;
    while (!flag) {
      lp_store(q1, lane, limb);
      __syncwarp();
      const u32 m2 = lp_mul<false>(q1, q1, c, hbuf2, lane);
      lp_store(q2, lane, m2);
      __syncwarp();
      const u32 m4 = lp_mul<false>(q2, q2, c, hbuf2, lane);
      lp_store(q4, lane, m4);
      __syncwarp();
      limb = lp_mul<false>(q4, q1, c, hbuf2, lane, lane < 9 ? 3u + lane : 0u);
    }
    if (limb == 0xffffffffu) out[64 + lane] = limb;
  } else if (warp == 0) {
    const LpLane c = lp_lane_consts<0>(lane);
    u32 limb = lane < 10 ? in[lane] : 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
      lp_store(p1, lane, limb);
      __syncwarp();
      const u32 m2 = lp_mul<false>(p1, p1, c, hbuf, lane);
      lp_store(p2, lane, m2);
      __syncwarp();
      const u32 m4 = lp_mul<false>(p2, p2, c, hbuf, lane);
      lp_store(p4, lane, m4);
      __syncwarp();
      limb = lp_mul<false>(p4, p1, c, hbuf, lane, lane < 9 ? 3u + lane : 0u);
    }
    const long long t1 = clock64();
    if (lane < 10) out[lane] = limb;
    if (lane == 0) cycles[0] = (t1 - t0) / iters;
    __threadfence_block();
    if (lane == 0) flag = 1;
  } else if (mode == 1 || (mode == 4 && warp == 4)) {
    while (!flag) {
    }
  } else if (mode == 2) {
    while (!flag) __nanosleep(40);
  } else if (mode == 3 || (mode == 5 && warp == 4) || (mode == 6 && warp >= 1 && warp <= 3) || (mode == 7 && warp >= 5) ||
             (mode == 8 && (warp == 2 || warp == 3 || warp == 6 || warp == 7)) || (mode == 9 && warp == 1)) {
    F29 x;
    for (int k = 0; k < 9; k++) x.l[k] = in[k] + warp;
    while (!flag) x = sqr29<FqCfg>(x);
    if (x.l[0] == 0xffffffffu) out[32 + threadIdx.x] = x.l[1];
  }
}

__global__ void __launch_bounds__(32) k_sqr29_chain(const u32* in, u32* out, int iters, long long* cycles) {
  F29 x;
  for (int k = 0; k < 9; k++) x.l[k] = in[k];
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) x = sqr29<FqCfg>(x);
  const long long t1 = clock64();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 9; k++) out[k] = x.l[k];
    cycles[0] = (t1 - t0) / iters;
  }
}

static Fq host_from_limbs(const u32* l) { F29 t = lp10_to_f29<0>(l); return lp_to_canonical<FqCfg>(t.l); }

int main() {
  const int iters = 2000;
  u32 h_in[10] = {0x12345678u & M29, 0x0badf00du & M29, 0x1ee7c0deu & M29, 0x11111111u, 0x02222222u, 0x13333333u, 0x04444444u, 0x15555555u, 0x00123456u, 0u};
  u32 *d_in, *d_out;
  long long* d_cyc;
  cudaMalloc(&d_in, 64); cudaMalloc(&d_out, 4096); cudaMalloc(&d_cyc, 64);
  cudaMemcpy(d_in, h_in, 40, cudaMemcpyHostToDevice);
  u32 h_out[10];
  long long cyc;
  // expected: x^(2^iters) via the host Montgomery code
  Fq x = host_from_limbs(h_in);
  Fq xm = to_mont<FqCfg>(x);
  Fq e = xm;
  for (int i = 0; i < iters; i++) e = mont_sqr<FqCfg>(e);
  e = from_mont<FqCfg>(e);
  for (int rep = 0; rep < 2; rep++) {
    k_lp_sqr_chain<<<1, 32>>>(d_in, d_out, iters, d_cyc);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h_out, d_out, 40, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  Fq g = host_from_limbs(h_out);
  bool ok = true;
  for (int i = 0; i < 8; i++) ok = ok && g.v[i] == e.v[i];
  printf("lp squaring chain : %lld cycles per multiplication, result %s (%s)\n", cyc, ok ? "OK" : "MISMATCH", cudaGetErrorString(cudaGetLastError()));
  // quintic chain: w <- w^5 + (3 + k per column), iters times
  Fq w = xm;
  Fq add = fe_zero<FqCfg>();
  {
    u32 al[10];
    for (int k = 0; k < 9; k++) al[k] = 3u + k;
    al[9] = 0;
    add = to_mont<FqCfg>(host_from_limbs(al));
  }
  const int it2 = 500;
  for (int i = 0; i < it2; i++) {
    Fq w2 = mont_sqr<FqCfg>(w), w4 = mont_sqr<FqCfg>(w2);
    w = fe_add<FqCfg>(mont_mul<FqCfg>(w4, w), add);
  }
  w = from_mont<FqCfg>(w);
  for (int rep = 0; rep < 2; rep++) {
    k_lp_quintic_chain<<<1, 32>>>(d_in, d_out, it2, d_cyc);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h_out, d_out, 40, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  g = host_from_limbs(h_out);
  ok = true;
  for (int i = 0; i < 8; i++) ok = ok && g.v[i] == w.v[i];
  printf("lp quintic chain  : %lld cycles per round (3 multiplications + addend), result %s (%s)\n", cyc, ok ? "OK" : "MISMATCH", cudaGetErrorString(cudaGetLastError()));
  for (int rep = 0; rep < 2; rep++) {
    k_lp_quintic_chain_sm<<<1, 32>>>(d_in, d_out, it2, d_cyc);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h_out, d_out, 40, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  g = host_from_limbs(h_out);
  ok = true;
  for (int i = 0; i < 8; i++) ok = ok && g.v[i] == w.v[i];
  printf("lp quintic chain, fold A through shared pieces: %lld cycles per round, result %s (%s)\n", cyc, ok ? "OK" : "MISMATCH", cudaGetErrorString(cudaGetLastError()));
  for (int mode = 0; mode < 12; mode++) {
    k_lp_quintic_interf<<<1, 256>>>(d_in, d_out, it2, d_cyc, mode);
    cudaDeviceSynchronize();
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("lp quintic chain beside 7 other warps, mode %d: %lld cycles per round (%s)\n", mode, cyc, cudaGetErrorString(cudaGetLastError()));
  }
  k_sqr29_chain<<<1, 32>>>(d_in, d_out, iters, d_cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  printf("sqr29 (one thread): %lld cycles per squaring\n", cyc);
  return 0;
}
