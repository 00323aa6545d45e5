// Microbenchmarks of the field arithmetic on the GPU:
//   (1) single-warp dependent-chain latency (cycles per op) -- what the Fiat-Shamir sponge sees
//   (2) whole-chip throughput of independent Montgomery multiplications -- the MSM / Poseidon
//       roofline denominator ("peak_modmul_per_s").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o build/bench_fp tools/bench_fp.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../reef_b200/csrc/fp.cuh"
#include "../reef_b200/csrc/fp29.cuh"
#include "../reef_b200/csrc/ec.cuh"
using namespace reef;
typedef Fe<FqCfg> Fq;
typedef Fe<FpCfg> Fp;

template <int MODE>
__global__ void k_latency(Fq* io, int iters, long long* cycles) {
  Fq x = io[threadIdx.x], y = io[32 + threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) x = mont_mul<FqCfg>(x, y);
    if (MODE == 1) x = mont_sqr<FqCfg>(x);
    if (MODE == 2) x = fe_add<FqCfg>(x, y);
    if (MODE == 3) { u32 T[16]; mul_wide(T, x.v, y.v); for (int k = 0; k < 8; k++) x.v[k] = T[k] ^ T[8 + k]; }
    if (MODE == 5) { F29 t = mul29<FqCfg>(f29_from_words(x.v), f29_from_words(y.v)); f29_normalize(t); f29_to_words(x.v, t); x.v[7] &= 0x3fffffff; }
    if (MODE == 6) { x = fe_inv<FqCfg>(x); x = fe_add<FqCfg>(x, y); }
    if (MODE == 7) x = fe_pow_pm2<FqCfg>(x);
    if (MODE == 4) { u32 T[16]; for (int k = 0; k < 8; k++) { T[k] = x.v[k]; T[8 + k] = y.v[k]; } mont_reduce<FqCfg>(x.v, T); }
  }
  long long t1 = clock64();
  io[threadIdx.x] = x;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

// ILP independent chains per thread
template <int ILP>
__global__ void __launch_bounds__(256) k_throughput(Fp* io, int iters) {
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x[ILP], y = io[gid];
  for (int k = 0; k < ILP; k++) { x[k] = io[gid]; x[k].v[0] += k; x[k].v[7] &= 0x3fffffff; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = mont_mul<FpCfg>(x[k], y);
  }
  Fp s = x[0];
  for (int k = 1; k < ILP; k++) s = fe_add<FpCfg>(s, x[k]);
  io[gid] = s;
}

int main() {
  Fq* io; long long* cyc;
  cudaMalloc(&io, 1 << 26); cudaMalloc(&cyc, 8);
  cudaMemset(io, 0x11, 1 << 26);
  const char* names[8] = {"mont_mul", "mont_sqr", "fe_add", "mul_wide", "mont_reduce", "mul29+convert", "fe_inv(bingcd)+add", "fe_pow_pm2"};
  for (int mode = 0; mode < 8; mode++) {
    const int iters = mode >= 6 ? 20 : 2000;
    for (int rep = 0; rep < 2; rep++) {
      if (mode == 0) k_latency<0><<<1, 32>>>(io, iters, cyc);
      if (mode == 1) k_latency<1><<<1, 32>>>(io, iters, cyc);
      if (mode == 2) k_latency<2><<<1, 32>>>(io, iters, cyc);
      if (mode == 3) k_latency<3><<<1, 32>>>(io, iters, cyc);
      if (mode == 4) k_latency<4><<<1, 32>>>(io, iters, cyc);
      if (mode == 5) k_latency<5><<<1, 32>>>(io, iters, cyc);
      if (mode == 6) k_latency<6><<<1, 32>>>(io, iters, cyc);
      if (mode == 7) k_latency<7><<<1, 32>>>(io, iters, cyc);
      cudaDeviceSynchronize();
    }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("latency  %-12s %8.1f cycles/op (single warp, dependent chain)\n", names[mode], (double)h / iters);
  }
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  for (int ilp = 1; ilp <= 2; ilp++) {
    for (int bps = 1; bps <= 4; bps *= 2) {
      const int iters = 4000;
      int blocks = sms * bps;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float best = 1e30f;
      for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        if (ilp == 1) k_throughput<1><<<blocks, 256>>>((Fp*)io, iters);
        else k_throughput<2><<<blocks, 256>>>((Fp*)io, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      double muls = (double)blocks * 256 * iters * ilp;
      printf("throughput ilp=%d blocks/SM=%d threads/SM=%4d : %7.2f G modmul/s  (%.3f ms)\n", ilp, bps, bps * 256, muls / best / 1e6, best);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
