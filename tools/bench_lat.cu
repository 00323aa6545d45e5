// Instruction-level microbenchmarks behind the low-latency field multiplication
// (reef_b200/csrc/fp_lat.cuh): dependent-chain latency and single-warp issue rate of
// IMAD.WIDE.U32 (64-bit accumulate), IMAD.LO, IADD3, SHFL and REDUX on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/bench_lat tools/bench_lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef uint32_t u32;

// MODE 0: one dependent chain of mad.wide.u32 (latency)
// MODE 1: NCH independent chains (issue rate)
template <int NCH>
__global__ void k_madwide(u64* io, int iters, long long* cycles) {
  u64 acc[NCH];
  u32 a = (u32)io[threadIdx.x], b = (u32)io[32 + threadIdx.x];
#pragma unroll
  for (int k = 0; k < NCH; k++) acc[k] = io[64 + k];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
  }
  long long t1 = clock64();
  u64 s = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++) s ^= acc[k];
  io[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// operand-reuse variants: MODE 0 acc[k] += a*b[k] (a shared); 1 acc[k] += a[k]*b[k]; 2 acc[k] += a[k&3]*b[k>>2] (4x4 block);
// 3 mul.wide only (no accumulate) xor-folded
template <int NCH, int MODE>
__global__ void k_madwide_ops(u64* io, int iters, long long* cycles) {
  u64 acc[NCH];
  u32 a[NCH], b[NCH];
#pragma unroll
  for (int k = 0; k < NCH; k++) { acc[k] = io[64 + k]; a[k] = (u32)io[128 + k + threadIdx.x]; b[k] = (u32)io[256 + k + threadIdx.x]; }
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) {
      if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a[0]), "r"(b[k]));
      if (MODE == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a[k]), "r"(b[k]));
      if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a[k & 3]), "r"(b[k >> 2]));
      if (MODE == 3) { u64 t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[k]), "r"(b[k])); acc[k] ^= t; }
    }
  }
  long long t1 = clock64();
  u64 s = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++) s ^= acc[k];
  io[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int NCH>
__global__ void k_madlo(u64* io, int iters, long long* cycles) {
  u32 acc[NCH];
  u32 a = (u32)io[threadIdx.x], b = (u32)io[32 + threadIdx.x];
#pragma unroll
  for (int k = 0; k < NCH; k++) acc[k] = (u32)io[64 + k];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(a), "r"(b));
  }
  long long t1 = clock64();
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++) s ^= acc[k];
  io[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int NCH>
__global__ void k_add(u64* io, int iters, long long* cycles) {
  u32 acc[NCH];
  u32 a = (u32)io[threadIdx.x];
#pragma unroll
  for (int k = 0; k < NCH; k++) acc[k] = (u32)io[64 + k];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[k]) : "r"(a));
  }
  long long t1 = clock64();
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++) s ^= acc[k];
  io[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// carry chain: add.cc / addc.cc pairs (dependent through the carry flag)
__global__ void k_carry(u64* io, int iters, long long* cycles) {
  u32 x[8];
  u32 a = (u32)io[threadIdx.x];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = (u32)io[64 + k];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %8;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %8;\n\t"
                 "addc.cc.u32 %4, %4, %8;\n\taddc.cc.u32 %5, %5, %8;\n\taddc.cc.u32 %6, %6, %8;\n\taddc.u32 %7, %7, %0;"
                 : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]) : "r"(a));
  }
  long long t1 = clock64();
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= x[k];
  io[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int NCH>
__global__ void k_shfl(u64* io, int iters, long long* cycles) {
  u32 acc[NCH];
#pragma unroll
  for (int k = 0; k < NCH; k++) acc[k] = (u32)io[64 + k + threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) acc[k] = __shfl_sync(0xffffffffu, acc[k], (threadIdx.x + 1) & 31);
  }
  long long t1 = clock64();
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++) s ^= acc[k];
  io[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int NCH>
__global__ void k_redux(u64* io, int iters, long long* cycles) {
  u32 acc[NCH];
#pragma unroll
  for (int k = 0; k < NCH; k++) acc[k] = (u32)io[64 + k + threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) acc[k] = __reduce_add_sync(0xffffffffu, acc[k]) + threadIdx.x;
  }
  long long t1 = clock64();
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++) s ^= acc[k];
  io[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

#define RUN(name, kern, nch, threads)                                                  \
  do {                                                                                 \
    for (int rep = 0; rep < 2; rep++) { kern<<<1, threads>>>(io, iters, cyc); cudaDeviceSynchronize(); } \
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                       \
    printf("%-28s chains=%2d threads=%4d : %7.2f cycles/iter  = %6.2f cycles/instr\n", name, nch, threads, \
           (double)h / iters, (double)h / iters / nch);                                \
  } while (0)

int main() {
  u64* io; long long* cyc;
  cudaMalloc(&io, 1 << 20); cudaMalloc(&cyc, 8);
  cudaMemset(io, 0x13, 1 << 20);
  const int iters = 4000;
  RUN("mad.wide.u32 (IMAD.WIDE)", k_madwide<1>, 1, 32);
  RUN("mad.wide.u32 (IMAD.WIDE)", k_madwide<2>, 2, 32);
  RUN("mad.wide.u32 (IMAD.WIDE)", k_madwide<4>, 4, 32);
  RUN("mad.wide.u32 (IMAD.WIDE)", k_madwide<8>, 8, 32);
  RUN("mad.wide.u32 (IMAD.WIDE)", k_madwide<16>, 16, 32);
  RUN("mad.wide.u32 4 warps/SMSP", k_madwide<8>, 8, 512);
  RUN("mad.wide a shared, b[k]", (k_madwide_ops<16, 0>), 16, 32);
  RUN("mad.wide a[k], b[k]", (k_madwide_ops<16, 1>), 16, 32);
  RUN("mad.wide a[k&3], b[k>>2]", (k_madwide_ops<16, 2>), 16, 32);
  RUN("mul.wide a[k], b[k] + xor", (k_madwide_ops<16, 3>), 16, 32);
  RUN("mad.wide a[k], b[k] 4w/SMSP", (k_madwide_ops<16, 1>), 16, 512);
  RUN("mad.wide a shared 4w/SMSP", (k_madwide_ops<16, 0>), 16, 512);
  RUN("mad.lo.u32 (IMAD)", k_madlo<1>, 1, 32);
  RUN("mad.lo.u32 (IMAD)", k_madlo<8>, 8, 32);
  RUN("mad.lo.u32 (IMAD)", k_madlo<16>, 16, 32);
  RUN("mad.lo.u32 4 warps/SMSP", k_madlo<8>, 8, 512);
  RUN("add.u32 (IADD3)", k_add<1>, 1, 32);
  RUN("add.u32 (IADD3)", k_add<8>, 8, 32);
  RUN("add.u32 (IADD3)", k_add<16>, 16, 32);
  RUN("addc chain of 8", k_carry, 8, 32);
  RUN("shfl.sync", k_shfl<1>, 1, 32);
  RUN("shfl.sync", k_shfl<8>, 8, 32);
  RUN("redux.sync.add", k_redux<1>, 1, 32);
  RUN("redux.sync.add", k_redux<8>, 8, 32);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
