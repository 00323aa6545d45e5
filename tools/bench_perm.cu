// Latency of ONE warp-cooperative Poseidon permutation (the Fiat-Shamir critical path):
// cold first call vs steady state, measured with clock64 inside a single-warp kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o build/bench_perm tools/bench_perm.cu
#define REEF_PERM_TIMING 1
#include "../reef_b200/csrc/poseidon.cu"
namespace reef {
std::atomic<unsigned long long> g_launches{0};
void set_error(const std::string&) {}
int fail(int code, const std::string& m) { fprintf(stderr, "%s\n", m.c_str()); return code; }
int ctx_scratch(reef_ctx*, size_t, void**) { return 1; }
int ctx_scratch2(reef_ctx*, size_t, void**) { return 1; }
int ctx_stage(reef_ctx*, size_t, void**) { return 1; }
}
using namespace reef;

__global__ void __launch_bounds__(32) k_perm_lat(Fq* io, int iters, const PoseidonTables* K, long long* cycles) {
  const int lane = threadIdx.x;
  Fq s = io[lane & 7];
  long long t0 = clock64();
  poseidon_permute_warp5(s, K);
  long long t1 = clock64();
  for (int i = 0; i < iters; i++) poseidon_permute_warp5(s, K);
  long long t2 = clock64();
  if (lane < 5) io[lane] = s;
  if (lane == 0) { cycles[0] = t1 - t0; cycles[1] = (t2 - t1) / (iters > 0 ? iters : 1); }
}

// two-warp permutation (warp 0 owns the state, warp 1 carries the side lanes)
__global__ void __launch_bounds__(64) k_perm_pair_lat(Fq* io, int iters, const PoseidonTables* K, long long* cycles) {
  const int lane = threadIdx.x & 31;
  Fq s = io[lane & 7];
  long long t0 = clock64();
  poseidon_permute_pair(s, K);
  long long t1 = clock64();
  for (int i = 0; i < iters; i++) poseidon_permute_pair(s, K);
  long long t2 = clock64();
  if (threadIdx.x < 5) io[16 + lane] = s;
  if (threadIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = (t2 - t1) / (iters > 0 ? iters : 1); }
}

__global__ void __launch_bounds__(32) k_sqr29_lat(Fq* io, int iters, long long* cycles) {
  F29 x = f29_from_words(io[threadIdx.x & 7].v);
  x.l[8] &= 0xffff;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) x = sqr29<FqCfg>(x);
  long long t1 = clock64();
  for (int k = 0; k < 8; k++) io[threadIdx.x].v[k] = x.l[k];
  if (threadIdx.x == 0) cycles[0] = (t1 - t0) / iters;
}

__global__ void __launch_bounds__(32) k_mul29_lat(Fq* io, int iters, long long* cycles) {
  F29 x = f29_from_words(io[threadIdx.x & 7].v), y = f29_from_words(io[8].v);
  x.l[8] &= 0xffff; y.l[8] &= 0xffff;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) x = mul29<FqCfg>(x, y);
  long long t1 = clock64();
  for (int k = 0; k < 8; k++) io[threadIdx.x].v[k] = x.l[k];
  if (threadIdx.x == 0) cycles[0] = (t1 - t0) / iters;
}

// product phase only: 81 IMAD.WIDE into 17 independent columns, folded back cheaply
__global__ void __launch_bounds__(32) k_cols_lat(Fq* io, int iters, long long* cycles) {
  F29 x = f29_from_words(io[threadIdx.x & 7].v), y = f29_from_words(io[8].v);
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    u64 col[18];
    mul29_cols<false>(col, x, y);
#pragma unroll
    for (int k = 0; k < 9; k++) x.l[k] = ((u32)col[k] ^ (u32)(col[k + 8] >> 29)) & M29;
  }
  long long t1 = clock64();
  for (int k = 0; k < 8; k++) io[threadIdx.x].v[k] = x.l[k];
  if (threadIdx.x == 0) cycles[0] = (t1 - t0) / iters;
}

// reduction only
__global__ void __launch_bounds__(32) k_redc_lat(Fq* io, int iters, long long* cycles) {
  F29 x = f29_from_words(io[threadIdx.x & 7].v);
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    u64 col[18];
#pragma unroll
    for (int k = 0; k < 17; k++) col[k] = ((u64)x.l[k % 9] << 20) + x.l[(k + 3) % 9];
    col[17] = 0;
    x = redc29<FqCfg>(col);
  }
  long long t1 = clock64();
  for (int k = 0; k < 8; k++) io[threadIdx.x].v[k] = x.l[k];
  if (threadIdx.x == 0) cycles[0] = (t1 - t0) / iters;
}

// 8 warps per SMSP running independent mul29 chains: throughput per SM
__global__ void __launch_bounds__(1024) k_mul29_tput(Fq* io, int iters) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  F29 x = f29_from_words(io[gid & 1023].v), y = f29_from_words(io[(gid + 5) & 1023].v);
  x.l[8] &= 0xffff; y.l[8] &= 0xffff;
  for (int i = 0; i < iters; i++) x = mul29<FqCfg>(x, y);
  for (int k = 0; k < 8; k++) io[gid].v[k] = x.l[k];
}

int main() {
  PoseidonTables* h = new PoseidonTables;
  poseidon_tables_host(h);
  PoseidonTables* d; cudaMalloc(&d, sizeof(PoseidonTables));
  cudaMemcpy(d, h, sizeof(PoseidonTables), cudaMemcpyHostToDevice);
  Fq* io; cudaMalloc(&io, 1 << 16); cudaMemset(io, 0x05, 1 << 16);
  long long* cyc; cudaMalloc(&cyc, 64);
  long long hc[2];
  for (int rep = 0; rep < 3; rep++) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_perm_lat<<<1, 32>>>(io, 20, d, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
    printf("permutation: first call %lld cycles, steady %lld cycles/perm (kernel of 21 perms: %.1f us)\n", hc[0], hc[1], ms * 1e3);
  }
  for (int rep = 0; rep < 3; rep++) {
    k_perm_pair_lat<<<1, 64>>>(io, 20, d, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
    long long tp[4]; cudaMemcpyFromSymbol(tp, reef::reef_perm_timing, sizeof(tp));
    printf("two-warp permutation: first call %lld cycles, steady %lld cycles/perm (56 partial rounds on warp A: %lld cycles = %lld per round)\n", hc[0], hc[1], tp[0], tp[0] / 56);
  }
  {
    // same input through both implementations must give the same state
    Fq h1[5], h2[5];
    cudaMemset(io, 0x05, 1 << 16);
    k_perm_lat<<<1, 32>>>(io, 0, d, cyc); cudaDeviceSynchronize();
    cudaMemcpy(h1, io, sizeof(h1), cudaMemcpyDeviceToHost);
    cudaMemset(io, 0x05, 1 << 16);
    k_perm_pair_lat<<<1, 64>>>(io, 0, d, cyc); cudaDeviceSynchronize();
    cudaMemcpy(h2, io + 16, sizeof(h2), cudaMemcpyDeviceToHost);
    printf("single-warp vs two-warp state: %s\n", memcmp(h1, h2, sizeof(h1)) == 0 ? "IDENTICAL" : "DIFFERENT");
  }
  k_sqr29_lat<<<1, 32>>>(io, 2000, cyc); cudaDeviceSynchronize();
  k_sqr29_lat<<<1, 32>>>(io, 2000, cyc); cudaDeviceSynchronize();
  cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost);
  printf("sqr29 dependent latency: %lld cycles/op\n", hc[0]);
  for (int rep = 0; rep < 3; rep++) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_perm_lat<<<1, 32>>>(io, 0, d, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
    printf("single-perm kernel: %lld cycles inside, %.1f us by events\n", hc[0], ms * 1e3);
  }
  k_mul29_lat<<<1, 32>>>(io, 2000, cyc); cudaDeviceSynchronize();
  k_mul29_lat<<<1, 32>>>(io, 2000, cyc); cudaDeviceSynchronize();
  cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mul29 dependent latency: %lld cycles/op\n", hc[0]);
  k_cols_lat<<<1, 32>>>(io, 2000, cyc); cudaDeviceSynchronize();
  cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mul29_cols (product phase only): %lld cycles/op\n", hc[0]);
  k_redc_lat<<<1, 32>>>(io, 2000, cyc); cudaDeviceSynchronize();
  cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost);
  printf("redc29 (reduction only): %lld cycles/op\n", hc[0]);
  {
    Fq* big; cudaMalloc(&big, 148 * 1024 * 32 + 65536); cudaMemset(big, 0x07, 148 * 1024 * 32);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      k_mul29_tput<<<148, 1024>>>(big, 2000);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("mul29 throughput, 148 x 1024 threads: %.2f G modmul/s\n", 148.0 * 1024 * 2000 / ms / 1e6);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
