// Development aid (not part of the product): what does an extra instruction cost next to the IMAD.WIDE
// chains the field multiplier is made of?  Each kernel runs the chained mad.wide loop of k_peak_imad_wide
// (64 IMAD.WIDE.U32.X per iteration) plus K extra instructions of one kind per 8 IMAD.WIDE, at the occupancy
// of k_var_base (6 blocks of 128 threads per SM).  The extra instructions form their own dependency chains
// (4 independent accumulators) so that they are neither eliminated nor on the multiply chains' critical path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_model pipe_model.cu && ./pipe_model
// Output: cycles per warp per group of 8 IMAD.WIDE for every (kind, K); the slope over K is the marginal cost
// of one instruction of that kind.  Check the SASS (cuobjdump -sass) for what ptxas made of each kind.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32;

#define PKA(lo, hi, m) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(m), "r"(b));
#define PKB(lo, hi, m) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(m), "r"(b));
#define CHAIN4A PKA(l0, h0, l1) PKB(l2, h2, l3) PKB(l4, h4, l5) PKB(l6, h6, l7)
#define CHAIN4B PKA(l1, h1, l2) PKB(l3, h3, l4) PKB(l5, h5, l6) PKB(l7, h7, l0)

// kinds: 0 none, 1 add.u32 (IADD3), 2 mad.lo.u32 (IMAD), 3 xor (LOP3), 4 shf (SHF), 5 addc chain (IADD3.X), 6 prmt (PRMT / mov-like)
template <int KIND> __device__ __forceinline__ void extra(u32 &x, u32 y) {
  if (KIND == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
  if (KIND == 2) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(y));
  if (KIND == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(y), "r"(0x5A5A5A5Au));
  if (KIND == 4) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x) : "r"(y));
  if (KIND == 5) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
  if (KIND == 6) asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(x) : "r"(y));
}
template <int KIND, int K> __device__ __forceinline__ void extras(u32 &e0, u32 &e1, u32 &e2, u32 &e3, u32 y) {
  // e0 op= e1, e1 op= e2, e2 op= e3, e3 op= e0: values keep changing, nothing folds into a constant
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if ((k & 3) == 0) extra<KIND>(e0, e1);
    if ((k & 3) == 1) extra<KIND>(e1, e2);
    if ((k & 3) == 2) extra<KIND>(e2, e3);
    if ((k & 3) == 3) extra<KIND>(e3, e0);
  }
}

template <int KIND, int K>
__global__ void __launch_bounds__(128, 6) k_mix(u32 *sink, int iters, u32 b) {
  u32 l0 = threadIdx.x * 2654435761u + 1u, l1 = l0 * 3u + 1u, l2 = l1 * 3u + 1u, l3 = l2 * 3u + 1u, l4 = l3 * 3u + 1u,
      l5 = l4 * 3u + 1u, l6 = l5 * 3u + 1u, l7 = l6 * 3u + 1u;
  u32 h0 = 0, h1 = 1, h2 = 2, h3 = 3, h4 = 4, h5 = 5, h6 = 6, h7 = 7;
  u32 e0 = threadIdx.x, e1 = e0 + 11, e2 = e0 + 22, e3 = e0 + 33, y = b | 5u;
  b |= 0x80000001u;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      // the extras go between the two 4-multiply chains so that they never sit inside a carry chain
      CHAIN4A
      extras<KIND, K / 2>(e0, e1, e2, e3, y);
      CHAIN4B
      extras<KIND, K - K / 2>(e0, e1, e2, e3, y);
    }
  }
  u32 s = l0 ^ l1 ^ l2 ^ l3 ^ l4 ^ l5 ^ l6 ^ l7 ^ h0 ^ h1 ^ h2 ^ h3 ^ h4 ^ h5 ^ h6 ^ h7 ^ e0 ^ e1 ^ e2 ^ e3;
  if (s == 0x12345u) sink[0] = s;
}

template <int KIND, int K> static void run(const char *name, u32 *sink, double clk_ghz) {
  const int blocks = 148 * 6, threads = 128, iters = 2048;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k_mix<KIND, K><<<blocks, threads>>>(sink, iters, 3u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  // per SM sub-partition: 6 warps resident... every warp runs iters * 8 groups of 8 IMAD.WIDE; 6 warps share one scheduler
  double warps_per_smsp = 6.0 * 4 / 4;
  double groups = (double)iters * 8;
  double cyc_per_group = best * 1e-3 * clk_ghz * 1e9 / (groups * warps_per_smsp);
  printf("%-10s K=%2d  %8.3f ms  %7.2f cycles per 8 IMAD.WIDE (+%d extra)  -> %5.2f per IMAD.WIDE\n", name, K, best, cyc_per_group, K, cyc_per_group / 8);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

int main() {
  u32 *sink;
  cudaMalloc(&sink, 1024);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz * 1e-6;
  printf("SM clock (attr) %.3f GHz\n", ghz);
  run<0, 0>("none", sink, ghz);
#define ALLK(kind, name) run<kind, 2>(name, sink, ghz); run<kind, 4>(name, sink, ghz); run<kind, 8>(name, sink, ghz); run<kind, 16>(name, sink, ghz); run<kind, 24>(name, sink, ghz);
  ALLK(1, "add")
  ALLK(2, "mad.lo")
  ALLK(3, "xor")
  ALLK(4, "shf")
  ALLK(5, "addc.cc")
  ALLK(6, "prmt")
  cudaFree(sink);
  return 0;
}
