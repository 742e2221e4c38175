// INT/DPX issue-rate microbenchmark for sm_100a (B200).
//
// MEASURED_PEAKS.json (driver-written) only carries HBM GB/s and bf16 TFLOP/s; the
// Smith-Waterman extension kernel is bound by the integer/DPX issue rate instead
// (SURVEY.md §8d), so this program measures that denominator: warp-instructions
// issued per clock per SM for the packed-int16 DPX ops the kernel is made of, and
// their mixes with the FMA-pipe integer ops and shared-memory loads.
//
// Output: one JSON object on stdout.  Rates are lane-ops per clock per SM
// (32 x warp instructions / elapsed SM clocks), measured with clock64() inside the
// kernel (max over CTAs), and Gops/s from CUDA events.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

enum Op { VIADDMNMX16 = 0, VIMNMX3_16, VIMNMX16, VIADD16, IADD32, IMAD32, LOP3, PRMT_OP,
          MIX_DPX_IMAD, MIX_DPX_LDS, VIADDMNMX32, MIX_SWCELL, HMNMX2_OP, MIX_DPX_HMNMX2, HADD2_OP, MIX_DPX_HADD2, NUM_OPS };

static const char *kNames[NUM_OPS] = {
  "viaddmnmx_s16x2", "vimnmx3_s16x2", "vimnmx_s16x2", "viadd_16x2", "iadd3_s32", "imad_s32",
  "lop3", "prmt", "mix_viaddmnmx16_imad_1to1", "mix_viaddmnmx16_lds_4to1", "viaddmnmx_s32",
  "mix_swcell_5dpx_1prmt", "hmnmx2", "mix_viaddmnmx16_hmnmx2_1to1", "hadd2", "mix_viaddmnmx16_hadd2_1to1" };
// instructions counted per chain step
static const int kInstrPerStep[NUM_OPS] = {1, 1, 1, 1, 1, 1, 1, 1, 2, 5, 1, 7, 1, 2, 1, 2};

template <int OP>
__global__ void __launch_bounds__(1024) rate_kernel(unsigned *out, unsigned a, unsigned b, unsigned c,
                                                    long long *cycles) {
  __shared__ unsigned tab[1024];
  tab[threadIdx.x & 1023] = threadIdx.x * a;
  __syncthreads();
  unsigned acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * 2654435761u + i * b;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == VIADDMNMX16) acc[i] = __viaddmax_s16x2(acc[i], a, b);
      else if (OP == VIMNMX3_16) acc[i] = __vimax3_s16x2(acc[i], a, acc[(i + 1) % CHAINS] ^ c);
      else if (OP == VIMNMX16) acc[i] = __vmaxs2(acc[i] ^ 0, acc[(i + 3) % CHAINS]) + 0;
      else if (OP == VIADD16) acc[i] = __vadd2(acc[i], a);
      else if (OP == IADD32) asm volatile("add.s32 %0, %0, %1;" : "+r"(acc[i]) : "r"(a));
      else if (OP == IMAD32) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b));
      else if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[i]) : "r"(a), "r"(b));
      else if (OP == PRMT_OP) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(c));
      else if (OP == VIADDMNMX32) acc[i] = (unsigned)__viaddmax_s32((int)acc[i], (int)a, (int)b);
      else if (OP == MIX_DPX_IMAD) {
        acc[i] = __viaddmax_s16x2(acc[i], a, b);
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(acc[(i + 4) % CHAINS]) : "r"(a), "r"(b));
      } else if (OP == MIX_DPX_LDS) {
        unsigned v = tab[(acc[i] >> 3) & 1023];
        acc[i] = __viaddmax_s16x2(acc[i], a, v);
        acc[i] = __viaddmax_s16x2(acc[i], b, c);
        acc[i] = __viaddmax_s16x2(acc[i], c, a);
        acc[i] = __viaddmax_s16x2(acc[i], a, b);
      } else if (OP == HMNMX2_OP) {
        __half2 x = *reinterpret_cast<__half2 *>(&acc[i]), y = *reinterpret_cast<__half2 *>(&acc[(i + 3) % CHAINS]);
        __half2 r = __hmax2(x, y);
        acc[i] = *reinterpret_cast<unsigned *>(&r) + 1u;
      } else if (OP == MIX_DPX_HMNMX2) {
        acc[i] = __viaddmax_s16x2(acc[i], a, b);
        __half2 x = *reinterpret_cast<__half2 *>(&acc[(i + 4) % CHAINS]), y = *reinterpret_cast<__half2 *>(&c);
        __half2 r = __hmax2(x, y);
        acc[(i + 4) % CHAINS] = *reinterpret_cast<unsigned *>(&r);
      } else if (OP == HADD2_OP) {
        __half2 x = *reinterpret_cast<__half2 *>(&acc[i]), y = *reinterpret_cast<__half2 *>(&a);
        __half2 r = __hadd2(x, y);
        acc[i] = *reinterpret_cast<unsigned *>(&r);
      } else if (OP == MIX_DPX_HADD2) {
        acc[i] = __viaddmax_s16x2(acc[i], a, b);
        __half2 x = *reinterpret_cast<__half2 *>(&acc[(i + 4) % CHAINS]), y = *reinterpret_cast<__half2 *>(&c);
        __half2 r = __hadd2(x, y);
        acc[(i + 4) % CHAINS] = *reinterpret_cast<unsigned *>(&r);
      } else if (OP == MIX_SWCELL) {
        // the dependency shape of one packed SW cell: E, a, m, F(chain), H, colmax + 1 PRMT
        unsigned e = __viaddmax_s16x2(acc[i], a, acc[(i + 1) % CHAINS]);
        unsigned t = __viaddmax_s16x2(e, b, 0u);
        unsigned s; asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(s) : "r"(acc[(i + 2) % CHAINS]), "r"(c));
        unsigned m = __viaddmax_s16x2(acc[(i + 3) % CHAINS], s, t);
        unsigned f = __viaddmax_s16x2(acc[(i + 5) % CHAINS], a, m);
        unsigned h = __viaddmax_s16x2(f, b, m);
        acc[i] = __vmaxs2(h, e);
      }
    }
  }
  long long t1 = clock64();
  unsigned r = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) r ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(int sms, int threads, int ctas_per_sm, unsigned *d_out, long long *d_cyc, double clock_ghz_hint) {
  int grid = sms * ctas_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) rate_kernel<OP><<<grid, threads>>>(d_out, 3u + w, 0x00050007u, 0x3210u, d_cyc);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  rate_kernel<OP><<<grid, threads>>>(d_out, 3u, 0x00050007u, 0x3210u, d_cyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0; for (long long v : cyc) mx = v > mx ? v : mx;
  double lane_ops_per_sm = (double)ITERS * CHAINS * kInstrPerStep[OP] * threads * ctas_per_sm;
  double per_clk = lane_ops_per_sm / (double)mx;
  double gops = lane_ops_per_sm * sms / (ms * 1e-3) / 1e9;
  printf("  \"%s@%d\": {\"lane_ops_per_clk_per_sm\": %.2f, \"giga_lane_ops_per_s\": %.1f, \"ms\": %.4f, "
         "\"threads\": %d, \"ctas_per_sm\": %d, \"eff_clock_ghz\": %.3f},\n",
         kNames[OP], threads, per_clk, gops, ms, threads, ctas_per_sm, (double)mx / (ms * 1e-3) / 1e9);
  (void)clock_ghz_hint;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  unsigned *d_out; long long *d_cyc;
  CK(cudaMalloc(&d_out, (size_t)sms * 2 * 1024 * sizeof(unsigned)));
  CK(cudaMalloc(&d_cyc, (size_t)sms * 2 * sizeof(long long)));
  printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"clock_khz_prop\": %d,\n", prop.name, sms, prop.clockRate);
  const int T = 1024, C = 1;
  run<VIADDMNMX16>(sms, T, C, d_out, d_cyc, 0);
  run<VIMNMX3_16>(sms, T, C, d_out, d_cyc, 0);
  run<VIMNMX16>(sms, T, C, d_out, d_cyc, 0);
  run<VIADD16>(sms, T, C, d_out, d_cyc, 0);
  run<VIADDMNMX32>(sms, T, C, d_out, d_cyc, 0);
  run<IADD32>(sms, T, C, d_out, d_cyc, 0);
  run<IMAD32>(sms, T, C, d_out, d_cyc, 0);
  run<LOP3>(sms, T, C, d_out, d_cyc, 0);
  run<PRMT_OP>(sms, T, C, d_out, d_cyc, 0);
  run<MIX_DPX_IMAD>(sms, T, C, d_out, d_cyc, 0);
  run<MIX_DPX_LDS>(sms, T, C, d_out, d_cyc, 0);
  run<MIX_SWCELL>(sms, T, C, d_out, d_cyc, 0);
  run<HMNMX2_OP>(sms, T, C, d_out, d_cyc, 0);
  run<MIX_DPX_HMNMX2>(sms, T, C, d_out, d_cyc, 0);
  run<HADD2_OP>(sms, T, C, d_out, d_cyc, 0);
  run<MIX_DPX_HADD2>(sms, T, C, d_out, d_cyc, 0);
  // low-occupancy points (what a 200-register kernel gets): 256 threads/SM
  run<VIADDMNMX16>(sms, 256, 1, d_out, d_cyc, 0);
  run<MIX_SWCELL>(sms, 256, 1, d_out, d_cyc, 0);
  printf("  \"iters\": %d, \"chains\": %d\n}\n", ITERS, CHAINS);
  return 0;
}
