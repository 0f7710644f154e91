// Measured denominators for the rooflines that are not HBM: the integer-issue peak of the ALU pipe (LOP3 / SHF / IADD3),
// which bounds k_sha256_validate ("the fraction of a measured INT32 issue-rate microbenchmark").
#include "common.h"

namespace fei {

// 8 independent chains per thread, each step one LOP3 (3-input xor) and one SHF (rotate): the two instruction kinds SHA-256's
// sigma / ch / maj are made of.  kSteps * 8 * 2 ALU-pipe instructions per loop iteration, nothing else but the loop branch.
constexpr int kAluSteps = 32;
constexpr int kAluInstrPerIter = kAluSteps * 8 * 2;

__global__ void __launch_bounds__(256) k_alu_peak(uint32_t* __restrict__ out, uint32_t seed, int iters) {
  uint32_t x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = seed * (2 * k + 1) + threadIdx.x;
  uint32_t y = seed ^ 0x9E3779B9u, z = seed + blockIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int st = 0; st < kAluSteps; ++st) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(y), "r"(z));
        asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[k]));
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc ^= x[k];
  if (acc == 0x12345678u) out[0] = acc;                     // keeps the chains alive; practically never taken
}

}  // namespace fei

using namespace fei;

/* Integer-issue peak: tera lane-operations per second of LOP3 / SHF over the whole chip (best of `reps`), and the
 * instructions per thread the figure is computed from.  instr_per_sm_clk = lane-ops per SM per clock at `sm_mhz`. */
extern "C" int fei_microbench_alu(int reps, float* tera_lane_ops, float* ms_out) {
  FEI_TRY(require_ready());
  Context& cx = ctx();
  cudaStream_t s = cx.stream;
  DevBuf out;
  FEI_TRY(out.alloc(64));
  const int iters = 256;
  const unsigned grid = (unsigned)cx.sm_count * 8;
  cudaEvent_t e0, e1;
  FEI_CUDA(cudaEventCreate(&e0)); FEI_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps + 2; ++r) {
    FEI_CUDA(cudaEventRecord(e0, s));
    k_alu_peak<<<grid, 256, 0, s>>>(out.as<uint32_t>(), 0xC0FFEEu + r, iters);
    FEI_CUDA(cudaEventRecord(e1, s));
    FEI_CUDA(cudaStreamSynchronize(s));
    float ms = 0; FEI_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 2 && ms < best) best = ms;
  }
  FEI_CUDA(cudaGetLastError());
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  const double ops = (double)grid * 256.0 * (double)iters * kAluInstrPerIter;
  if (tera_lane_ops) *tera_lane_ops = (float)(ops / (best * 1e-3) / 1e12);
  if (ms_out) *ms_out = best;
  return FEI_OK;
}
