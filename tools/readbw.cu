// Read-only streaming micro-benchmark: what does HBM deliver for the access patterns the scan kernels use?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/readbw tools/readbw.cu ; tools/readbw [GiB]
// (a measuring tool, not part of the product library)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t fold(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }

// 1. grid-stride, every warp touches 512 contiguous bytes per step, consecutive warps consecutive addresses
__global__ void k_gridstride(const uint4* p, size_t n16, uint32_t* out) {
  uint32_t acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) acc ^= fold(ldg_stream16(p + i));
  if (acc == 0x12345678u) *out = acc;
}

// 2. persistent: every warp owns chunks of `chunk` bytes (a "group"), taken from an atomic counter, and walks a chunk
//    512 B per step with U rows in flight
template <int U>
__global__ void __launch_bounds__(1024, 1) k_warpstream(const uint8_t* p, size_t bytes, size_t chunk, unsigned long long* counter, uint32_t* out) {
  const int lane = threadIdx.x & 31;
  uint32_t acc = 0;
  const size_t n_chunks = bytes / chunk;
  for (;;) {
    unsigned long long g = 0;
    if (lane == 0) g = atomicAdd(counter, 1ull);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= n_chunks) break;
    const uint8_t* q = p + g * chunk + lane * 16;
    const size_t rows = chunk / 512;
    uint4 buf[U];
#pragma unroll
    for (int u = 0; u < U; ++u) buf[u] = ldg_stream16(q + (size_t)u * 512);
    size_t r = 0;
    for (; r + 2 * U <= rows; r += U) {
#pragma unroll
      for (int u = 0; u < U; ++u) { acc ^= fold(buf[u]); buf[u] = ldg_stream16(q + (r + U + u) * 512); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= fold(buf[u]);
  }
  if (acc == 0x12345678u) *out = acc;
}

template <typename F>
static void timeit(const char* name, size_t bytes, F launch) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int it = 0; it < 6; ++it) {
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (it && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  printf("%-44s %8.3f ms  %7.1f GB/s %s\n", name, best, bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char** argv) {
  size_t gib = argc > 1 ? atoi(argv[1]) : 16;
  size_t bytes = gib << 30;
  uint8_t* p; cudaMalloc(&p, bytes); cudaMemset(p, 1, bytes);
  uint32_t* out; cudaMalloc(&out, 4);
  unsigned long long* ctr; cudaMalloc(&ctr, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int mult : {8, 16, 32}) for (int bs : {256, 512, 1024}) {
    char nm[96]; snprintf(nm, sizeof nm, "gridstride grid=%dxSM block=%d", mult, bs);
    timeit(nm, bytes, [&] { k_gridstride<<<sms * mult, bs>>>((const uint4*)p, bytes / 16, out); });
  }
  for (size_t chunk : {(size_t)32 << 10, (size_t)128 << 10, (size_t)1 << 20}) {
    char nm[96];
    snprintf(nm, sizeof nm, "warpstream U=1 chunk=%zuK", chunk >> 10); timeit(nm, bytes, [&] { cudaMemsetAsync(ctr, 0, 8); k_warpstream<1><<<sms, 1024>>>(p, bytes, chunk, ctr, out); });
    snprintf(nm, sizeof nm, "warpstream U=2 chunk=%zuK", chunk >> 10); timeit(nm, bytes, [&] { cudaMemsetAsync(ctr, 0, 8); k_warpstream<2><<<sms, 1024>>>(p, bytes, chunk, ctr, out); });
    snprintf(nm, sizeof nm, "warpstream U=4 chunk=%zuK", chunk >> 10); timeit(nm, bytes, [&] { cudaMemsetAsync(ctr, 0, 8); k_warpstream<4><<<sms, 1024>>>(p, bytes, chunk, ctr, out); });
    snprintf(nm, sizeof nm, "warpstream U=8 chunk=%zuK", chunk >> 10); timeit(nm, bytes, [&] { cudaMemsetAsync(ctr, 0, 8); k_warpstream<8><<<sms, 1024>>>(p, bytes, chunk, ctr, out); });
  }
  return 0;
}
