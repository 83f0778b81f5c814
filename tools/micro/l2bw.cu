// Microbenchmark: L2 -> SM bandwidth of 1-D bulk copies (cp.async.bulk) on B200, as the fused encoder would stream its
// weight images.  mode 0: every CTA streams the SAME `bytes` buffer (weights);  mode 1: every CTA streams its own slice
// of a `bytes * grid` buffer (HBM/L2 stream).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
constexpr int kStages = 6, kChunk = 32768;
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, size_t bytes, int iters, int mode, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[kStages];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* base = mode == 0 ? src : src + (size_t)blockIdx.x * bytes;
    const int nchunk = (int)(bytes / kChunk);
    const long total = (long)nchunk * iters;
    long issued = 0;
    for (; issued < kStages && issued < total; ++issued) {
      mbar_expect(smem_u32(&bars[issued % kStages]), kChunk);
      bulk_g2s(smem_u32(smem + (issued % kStages) * kChunk), base + (size_t)((issued + blockIdx.x * 7) % nchunk) * kChunk, kChunk, smem_u32(&bars[issued % kStages]));
    }
    unsigned long long acc = 0;
    for (long c = 0; c < total; ++c) {
      const int s = c % kStages;
      mbar_wait(smem_u32(&bars[s]), (c / kStages) & 1);
      acc += *reinterpret_cast<volatile uint32_t*>(smem + s * kChunk);
      if (issued < total) {
        mbar_expect(smem_u32(&bars[s]), kChunk);
        bulk_g2s(smem_u32(smem + s * kChunk), base + (size_t)((issued + blockIdx.x * 7) % nchunk) * kChunk, kChunk, smem_u32(&bars[s]));
        ++issued;
      }
    }
    sink[blockIdx.x] = acc;
  }
}
int main(int argc, char** argv) {
  int dev = 0;
  cudaSetDevice(dev);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, dev);
  const int grid = p.multiProcessorCount;
  unsigned long long* sink;
  cudaMalloc(&sink, grid * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kChunk);
  struct { const char* name; int mode; size_t bytes; int iters; } cases[] = {
      {"same 1 MiB buffer, all CTAs (weights)", 0, 1u << 20, 400},
      {"same 8 MiB buffer, all CTAs", 0, 8u << 20, 50},
      {"own 256 KiB slice per CTA, re-read (37 MiB footprint, L2 resident)", 1, 256u << 10, 1600},
      {"own 64 MiB slice per CTA, once (9.25 GiB, HBM stream)", 1, 64u << 20, 1},
  };
  for (auto& c : cases) {
    const size_t total = c.mode == 0 ? c.bytes : c.bytes * grid;
    uint8_t* buf;
    if (cudaMalloc(&buf, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 1, total);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(a);
      k<<<grid, 128, kStages * kChunk>>>(buf, c.bytes, c.iters, c.mode, sink);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (rep > 0 && ms < best) best = ms;
    }
    const double moved = (double)c.bytes * c.iters * grid;
    printf("L2BW %-70s %8.3f ms  %8.1f GB/s  (%.1f B/clk/SM at 1.965 GHz)\n", c.name, best, moved / best / 1e6, moved / best / 1e6 / grid / 1.965);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
    cudaFree(buf);
  }
  return 0;
}
