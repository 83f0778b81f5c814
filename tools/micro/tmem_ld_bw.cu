// Microbenchmark: tensor-memory -> register read throughput of one SM (tcgen05.ld), the roofline that binds the quantiser's
// argmax epilogue and the encoder's activation epilogues.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tmem_ld_bw.cu
//   ./tmem_ld_bw            prints cycles and bytes/cycle/SM for 4 / 8 / 16 warps, plain and .pack::16b loads
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 64 columns of 16-bit data (low halves) packed two per register
__device__ __forceinline__ void ld_x32_pack(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

template <bool PACK>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* out_cycles, uint32_t* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[32];
    const uint32_t col = (uint32_t)((i * (PACK ? 64 : 32) + (warp >> 2) * 64) & 447);
    if (PACK) ld_x32_pack(base + col, v); else ld_x32(base + col, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= v[j];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
}

int main() {
  long long* d_c; uint32_t* d_s;
  cudaMalloc(&d_c, 1024 * sizeof(long long)); cudaMalloc(&d_s, 4);
  const int iters = 4096;
  for (int pack = 0; pack < 2; ++pack)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (pack) k<true><<<148, warps * 32>>>(iters, d_c, d_s); else k<false><<<148, warps * 32>>>(iters, d_c, d_s);
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, d_c, sizeof(c), cudaMemcpyDeviceToHost);
      const double regs_bytes = (double)warps * iters * 32 * 32 * 4;           // bytes delivered to registers per SM
      const double cells = (double)warps * iters * 32 * (pack ? 64 : 32) * 4;  // TMEM cells (32-bit) read per SM, in bytes
      printf("%s warps %2d: %lld cycles, %.1f register bytes/cycle/SM, %.1f TMEM cell bytes/cycle/SM  (%s)\n", pack ? "pack::16b" : "plain    ",
             warps, c, regs_bytes / c, cells / c, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
