// Microbenchmark: tcgen05.ld throughput of one SM by load shape (32x32b .x8 / .x16 / .x32), warps per SM and software
// pipelining (one load in flight while the previous one is consumed), and the issue rate of the 2- and 3-input fp32 maxima
// the argmax scan is made of.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tmem_ld_shapes.cu -o tmem_ld_shapes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&v)[X]);
template <>
__device__ __forceinline__ void ld<8>(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
template <int X>
__device__ __forceinline__ void wait(uint32_t (&v)[X]) {  // ties the registers to the wait
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < X; ++j) asm volatile("" : "+r"(v[j]));
}

template <int X, bool PIPE>
__global__ void __launch_bounds__(1024, 1) k_ld(int iters, long long* out_cycles, uint32_t* sink) {
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
  uint32_t v[2][X];
  const long long t0 = clock64();
  if (PIPE) {
    ld<X>(base + ((warp >> 2) * 64 & 447), v[0]);
    wait<X>(v[0]);
#pragma unroll 2
    for (int i = 0; i < iters; i += 2) {
      ld<X>(base + (((i + 1) * X + (warp >> 2) * 64) & 447), v[1]);
#pragma unroll
      for (int j = 0; j < X; ++j) acc = max(acc, v[0][j]);
      wait<X>(v[1]);
      ld<X>(base + (((i + 2) * X + (warp >> 2) * 64) & 447), v[0]);
#pragma unroll
      for (int j = 0; j < X; ++j) acc = max(acc, v[1][j]);
      wait<X>(v[0]);
    }
  } else {
    for (int i = 0; i < iters; ++i) {
      ld<X>(base + ((i * X + (warp >> 2) * 64) & 447), v[0]);
      wait<X>(v[0]);
#pragma unroll
      for (int j = 0; j < X; ++j) acc = max(acc, v[0][j]);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
}

// 16 independent chains of fp32 maxima per thread: MODE 2 = max.f32 a, b; MODE 3 = max.f32 a, b, c; MODE 0 = fma (the FMA pipe, for scale)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_alu(int iters, long long* out_cycles, float* sink, float seed) {
  float g[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) g[j] = seed * (threadIdx.x + j);
  float a = seed * 3.f, b = seed * 5.f;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MODE == 2) g[j] = fmaxf(g[j], a);
      if (MODE == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(g[j]) : "f"(a), "f"(b));
      if (MODE == 0) g[j] = fmaf(g[j], a, b);
    }
    a += 1.0f;
    b -= 1.0f;
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += g[j];
  if (s == 1.2345f) sink[0] = s;
}

template <int X, bool PIPE>
void run_ld(long long* d_c, uint32_t* d_s) {
  const int iters = 4096;
  for (int warps : {4, 8, 12, 16, 24, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      k_ld<X, PIPE><<<148, warps * 32>>>(iters, d_c, d_s);
      cudaDeviceSynchronize();
    }
    long long c;
    cudaMemcpy(&c, d_c, sizeof(c), cudaMemcpyDeviceToHost);
    const double bytes = (double)warps * iters * 32 * X * 4;
    printf("ld x%-2d %s warps %2d: %9lld cycles, %6.1f B/cycle/SM, %6.1f cycles per load and warp  (%s)\n", X, PIPE ? "pipelined" : "serial   ", warps, c,
           bytes / c, (double)c / iters, cudaGetErrorString(cudaGetLastError()));
  }
}
template <int MODE>
void run_alu(long long* d_c, float* d_f) {
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      k_alu<MODE><<<148, warps * 32>>>(iters, d_c, d_f, 1.0f);
      cudaDeviceSynchronize();
    }
    long long c;
    cudaMemcpy(&c, d_c, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%s warps %2d: %9lld cycles, %6.2f warp-instructions/cycle/SM  (%s)\n", MODE == 2 ? "max2" : MODE == 3 ? "max3" : "fma ", warps, c,
           (double)warps * iters * 16 / c, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  long long* d_c; uint32_t* d_s;
  cudaMalloc(&d_c, 1024 * sizeof(long long)); cudaMalloc(&d_s, 16);
  run_ld<8, false>(d_c, d_s);  run_ld<8, true>(d_c, d_s);
  run_ld<16, false>(d_c, d_s); run_ld<16, true>(d_c, d_s);
  run_ld<32, false>(d_c, d_s);
  run_alu<2>(d_c, (float*)d_s); run_alu<3>(d_c, (float*)d_s); run_alu<0>(d_c, (float*)d_s);
  return 0;
}
