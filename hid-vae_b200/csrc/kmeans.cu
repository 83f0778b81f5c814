// K-means codebook init, the update half of one Lloyd iteration (init/kmeans.py:52-61) -- HBM/L2-bound.
//
// The assignment half is hv_rq_forward with n_levels = 1 (HV_ALGO_SIMT_DIFF reproduces the reference's
// difference-form table, init/kmeans.py:44-47).  Here:
//   hv_kmeans_accumulate  per-cluster sums / counts of the assigned rows.  One CTA per cluster walks the
//                         assignment vector (coalesced int64 reads, L2-resident after the first CTA) and adds its
//                         members in ROW ORDER per thread, then a fixed-shape shuffle/shared-memory tree: the
//                         result is bit-reproducible from run to run.  That matters: the reference stops when
//                         max ||c_new - c_old|| < 1e-10 (init/kmeans.py:68), i.e. when the update reproduces the
//                         centroids exactly, which an atomics-ordered float sum never does.
//   hv_kmeans_finalize    means, empty-cluster reseed, max centroid shift, in one CTA.
#include "common.cuh"

namespace hv {
namespace {

constexpr int kAccThreads = 256;

template <int D>
__global__ void __launch_bounds__(kAccThreads) kmeans_accumulate_kernel(const float* __restrict__ x, int64_t n,
                                                                        const int64_t* __restrict__ assign,
                                                                        const int64_t* __restrict__ prev_assign, int k,
                                                                        float* __restrict__ sums, float* __restrict__ counts,
                                                                        unsigned long long* __restrict__ n_changed) {
  __shared__ float s_part[kAccThreads / 32][D + 1];
  const int c = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  float acc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) acc[i] = 0.f;
  float cnt = 0.f;
  for (int64_t row = tid; row < n; row += kAccThreads) {
    if (assign[row] == c) {
      const float4* src = reinterpret_cast<const float4*>(x + row * D);
#pragma unroll
      for (int i = 0; i < D / 4; ++i) {
        const float4 v = __ldg(src + i);
        acc[4 * i] += v.x, acc[4 * i + 1] += v.y, acc[4 * i + 2] += v.z, acc[4 * i + 3] += v.w;
      }
      cnt += 1.f;
    }
  }
  // fixed-shape reduction: xor tree inside the warp, then the 8 warp partials in warp order
#pragma unroll
  for (int i = 0; i < D; ++i) {
    float v = acc[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) s_part[warp][i] = v;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  if (lane == 0) s_part[warp][D] = cnt;
  __syncthreads();
  for (int i = tid; i <= D; i += kAccThreads) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kAccThreads / 32; ++w) v += s_part[w][i];
    if (i < D)
      sums[static_cast<int64_t>(c) * D + i] = v;
    else
      counts[c] = v;
  }

  if (n_changed != nullptr) {
    // every CTA also counts the changed assignments of its own slice of the rows (integer sum: order-free)
    const int64_t per = (n + k - 1) / k;
    const int64_t lo = static_cast<int64_t>(c) * per;
    const int64_t hi = lo + per < n ? lo + per : n;
    unsigned int changed = 0;
    for (int64_t row = lo + tid; row < hi; row += kAccThreads)
      changed += (prev_assign == nullptr || prev_assign[row] != assign[row]) ? 1u : 0u;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, off);
    if (lane == 0 && changed) atomicAdd(n_changed, static_cast<unsigned long long>(changed));
  }
}

__global__ void __launch_bounds__(256) kmeans_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ counts,
                                                             const float* __restrict__ reseed_rows, int k, int d,
                                                             float* __restrict__ centroids, float* __restrict__ stats) {
  __shared__ float s_max[8];
  __shared__ float s_empty[8];
  float worst = 0.f, n_empty = 0.f;
  for (int c = threadIdx.x; c < k; c += blockDim.x) {
    const float cnt = counts[c];
    const bool empty = !(cnt > 0.f);
    if (empty) n_empty += 1.f;
    float shift2 = 0.f;
    for (int i = 0; i < d; ++i) {
      const int64_t o = static_cast<int64_t>(c) * d + i;
      const float old = centroids[o];
      float nv;
      if (!empty)
        nv = sums[o] / cnt;  // x[members].mean(axis=0), init/kmeans.py:60
      else
        nv = reseed_rows != nullptr ? reseed_rows[o] : old;  // init/kmeans.py:56
      const float df = nv - old;
      shift2 = fmaf(df, df, shift2);
      centroids[o] = nv;
    }
    worst = fmaxf(worst, sqrtf(shift2));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, off));
    n_empty += __shfl_xor_sync(0xffffffffu, n_empty, off);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_max[warp] = worst, s_empty[warp] = n_empty;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f, e = 0.f;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) m = fmaxf(m, s_max[w]), e += s_empty[w];
    stats[0] = m;
    stats[1] = e;
  }
}

template <int D>
int launch_acc(const float* x, int64_t n, const int64_t* assign, const int64_t* prev, int k, float* sums, float* counts,
               int64_t* n_changed, cudaStream_t s) {
  kmeans_accumulate_kernel<D><<<k, kAccThreads, 0, s>>>(x, n, assign, prev, k, sums, counts,
                                                        reinterpret_cast<unsigned long long*>(n_changed));
  HV_CUDA_CHECK(cudaGetLastError());
  return HV_OK;
}

}  // namespace
}  // namespace hv

extern "C" int hv_kmeans_accumulate(const float* x, int64_t n, int d, const int64_t* assign, const int64_t* prev_assign,
                                    int k, float* sums, float* counts, int64_t* n_changed, void* stream) {
  using namespace hv;
  if (n < 0 || d <= 0 || k <= 0) {
    set_error("hv_kmeans_accumulate: bad shape n=%lld d=%d k=%d", (long long)n, d, k);
    return HV_ERR_BAD_SHAPE;
  }
  if (!sums || !counts || (n > 0 && (!x || !assign))) {
    set_error("hv_kmeans_accumulate: null pointer");
    return HV_ERR_NULL;
  }
  if (n > 0 && !aligned16(x)) {
    set_error("hv_kmeans_accumulate: x must be 16-byte aligned");
    return HV_ERR_MISALIGNED;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_changed != nullptr) HV_CUDA_CHECK(cudaMemsetAsync(n_changed, 0, sizeof(int64_t), s));
  switch (d) {
    case 4: return launch_acc<4>(x, n, assign, prev_assign, k, sums, counts, n_changed, s);
    case 8: return launch_acc<8>(x, n, assign, prev_assign, k, sums, counts, n_changed, s);
    case 16: return launch_acc<16>(x, n, assign, prev_assign, k, sums, counts, n_changed, s);
    case 32: return launch_acc<32>(x, n, assign, prev_assign, k, sums, counts, n_changed, s);
    case 64: return launch_acc<64>(x, n, assign, prev_assign, k, sums, counts, n_changed, s);
    case 128: return launch_acc<128>(x, n, assign, prev_assign, k, sums, counts, n_changed, s);
    default:
      set_error("hv_kmeans_accumulate: embed dim %d has no instantiation (supported: 4, 8, 16, 32, 64, 128)", d);
      return HV_ERR_UNSUPPORTED;
  }
}

extern "C" int hv_kmeans_finalize(const float* sums, const float* counts, const float* reseed_rows, int k, int d,
                                  float* centroids, float* stats, void* stream) {
  using namespace hv;
  if (d <= 0 || k <= 0) {
    set_error("hv_kmeans_finalize: bad shape d=%d k=%d", d, k);
    return HV_ERR_BAD_SHAPE;
  }
  if (!sums || !counts || !centroids || !stats) {
    set_error("hv_kmeans_finalize: null pointer");
    return HV_ERR_NULL;
  }
  kmeans_finalize_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(sums, counts, reseed_rows, k, d, centroids, stats);
  HV_CUDA_CHECK(cudaGetLastError());
  return HV_OK;
}
