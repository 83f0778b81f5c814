// Exact-fp32 CUDA-core variant of the fused L-level residual quantiser (HV_ALGO_SIMT / HV_ALGO_SIMT_DIFF).
//
// One thread owns one row for all L levels (rows are independent, levels are sequential per row).  The level's
// codebook is streamed through shared memory in chunks of KC codes and read back as warp-wide broadcasts; the
// [N, K] distance table only ever exists one scalar at a time in a register.
//   GEMM form  d = (|x|^2 + |c|^2) - 2 x.c      modules/quantize.py:108-113
//   DIFF form  d = sum_d (x_d - c_d)^2          init/kmeans.py:44-47
// First index wins exact ties (strict <, ascending k) like torch.min (modules/quantize.py:122).
// This variant serves shapes the tcgen05 kernel has no instantiation for, the k-means assignment in its
// reference difference form, and the on-device cross-check of the tensor-core kernel.
#include "common.cuh"

namespace hv {
namespace {

constexpr int kRowsPerCta = 128;
constexpr int kCodesPerChunk = 128;

template <int D, bool ROT, bool DIFF>
__global__ void __launch_bounds__(kRowsPerCta) rq_fwd_simt_kernel(RqFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* s_code = smem;                          // [KC][D]
  float* s_cc = smem + kCodesPerChunk * D;       // [KC]  |c|^2

  const int tid = threadIdx.x;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowsPerCta + tid;
  const bool valid = row < a.n;

  float r[D];
  if (valid) {
    load_row<D>(r, a.x + row * D);
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) r[i] = 0.f;
  }

  float total_loss = 0.f;
  for (int l = 0; l < a.n_levels; ++l) {
    const float* cb = a.codebooks + static_cast<int64_t>(l) * a.k * D;
    if (valid && a.residuals != nullptr) store_row<D>(a.residuals + (static_cast<int64_t>(l) * a.n + row) * D, r);

    float xx = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) xx = fmaf(r[i], r[i], xx);

    float best = INFINITY;
    int best_k = 0;
    for (int k0 = 0; k0 < a.k; k0 += kCodesPerChunk) {
      const int kc = min(kCodesPerChunk, a.k - k0);
      __syncthreads();  // previous chunk fully consumed
      for (int i = tid; i < kc * (D / 4); i += kRowsPerCta)
        reinterpret_cast<float4*>(s_code)[i] = __ldg(reinterpret_cast<const float4*>(cb + static_cast<int64_t>(k0) * D) + i);
      __syncthreads();
      if (!DIFF) {
        for (int c = tid; c < kc; c += kRowsPerCta) {
          float cc = 0.f;
#pragma unroll
          for (int i = 0; i < D; ++i) cc = fmaf(s_code[c * D + i], s_code[c * D + i], cc);
          s_cc[c] = cc;
        }
        __syncthreads();
      }
      for (int c = 0; c < kc; ++c) {
        const float4* code = reinterpret_cast<const float4*>(s_code + c * D);
        float dist;
        if (DIFF) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < D / 4; ++i) {
            const float4 v = code[i];
            const float d0 = r[4 * i] - v.x, d1 = r[4 * i + 1] - v.y, d2 = r[4 * i + 2] - v.z, d3 = r[4 * i + 3] - v.w;
            acc = fmaf(d0, d0, acc), acc = fmaf(d1, d1, acc), acc = fmaf(d2, d2, acc), acc = fmaf(d3, d3, acc);
          }
          dist = acc;
        } else {
          float dot = 0.f;
#pragma unroll
          for (int i = 0; i < D / 4; ++i) {
            const float4 v = code[i];
            dot = fmaf(r[4 * i], v.x, dot), dot = fmaf(r[4 * i + 1], v.y, dot);
            dot = fmaf(r[4 * i + 2], v.z, dot), dot = fmaf(r[4 * i + 3], v.w, dot);
          }
          dist = (xx + s_cc[c]) - 2.0f * dot;
        }
        if (dist < best) {
          best = dist;
          best_k = k0 + c;
        }
      }
    }

    if (valid) {
      a.ids[row * a.ids_row_stride + l * a.ids_level_stride] = best_k;
      float e[D];
      load_row<D>(e, cb + static_cast<int64_t>(best_k) * D);
      float* o_out = a.emb_out != nullptr ? a.emb_out + (static_cast<int64_t>(l) * a.n + row) * D : nullptr;
      const float ll = rq_level_tail<D, ROT>(r, e, a.beta, o_out);
      total_loss += ll;
      if (a.level_loss != nullptr) a.level_loss[static_cast<int64_t>(l) * a.n + row] = ll;
    }
  }
  if (valid) {
    if (a.loss != nullptr) a.loss[row] = total_loss;
    if (a.final_residual != nullptr) store_row<D>(a.final_residual + row * D, r);
  }
}

template <int D>
int launch_d(const RqFwdArgs& a, bool rot, bool diff, cudaStream_t stream) {
  const unsigned grid = static_cast<unsigned>((a.n + kRowsPerCta - 1) / kRowsPerCta);
  const size_t smem = sizeof(float) * (kCodesPerChunk * D + kCodesPerChunk);
  auto go = [&](auto kernel) -> int {
    if (smem > 48 * 1024) HV_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, kRowsPerCta, smem, stream>>>(a);
    HV_CUDA_CHECK(cudaGetLastError());
    return HV_OK;
  };
  if (rot) return diff ? go(rq_fwd_simt_kernel<D, true, true>) : go(rq_fwd_simt_kernel<D, true, false>);
  return diff ? go(rq_fwd_simt_kernel<D, false, true>) : go(rq_fwd_simt_kernel<D, false, false>);
}

}  // namespace

int launch_rq_fwd_simt(const RqFwdArgs& a, int d, bool rot, bool diff_form, cudaStream_t stream) {
  if (a.n == 0) return HV_OK;
  switch (d) {
    case 4: return launch_d<4>(a, rot, diff_form, stream);
    case 8: return launch_d<8>(a, rot, diff_form, stream);
    case 16: return launch_d<16>(a, rot, diff_form, stream);
    case 32: return launch_d<32>(a, rot, diff_form, stream);
    case 64: return launch_d<64>(a, rot, diff_form, stream);
    case 128: return launch_d<128>(a, rot, diff_form, stream);
    default:
      set_error("hv_rq_forward: embed dim %d has no SIMT instantiation (supported: 4, 8, 16, 32, 64, 128)", d);
      return HV_ERR_UNSUPPORTED;
  }
}

}  // namespace hv
