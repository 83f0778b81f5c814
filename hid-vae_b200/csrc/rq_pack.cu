// Operand images of the tcgen05 quantiser kernels and the dispatcher behind HV_ALGO_TCGEN05 (include/hidvae_b200.h).
//
// hv_rq_pack_codebooks writes, per (level, 256-code N tile), the bf16 hi/lo split of the effective codebooks in the
// UMMA K-major core-matrix layout plus the -|c|^2/2 terms, so that one 1-D bulk TMA copy lands an image MMA-ready
// (modules/quantize.py:106-113: the distance table the reference materialises).  Two kernels consume the images:
//   rq_fwd_tc_v11.cu   "row owners": D = 32, K <= 256, L <= 3, everything resident in shared memory (C1, C2, C3, C5)
//   rq_fwd_tc_v4.cu    streamed images through a TMA ring: D = 16 / 32 / 64, any K (C4)
#include "common.cuh"
#include "ptx.cuh"

namespace hv {
namespace {

constexpr int kNTile = 256;  // codes per operand image

int image_bytes(int d) { return kNTile * (4 * d + 32); }
int images_per_level(int k) { return (k + kNTile - 1) / kNTile; }

// ---------------------------------------------------------------------------------------------------------
// Pack kernel: fp32 [L, K, D] -> per (level, 256-code image)
//   [c_hi : D/8 chunks][c_lo : D/8 chunks][norm : 2 chunks], chunk = [256 codes][8 bf16] (16 B per code)
// Padded codes (index >= K) get zero vectors and a -1e30 norm term so they can never win the argmax.
// ---------------------------------------------------------------------------------------------------------
// `cb32` (optional, D = 32 with one image per level): the swizzled fp32 copy generation 11 gathers from -- chunk c
// (16 bytes) of code k at chunk c ^ (k & 7) of its 128-byte row, [L][256 codes][32 floats].
template <int D>
__global__ void rq_pack_codebooks_kernel(const float* __restrict__ codebooks, int n_levels, int k, int n_ktiles,
                                         int tile_bytes, uint8_t* __restrict__ packed, uint8_t* __restrict__ cb32) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (level, image, code in image)
  const int total = n_levels * n_ktiles * kNTile;
  if (idx >= total) return;
  const int c = idx % kNTile;
  const int tile = idx / kNTile;  // level * n_ktiles + t
  const int t = tile % n_ktiles;
  const int level = tile / n_ktiles;
  const int code = t * kNTile + c;
  uint8_t* img = packed + static_cast<size_t>(tile) * tile_bytes;
  constexpr size_t chunk_stride = static_cast<size_t>(kNTile) * 16;
  uint8_t* hi_base = img;
  uint8_t* lo_base = img + (D / 8) * chunk_stride;
  uint8_t* nrm_base = img + 2 * (D / 8) * chunk_stride;

  float v[D];
  const bool real = code < k;
  if (real) {
    load_row<D>(v, codebooks + (static_cast<int64_t>(level) * k + code) * D);
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = 0.f;
  }
  if constexpr (D == 32) {
    if (cb32 != nullptr) {
      uint8_t* row = cb32 + (static_cast<size_t>(level) * kNTile + c) * (D * 4);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<float4*>(row + ((ch ^ (c & 7)) << 4)) = make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
    }
  }
  float cc = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) cc = fmaf(v[i], v[i], cc);
#pragma unroll
  for (int kc = 0; kc < D / 8; ++kc) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x0 = v[kc * 8 + 2 * j], x1 = v[kc * 8 + 2 * j + 1];
      const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
      const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __low2float(h), x1 - __high2float(h));
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(hi_base + kc * chunk_stride + c * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(lo_base + kc * chunk_stride + c * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  const float nv = real ? -0.5f * cc : -1e30f;
  const __nv_bfloat16 n1 = __float2bfloat16_rn(nv);
  const float rem1 = nv - __bfloat162float(n1);
  const __nv_bfloat16 n2 = __float2bfloat16_rn(rem1);
  const __nv_bfloat16 n3 = __float2bfloat16_rn(rem1 - __bfloat162float(n2));
  const uint32_t w0 = static_cast<uint32_t>(__bfloat16_as_ushort(n1)) | (static_cast<uint32_t>(__bfloat16_as_ushort(n2)) << 16);
  const uint32_t w1 = static_cast<uint32_t>(__bfloat16_as_ushort(n3));
  *reinterpret_cast<uint4*>(nrm_base + c * 16) = make_uint4(w0, w1, 0u, 0u);
  *reinterpret_cast<uint4*>(nrm_base + chunk_stride + c * 16) = make_uint4(0u, 0u, 0u, 0u);
}

template <int D>
int pack_d(const float* codebooks, int n_levels, int k, uint8_t* packed, cudaStream_t stream, uint8_t* cb32 = nullptr) {
  const int total_codes = n_levels * images_per_level(k) * kNTile;
  rq_pack_codebooks_kernel<D><<<(total_codes + 127) / 128, 128, 0, stream>>>(codebooks, n_levels, k, images_per_level(k),
                                                                           image_bytes(D), packed, cb32);
  HV_CUDA_CHECK(cudaGetLastError());
  return HV_OK;
}

}  // namespace

bool rq_fwd_tc_supported(int d, int k, int n_levels) {
  return rq_fwd_tc_v11_supported(d, k, n_levels) || rq_fwd_tc_v4_supported(d, k, n_levels);
}

size_t rq_fwd_tc_workspace_bytes(int d, int k, int n_levels) {
  if (!rq_fwd_tc_supported(d, k, n_levels)) return 0;
  // [operand images | swizzled fp32 copy of the codebooks for generation 11 (when the shape is served)]
  return static_cast<size_t>(n_levels) * images_per_level(k) * image_bytes(d) + rq_fwd_tc_v11_extra_bytes(d, k, n_levels);
}

int launch_rq_pack(const float* codebooks, int n_levels, int k, int d, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream) {
  if (!rq_fwd_tc_supported(d, k, n_levels)) {
    set_error("hv_rq_pack_codebooks: no tcgen05 instantiation for D=%d K=%d L=%d", d, k, n_levels);
    return HV_ERR_UNSUPPORTED;
  }
  const size_t images = static_cast<size_t>(n_levels) * images_per_level(k) * image_bytes(d);
  const size_t need = images + rq_fwd_tc_v11_extra_bytes(d, k, n_levels);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("hv_rq_pack_codebooks: needs a %zu-byte workspace (got %zu)", need, workspace_bytes);
    return HV_ERR_WORKSPACE;
  }
  if (!aligned16(workspace) || !aligned16(codebooks)) {
    set_error("hv_rq_pack_codebooks: codebooks and workspace must be 16-byte aligned");
    return HV_ERR_MISALIGNED;
  }
  uint8_t* packed = static_cast<uint8_t*>(workspace);
  switch (d) {
    case 16: return pack_d<16>(codebooks, n_levels, k, packed, stream);
    case 32: return pack_d<32>(codebooks, n_levels, k, packed, stream, need > images ? packed + images : nullptr);
    case 64: return pack_d<64>(codebooks, n_levels, k, packed, stream);
    default: return HV_ERR_UNSUPPORTED;
  }
}

int launch_rq_fwd_tc(const RqFwdArgs& a, int d, bool rot, void* workspace, size_t workspace_bytes, bool prepacked,
                     cudaStream_t stream) {
  if (!rq_fwd_tc_supported(d, a.k, a.n_levels)) {
    set_error("hv_rq_forward: no tcgen05 instantiation for D=%d K=%d L=%d", d, a.k, a.n_levels);
    return HV_ERR_UNSUPPORTED;
  }
  const size_t images = static_cast<size_t>(a.n_levels) * images_per_level(a.k) * image_bytes(d);
  const size_t need = images + rq_fwd_tc_v11_extra_bytes(d, a.k, a.n_levels);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("hv_rq_forward: tcgen05 path needs a %zu-byte workspace (got %zu)", need, workspace_bytes);
    return HV_ERR_WORKSPACE;
  }
  if (!aligned16(workspace)) {
    set_error("hv_rq_forward: workspace must be 16-byte aligned");
    return HV_ERR_MISALIGNED;
  }
  if (a.n == 0) return HV_OK;
  if (!prepacked)
    if (int st = launch_rq_pack(a.codebooks, a.n_levels, a.k, d, workspace, workspace_bytes, stream)) return st;
  // Generation 11 serves every shape it supports (measured faster from 12 K to 4 Mi rows, profiles/README.md);
  // everything else streams its images through the ring of rq_fwd_tc_v4.cu.
  if (need > images) return launch_rq_fwd_tc_v11(a, rot, workspace, static_cast<const uint8_t*>(workspace) + images, stream);
  return launch_rq_fwd_tc_v4(a, d, rot, workspace, stream);
}

}  // namespace hv
