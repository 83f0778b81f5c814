// Fused encoder MLP in front of the quantiser for sm_100a:
//     z = [l2norm]( W4 silu( W3 silu( W2 silu( W1 x ))))        modules/encoder.py:23-36, called at modules/h_rqvae.py:599
// for the reference's encoder shape 768 -> 512 -> 256 -> 128 -> 32 (bias-free Linear + SiLU, optional L2-norm tail,
// configs/h_rqvae_*.gin).  One persistent CTA per SM walks 128-row tiles of x; the four GEMMs of a tile run back to
// back on tcgen05 and NO intermediate activation ever leaves the SM:
//
//   PRECISION  operands are rounded to fp16 (11-bit significand = TF32's, the precision of the reference's own GPU path:
//              torch.set_float32_matmul_precision('high'), modules/h_rqvae.py:21), products accumulate in fp32 in
//              tensor memory, activations are evaluated in fp32.  kind::f16 runs at twice the kind::tf32 rate.
//   LAYER 1    A = the x tile: the 8 epilogue warps read fp32 rows (coalesced 16-byte loads of L2-prefetched lines),
//              round to fp16 and write 64-column chunks straight into the UMMA K-major core-matrix layout (ring of
//              4 stages; the K-group stride is padded by 16 bytes so the stores are conflict free).  B = W1 as a
//              pre-packed fp16 image streamed from L2 by 1-D bulk TMA copies in 32 KB blocks (ring of 4 stages).
//              D1 = 128 x 512 fp32 fills ALL 512 TMEM columns.
//   LAYERS 2-4 the epilogue of layer l (tcgen05.ld -> SiLU -> fp16 pairs -> tcgen05.st) packs the activations IN PLACE
//              into tensor memory and the next layer's MMAs take their A operand from there
//              (tcgen05.mma [d], [a_tmem], b_desc): no shared-memory round trip, no proxy fence.  Column plan:
//                  D1 [0,512)                 h1 [0,128) u [384,512)   (half 0 packs upwards, half 1 downwards)
//                  D2 [128,384)               h2 [0,128)
//                  D3 [384,512)               h3 [128,192)
//                  D4 [256,288)
//   SiLU       x sigmoid(x) = h + h tanh(h), h = x/2: ONE MUFU op per element (B200 has 16 MUFU lanes per SM, and
//              896 activations per row make the epilogues MUFU-bound); `precise` selects ex2 + rcp instead.
//   OUTPUT     z [N, 32] fp32 (row-normalised when the encoder was built with normalize=True), staged through
//              shared memory so that a warp stores four full 128-byte rows per instruction.
// Because D1 needs the whole of tensor memory the layers of a tile are serial; the next tile's first four x chunks are
// converted while layer 2 runs, and the weight ring runs ahead across tiles.
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace hv {
namespace {

constexpr int kIn = 768, kH1 = 512, kH2 = 256, kH3 = 128, kOut = 32;
constexpr int kTileRows = 128;
constexpr int kThreads = 384;          // warp 0 weight producer, 1 MMA issuer, 2 TMEM allocator, 4..11 convert + epilogue
constexpr int kEpiWarp0 = 4;
constexpr int kEpiThreads = 256;
constexpr int kWStageBytes = 32768;    // one weight block
constexpr int kWStages = 4;
constexpr int kALbo = 128 * 16 + 16;   // bytes between K groups (8 fp16) of the A chunk: padded, see convert_chunk
constexpr int kAStageBytes = 8 * kALbo;  // one 128 x 64 fp16 chunk of x
constexpr int kAStages = 4;
constexpr int kChunks = kIn / 64;      // x chunks per tile
constexpr int kBlocksL1 = kIn / 32, kBlocksL2 = kH1 / 64, kBlocksL3 = kH2 / 128;
constexpr int kBlocks = kBlocksL1 + kBlocksL2 + kBlocksL3 + 1;   // weight blocks per tile (the last one is 8 KB)
constexpr int kLastBlockBytes = kH3 * kOut * 2;
constexpr size_t kImageBytes = static_cast<size_t>(kBlocks - 1) * kWStageBytes + kLastBlockBytes;
constexpr int kZStageBytes = kTileRows * kOut * 4;
constexpr int kBarBytes = 256;
constexpr int kSmemBytes = kWStages * kWStageBytes + kAStages * kAStageBytes + kZStageBytes + kBarBytes;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(kImageBytes == 2ull * (kIn * kH1 + kH1 * kH2 + kH2 * kH3 + kH3 * kOut), "image = every weight once, fp16");

// tensor-memory columns (see the plan above)
constexpr uint32_t kColD2 = 128, kColD3 = 384, kColD4 = 256, kColH1b = 384, kColH3 = 128;

// kind::f16 instruction descriptor with fp16 operands (a_format = b_format = 0), fp32 accumulate, both K-major
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t m, uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24); }

__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

template <bool PRECISE>
__device__ __forceinline__ float silu(float x) {
  if constexpr (PRECISE) {
    return __fdividef(x, 1.0f + __expf(-x));
  } else {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  }
}

__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// Epilogue of a hidden layer for one warp: NCH chunks of 32 accumulator columns starting at `src` -> SiLU -> fp16 pairs
// -> 16 columns each starting at `dst`.  DESC walks the chunks downwards (the in-place packing of layer 1's upper half).
template <int NCH, bool DESC, bool PRECISE>
__device__ __forceinline__ void epilogue_pack(uint32_t src, uint32_t dst) {
  uint32_t v[2][32];
  ptx::tmem_ld_32x32(src + 32 * (DESC ? NCH - 1 : 0), v[0]);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int j = DESC ? NCH - 1 - i : i;
    ptx::tmem_wait_ld(v[i & 1]);
    if (i + 1 < NCH) ptx::tmem_ld_32x32(src + 32 * (DESC ? j - 1 : j + 1), v[(i + 1) & 1]);
    uint32_t p[16];
#pragma unroll
    for (int c = 0; c < 16; ++c)
      p[c] = pack_f16(silu<PRECISE>(__uint_as_float(v[i & 1][2 * c])), silu<PRECISE>(__uint_as_float(v[i & 1][2 * c + 1])));
    ptx::tmem_st_32x16(dst + 16 * j, p);
  }
  ptx::tmem_wait_st();
}

struct EncArgs {
  const void* x;           // [N, 768] fp32, or fp16 for the HALF_IN instantiation
  const uint8_t* image;
  float* z;
  int64_t n;
  int normalize;
};

// HALF_IN: the items are already fp16 (a catalogue kept in half precision: half the HBM / PCIe bytes).  The fp32 path
// rounds x to fp16 (round to nearest even) before the first GEMM, so fp16 items that are the rounded fp32 items give
// bit-identical z; the converter warps then only move 16-byte pieces into the operand layout.
template <bool PRECISE, bool HALF_IN>
__global__ void __launch_bounds__(kThreads, 1) enc_mlp_kernel(EncArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_a = s_w + kWStages * kWStageBytes;
  uint8_t* s_z = s_a + kAStages * kAStageBytes;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_z + kZStageBytes);
  uint64_t* w_full = s_bar;               // [kWStages]  bulk copy landed
  uint64_t* w_empty = w_full + kWStages;  // [kWStages]  the MMAs that read the block have completed
  uint64_t* a_full = w_empty + kWStages;  // [kAStages]  256 converter threads stored + fenced
  uint64_t* a_empty = a_full + kAStages;  // [kAStages]
  uint64_t* d_ready = a_empty + kAStages;  // a layer's accumulator is complete (4 completions per tile)
  uint64_t* h_ready = d_ready + 1;         // the 8 epilogue warps have packed / drained it (4 completions per tile)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(h_ready + 1);

  const int warp = ptx::warp_index();
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWStages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&w_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&w_empty[i]), 1);
    }
    for (int i = 0; i < kAStages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&a_full[i]), kEpiThreads);
      ptx::mbar_init(ptx::smem_u32(&a_empty[i]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(d_ready), 1);
    ptx::mbar_init(ptx::smem_u32(h_ready), kEpiThreads / 32);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(s_tmem), 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;

  const int64_t n_tiles = (a.n + kTileRows - 1) / kTileRows;
  const int my_tiles = static_cast<int>((n_tiles - 1 - blockIdx.x) / gridDim.x + 1);  // grid <= n_tiles
  auto tile_row0 = [&](int i) -> int64_t { return (static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x) * kTileRows; };

  if (warp == 0) {
    // ================================= weight producer (+ L2 prefetch of the next x tile) =========================
    if (ptx::elect_one()) {
      uint32_t seq = 0;
      for (int i = 0; i < my_tiles; ++i) {
        if (i + 1 < my_tiles) {
          const int64_t r0 = tile_row0(i + 1);
          const int64_t rows = a.n - r0 < kTileRows ? a.n - r0 : kTileRows;
          constexpr int kElem = HALF_IN ? 2 : 4;
          const uint8_t* p = static_cast<const uint8_t*>(a.x) + r0 * kIn * kElem;
          const int64_t bytes = rows * kIn * kElem;
          for (int64_t off = 0; off < bytes; off += 32768)
            ptx::bulk_prefetch_l2(p + off, static_cast<uint32_t>(bytes - off < 32768 ? bytes - off : 32768));
        }
        for (int b = 0; b < kBlocks; ++b, ++seq) {
          const uint32_t s = seq % kWStages, use = seq / kWStages;
          if (use > 0) ptx::mbar_wait(ptx::smem_u32(&w_empty[s]), (use - 1) & 1u);
          const uint32_t bytes = b == kBlocks - 1 ? kLastBlockBytes : kWStageBytes;
          const uint32_t bar = ptx::smem_u32(&w_full[s]);
          ptx::mbar_arrive_expect_tx(bar, bytes);
          ptx::bulk_g2s(ptx::smem_u32(s_w + s * kWStageBytes), a.image + static_cast<size_t>(b) * kWStageBytes, bytes, bar);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================ MMA issuer ===============================================
    if (ptx::elect_one()) {
      const uint32_t sbo_hi = ptx::umma_desc_hi(128);
      uint32_t wseq = 0, aseq = 0, hcnt = 0;
      auto wait_w = [&]() -> uint32_t {
        const uint32_t s = wseq % kWStages;
        ptx::mbar_wait(ptx::smem_u32(&w_full[s]), (wseq / kWStages) & 1u);
        return s;
      };
      auto free_w = [&](uint32_t s) {
        ptx::umma_commit(ptx::smem_u32(&w_empty[s]));
        ++wseq;
      };
      auto wait_h = [&]() {
        ptx::mbar_wait(ptx::smem_u32(h_ready), hcnt & 1u);
        ++hcnt;
        ptx::tc_fence_after_sync();
      };
      for (int i = 0; i < my_tiles; ++i) {
        if (i > 0) wait_h();  // the previous tile's last accumulator has been read: tensor memory is free
        // ---- layer 1: D1[128 x 512] = x[128 x 768] W1^T, A chunks from shared memory ----
        uint32_t as = 0;
        for (int b = 0; b < kBlocksL1; ++b) {
          const uint32_t s = wait_w();
          if ((b & 1) == 0) {
            as = aseq % kAStages;
            ptx::mbar_wait(ptx::smem_u32(&a_full[as]), (aseq / kAStages) & 1u);
          }
          ptx::tc_fence_after_sync();
          const uint32_t a_base = ptx::smem_u32(s_a + as * kAStageBytes) + (b & 1) * 4 * kALbo;
          const uint32_t w_base = ptx::smem_u32(s_w + s * kWStageBytes);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t ad = ptx::umma_desc(ptx::umma_desc_lo(a_base + ks * 2 * kALbo, kALbo), sbo_hi);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t bd = ptx::umma_desc(ptx::umma_desc_lo(w_base + ks * (2 * kH1 * 16) + h * (256 * 16), kH1 * 16), sbo_hi);
              ptx::umma_bf16(tmem + h * 256, ad, bd, idesc_f16(128, 256), (b | ks) > 0 ? 1u : 0u);
            }
          }
          free_w(s);
          if (b & 1) {
            ptx::umma_commit(ptx::smem_u32(&a_empty[as]));
            ++aseq;
          }
        }
        ptx::umma_commit(ptx::smem_u32(d_ready));
        // ---- layer 2: D2[128 x 256] = h1[128 x 512] W2^T, A from tensor memory ----
        wait_h();
        for (int b = 0; b < kBlocksL2; ++b) {
          const uint32_t s = wait_w();
          ptx::tc_fence_after_sync();
          const uint32_t w_base = ptx::smem_u32(s_w + s * kWStageBytes);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const int j = 4 * b + ks;
            const uint32_t at = tmem + (j < 16 ? 8 * j : kColH1b + 8 * (j - 16));
            const uint64_t bd = ptx::umma_desc(ptx::umma_desc_lo(w_base + ks * (2 * kH2 * 16), kH2 * 16), sbo_hi);
            ptx::umma_bf16_ts(tmem + kColD2, at, bd, idesc_f16(128, 256), j > 0 ? 1u : 0u);
          }
          free_w(s);
        }
        ptx::umma_commit(ptx::smem_u32(d_ready));
        // ---- layer 3: D3[128 x 128] = h2[128 x 256] W3^T ----
        wait_h();
        for (int b = 0; b < kBlocksL3; ++b) {
          const uint32_t s = wait_w();
          ptx::tc_fence_after_sync();
          const uint32_t w_base = ptx::smem_u32(s_w + s * kWStageBytes);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const int j = 8 * b + ks;
            const uint64_t bd = ptx::umma_desc(ptx::umma_desc_lo(w_base + ks * (2 * kH3 * 16), kH3 * 16), sbo_hi);
            ptx::umma_bf16_ts(tmem + kColD3, tmem + 8 * j, bd, idesc_f16(128, 128), j > 0 ? 1u : 0u);
          }
          free_w(s);
        }
        ptx::umma_commit(ptx::smem_u32(d_ready));
        // ---- layer 4: D4[128 x 32] = h3[128 x 128] W4^T ----
        wait_h();
        {
          const uint32_t s = wait_w();
          ptx::tc_fence_after_sync();
          const uint32_t w_base = ptx::smem_u32(s_w + s * kWStageBytes);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t bd = ptx::umma_desc(ptx::umma_desc_lo(w_base + ks * (2 * kOut * 16), kOut * 16), sbo_hi);
            ptx::umma_bf16_ts(tmem + kColD4, tmem + kColH3 + 8 * ks, bd, idesc_f16(128, 32), ks > 0 ? 1u : 0u);
          }
          free_w(s);
        }
        ptx::umma_commit(ptx::smem_u32(d_ready));
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // ================================== x converter + epilogues (8 warps) ======================================
    const int we = warp - kEpiWarp0;      // 0..7
    const int q = we & 3;                  // TMEM lane quarter of this warp (== warp % 4)
    const int hf = we >> 2;                // column half
    const uint32_t tl = tmem + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t aseq = 0, dcnt = 0;

    // one 128 x 64 chunk of x: instruction i of a warp reads rows 16 we + 2 i + (lane >> 4), 16-byte piece lane & 15 --
    // two fully coalesced 256-byte row segments -- and stores 4 fp16 (8 bytes) at K group (piece >> 1), row, half
    // (piece & 1).  With the K-group stride padded to 2064 bytes the 16 lanes of a row cover 128 consecutive bytes of
    // bank space: conflict free.
    // (fp32 items) 8 x 16-byte loads of 4 floats per lane and chunk; (fp16 items) instruction k of a warp reads rows
    // 16 we + 4 k + (lane >> 3), 16-byte piece lane & 7 = one whole K group of 8 fp16: four 128-byte row segments per
    // instruction, stored as they are.
    constexpr int kLoads = HALF_IN ? 4 : 8;
    using Piece = typename std::conditional<HALF_IN, uint4, float4>::type;
    auto load_chunk = [&](int i, int c, Piece (&v)[kLoads]) {
      if constexpr (HALF_IN) {
        const int64_t r0 = tile_row0(i) + 16 * we + (lane >> 3);
        const __half* p = static_cast<const __half*>(a.x) + r0 * kIn + c * 64 + (lane & 7) * 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (r0 + 4 * k < a.n) {
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w)
                         : "l"(p + static_cast<int64_t>(4 * k) * kIn));
          } else {
            v[k] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      } else {
        const int64_t r0 = tile_row0(i) + 16 * we + (lane >> 4);
        const float* p = static_cast<const float*>(a.x) + r0 * kIn + c * 64 + (lane & 15) * 4;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (r0 + 2 * k < a.n) {
            v[k] = ldg_stream(p + static_cast<int64_t>(2 * k) * kIn);
          } else {
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    };
    auto store_chunk = [&](const Piece (&v)[kLoads]) {
      const uint32_t s = aseq % kAStages, use = aseq / kAStages;
      if (use > 0) ptx::mbar_wait(ptx::smem_u32(&a_empty[s]), (use - 1) & 1u);
      if constexpr (HALF_IN) {
        const uint32_t base = ptx::smem_u32(s_a + s * kAStageBytes) + (lane & 7) * kALbo + (16 * we + (lane >> 3)) * 16;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + 4 * k * 16), "r"(v[k].x), "r"(v[k].y), "r"(v[k].z),
                       "r"(v[k].w)
                       : "memory");
      } else {
        const uint32_t base = ptx::smem_u32(s_a + s * kAStageBytes) + ((lane & 15) >> 1) * kALbo + (16 * we + (lane >> 4)) * 16 + (lane & 1) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t lo = pack_f16(v[k].x, v[k].y), hi = pack_f16(v[k].z, v[k].w);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base + 2 * k * 16), "r"(lo), "r"(hi) : "memory");
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(ptx::smem_u32(&a_full[s]));
      ++aseq;
    };
    auto convert = [&](int i, int c0, int c1) {  // chunks [c0, c1) of tile i (an even count), loads one chunk ahead of the stores
      Piece va[kLoads], vb[kLoads];
      load_chunk(i, c0, va);
#pragma unroll 1
      for (int c = c0; c < c1; c += 2) {
        load_chunk(i, c + 1, vb);
        store_chunk(va);
        if (c + 2 < c1) load_chunk(i, c + 2, va);
        store_chunk(vb);
      }
    };
    auto wait_d = [&]() {
      ptx::mbar_wait(ptx::smem_u32(d_ready), dcnt & 1u);
      ++dcnt;
      ptx::tc_fence_after_sync();
    };
    auto signal_h = [&]() {
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(h_ready));
    };

    // instrumented builds (-DHV_TC_INSTRUMENT): mean cycles per tile of every phase as seen by epilogue warp 0 of CTA 0,
    // accumulated in registers and printed once after the last tile (a printf inside the loop perturbs the pipeline)
#ifdef HV_TC_INSTRUMENT
    long long ts[12], ts_sum[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define ENC_STAMP(j) ts[j] = clock64()
#else
#define ENC_STAMP(j)
#endif
    int pre = 0;  // chunks of the current tile converted ahead of time
    for (int i = 0; i < my_tiles; ++i) {
      ENC_STAMP(0);
      convert(i, pre, kChunks);
      ENC_STAMP(1);
      // ---- layer 1 -> h1 (in place: half 0 packs [0,256) upwards into [0,128), half 1 packs [256,512) downwards into [384,512)) ----
      wait_d();
      ENC_STAMP(2);
      if (hf == 0) epilogue_pack<8, false, PRECISE>(tl, tl);
      else epilogue_pack<8, true, PRECISE>(tl + 256, tl + kColH1b);
      signal_h();
      ENC_STAMP(3);
      // the next tile's first chunks while layer 2 runs (their stages were freed by layer 1's commits)
      pre = 0;
      if (i + 1 < my_tiles) {
        convert(i + 1, 0, kAStages);
        pre = kAStages;
      }
      // ---- layer 2 -> h2 [0,128) ----
      ENC_STAMP(4);
      wait_d();
      ENC_STAMP(5);
      epilogue_pack<4, false, PRECISE>(tl + kColD2 + 128 * hf, tl + 64 * hf);
      signal_h();
      ENC_STAMP(6);
      // ---- layer 3 -> h3 [128,192) ----
      wait_d();
      ENC_STAMP(7);
      epilogue_pack<2, false, PRECISE>(tl + kColD3 + 64 * hf, tl + kColH3 + 32 * hf);
      signal_h();
      ENC_STAMP(8);
      // ---- layer 4 -> z ----
      wait_d();
      ENC_STAMP(9);
      if (hf == 0) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tl + kColD4, v);
        ptx::tmem_wait_ld(v);
        signal_h();  // tensor memory is free for the next tile's layer 1
        float zr[kOut];
#pragma unroll
        for (int d = 0; d < kOut; ++d) zr[d] = __uint_as_float(v[d]);
        if (a.normalize) {  // F.normalize(x, p=2, dim=-1, eps=1e-12), modules/normalize.py:7-8
          float ss = 0.f;
#pragma unroll
          for (int d = 0; d < kOut; ++d) ss = fmaf(zr[d], zr[d], ss);
          const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
          for (int d = 0; d < kOut; ++d) zr[d] *= inv;
        }
        // own row -> shared memory (16-byte chunks XOR-swizzled by row), then four full rows per warp instruction
        const int t = q * 32 + lane;
        const uint32_t zs = ptx::smem_u32(s_z);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(zs + t * 128 + ((c ^ (t & 7)) << 4)), "f"(zr[4 * c]),
                       "f"(zr[4 * c + 1]), "f"(zr[4 * c + 2]), "f"(zr[4 * c + 3])
                       : "memory");
        __syncwarp();
        const int64_t row0 = tile_row0(i);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int row = q * 32 + 4 * k + (lane >> 3), c = lane & 7;
          const float4 val = ptx::lds128(zs + row * 128 + ((c ^ (row & 7)) << 4));
          if (row0 + row < a.n) *reinterpret_cast<float4*>(a.z + (row0 + row) * kOut + c * 4) = val;
        }
        __syncwarp();
      } else {
        signal_h();
      }
#ifdef HV_TC_INSTRUMENT
      ENC_STAMP(10);
      if (i > 0)
        for (int j = 0; j < 10; ++j) ts_sum[j] += ts[j + 1] - ts[j];
#endif
    }
#ifdef HV_TC_INSTRUMENT
    if (blockIdx.x == 0 && threadIdx.x == kEpiWarp0 * 32 && my_tiles > 1) {
      const long long t = my_tiles - 1;
      printf("ENC mean cycles per tile over %lld tiles: convert %lld | wait L1 %lld | epi1 %lld | preconvert %lld | wait L2 %lld | epi2 %lld | wait L3 %lld | epi3 %lld | wait L4 %lld | out %lld\n",
             t, ts_sum[0] / t, ts_sum[1] / t, ts_sum[2] / t, ts_sum[3] / t, ts_sum[4] / t, ts_sum[5] / t, ts_sum[6] / t, ts_sum[7] / t,
             ts_sum[8] / t, ts_sum[9] / t);
    }
#endif
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// fp32 weights [out, in] (nn.Linear) -> the fp16 image the kernel streams.  Block b of layer l covers KC consecutive
// input features; inside a block  [K group of 8][output feature n][8 fp16]  -- the UMMA K-major core-matrix layout
// (LBO = N * 16 bytes between K groups, SBO = 128 bytes between 8-row groups), so a block lands MMA-ready.
__global__ void enc_pack_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ w3,
                                        const float* __restrict__ w4, uint8_t* __restrict__ image) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte unit (8 fp16) per thread
  constexpr int kUnits = static_cast<int>(kImageBytes / 16);
  if (u >= kUnits) return;
  constexpr int u1 = kIn * kH1 / 8, u2 = u1 + kH1 * kH2 / 8, u3 = u2 + kH2 * kH3 / 8;
  const float* w;
  int n_out, n_in, kc, v;
  if (u < u1) w = w1, n_out = kH1, n_in = kIn, kc = 32, v = u;
  else if (u < u2) w = w2, n_out = kH2, n_in = kH1, kc = 64, v = u - u1;
  else if (u < u3) w = w3, n_out = kH3, n_in = kH2, kc = 128, v = u - u2;
  else w = w4, n_out = kOut, n_in = kH3, kc = 128, v = u - u3;
  const int per_block = kc / 8 * n_out;
  const int b = v / per_block, in_block = v % per_block;
  const int kg = in_block / n_out, n = in_block % n_out;
  const float* src = w + static_cast<size_t>(n) * n_in + b * kc + kg * 8;
  const float4 lo = __ldg(reinterpret_cast<const float4*>(src)), hi = __ldg(reinterpret_cast<const float4*>(src) + 1);
  *reinterpret_cast<uint4*>(image + static_cast<size_t>(u) * 16) =
      make_uint4(pack_f16(lo.x, lo.y), pack_f16(lo.z, lo.w), pack_f16(hi.x, hi.y), pack_f16(hi.z, hi.w));
}

bool shape_ok(int n_layers, const int* dims) {
  return n_layers == 4 && dims != nullptr && dims[0] == kIn && dims[1] == kH1 && dims[2] == kH2 && dims[3] == kH3 && dims[4] == kOut;
}

}  // namespace
}  // namespace hv

extern "C" {

size_t hv_encoder_workspace_bytes(int n_layers, const int* dims) { return hv::shape_ok(n_layers, dims) ? hv::kImageBytes : 0; }

int hv_encoder_pack_weights(const float* const* weights, int n_layers, const int* dims, void* workspace, size_t workspace_bytes,
                            void* stream) {
  using namespace hv;
  if (!shape_ok(n_layers, dims)) {
    set_error("hv_encoder_pack_weights: no fused instantiation for this MLP shape (served: 768-512-256-128-32)");
    return HV_ERR_UNSUPPORTED;
  }
  if (!weights || !workspace) {
    set_error("hv_encoder_pack_weights: null pointer");
    return HV_ERR_NULL;
  }
  for (int l = 0; l < 4; ++l)
    if (!weights[l] || !aligned16(weights[l])) {
      set_error("hv_encoder_pack_weights: weight %d is null or not 16-byte aligned", l);
      return weights[l] ? HV_ERR_MISALIGNED : HV_ERR_NULL;
    }
  if (workspace_bytes < kImageBytes || !aligned16(workspace)) {
    set_error("hv_encoder_pack_weights: needs a 16-byte aligned workspace of %zu bytes (got %zu)", kImageBytes, workspace_bytes);
    return HV_ERR_WORKSPACE;
  }
  constexpr int units = static_cast<int>(kImageBytes / 16);
  enc_pack_weights_kernel<<<(units + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      weights[0], weights[1], weights[2], weights[3], static_cast<uint8_t*>(workspace));
  HV_CUDA_CHECK(cudaGetLastError());
  return HV_OK;
}

static int encoder_forward_impl(const char* who, const void* x, bool half_in, int64_t n, int n_layers, const int* dims,
                                const void* workspace, size_t workspace_bytes, int normalize, int precise_silu, float* z, void* stream) {
  using namespace hv;
  if (n < 0) {
    set_error("%s: bad row count %lld", who, static_cast<long long>(n));
    return HV_ERR_BAD_SHAPE;
  }
  if (!shape_ok(n_layers, dims)) {
    set_error("%s: no fused instantiation for this MLP shape (served: 768-512-256-128-32)", who);
    return HV_ERR_UNSUPPORTED;
  }
  if (n == 0) return HV_OK;
  if (!x || !workspace || !z) {
    set_error("%s: null pointer", who);
    return HV_ERR_NULL;
  }
  if (!aligned16(x) || !aligned16(z) || !aligned16(workspace)) {
    set_error("%s: x, z and the weight image must be 16-byte aligned", who);
    return HV_ERR_MISALIGNED;
  }
  if (workspace_bytes < kImageBytes) {
    set_error("%s: the weight image is %zu bytes (got %zu)", who, kImageBytes, workspace_bytes);
    return HV_ERR_WORKSPACE;
  }
  DeviceProps props;
  if (int st = device_props(&props)) return st;
  if (props.cc_major != 10) {
    set_error("%s: built for sm_100a, current device is sm_%d%d", who, props.cc_major, props.cc_minor);
    return HV_ERR_UNSUPPORTED;
  }
  const int64_t n_tiles = (n + kTileRows - 1) / kTileRows;
  const unsigned grid = static_cast<unsigned>(n_tiles < props.sm_count ? n_tiles : props.sm_count);
  EncArgs a{x, static_cast<const uint8_t*>(workspace), z, n, normalize};
  auto go = [&](auto kernel) -> int {
    if (int st = prepare_kernel(kernel, 0, kSmemBytes)) return st;
    kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(a);
    HV_CUDA_CHECK(cudaGetLastError());
    return HV_OK;
  };
  if (half_in) return precise_silu ? go(enc_mlp_kernel<true, true>) : go(enc_mlp_kernel<false, true>);
  return precise_silu ? go(enc_mlp_kernel<true, false>) : go(enc_mlp_kernel<false, false>);
}

int hv_encoder_forward(const float* x, int64_t n, int n_layers, const int* dims, const void* workspace, size_t workspace_bytes,
                       int normalize, int precise_silu, float* z, void* stream) {
  return encoder_forward_impl("hv_encoder_forward", x, false, n, n_layers, dims, workspace, workspace_bytes, normalize, precise_silu, z,
                              stream);
}

int hv_encoder_forward_f16(const void* x_f16, int64_t n, int n_layers, const int* dims, const void* workspace, size_t workspace_bytes,
                           int normalize, int precise_silu, float* z, void* stream) {
  return encoder_forward_impl("hv_encoder_forward_f16", x_f16, true, n, n_layers, dims, workspace, workspace_bytes, normalize,
                              precise_silu, z, stream);
}

}  // extern "C"
