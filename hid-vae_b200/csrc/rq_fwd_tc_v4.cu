// tcgen05 kernel of the fused L-level residual quantiser with STREAMED operand images (generation 4): D = 16 / 32 / 64,
// any K -- every shape generation 11 (rq_fwd_tc_v11.cu: D = 32, K <= 256, resident images) does not serve, notably config
// C4 (D = 64, K = 4096).  Consumes the operand images written by rq_pack.cu (256-code N tiles).
//
// Mapping (re-derived for B200 and revised after the first ncu captures, profiles/):
//   GEMM  M = 128 rows of one row tile (= the 128 TMEM lanes), N = up to 256 codes, reduction = D.
//   score[row, k] = r.c_k - |c_k|^2 / 2   (argmax == argmin of |r|^2 + |c_k|^2 - 2 r.c_k; |r|^2 is row-constant)
//   fp32-grade scores from bf16 tensor cores: r = r_hi + r_lo, c = c_hi + c_lo (bf16 each), three products
//   r_hi.c_hi + r_lo.c_hi + r_hi.c_lo accumulated in the fp32 TMEM accumulator, and -|c|^2/2 folded in as one
//   more K=16 step (A = [1,1,1,0..], B = the norm split in three bf16 pieces).  3*D/16 + 1 tcgen05.mma per unit.
//   The codebooks are pre-packed once (hv_rq_pack_codebooks) into the UMMA K-major core-matrix layout and staged
//   by 1-D bulk TMA copies: resident in shared memory for all L levels when they fit (K=256, D=32, L=3: 120 KB),
//   otherwise streamed through a ring of stages (K=4096, D=64).
//
//   A row's L levels are a strictly serial chain (stage A -> MMA -> argmax scan -> code gather -> residual), each
//   link latency-bound, so throughput comes from the NUMBER OF ROW TILES IN FLIGHT per SM: the persistent CTA runs
//   NWG epilogue warpgroups (4 where shared memory allows, else 2), each owning one 128-row tile and one
//   512/NWG-column fp32 accumulator in TMEM; a level's N tile is processed in units of at most that many columns.
//   One more warp is the TMA producer, one allocates TMEM and issues the tcgen05.mma -- resident images: serving whichever
//   warpgroup is ready first; streamed images (C4): one issuer warp PER warpgroup, taking turns, so that an issuer's
//   bookkeeping between two units overlaps the other issuer's MMAs (tcgen05.mma blocks its thread while the queue is
//   full) and the scan of one warpgroup's unit always hides behind the other warpgroup's MMAs.
//   Epilogue thread t owns row t of its tile for all L levels: a whole 256-column unit is scanned by the branch-free
//   2-D fold of rq_rows.cuh (other unit widths: chunk by chunk with tcgen05.ld 32x32b), the running (max, argmax) stays
//   in registers, then the warp gathers the fp32 code rows,
//   forms emb_out / loss / the next residual in registers and re-stages the residual (bf16 hi/lo) as the next
//   level's A operand.  The [N, K] score matrix never leaves the SM.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "rq_rows.cuh"

namespace hv {
namespace {

constexpr int kTileRows = 128;
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 16;
constexpr int kMaxWg = 4;
constexpr int kOnesBytes = 2 * kTileRows * 16;  // one K=16 step of the A operand: [2 chunks][128 rows][8 bf16]
constexpr int kSmemLimit = 227 * 1024;

struct TcPlan {
  int ntile;       // codes per N tile (multiple of 32, <= 256)
  int n_ktiles;    // N tiles per level
  int tile_bytes;  // packed image of one (level, N tile)
  int stages;      // shared-memory stages for packed images
  int resident;    // 1: every (level, tile) image has its own stage and is loaded once
  int smem_bytes;
  int a_bytes;     // one warpgroup's A operand (hi + lo) == one fp32 row tile
  int n_wg;        // epilogue warpgroups = row tiles in flight per CTA
};

bool plan_for(int d, int k, int n_levels, int n_wg, TcPlan* p) {
  p->ntile = 256;  // image format of rq_pack.cu: every N tile holds 256 codes (padded codes can never win)
  p->n_ktiles = (k + p->ntile - 1) / p->ntile;
  p->tile_bytes = p->ntile * (4 * d + 32);
  p->a_bytes = kTileRows * d * 4;
  p->n_wg = n_wg;
  const int fixed = n_wg * p->a_bytes + kOnesBytes + 1024;
  const int budget = kSmemLimit - fixed;
  const int total_tiles = n_levels * p->n_ktiles;
  if (total_tiles <= kMaxStages && static_cast<long long>(total_tiles) * p->tile_bytes <= budget) {
    p->resident = 1;
    p->stages = total_tiles;
  } else {
    p->resident = 0;
    p->stages = budget / p->tile_bytes;
    if (p->stages > 4) p->stages = 4;
    if (p->stages < 2) return false;
  }
  p->smem_bytes = fixed + p->stages * p->tile_bytes;
  return true;
}

// The packed image (ntile, tile_bytes) does not depend on n_wg, so pack and forward always agree.
bool make_plan(int d, int k, int n_levels, TcPlan* p) {
  if ((d != 16 && d != 32 && d != 64) || k < 1 || n_levels < 1) return false;
  // four tiles in flight when the whole operand image stays resident beside four A buffers; else two.
  if (plan_for(d, k, n_levels, 4, p) && p->resident) return true;
  return plan_for(d, k, n_levels, 2, p);
}

struct TcParams {
  const uint8_t* packed;
  int ntile;
  int n_ktiles;
  int tile_bytes;
  int stages;
  int resident;
  int a_bytes;
  int tiles_per_cta;  // active warpgroups (1..NWG): fewer when there are not enough row tiles to fill the SMs
};

// bf16 hi/lo split of one row into the K-major core-matrix layout: chunk kc of row `row` lives at
// base + kc * (128 rows * 16 B) + row * 16.
template <int D>
__device__ __forceinline__ void stage_a_operand(uint8_t* a_hi, uint8_t* a_lo, int row, const float (&r)[D]) {
#pragma unroll
  for (int kc = 0; kc < D / 8; ++kc) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x0 = r[kc * 8 + 2 * j], x1 = r[kc * 8 + 2 * j + 1];
      const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
      const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __low2float(h), x1 - __high2float(h));
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(a_hi + kc * (kTileRows * 16) + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a_lo + kc * (kTileRows * 16) + row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));  // FMNMX3: one ALU-pipe op for two compares
  return d;
}

// max of 32 floats as a 3-ary tree: 17 independent-ish FMNMX3 instead of a 32-long dependent compare/select chain
__device__ __forceinline__ float max32(const float (&f)[32]) {
  float a[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) a[i] = max3(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
  a[10] = fmaxf(f[30], f[31]);
  const float b0 = max3(a[0], a[1], a[2]), b1 = max3(a[3], a[4], a[5]), b2 = max3(a[6], a[7], a[8]);
  const float b3 = fmaxf(a[9], a[10]);
  return fmaxf(max3(b0, b1, b2), b3);
}

// Running (max, argmax) over one 32-column chunk of scores held in registers; the first (lowest) index wins exact
// ties, like torch.min on the distances (modules/quantize.py:122).
//   phase A  m = max of the chunk (FMNMX3 tree, ALU pipe)
//   phase B  only if some row of the warp improves: position of the first element equal to m, computed on the FMA
//            pipe so it does not compete with phase A:  t_j = (f_j - m) * 2^120 + (32 - j)  is (32 - j) where
//            f_j == m and hugely negative elsewhere; the max of t_j therefore names the first maximiser.
//            (exact for any two scores that differ by at least 2^-114.)  t overwrites f: no extra registers.
__device__ __forceinline__ void scan_chunk(uint32_t (&v)[32], int base, float& best, int& best_k) {
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  const float m = max32(f);
  const bool better = m > best;  // strict: an earlier chunk keeps exact ties
  if (__any_sync(0xffffffffu, better)) {
    const float kBig = 1.329227995784916e36f;  // 2^120
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaf(f[j] - m, kBig, static_cast<float>(32 - j));
    const int loc = 32 - static_cast<int>(max32(f));
    if (better) {
      best = m;
      best_k = base + loc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Warp-cooperative row movers.  One thread owns one row (D fp32 = D/4 16-byte units), but a warp-wide 128-bit
// access in which every lane touches a different row costs 32 L1 wavefronts; so rows travel between global memory
// and their owner threads through a swizzled transpose in shared memory: global side = each row handled by D/4
// adjacent lanes (4 wavefronts per instruction at D = 32), owner side = conflict-free 128-bit shared accesses.
// The scratch is the warpgroup's own A-operand buffer (128 rows x D x 4 bytes), free whenever no MMA is reading
// it; warp q only touches the slots of its rows [32q, 32q+32), so __syncwarp is the only synchronisation.
// ---------------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ uint32_t xpose_addr(uint32_t base, int row, int u) {
  return base + u * (kTileRows * 16) + ((row & ~7) << 4) + (((row ^ u) & 7) << 4);
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// dst <- row `lane` of the warp's 32 rows; src_of(lr) is the global address of local row lr (nullptr = zeros).
template <int D, typename SrcOf>
__device__ __forceinline__ void warp_load_rows(float (&dst)[D], uint32_t scratch, int row0, int lane, SrcOf src_of) {
  constexpr int U = D / 4, RPI = 32 / U;
  static_assert(U <= 32 && 32 % U == 0, "row must be 1..32 16-byte units");
#pragma unroll
  for (int g = 0; g < 32 / RPI; ++g) {
    const int lr = g * RPI + lane / U, u = lane % U;
    const float* src = src_of(lr);
    const float4 v = src != nullptr ? __ldg(reinterpret_cast<const float4*>(src) + u) : make_float4(0.f, 0.f, 0.f, 0.f);
    sts128(xpose_addr<D>(scratch, row0 + lr, u), v);
  }
  __syncwarp();
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float4 v = lds128(xpose_addr<D>(scratch, row0 + lane, u));
    dst[4 * u] = v.x, dst[4 * u + 1] = v.y, dst[4 * u + 2] = v.z, dst[4 * u + 3] = v.w;
  }
  __syncwarp();
}

// row `lane` (src) -> global; dst_of(lr) is the global address of local row lr (nullptr = skip).
template <int D, typename DstOf>
__device__ __forceinline__ void warp_store_rows(const float (&src)[D], uint32_t scratch, int row0, int lane, DstOf dst_of) {
  constexpr int U = D / 4, RPI = 32 / U;
#pragma unroll
  for (int u = 0; u < U; ++u)
    sts128(xpose_addr<D>(scratch, row0 + lane, u), make_float4(src[4 * u], src[4 * u + 1], src[4 * u + 2], src[4 * u + 3]));
  __syncwarp();
#pragma unroll
  for (int g = 0; g < 32 / RPI; ++g) {
    const int lr = g * RPI + lane / U, u = lane % U;
    float* dst = dst_of(lr);
    const float4 v = lds128(xpose_addr<D>(scratch, row0 + lr, u));
    if (dst != nullptr) reinterpret_cast<float4*>(dst)[u] = v;
  }
  __syncwarp();
}

#ifdef HV_TC_INSTRUMENT
// clock64 timeline of block 0 (tools/ts_show.py): who 0 = lane 0 of the first epilogue warp, who 1 = the MMA issuer
__device__ uint32_t g_ts4[2][2 * 512];
__device__ uint32_t g_ts4_n[2];
#define HV4_STAMP(who, code)                                              \
  do {                                                                    \
    if (blockIdx.x == 0 && ts_n < 500) {                                  \
      g_ts4[who][2 * ts_n] = (code);                                      \
      g_ts4[who][2 * ts_n + 1] = static_cast<uint32_t>(clock64());        \
      ++ts_n;                                                             \
    }                                                                     \
  } while (0)
#else
#define HV4_STAMP(who, code) do { } while (0)
#endif

template <int NWG>
struct Roles {
  static constexpr int kEpiWarps = NWG * 4;
  static constexpr int kProducerWarp = NWG * 4;
  static constexpr int kMmaWarp = NWG * 4 + 1;
  // the helper warpgroup (TMA producer, MMA issuer, two idle warps) gives its registers to the epilogue warpgroups
  static constexpr int kThreads = (NWG * 4 + 4) * 32;
  // setmaxnreg only MOVES registers inside the CTA's launch allocation (kLaunchRegs per thread, what ptxas assigns
  // under __launch_bounds__(kThreads, 1)); asking for more leaves some warps spinning in the allocation forever.
  // launch_wg() checks kLaunchRegs against cudaFuncGetAttributes before every launch.
  static constexpr int kWarps = NWG * 4 + 4;
  static constexpr int kLaunchRegs = NWG == 4 ? 96 : 168;
  static constexpr int kHelperRegs = 56;
  static constexpr int kEpiRegs = ((kLaunchRegs * kWarps - 4 * kHelperRegs) / (NWG * 4)) / 8 * 8;
  static_assert(kEpiRegs * NWG * 4 + kHelperRegs * 4 <= kLaunchRegs * kWarps, "register hand-over exceeds the launch allocation");
  static constexpr int kAccCols = kTmemCols / NWG;  // fp32 accumulator columns of one warpgroup
};

// issue the 3*D/16 + 1 MMAs of one unit (`ncols` codes starting at code `col0` of the staged N tile).
// Called by ONE elected thread.  Descriptors are assembled from 32-bit words so that stepping through K chunks is
// one add per operand (address field in 16-byte units: A chunk = 128 rows x 16 B = 128 units, B chunk = ntile).
template <int D>
__device__ __forceinline__ void issue_unit(uint32_t acc, uint32_t a_hi, uint32_t a_lo, uint32_t ones, uint32_t b_tile,
                                           int ntile, int col0, int ncols, uint32_t bar_full) {
  const uint32_t idesc = ptx::umma_idesc_bf16(kTileRows, ncols);
  const uint32_t hi = ptx::umma_desc_hi(128);
  const uint32_t chunk_b = ntile * 16;                            // bytes between K chunks of the B image
  const uint32_t a_step = 2 * kTileRows, b_step = 2 * ntile;     // one K=16 step = two chunks, in 16-byte units
  const uint32_t d_ahi = ptx::umma_desc_lo(a_hi, kTileRows * 16), d_alo = ptx::umma_desc_lo(a_lo, kTileRows * 16);
  const uint32_t d_one = ptx::umma_desc_lo(ones, kTileRows * 16);
  const uint32_t b_hi = b_tile + col0 * 16;
  const uint32_t d_bhi = ptx::umma_desc_lo(b_hi, chunk_b);
  const uint32_t d_blo = ptx::umma_desc_lo(b_hi + (D / 8) * chunk_b, chunk_b);
  const uint32_t d_bnrm = ptx::umma_desc_lo(b_hi + 2 * (D / 8) * chunk_b, chunk_b);
#pragma unroll
  for (int j = 0; j < D / 16; ++j)  // r_hi . c_hi
    ptx::umma_bf16(acc, ptx::umma_desc(d_ahi + j * a_step, hi), ptx::umma_desc(d_bhi + j * b_step, hi), idesc, j > 0 ? 1u : 0u);
#pragma unroll
  for (int j = 0; j < D / 16; ++j)  // r_lo . c_hi
    ptx::umma_bf16(acc, ptx::umma_desc(d_alo + j * a_step, hi), ptx::umma_desc(d_bhi + j * b_step, hi), idesc, 1u);
#pragma unroll
  for (int j = 0; j < D / 16; ++j)  // r_hi . c_lo
    ptx::umma_bf16(acc, ptx::umma_desc(d_ahi + j * a_step, hi), ptx::umma_desc(d_blo + j * b_step, hi), idesc, 1u);
  ptx::umma_bf16(acc, ptx::umma_desc(d_one, hi), ptx::umma_desc(d_bnrm, hi), idesc, 1u);  // 1 * (-|c|^2 / 2)
  ptx::umma_commit(bar_full);
}

template <int D, bool ROT, int NWG>
__global__ void __launch_bounds__(Roles<NWG>::kThreads, 1) rq_fwd_tc_kernel(RqFwdArgs a, TcParams p) {
  using R = Roles<NWG>;
  extern __shared__ __align__(1024) uint8_t smem[];
  // [A wg0 (hi | lo) | ... | A wg(NWG-1) | ones | barriers (1 KB) | B stages ...]
  uint8_t* s_a = smem;
  uint8_t* s_ones = smem + NWG * p.a_bytes;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_ones + kOnesBytes);
  uint8_t* s_b = s_ones + kOnesBytes + 1024;

  uint64_t* bar_b_full = s_bar;                       // [kMaxStages]
  uint64_t* bar_b_empty = s_bar + kMaxStages;         // [kMaxStages]
  uint64_t* bar_acc_full = s_bar + 2 * kMaxStages;    // [kMaxWg]  MMA -> epilogue (tcgen05.commit needs an mbarrier)
  // epilogue -> MMA issuer progress counters, polled by ONE thread: a plain LDS answers in ~30 cycles where an
  // mbarrier probe takes ~150, and the issuer has up to 2 x NWG conditions to watch
  uint32_t* cnt_a_ready = reinterpret_cast<uint32_t*>(bar_acc_full + kMaxWg);  // [kMaxWg] warp arrivals: 4 per staged level
  uint32_t* cnt_acc_empty = cnt_a_ready + kMaxWg;                              // [kMaxWg] warp arrivals: 4 per drained unit
  uint32_t* s_tmem = cnt_acc_empty + kMaxWg;
  uint32_t* s_turn = s_tmem + 1;  // streamed images: units issued so far, in the global order tile 0 wg 0, tile 0 wg 1, tile 1 wg 0, ...

  const int warp = ptx::warp_index();
  const int lane = threadIdx.x & 31;

  if (warp == R::kProducerWarp && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bar_b_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_b_empty[s]), p.resident ? 1 : p.tiles_per_cta);  // (streamed: one commit per issuer)
    }
    *s_turn = 0;
    for (int w = 0; w < kMaxWg; ++w) {
      ptx::mbar_init(ptx::smem_u32(&bar_acc_full[w]), 1);
      cnt_a_ready[w] = 0;
      cnt_acc_empty[w] = 0;
    }
    ptx::fence_mbar_init();
  }
  if (warp == R::kMmaWarp) {
    ptx::tmem_alloc(ptx::smem_u32(s_tmem), kTmemCols);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x < kTileRows) {
    // constant A block that multiplies the norm pieces: row -> [1, 1, 1, 0, 0, 0, 0, 0 | 0 x 8]
    const uint32_t one2 = 0x3F803F80u;  // bf16 (1.0, 1.0)
    *reinterpret_cast<uint4*>(s_ones + threadIdx.x * 16) = make_uint4(one2, 0x00003F80u, 0u, 0u);
    *reinterpret_cast<uint4*>(s_ones + kTileRows * 16 + threadIdx.x * 16) = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  const int64_t n_row_tiles = (a.n + kTileRows - 1) / kTileRows;
  const int tpc = p.tiles_per_cta;
  const int64_t n_groups = (n_row_tiles + tpc - 1) / tpc;  // one CTA iteration handles a group of tpc row tiles
  const int total_tiles = a.n_levels * p.n_ktiles;
  const int units_per_tile = (p.ntile + R::kAccCols - 1) / R::kAccCols;

  if (warp < R::kEpiWarps) {
    ptx::setmaxnreg_inc<R::kEpiRegs>();  // registers handed over by the helper warpgroup below
    if ((warp >> 2) < tpc) {
    // ===================================== epilogue warpgroups ============================================
    const int w = warp >> 2;                       // warpgroup = which row tile of the group / which accumulator
    const int row_in_tile = threadIdx.x - w * kTileRows;
    const int quarter = warp & 3;                  // TMEM lanes [32*quarter, 32*quarter + 32)
    uint8_t* a_hi = s_a + w * p.a_bytes;
    uint8_t* a_lo = a_hi + p.a_bytes / 2;
    const uint32_t acc_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + w * R::kAccCols;
    const uint32_t cnt_ready = ptx::smem_u32(&cnt_a_ready[w]);
    const uint32_t bar_full = ptx::smem_u32(&bar_acc_full[w]);
    const uint32_t cnt_empty = ptx::smem_u32(&cnt_acc_empty[w]);
    uint32_t acc_phase = 0;
    const uint32_t scratch = ptx::smem_u32(a_hi);   // transpose scratch = this warpgroup's A buffer (see above)
    const int row0 = quarter * 32;                  // first tile row of this warp

    [[maybe_unused]] int ts_n = threadIdx.x == 0 ? 0 : 100000;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
      HV4_STAMP(0, 1);
      const int64_t warp_row0 = (tpc * grp + w) * kTileRows + row0;  // global row of the warp's local row 0
      const int64_t row = warp_row0 + lane;
      const bool valid = row < a.n;
      float r[D];
      warp_load_rows<D>(r, scratch, row0, lane,
                        [&](int lr) { return warp_row0 + lr < a.n ? a.x + (warp_row0 + lr) * D : nullptr; });
      float total_loss = 0.f;
      for (int l = 0; l < a.n_levels; ++l) {
        if (a.residuals != nullptr) {
          float* base = a.residuals + static_cast<int64_t>(l) * a.n * D;
          warp_store_rows<D>(r, scratch, row0, lane,
                             [&](int lr) { return warp_row0 + lr < a.n ? base + (warp_row0 + lr) * D : nullptr; });
        }
        HV4_STAMP(0, 2);
        stage_a_operand<D>(a_hi, a_lo, row_in_tile, r);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::counter_add_release(cnt_ready, 1);
        HV4_STAMP(0, 3);

        float best = -INFINITY;
        int best_k = 0;
        for (int t = 0; t < p.n_ktiles; ++t) {
          for (int u = 0; u < units_per_tile; ++u) {
            const int col0 = u * R::kAccCols;
            const int n_chunks = min(R::kAccCols, p.ntile - col0) / 32;
            ptx::mbar_wait(bar_full, acc_phase);
            acc_phase ^= 1;
            ptx::tc_fence_after_sync();
            if (t < 3 || t + 1 == p.n_ktiles) HV4_STAMP(0, 5);
            if (R::kAccCols == 256 && n_chunks == 8) {
              // A whole 256-column unit in one branch-free 2-D fold (rq_rows.cuh: pipelined 8-column loads, maxima per column
              // class and per chunk, the maximiser located once per unit).  The chunk-by-chunk form below votes and, early in a
              // level, re-scans most chunks: 2100 cycles per unit against the 1664 the other warpgroup's MMAs take, which left
              // the tensor pipe idle for 30 % of a tile (clock64 timeline, tools/v4_timeline.sh).
              float m;
              int c;
              rows::scan_x8_pairs<32>(acc_addr, m, c);
              if (m > best) {  // strict: an earlier unit keeps exact ties
                best = m;
                best_k = t * p.ntile + col0 + c;
              }
            } else {
              for (int c = 0; c < n_chunks; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(acc_addr + c * 32, v);
                ptx::tmem_wait_ld(v);
                scan_chunk(v, t * p.ntile + col0 + c * 32, best, best_k);
              }
            }
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::counter_add_release(cnt_empty, 1);
            if (t < 3 || t + 1 == p.n_ktiles) HV4_STAMP(0, 6);
          }
        }
        best_k = min(best_k, a.k - 1);

        // every MMA of this level has completed (last acc_full observed): the A buffer is free to be the scratch
        const float* cb = a.codebooks + static_cast<int64_t>(l) * a.k * D;
        float e[D], o[D];
        warp_load_rows<D>(e, scratch, row0, lane, [&](int lr) {
          return cb + static_cast<int64_t>(__shfl_sync(0xffffffffu, best_k, lr)) * D;
        });
        HV4_STAMP(0, 7);
        const float ll = rq_level_tail_o<D, ROT>(r, e, a.beta, o);
        if (a.emb_out != nullptr) {
          float* base = a.emb_out + static_cast<int64_t>(l) * a.n * D;
          warp_store_rows<D>(o, scratch, row0, lane,
                             [&](int lr) { return warp_row0 + lr < a.n ? base + (warp_row0 + lr) * D : nullptr; });
        }
        total_loss += ll;
        if (valid) {
          a.ids[row * a.ids_row_stride + l * a.ids_level_stride] = best_k;
          if (a.level_loss != nullptr) a.level_loss[static_cast<int64_t>(l) * a.n + row] = ll;
        }
      }
      HV4_STAMP(0, 8);
      if (valid && a.loss != nullptr) a.loss[row] = total_loss;
      if (a.final_residual != nullptr)
        warp_store_rows<D>(r, scratch, row0, lane,
                           [&](int lr) { return warp_row0 + lr < a.n ? a.final_residual + (warp_row0 + lr) * D : nullptr; });
    }
#ifdef HV_TC_INSTRUMENT
    if (blockIdx.x == 0 && threadIdx.x == 0) g_ts4_n[0] = ts_n;
#endif
    }  // else: warpgroup without a row tile (small N: fewer tiles per CTA so that more SMs work)
  } else {
    ptx::setmaxnreg_dec<R::kHelperRegs>();
    if (warp == R::kProducerWarp) {
    // ===================================== TMA producer ===================================================
    if (lane == 0) {
      if (p.resident) {
        if (static_cast<int64_t>(blockIdx.x) < n_groups) {
          for (int s = 0; s < total_tiles; ++s) {
            const uint32_t bar = ptx::smem_u32(&bar_b_full[s]);
            ptx::mbar_arrive_expect_tx(bar, p.tile_bytes);
            ptx::bulk_g2s(ptx::smem_u32(s_b + static_cast<size_t>(s) * p.tile_bytes),
                          p.packed + static_cast<size_t>(s) * p.tile_bytes, p.tile_bytes, bar);
          }
        }
      } else {
        uint32_t it = 0;
        for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
          for (int tile = 0; tile < total_tiles; ++tile, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            ptx::mbar_wait(ptx::smem_u32(&bar_b_empty[s]), ph ^ 1);
            const uint32_t bar = ptx::smem_u32(&bar_b_full[s]);
            ptx::mbar_arrive_expect_tx(bar, p.tile_bytes);
            ptx::bulk_g2s(ptx::smem_u32(s_b + static_cast<size_t>(s) * p.tile_bytes),
                          p.packed + static_cast<size_t>(tile) * p.tile_bytes, p.tile_bytes, bar);
          }
        }
      }
    }
    } else if (warp == R::kMmaWarp || (!p.resident && warp == R::kMmaWarp + 1 && p.tiles_per_cta > 1)) {
    // ===================================== MMA issuer =====================================================
    const uint32_t ones = ptx::smem_u32(s_ones);
    if (p.resident) {
      // Every operand image is resident: the warpgroups are independent, so serve whichever one is ready
      // (its residual staged and its accumulator drained) instead of a fixed round that blocks on the slowest.
      if (static_cast<int64_t>(blockIdx.x) < n_groups)
        for (int s = 0; s < total_tiles; ++s) ptx::mbar_wait(ptx::smem_u32(&bar_b_full[s]), 0);
      const int64_t my_groups = static_cast<int64_t>(blockIdx.x) < n_groups
                                    ? (n_groups - 1 - blockIdx.x) / gridDim.x + 1 : 0;
      const int units_per_level = p.n_ktiles * units_per_tile;
      const uint32_t units_per_wg = static_cast<uint32_t>(my_groups * a.n_levels * units_per_level);
      {
        // The whole warp runs the scheduler with warp-uniform state (lane 0's view of the counters decides); the
        // MMAs and the commit are issued by one elected lane.
        uint32_t done[NWG];         // units issued so far
        uint32_t in_level[NWG];     // position of the next unit inside its level
        uint32_t level[NWG];        // level of the next unit
        uint32_t levels_seen[NWG];  // number of staged residuals already consumed
#pragma unroll
        for (int w = 0; w < NWG; ++w) done[w] = in_level[w] = level[w] = levels_seen[w] = 0;
        int remaining = tpc;
        long long idle_since = 0;
        while (remaining > 0) {
          bool progressed = false;
#pragma unroll
          for (int w = 0; w < NWG; ++w) {
            if (w >= tpc || done[w] >= units_per_wg) continue;
            // residual of this unit's level staged by all 4 warps?  accumulator drained of every earlier unit?
            const uint32_t ready = __shfl_sync(0xffffffffu, ptx::counter_ld_acquire(ptx::smem_u32(&cnt_a_ready[w])), 0);
            if (static_cast<int32_t>(ready - 4 * (levels_seen[w] + 1)) < 0) continue;
            const uint32_t drained = __shfl_sync(0xffffffffu, ptx::counter_ld_acquire(ptx::smem_u32(&cnt_acc_empty[w])), 0);
            if (static_cast<int32_t>(drained - 4 * done[w]) < 0) continue;
            ptx::tc_fence_after_sync();
            if (ptx::elect_one()) {
              const int t = in_level[w] / units_per_tile, u = in_level[w] % units_per_tile;
              const int col0 = u * R::kAccCols;
              const int ncols = min(R::kAccCols, p.ntile - col0);
              const uint32_t a_hi = ptx::smem_u32(s_a + w * p.a_bytes);
              issue_unit<D>(tmem_base + w * R::kAccCols, a_hi, a_hi + p.a_bytes / 2, ones,
                            ptx::smem_u32(s_b + static_cast<size_t>(level[w] * p.n_ktiles + t) * p.tile_bytes), p.ntile, col0,
                            ncols, ptx::smem_u32(&bar_acc_full[w]));
            }
            __syncwarp();
            done[w]++;
            if (++in_level[w] == static_cast<uint32_t>(units_per_level)) {  // next unit opens a new level
              in_level[w] = 0;
              level[w] = level[w] + 1 == static_cast<uint32_t>(a.n_levels) ? 0 : level[w] + 1;
              levels_seen[w]++;
            }
            if (done[w] >= units_per_wg) remaining--;
            progressed = true;
          }
          if (progressed) {
            idle_since = 0;
          } else {
            if (idle_since == 0) idle_since = clock64();
            if (clock64() - idle_since > 4000000000LL) {
              if (lane == 0) printf("hidvae_b200: MMA scheduler starved (block %d)\n", blockIdx.x);
              __trap();
            }
          }
        }
      }
      __syncwarp();
    } else {
      // Streamed operand images: all warpgroups consume the same stage in lock step (one load serves tpc tiles), unit after
      // unit in the order tile t wg 0, tile t wg 1, tile t + 1 wg 0 ...  ONE ISSUER WARP PER WARPGROUP: tcgen05.mma blocks its
      // thread while the queue is full, i.e. for about the execution time of the unit, and the bookkeeping between two units
      // (stage barrier, accumulator counter, descriptors, commit: ~1000 cycles measured) used to run when the queue had
      // drained -- the tensor pipe idled 22 % of a tile.  Now issuer w does its waiting while the other issuer's MMAs are
      // being pushed, and takes its turn (s_turn) the moment they are all queued.  (streamed plans run two warpgroups.)
      static_assert(kMaxWg >= 2, "two issuers");
      const int w = warp - R::kMmaWarp;          // this issuer's warpgroup
      const uint32_t n_issuers = static_cast<uint32_t>(tpc);
      uint32_t it = 0, levels_seen = 0, acc_uses = 0;
      [[maybe_unused]] int ts_n = (lane == 0 && w == 0) ? 0 : 100000;
      for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        for (int tile = 0; tile < total_tiles; ++tile, ++it) {
          const int t = tile % p.n_ktiles;
          const int s = it % p.stages;
          if (t < 3) HV4_STAMP(1, 20);
          ptx::mbar_wait(ptx::smem_u32(&bar_b_full[s]), (it / p.stages) & 1);
          if (t < 3) HV4_STAMP(1, 21);
          const uint32_t b_tile = ptx::smem_u32(s_b + static_cast<size_t>(s) * p.tile_bytes);
          if (t == 0) {
            levels_seen++;
            ptx::counter_wait(ptx::smem_u32(&cnt_a_ready[w]), 4 * levels_seen);
          }
          for (int u = 0; u < units_per_tile; ++u) {
            ptx::counter_wait(ptx::smem_u32(&cnt_acc_empty[w]), 4 * acc_uses);
            acc_uses++;
            ptx::counter_wait(ptx::smem_u32(s_turn), (it * units_per_tile + u) * n_issuers + w);  // my turn
            ptx::tc_fence_after_sync();
            if (t < 3) HV4_STAMP(1, 22 + w);
            if (ptx::elect_one()) {
              const int col0 = u * R::kAccCols;
              const uint32_t a_hi = ptx::smem_u32(s_a + w * p.a_bytes);
              issue_unit<D>(tmem_base + w * R::kAccCols, a_hi, a_hi + p.a_bytes / 2, ones, b_tile, p.ntile, col0,
                            min(R::kAccCols, p.ntile - col0), ptx::smem_u32(&bar_acc_full[w]));
            }
            __syncwarp();
            if (lane == 0) ptx::counter_add_release(ptx::smem_u32(s_turn), 1u);
          }
          if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(&bar_b_empty[s]));
          __syncwarp();
          if (t < 3) HV4_STAMP(1, 26);
        }
      }
#ifdef HV_TC_INSTRUMENT
      if (blockIdx.x == 0 && lane == 0 && w == 0) g_ts4_n[1] = ts_n;
#endif
    }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
#ifdef HV_TC_INSTRUMENT
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int who = 0; who < 2; ++who)
      for (uint32_t i = 0; i < g_ts4_n[who] && i < 500; ++i) printf("TS %d %u %u\n", who, g_ts4[who][2 * i], g_ts4[who][2 * i + 1]);
#endif
  if (warp == R::kMmaWarp) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int D, int NWG>
int launch_wg(const RqFwdArgs& a, bool rot, const TcPlan& plan, uint8_t* packed, const DeviceProps& props, cudaStream_t stream) {
  const int64_t n_row_tiles = (a.n + kTileRows - 1) / kTileRows;
  // as many row tiles per CTA as it takes to cover them with one CTA per SM, at most NWG
  int64_t tpc = (n_row_tiles + props.sm_count - 1) / props.sm_count;
  tpc = tpc < 1 ? 1 : (tpc > NWG ? NWG : tpc);
  const int64_t n_groups = (n_row_tiles + tpc - 1) / tpc;
  const unsigned grid = static_cast<unsigned>(n_groups < props.sm_count ? n_groups : props.sm_count);
  TcParams p{packed, plan.ntile, plan.n_ktiles, plan.tile_bytes, plan.stages, plan.resident, plan.a_bytes, static_cast<int>(tpc)};
  auto go = [&](auto kernel) -> int {
    if (int st = prepare_kernel(kernel, Roles<NWG>::kLaunchRegs, plan.smem_bytes)) return st;
    kernel<<<grid, Roles<NWG>::kThreads, plan.smem_bytes, stream>>>(a, p);
    HV_CUDA_CHECK(cudaGetLastError());
    return HV_OK;
  };
  return rot ? go(rq_fwd_tc_kernel<D, true, NWG>) : go(rq_fwd_tc_kernel<D, false, NWG>);
}

template <int D>
int launch_d(const RqFwdArgs& a, bool rot, const TcPlan& plan, uint8_t* packed, cudaStream_t stream) {
  DeviceProps props;
  if (int st = device_props(&props)) return st;
  return plan.n_wg == 4 ? launch_wg<D, 4>(a, rot, plan, packed, props, stream) : launch_wg<D, 2>(a, rot, plan, packed, props, stream);
}

}  // namespace

bool rq_fwd_tc_v4_supported(int d, int k, int n_levels) {
  TcPlan plan;
  return make_plan(d, k, n_levels, &plan);
}

// `packed` = image written by launch_rq_pack (rq_pack.cu)
int launch_rq_fwd_tc_v4(const RqFwdArgs& a, int d, bool rot, void* packed, cudaStream_t stream) {
  TcPlan plan;
  if (!make_plan(d, a.k, a.n_levels, &plan)) {
    set_error("hv_rq_forward: no v4 tcgen05 instantiation for D=%d K=%d L=%d", d, a.k, a.n_levels);
    return HV_ERR_UNSUPPORTED;
  }
  if (a.n == 0) return HV_OK;
  uint8_t* img = static_cast<uint8_t*>(packed);
  switch (d) {
    case 16: return launch_d<16>(a, rot, plan, img, stream);
    case 32: return launch_d<32>(a, rot, plan, img, stream);
    case 64: return launch_d<64>(a, rot, plan, img, stream);
    default: return HV_ERR_UNSUPPORTED;
  }
}

}  // namespace hv
