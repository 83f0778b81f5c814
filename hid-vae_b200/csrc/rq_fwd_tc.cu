// tcgen05 kernel of the fused L-level residual quantiser (HV_ALGO_TCGEN05) for sm_100a -- generation 5.
//
// What one (row tile, level) costs, measured on the previous generation (profiles/README.md): 7 tcgen05.mma of
// N = 256 (896 tensor cycles) against ~6200 warp instructions of epilogue (1550 issue cycles per SM sub-partition),
// and the epilogue warps spent a third of their time waiting for the accumulator.  This generation is built
// around those two numbers:
//
//   GEMM      M = 128 rows of a row tile (the 128 TMEM lanes), N = 256 codes (one operand image), reduction = D.
//             score[row, k] = r.c_k - |c_k|^2 / 2 (argmax == the reference's argmin, modules/quantize.py:108-122),
//             fp32-grade from bf16 tensor cores: r = r_hi + r_lo, c = c_hi + c_lo, products hi.hi + lo.hi + hi.lo and
//             the norm term (three bf16 pieces against a constant-ones A block) accumulate in fp32 TMEM:
//             3*D/16 + 1 tcgen05.mma (M128 N128 K16) per 128-code unit, two units per image.
//   SLOTS     the persistent CTA (one per SM) keeps TWO row tiles in flight ("slots"), each with a 256-column
//             accumulator (2 x 256 = the 512 TMEM columns) and EIGHT warps.  The slot issues its own MMAs (one
//             elected thread, straight after the slot-wide barrier that follows the A-operand staging; there is no
//             scheduler warp to poll), both units back to back with one tcgen05.commit each, so the scan of unit 0
//             overlaps the MMAs of unit 1 and the other slot's epilogue overlaps this slot's MMAs.
//   SCAN      thread (quarter q, half h, lane) owns accumulator row 32q + lane and 64 columns of each unit (four
//             32-column chunks, tcgen05.ld 32x32b).  Instead of an argmax per chunk it folds the chunks into 32
//             running column classes  g_j = max_c f[c][j]  (one FMNMX per score) and keeps the four chunk maxima
//             (3-ary FMNMX3 tree); the row maximum m is the largest chunk maximum, the chunk is the one whose maximum
//             equals m and the class is found on the FMA pipe:  sum_j sat((g_j - m) * 2^120 + 1) * (64 + j)  is
//             64 + j* when exactly one class attains m.  ~1.1 ALU-pipe and 0.8 FMA-pipe instructions per score
//             instead of 3.4.  If more than one class or chunk attains m (duplicate code rows, all-zero inputs) the
//             warp re-reads its chunks and takes the exact first-index path, so exact ties still resolve to the
//             lowest index like torch.min (modules/quantize.py:122).
//   ROW WORK  everything that is per row rather than per score (x load, code gather, STE / rotation value, loss,
//             residual update, bf16 hi/lo split of the next A operand, stores) runs in a row-cooperative layout: a
//             thread holds one 8-float K chunk of D/16 rows, so global accesses are coalesced 32-byte pieces, the A
//             operand is written with conflict-free 128-bit stores straight into the UMMA core-matrix layout and a
//             row costs D/8 lanes x 8 floats instead of one thread x D floats.  The two halves of a row meet through
//             a 4 KB candidate table in shared memory.
//   OPERANDS  codebook images are packed once (hv_rq_pack_codebooks) and staged by 1-D bulk TMA copies: resident in
//             shared memory for all levels when they fit (K = 256, D = 32, L = 3: 120 KB), otherwise streamed through
//             a ring of stages that both slots consume in lock step (K = 4096, D = 64).
//   The [N, K] score table never leaves the SM; only ids (and emb_out / loss in training) go to HBM.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace hv {
namespace {

constexpr int kTileRows = 128;
constexpr int kNTile = 256;                 // codes per operand image (= accumulator columns of a slot)
constexpr int kUnitCols = 128;              // N of one tcgen05.mma
constexpr int kMaxSlots = 4;                // row tiles in flight per CTA (Cfg<D>::NT of them)
#ifndef HV_TC_ROW_WARPS
#define HV_TC_ROW_WARPS 8
#endif
constexpr int kScanWarps = 8;               // scan group: 4 lane quarters x 2 column halves (warps 0..7)
constexpr int kRowWarps = HV_TC_ROW_WARPS;  // row group (warps 8..): 8, or 16 (more row tiles' worth of latency hiding)
constexpr int kThreads = (kScanWarps + kRowWarps) * 32;
// setmaxnreg split of the launch allocation (launch_bounds(kThreads, 1): 128 registers at 512 threads, 80 at 768)
constexpr int kLaunchRegs = kRowWarps == 8 ? 128 : 80;
constexpr int kRowRegs = kRowWarps == 8 ? 160 : 72, kScanRegs = 96;
static_assert(kRowRegs * kRowWarps + kScanRegs * kScanWarps <= kLaunchRegs * (kRowWarps + kScanWarps),
              "register hand-over must stay inside the launch allocation");
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 16;
constexpr int kOnesBytes = 2 * kTileRows * 16;              // one K=16 step of the A operand: [2 chunks][128 rows][8 bf16]
constexpr int kCandSlotBytes = 2 * kTileRows * 8;           // (max, column) per half, row of one slot
constexpr int kLossSlotBytes = kTileRows * 4;               // running loss per row of one slot
constexpr int kBarBytes = 4096;
constexpr int kSmemLimit = 227 * 1024;
#ifdef HV_TC_INSTRUMENT
constexpr bool kAblate = true;   // HIDVAE_TC_DEBUG bits 1/2/4 drop the scan / the gather / the MMAs (timing experiments)
#else
constexpr bool kAblate = false;
#endif

// row tiles in flight: four while their A operands fit beside the operand images, else two
template <int D>
struct Cfg {
  static constexpr int NT = D <= 32 ? 4 : 2;
};
inline int slots_for(int d) { return d <= 32 ? 4 : 2; }

struct TcPlan {
  int n_ktiles;    // operand images per level
  int tile_bytes;  // one packed image
  int stages;      // shared-memory stages for images
  int resident;    // 1: every (level, image) has its own stage and is loaded once
  int smem_bytes;
  int a_bytes;     // one slot's A operand (hi + lo) == one fp32 row tile
};

bool make_plan(int d, int k, int n_levels, TcPlan* p) {
  if ((d != 16 && d != 32 && d != 64) || k < 1 || n_levels < 1) return false;
  p->n_ktiles = (k + kNTile - 1) / kNTile;
  p->tile_bytes = kNTile * (4 * d + 32);
  p->a_bytes = kTileRows * d * 4;
  const int fixed = slots_for(d) * (p->a_bytes + kCandSlotBytes + kLossSlotBytes) + kOnesBytes + kBarBytes;
  const int budget = kSmemLimit - fixed;
  const long long total_tiles = static_cast<long long>(n_levels) * p->n_ktiles;
  if (total_tiles <= kMaxStages && total_tiles * p->tile_bytes <= budget) {
    p->resident = 1;
    p->stages = static_cast<int>(total_tiles);
  } else {
    p->resident = 0;
    p->stages = budget / p->tile_bytes;
    if (p->stages > 4) p->stages = 4;
    if (p->stages < 2) return false;
  }
  p->smem_bytes = fixed + p->stages * p->tile_bytes;
  return true;
}

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

struct TcParams {
  const uint8_t* packed;
  int n_ktiles;
  int tile_bytes;
  int stages;
  int resident;
  int a_bytes;
  int debug;  // HIDVAE_TC_DEBUG ablation mask (timing experiments only; results are then meaningless)
};

// ---------------------------------------------------------------------------------------------------------
// Row-cooperative layout: the 256 threads of a slot hold the slot's 128 x D fp32 rows as 8-float K chunks.
//   lane -> (j = lane % RPI, kc = lane / RPI);  warp wt owns rows [16 wt, 16 wt + 16);  the thread's g-th row is
//   16 wt + g * RPI + j.  A quarter warp (8 consecutive lanes) then shares kc and covers 8 consecutive rows, which
//   makes the 128-bit stores into the core-matrix layout (chunk kc of row `row` at kc * 2048 + row * 16) conflict free.
// ---------------------------------------------------------------------------------------------------------
template <int D>
struct Rc {
  static constexpr int KC = D / 8;     // 16-byte (8 x bf16) K chunks per row == lanes per row
  static constexpr int RPI = 32 / KC;  // rows per warp instruction
  static constexpr int RPT = KC / 2;   // rows per thread
  static_assert(KC >= 2 && KC <= 8 && RPI * RPT == 16, "row-cooperative layout covers 16 rows per warp");
};

// sum over the KC lanes that share a row (they differ in the lane bits above log2(RPI))
template <int D>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int m = Rc<D>::RPI; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// bf16 hi/lo split of 8 consecutive floats -> one 16-byte chunk entry each of the hi and the lo operand
__device__ __forceinline__ void split8(const float (&r)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float x0 = r[2 * j], x1 = r[2 * j + 1];
    const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
    const uint32_t hw = *reinterpret_cast<const uint32_t*>(&hh);
    const float h0 = __uint_as_float(hw << 16), h1 = __uint_as_float(hw & 0xFFFF0000u);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - h0, x1 - h1);
    h[j] = hw;
    l[j] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));  // FMNMX3: one ALU-pipe op for two compares
  return d;
}

// max of 32 floats as a 3-ary tree: 17 FMNMX3/FMNMX instead of a 32-long dependent chain
__device__ __forceinline__ float max32(const uint32_t (&v)[32]) {
  float a[11];
#pragma unroll
  for (int i = 0; i < 10; ++i)
    a[i] = max3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
  a[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
  const float b0 = max3(a[0], a[1], a[2]), b1 = max3(a[3], a[4], a[5]), b2 = max3(a[6], a[7], a[8]);
  const float b3 = fmaxf(a[9], a[10]);
  return fmaxf(max3(b0, b1, b2), b3);
}

constexpr float kBig = 1.329227995784916e36f;  // 2^120

// Exact first-index (max, argmax) of one 32-column chunk against the running pair -- the slow path, taken only by
// warps in which some row has more than one maximiser.  t_j = (f_j - m) * 2^120 + (32 - j) is 32 - j where f_j == m
// and hugely negative elsewhere, so its maximum names the first maximiser.
__device__ __forceinline__ void scan_chunk_exact(const uint32_t (&v)[32], int base, float& best, int& best_col) {
  const float m = max32(v);
  if (m > best) {  // strict: an earlier chunk keeps exact ties
    float t = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) t = fmaxf(t, fmaf(__uint_as_float(v[j]) - m, kBig, static_cast<float>(32 - j)));
    best = m;
    best_col = base + 32 - static_cast<int>(t);
  }
}

// issue the 3*D/16 + 1 MMAs of one unit (128 codes starting at code `col0` of the staged image) and commit them.
// Called by ONE elected thread.  Descriptors are assembled from 32-bit words so that stepping through K chunks is
// one add per operand (address field in 16-byte units: A chunk = 128 rows x 16 B = 128 units, B chunk = 256).
template <int D>
__device__ __forceinline__ void issue_unit(uint32_t acc, uint32_t a_hi, uint32_t a_lo, uint32_t ones, uint32_t b_tile,
                                           int col0, uint32_t bar_full) {
  constexpr uint32_t idesc = ptx::umma_idesc_bf16(kTileRows, kUnitCols);
  const uint32_t hi = ptx::umma_desc_hi(128);
  constexpr uint32_t chunk_b = kNTile * 16;                        // bytes between K chunks of the B image
  constexpr uint32_t a_step = 2 * kTileRows, b_step = 2 * kNTile;  // one K=16 step = two chunks, in 16-byte units
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

// Scan of one accumulator (256 columns, this thread's row, this thread's 64 columns of each unit); see the file
// header.  Returns the row maximum over the thread's 128 columns and its first column.
//   acc_addr  TMEM address of (lane quarter, slot accumulator column 0);  h = which 64 columns of each unit
__device__ __forceinline__ void scan_accumulator(uint32_t acc_addr, int h, uint32_t bar_u0, uint32_t bar_u1, uint32_t phase,
                                                 bool row_valid, float& m_out, int& col_out) {
  uint32_t g[32], v[32];
  const uint32_t c0 = acc_addr + 64 * h;
  ptx::mbar_wait(bar_u0, phase);
  ptx::tc_fence_after_sync();
  ptx::tmem_ld_32x32(c0, g);
  ptx::tmem_wait_ld(g);
  const float cm0 = max32(g);
  ptx::tmem_ld_32x32(c0 + 32, v);
  ptx::tmem_wait_ld(v);
  const float cm1 = max32(v);
#pragma unroll
  for (int j = 0; j < 32; ++j) g[j] = __float_as_uint(fmaxf(__uint_as_float(g[j]), __uint_as_float(v[j])));
  ptx::mbar_wait(bar_u1, phase);
  ptx::tc_fence_after_sync();
  ptx::tmem_ld_32x32(c0 + kUnitCols, v);
  ptx::tmem_wait_ld(v);
  const float cm2 = max32(v);
#pragma unroll
  for (int j = 0; j < 32; ++j) g[j] = __float_as_uint(fmaxf(__uint_as_float(g[j]), __uint_as_float(v[j])));
  ptx::tmem_ld_32x32(c0 + kUnitCols + 32, v);
  ptx::tmem_wait_ld(v);
  const float cm3 = max32(v);
#pragma unroll
  for (int j = 0; j < 32; ++j) g[j] = __float_as_uint(fmaxf(__uint_as_float(g[j]), __uint_as_float(v[j])));

  const float m = fmaxf(max3(cm0, cm1, cm2), cm3);
  // class of the maximiser on the FMA pipe: acc = sum_j [g_j == m] * (64 + j), four partial sums for ILP
  float acc4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float e = __saturatef(fmaf(__uint_as_float(g[j]) - m, kBig, 1.0f));
    acc4[j & 3] = fmaf(e, static_cast<float>(64 + j), acc4[j & 3]);
  }
  const float acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
  const bool e0 = cm0 == m, e1 = cm1 == m, e2 = cm2 == m, e3 = cm3 == m;
  const int n_chunks = static_cast<int>(e0) + static_cast<int>(e1) + static_cast<int>(e2) + static_cast<int>(e3);
  const int chunk_col = e0 ? 0 : e1 ? 32 : e2 ? kUnitCols : kUnitCols + 32;
  // exactly one class and one chunk attain m; a thread whose columns are all padding (m = -1e30) cannot win anyway
  const bool unique = (acc < 128.f && n_chunks == 1) || m < -1e29f;
  float best = m;
  int col = 64 * h + chunk_col + static_cast<int>(acc) - 64;
  if (__any_sync(0xffffffffu, !unique && row_valid)) {
    // more than one maximiser somewhere in this warp (duplicate code rows, all-zero input, ...): exact first index
    best = -INFINITY;
    col = 0;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int off = 64 * h + (c & 1) * 32 + (c >> 1) * kUnitCols;
      ptx::tmem_ld_32x32(acc_addr + off, v);
      ptx::tmem_wait_ld(v);
      scan_chunk_exact(v, off, best, col);
    }
  }
  m_out = best;
  col_out = col;
}

// ---------------------------------------------------------------------------------------------------------
// The kernel: a pipeline between two warp groups of the persistent CTA, all hand-overs are mbarriers.
//
//   row group  (warps 0-7)   split in NSLOT sub-groups of 8 / NSLOT warps; a sub-group owns one row tile ("slot") at a
//                            time in the row-cooperative layout and runs, per level:
//                              stage the residual as the slot's A operand (bf16 hi | lo)  ->  its first warp issues the
//                              level's tcgen05.mma into the next free accumulator (2 x 256 TMEM columns used
//                              alternately; a turn counter keeps the sub-groups in round-robin order)  ->  wait for the
//                              slot's scan  ->  merge the two half-row candidates, write the id, gather the chosen code
//                              row  ->  value / loss / next residual.
//                            The sub-groups are independent instruction streams, so one slot's global-load and MMA
//                            latencies are covered by the others.
//   scan group (warps 8-15)  takes the accumulators in issue order: 2-D fold argmax (scan_accumulator), candidate
//                            (max, column) per half row into shared memory, scan_done[slot] / acc_free[acc].
// ---------------------------------------------------------------------------------------------------------
template <int D, bool ROT, int NSLOT>
__global__ void __launch_bounds__(kThreads, 1) rq_fwd_tc_kernel(RqFwdArgs a, TcParams p) {
  using L = Rc<D>;
  constexpr int WPS = kRowWarps / NSLOT > 8 ? 8 : kRowWarps / NSLOT;  // warps per slot (at most 8: two rows per thread)
  constexpr int TPS = WPS * 32;                            // threads per slot
  constexpr int RPT = kTileRows * L::KC / TPS;             // rows per thread
  constexpr int GC = kRowWarps == 16 ? (RPT < 2 ? RPT : 2) : (RPT < 4 ? RPT : 4);  // rows gathered at a time (register budget)
  extern __shared__ __align__(1024) uint8_t smem[];
  // [A slot 0 (hi | lo) | ... | ones | candidates | loss | barriers | B stages ...]   (sized for kMaxSlots slots)
  constexpr int n_slots_smem = Cfg<D>::NT;
  uint8_t* s_a = smem;
  uint8_t* s_ones = smem + n_slots_smem * p.a_bytes;
  uint8_t* s_cand = s_ones + kOnesBytes;
  float* s_loss = reinterpret_cast<float*>(s_cand + n_slots_smem * kCandSlotBytes);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_cand + n_slots_smem * (kCandSlotBytes + kLossSlotBytes));
  uint8_t* s_b = reinterpret_cast<uint8_t*>(s_bar) + kBarBytes;

  uint64_t* bar_b_full = s_bar;                      // [kMaxStages]  TMA -> MMA issuers
  uint64_t* bar_b_empty = bar_b_full + kMaxStages;   // [kMaxStages]  MMA completion (one commit per slot) -> TMA
  uint64_t* bar_mma_done = bar_b_empty + kMaxStages; // [2 acc][2 units]  MMA completion -> scan group
  uint64_t* bar_acc_free = bar_mma_done + 4;         // [2 acc]       scan group (8 warps) -> MMA issuers
  uint64_t* bar_scan_done = bar_acc_free + 2;        // [kMaxSlots]   scan group (8 warps) -> the slot's sub-group
  uint32_t* s_turn = reinterpret_cast<uint32_t*>(bar_scan_done + kMaxSlots);  // number of accumulators issued so far
  uint32_t* s_tmem = s_turn + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t* s_ts = s_tmem + 1;  // [2][384] timestamps (HIDVAE_TC_DEBUG & 64, block 0)
  int ts_n = 0;
#ifdef HV_TC_INSTRUMENT
  const bool ts_on = (p.debug & 64) && blockIdx.x == 0 && lane == 0;
#else
  constexpr bool ts_on = false;  // make EXTRA=-DHV_TC_INSTRUMENT: clock64 timeline of block 0 (HIDVAE_TC_DEBUG & 64)
#endif
  auto stamp = [&](int who, int code) { if (ts_on && ts_n < 190) { s_ts[who * 384 + 2 * ts_n] = code; s_ts[who * 384 + 2 * ts_n + 1] = static_cast<uint32_t>(clock64()); ++ts_n; } };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bar_b_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_b_empty[s]), NSLOT);
    }
    for (int i = 0; i < 4; ++i) ptx::mbar_init(ptx::smem_u32(&bar_mma_done[i]), 1);
    for (int i = 0; i < 2; ++i) ptx::mbar_init(ptx::smem_u32(&bar_acc_free[i]), kScanWarps);
    for (int i = 0; i < kMaxSlots; ++i) ptx::mbar_init(ptx::smem_u32(&bar_scan_done[i]), kScanWarps);
    *s_turn = 0;
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
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

  const int n_levels = a.n_levels, n_k = p.n_ktiles;
  const int64_t n_row_tiles = (a.n + kTileRows - 1) / kTileRows;
  const int64_t n_groups = (n_row_tiles + NSLOT - 1) / NSLOT;  // a group = NSLOT consecutive row tiles, one per slot
  const int my_groups = static_cast<int>((n_groups - 1 - blockIdx.x) / gridDim.x + 1);  // grid <= n_groups
  const int total_tiles = n_levels * n_k;
  // first row of slot `slot` in this CTA's k-th group (may lie beyond n: the slot then runs on zeros)
  auto tile_row0 = [&](int k, int slot) -> int64_t {
    return ((static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(k) * gridDim.x) * NSLOT + slot) * kTileRows;
  };

  const int rwarp = warp - kScanWarps;  // row group = the HIGHER warp ids: the scheduler favours them, and the
                                         // row work -> MMA issue chain is the critical path of the pipeline
  if (warp >= kScanWarps) {
    // =========================================== row group ====================================================
    if constexpr (kRowRegs > kLaunchRegs) ptx::setmaxnreg_inc<kRowRegs>(); else ptx::setmaxnreg_dec<kRowRegs>();
    if (rwarp < WPS * NSLOT) {  // (NSLOT = 1 with 16 row warps: the second half of the row group has no rows)
    const int slot = rwarp / WPS;
    const int ws = rwarp % WPS;  // warp inside the sub-group
    const uint32_t n_images = static_cast<uint32_t>(my_groups) * total_tiles;
    auto load_image = [&](uint32_t i) {  // image i of this CTA's sequence -> its stage (one thread)
      const int s = p.resident ? static_cast<int>(i) : static_cast<int>(i % p.stages);
      const uint32_t bar = ptx::smem_u32(&bar_b_full[s]);
      ptx::mbar_arrive_expect_tx(bar, p.tile_bytes);
      ptx::bulk_g2s(ptx::smem_u32(s_b + static_cast<size_t>(s) * p.tile_bytes),
                    p.packed + static_cast<size_t>(i % total_tiles) * p.tile_bytes, p.tile_bytes, bar);
    };
    if (rwarp == 0 && lane == 0) {
      // resident: every image once; streamed: fill all stages but one (slot 0's issuer adds one image per iteration)
      const uint32_t first = p.resident ? static_cast<uint32_t>(total_tiles)
                                        : (n_images < static_cast<uint32_t>(p.stages - 1) ? n_images : p.stages - 1);
      for (uint32_t i = 0; i < first; ++i) load_image(i);
    }
    const int j = lane % L::RPI, kc = lane / L::RPI;
    const int rc_row0 = ws * (RPT * L::RPI) + j;  // + g * RPI
    const uint32_t ones = ptx::smem_u32(s_ones);
    const uint32_t a_hi = ptx::smem_u32(s_a + slot * p.a_bytes);
    const uint32_t a_lo = a_hi + p.a_bytes / 2;
    const uint32_t bar_slot = 1 + slot;  // named barrier of the sub-group
    const uint32_t bar_scan = ptx::smem_u32(&bar_scan_done[slot]);
    const uint32_t turn_addr = ptx::smem_u32(s_turn);
    const uint2* cand = reinterpret_cast<const uint2*>(s_cand + slot * kCandSlotBytes);  // [half][row]
    float* my_loss = s_loss + slot * kTileRows;
    const bool want_loss = a.loss != nullptr || a.level_loss != nullptr;
    // the last level's code row is only needed when something other than ids is asked for
    const bool tail_last = a.emb_out != nullptr || want_loss || a.final_residual != nullptr;
    uint32_t scan_phase = 0;

    for (int k = 0; k < my_groups; ++k) {
      const int64_t row0 = tile_row0(k, slot);
      if (ws == 0 && lane == 0 && k + 1 < my_groups) {  // pull the slot's next row tile into L2
        const int64_t next0 = tile_row0(k + 1, slot);
        if (next0 < a.n) {
          const int64_t rows = a.n - next0 < kTileRows ? a.n - next0 : kTileRows;
          ptx::bulk_prefetch_l2(a.x + next0 * D, static_cast<uint32_t>(rows * D * 4));
        }
      }
      float r[RPT][8];
      if (slot == 0 && ws == 0) stamp(0, 1);
#pragma unroll
      for (int g = 0; g < RPT; ++g) {
        const int64_t grow = row0 + rc_row0 + g * L::RPI;
        if (grow < a.n) {
          const float4* src = reinterpret_cast<const float4*>(a.x + grow * D + kc * 8);
          const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
          r[g][0] = v0.x, r[g][1] = v0.y, r[g][2] = v0.z, r[g][3] = v0.w;
          r[g][4] = v1.x, r[g][5] = v1.y, r[g][6] = v1.z, r[g][7] = v1.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) r[g][i] = 0.f;
        }
        if (want_loss && kc == 0) my_loss[rc_row0 + g * L::RPI] = 0.f;
      }

      for (int l = 0; l < n_levels; ++l) {
        if (slot == 0 && ws == 0) stamp(0, 2);
        // ---- stage the residual as the A operand (bf16 hi | lo), optionally store it ----
#pragma unroll
        for (int g = 0; g < RPT; ++g) {
          const int row = rc_row0 + g * L::RPI;
          uint4 hi, lo;
          split8(r[g], hi, lo);
          sts128(a_hi + kc * (kTileRows * 16) + row * 16, hi);
          sts128(a_lo + kc * (kTileRows * 16) + row * 16, lo);
          if (a.residuals != nullptr) {
            const int64_t grow = row0 + row;
            if (grow < a.n) {
              float4* dst = reinterpret_cast<float4*>(a.residuals + (static_cast<int64_t>(l) * a.n + grow) * D + kc * 8);
              dst[0] = make_float4(r[g][0], r[g][1], r[g][2], r[g][3]);
              dst[1] = make_float4(r[g][4], r[g][5], r[g][6], r[g][7]);
            }
          }
        }
        ptx::fence_proxy_async_smem();
        if (ws == 0) {
          ptx::named_bar_sync(bar_slot, TPS);
          if (slot == 0) stamp(0, 3);
          // ---- issue the level's MMAs, one accumulator per operand image, in round-robin order over the slots ----
          for (int t = 0; t < n_k; ++t) {
            const uint32_t it = static_cast<uint32_t>((k * n_levels + l) * n_k + t);  // image index in this CTA's sequence
            const uint32_t s_issue = it * NSLOT + slot;                               // accumulator sequence number
            const uint32_t acc = s_issue & 1u;
            if (NSLOT > 1) ptx::counter_wait(turn_addr, s_issue);
            ptx::mbar_wait(ptx::smem_u32(&bar_acc_free[acc]), ((s_issue >> 1) & 1u) ^ 1u);
            const int s = p.resident ? l * n_k + t : static_cast<int>(it % p.stages);
            ptx::mbar_wait(ptx::smem_u32(&bar_b_full[s]), p.resident ? 0u : (it / p.stages) & 1u);
            ptx::tc_fence_after_sync();
            if (ptx::elect_one()) {
              const uint32_t b_tile = ptx::smem_u32(s_b + static_cast<size_t>(s) * p.tile_bytes);
              const uint32_t acc_col = tmem_base + acc * kNTile;
              if (kAblate && (p.debug & 4)) {
                ptx::umma_commit(ptx::smem_u32(&bar_mma_done[2 * acc]));
                ptx::umma_commit(ptx::smem_u32(&bar_mma_done[2 * acc + 1]));
              } else {
              issue_unit<D>(acc_col, a_hi, a_lo, ones, b_tile, 0, ptx::smem_u32(&bar_mma_done[2 * acc]));
              issue_unit<D>(acc_col + kUnitCols, a_hi, a_lo, ones, b_tile, kUnitCols, ptx::smem_u32(&bar_mma_done[2 * acc + 1]));
              }
              if (NSLOT > 1) ptx::counter_add_release(turn_addr, 1);
              if (slot == 0) stamp(0, 4);
              if (!p.resident) {
                ptx::umma_commit(ptx::smem_u32(&bar_b_empty[s]));
                if (slot == 0) {
                  // refill: image it + stages - 1 goes where image it - 1 was, once every slot's MMAs on it are done
                  const uint32_t nxt = it + p.stages - 1;
                  if (nxt < n_images) {
                    if (it > 0) {
                      const uint32_t prev = it - 1;
                      ptx::mbar_wait(ptx::smem_u32(&bar_b_empty[prev % p.stages]), (prev / p.stages) & 1u);
                    }
                    load_image(nxt);
                  }
                }
              }
            }
            __syncwarp();
          }
        } else {
          ptx::named_bar_arrive(bar_slot, TPS);
        }

        // ---- wait for the scan, merge the two halves of every row, write ids, gather, value / loss / residual ----
        if (slot == 0 && ws == 0) stamp(0, 5);
        ptx::mbar_wait(bar_scan, scan_phase);
        scan_phase ^= 1;
        if (slot == 0 && ws == 0) stamp(0, 6);
        const float* cb = a.codebooks + static_cast<int64_t>(l) * a.k * D;
        const bool last = l + 1 == n_levels;
        const bool tail = !last || tail_last;
#pragma unroll
        for (int g0 = 0; g0 < RPT; g0 += GC) {
          float e[GC][8];
#pragma unroll
          for (int gg = 0; gg < GC; ++gg) {
            const int row = rc_row0 + (g0 + gg) * L::RPI;
            const uint2 c0 = cand[row], c1 = cand[kTileRows + row];
            const float m0 = __uint_as_float(c0.x), m1 = __uint_as_float(c1.x);
            const bool first = m0 > m1 || (m0 == m1 && c0.y < c1.y);  // lowest column wins exact ties
            uint32_t k_sel = first ? c0.y : c1.y;
            k_sel = k_sel < static_cast<uint32_t>(a.k) ? k_sel : static_cast<uint32_t>(a.k - 1);
            const int64_t grow = row0 + row;
            if (kc == 0 && grow < a.n) a.ids[grow * a.ids_row_stride + l * a.ids_level_stride] = k_sel;
            if (kAblate && tail && (p.debug & 2)) {
#pragma unroll
              for (int i = 0; i < 8; ++i) e[gg][i] = 0.25f * r[g0 + gg][i];
            } else if (tail) {
              const float4* src = reinterpret_cast<const float4*>(cb + static_cast<int64_t>(k_sel) * D + kc * 8);
              const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
              e[gg][0] = v0.x, e[gg][1] = v0.y, e[gg][2] = v0.z, e[gg][3] = v0.w;
              e[gg][4] = v1.x, e[gg][5] = v1.y, e[gg][6] = v1.z, e[gg][7] = v1.w;
            }
          }
          if (slot == 0 && ws == 0) stamp(0, 7);
          if (tail) {
#pragma unroll
            for (int gg = 0; gg < GC; ++gg) {
              const int g = g0 + gg;
              const int row = rc_row0 + g * L::RPI;
              const int64_t grow = row0 + row;
              const bool valid = grow < a.n;
              float o[8];
              float ll = 0.f;
              if (want_loss) {
                float sq = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float df = r[g][i] - e[gg][i];
                  sq = fmaf(df, df, sq);
                }
                sq = row_sum<D>(sq);
                ll = sq + a.beta * sq;  // (modules/loss.py:41-44)
              }
              if constexpr (!ROT) {
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = e[gg][i];
              } else {
                // modules/quantize.py:34-45,134-140:  o = r - 2 (r.w) w + 2 (r.u) q
                float rr = 0.f, ee = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  rr = fmaf(r[g][i], r[g][i], rr);
                  ee = fmaf(e[gg][i], e[gg][i], ee);
                }
                rr = row_sum<D>(rr);
                ee = row_sum<D>(ee);
                const float inv_r = 1.0f / (sqrtf(rr) + 1e-8f);  // u = r / (|r| + 1e-8)
                const float inv_e = 1.0f / (sqrtf(ee) + 1e-8f);  // q = e / (|e| + 1e-8)
                float ss = 0.f, ru = 0.f, rs = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float u = r[g][i] * inv_r;
                  const float qv = e[gg][i] * inv_e;
                  const float sv = u + qv;
                  ss = fmaf(sv, sv, ss);
                  ru = fmaf(r[g][i], u, ru);
                  rs = fmaf(r[g][i], sv, rs);
                }
                ss = row_sum<D>(ss);
                ru = row_sum<D>(ru);
                rs = row_sum<D>(rs);
                const float inv_s = 1.0f / fmaxf(sqrtf(ss), 1e-6f);  // w = (u+q) / max(|u+q|, 1e-6)
                const float rw2 = 2.0f * (rs * inv_s);               // 2 (r.w)
                const float ru2 = 2.0f * ru;                         // 2 (r.u)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float u = r[g][i] * inv_r;
                  const float qv = e[gg][i] * inv_e;
                  const float w = (u + qv) * inv_s;
                  o[i] = r[g][i] - rw2 * w + ru2 * qv;
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) r[g][i] = r[g][i] - o[i];
              if (valid && a.emb_out != nullptr) {
                float4* dst = reinterpret_cast<float4*>(a.emb_out + (static_cast<int64_t>(l) * a.n + grow) * D + kc * 8);
                dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                dst[1] = make_float4(o[4], o[5], o[6], o[7]);
              }
              if (want_loss && kc == 0) {
                const float tot = my_loss[row] + ll;  // private to this thread
                my_loss[row] = tot;
                if (valid) {
                  if (a.level_loss != nullptr) a.level_loss[static_cast<int64_t>(l) * a.n + grow] = ll;
                  if (last && a.loss != nullptr) a.loss[grow] = tot;
                }
              }
              if (last && valid && a.final_residual != nullptr) {
                float4* dst = reinterpret_cast<float4*>(a.final_residual + grow * D + kc * 8);
                dst[0] = make_float4(r[g][0], r[g][1], r[g][2], r[g][3]);
                dst[1] = make_float4(r[g][4], r[g][5], r[g][6], r[g][7]);
              }
            }
          }
        }
      }
    }
    }
  } else {
    // =========================================== scan group ===================================================
    if constexpr (kScanRegs > kLaunchRegs) ptx::setmaxnreg_inc<kScanRegs>(); else ptx::setmaxnreg_dec<kScanRegs>();
    const int sw = warp;
    const int q = sw & 3;   // TMEM lane quarter
    const int h = sw >> 2;  // which 64 columns of each unit this thread scans
    const int scan_row = q * 32 + lane;
    uint32_t s = 0;  // accumulators scanned so far
    for (int k = 0; k < my_groups; ++k) {
      for (int lt = 0; lt < total_tiles; ++lt) {
        const int t = lt % n_k;
#pragma unroll 1
        for (int slot = 0; slot < NSLOT; ++slot, ++s) {
          const uint32_t acc = s & 1u, par = (s >> 1) & 1u;
          const uint32_t acc_addr = tmem_base + acc * kNTile + (static_cast<uint32_t>(q * 32) << 16);
          float m;
          int col;
          if (sw == 0) stamp(1, 10 + slot);
          if (kAblate && (p.debug & 1)) {
            ptx::mbar_wait(ptx::smem_u32(&bar_mma_done[2 * acc]), par);
            ptx::mbar_wait(ptx::smem_u32(&bar_mma_done[2 * acc + 1]), par);
            m = 0.f, col = scan_row;
          } else
          scan_accumulator(acc_addr, h, ptx::smem_u32(&bar_mma_done[2 * acc]), ptx::smem_u32(&bar_mma_done[2 * acc + 1]), par,
                           tile_row0(k, slot) + scan_row < a.n, m, col);
          if (sw == 0) stamp(1, 20 + slot);
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_acc_free[acc]));
          // running (max, column) over the level's operand images lives in the thread's own candidate entry
          uint2* c = reinterpret_cast<uint2*>(s_cand + slot * kCandSlotBytes) + h * kTileRows + scan_row;
          col += t * kNTile;
          if (t > 0) {
            const uint2 prev = *c;
            if (!(m > __uint_as_float(prev.x))) {  // strict: an earlier image keeps exact ties
              m = __uint_as_float(prev.x);
              col = static_cast<int>(prev.y);
            }
          }
          *c = make_uint2(__float_as_uint(m), static_cast<uint32_t>(col));
          if (t == n_k - 1) {
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bar_scan_done[slot]));
          }
        }
      }
    }
  }

  if (ts_on && (warp == 0 || warp == kScanWarps)) { const int who = warp == 0 ? 1 : 0; s_ts[who * 384 + 380] = ts_n; }
  ptx::tc_fence_before_sync();
  __syncthreads();
#ifdef HV_TC_INSTRUMENT
  if ((p.debug & 64) && blockIdx.x == 0 && threadIdx.x == 0) {
    for (int who = 0; who < 2; ++who) {
      const int n = s_ts[who * 384 + 380];
      for (int i = 0; i < n; ++i) printf("TS %d %u %u\n", who, s_ts[who * 384 + 2 * i], s_ts[who * 384 + 2 * i + 1]);
    }
  }
#endif
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int D>
int pack_d(const float* codebooks, int n_levels, int k, const TcPlan& plan, uint8_t* packed, cudaStream_t stream,
           uint8_t* cb32 = nullptr) {
  const int total_codes = n_levels * plan.n_ktiles * kNTile;
  rq_pack_codebooks_kernel<D><<<(total_codes + 127) / 128, 128, 0, stream>>>(codebooks, n_levels, k, plan.n_ktiles,
                                                                           plan.tile_bytes, packed, cb32);
  HV_CUDA_CHECK(cudaGetLastError());
  return HV_OK;
}

template <int D, int NSLOT>
int launch_slots(const RqFwdArgs& a, bool rot, const TcPlan& plan, uint8_t* packed, const DeviceProps& props, cudaStream_t stream) {
  const int64_t n_row_tiles = (a.n + kTileRows - 1) / kTileRows;
  const int64_t n_groups = (n_row_tiles + NSLOT - 1) / NSLOT;
  const unsigned grid = static_cast<unsigned>(n_groups < props.sm_count ? n_groups : props.sm_count);
  static const int debug = [] {
    const char* e = getenv("HIDVAE_TC_DEBUG");
    return e != nullptr ? atoi(e) : 0;
  }();
  TcParams p{packed, plan.n_ktiles, plan.tile_bytes, plan.stages, plan.resident, plan.a_bytes, debug};
  auto go = [&](auto kernel) -> int {
    cudaFuncAttributes attr;
    HV_CUDA_CHECK(cudaFuncGetAttributes(&attr, kernel));
    if (attr.numRegs < kLaunchRegs) {  // setmaxnreg.inc would wait forever: refuse loudly instead
      set_error("hv_rq_forward: tcgen05 kernel was built with %d registers/thread, the register hand-over needs %d",
                attr.numRegs, kLaunchRegs);
      return HV_ERR_UNSUPPORTED;
    }
    HV_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes));
    kernel<<<grid, kThreads, plan.smem_bytes, stream>>>(a, p);
    HV_CUDA_CHECK(cudaGetLastError());
    return HV_OK;
  };
  return rot ? go(rq_fwd_tc_kernel<D, true, NSLOT>) : go(rq_fwd_tc_kernel<D, false, NSLOT>);
}

template <int D>
int launch_d(const RqFwdArgs& a, bool rot, const TcPlan& plan, uint8_t* packed, bool prepacked, cudaStream_t stream) {
  if (!prepacked)
    if (int st = pack_d<D>(a.codebooks, a.n_levels, a.k, plan, packed, stream)) return st;
  DeviceProps props;
  if (int st = device_props(&props)) return st;
  const int64_t n_row_tiles = (a.n + kTileRows - 1) / kTileRows;
  // as many row tiles in flight per CTA as it takes to cover them with one CTA per SM (1, 2 or Cfg<D>::NT)
  const int64_t per_sm = (n_row_tiles + props.sm_count - 1) / props.sm_count;
  static const int forced = [] {
    const char* e = getenv("HIDVAE_TC_SLOTS");
    return e != nullptr ? atoi(e) : 0;
  }();
  int nslot = per_sm <= 1 ? 1 : (per_sm == 2 ? 2 : Cfg<D>::NT);
  if (forced == 1 || forced == 2 || (forced == 4 && Cfg<D>::NT == 4)) nslot = forced;  // tuning experiments only
  if constexpr (Cfg<D>::NT == 4)
    if (nslot == 4) return launch_slots<D, 4>(a, rot, plan, packed, props, stream);
  return nslot == 1 ? launch_slots<D, 1>(a, rot, plan, packed, props, stream)
                    : launch_slots<D, 2>(a, rot, plan, packed, props, stream);
}

bool use_v7_only() {
  static const bool on = getenv("HIDVAE_TC_NEW_ONLY") != nullptr;
  return on;
}

bool use_v4() {
  static const bool v4 = [] {
    const char* e = getenv("HIDVAE_TC_IMPL");
    return e != nullptr && e[0] == 'v' && e[1] == '4';
  }();
  return v4;
}

// Generation 10 (rq_fwd_tc_v10.cu).  Measured (profiles/README.md): 8-10 % faster than generation 7 on encode-only
// launches of >= 256 Ki rows, slower on small launches and on training forwards.  HIDVAE_TC_IMPL=v10 forces it,
// v4 / v7 keep it out (A/B runs).
int v10_mode() {  // -1 never, 0 by size, 1 always
  static const int mode = [] {
    const char* e = getenv("HIDVAE_TC_IMPL");
    if (e == nullptr || e[0] != 'v') return 0;
    if (e[1] == '1' && e[2] == '0') return 1;
    return (e[1] == '4' || e[1] == '7' || e[1] == '1') ? -1 : 0;
  }();
  return mode;
}
int v11_mode() {  // generation 11 (rq_fwd_tc_v11.cu): -1 never, 0 by rule, 1 always
  static const int mode = [] {
    const char* e = getenv("HIDVAE_TC_IMPL");
    if (e == nullptr || e[0] != 'v') return 0;
    if (e[1] == '1' && e[2] == '1') return 1;
    return -1;
  }();
  return mode;
}

}  // namespace

bool rq_fwd_tc_supported(int d, int k, int n_levels) {
  TcPlan plan;
  return make_plan(d, k, n_levels, &plan);
}

size_t rq_fwd_tc_workspace_bytes(int d, int k, int n_levels) {
  TcPlan plan;
  if (!make_plan(d, k, n_levels, &plan)) return 0;
  // [operand images | swizzled fp32 copy of the codebooks for generation 11 (when the shape is served)]
  return static_cast<size_t>(n_levels) * plan.n_ktiles * plan.tile_bytes + rq_fwd_tc_v11_extra_bytes(d, k, n_levels);
}

int launch_rq_pack(const float* codebooks, int n_levels, int k, int d, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream) {
  TcPlan plan;
  if (!make_plan(d, k, n_levels, &plan)) {
    set_error("hv_rq_pack_codebooks: no tcgen05 instantiation for D=%d K=%d L=%d", d, k, n_levels);
    return HV_ERR_UNSUPPORTED;
  }
  const size_t images = static_cast<size_t>(n_levels) * plan.n_ktiles * plan.tile_bytes;
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
    case 16: return pack_d<16>(codebooks, n_levels, k, plan, packed, stream);
    case 32: return pack_d<32>(codebooks, n_levels, k, plan, packed, stream, need > images ? packed + images : nullptr);
    case 64: return pack_d<64>(codebooks, n_levels, k, plan, packed, stream);
    default: return HV_ERR_UNSUPPORTED;
  }
}

int launch_rq_fwd_tc(const RqFwdArgs& a, int d, bool rot, void* workspace, size_t workspace_bytes, bool prepacked,
                     cudaStream_t stream) {
  TcPlan plan;
  if (!make_plan(d, a.k, a.n_levels, &plan)) {
    set_error("hv_rq_forward: no tcgen05 instantiation for D=%d K=%d L=%d", d, a.k, a.n_levels);
    return HV_ERR_UNSUPPORTED;
  }
  const size_t images = static_cast<size_t>(a.n_levels) * plan.n_ktiles * plan.tile_bytes;
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
  uint8_t* packed = static_cast<uint8_t*>(workspace);
  const bool outputs = a.emb_out != nullptr || a.loss != nullptr || a.level_loss != nullptr || a.residuals != nullptr;
  // Generation 11 (rq_fwd_tc_v11.cu) serves every shape it supports (D = 32, K <= 256, L <= 3): measured faster than
  // generations 4 / 7 / 10 on encode and training forwards from 12 K to 4 Mi rows (profiles/README.md).
  if (need > images && v11_mode() >= 0) {
    if (!prepacked)
      if (int st = launch_rq_pack(a.codebooks, a.n_levels, a.k, d, workspace, workspace_bytes, stream)) return st;
    return launch_rq_fwd_tc_v11(a, rot, workspace, static_cast<const uint8_t*>(workspace) + images, stream);
  }
  // (generation 10 addresses a tile's ids with 32-bit offsets: row stride * 128 must fit)
  const bool v10_ok = rq_fwd_tc_v10_supported(d, a.k, a.n_levels) && a.ids_row_stride < (1 << 23) && a.ids_row_stride > -(1 << 23);
  if (v10_ok && (v10_mode() == 1 || (v10_mode() == 0 && !outputs && a.final_residual == nullptr && a.n >= (1 << 17)))) {
    if (!prepacked)
      if (int st = launch_rq_pack(a.codebooks, a.n_levels, a.k, d, workspace, workspace_bytes, stream)) return st;
    return launch_rq_fwd_tc_v10(a, d, rot, workspace, stream);
  }
  // Measured (profiles/README.md): the previous generation's lock-step warpgroups are still faster for D = 64
  // (streamed operand images, MMA-bound) and for large training forwards, whose heavier row work (rotation value,
  // emb_out / loss stores) is spread over 16 warps there instead of this kernel's 8 row warps.
  const bool big = a.n > static_cast<int64_t>(kTileRows) * 148;
  if (use_v4() || ((d == 64 || (outputs && big)) && !use_v7_only())) {
    if (!prepacked)
      if (int st = launch_rq_pack(a.codebooks, a.n_levels, a.k, d, workspace, workspace_bytes, stream)) return st;
    return launch_rq_fwd_tc_v4(a, d, rot, workspace, stream);
  }
  switch (d) {
    case 16: return launch_d<16>(a, rot, plan, packed, prepacked, stream);
    case 32: return launch_d<32>(a, rot, plan, packed, prepacked, stream);
    case 64: return launch_d<64>(a, rot, plan, packed, prepacked, stream);
    default: return HV_ERR_UNSUPPORTED;
  }
}

}  // namespace hv
