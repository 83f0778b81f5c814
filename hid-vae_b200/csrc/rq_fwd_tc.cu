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
constexpr int kSlotWarps = 8;               // a slot = 8 warps, one accumulator set, two row tiles in flight
constexpr int kSlotThreads = kSlotWarps * 32;
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 16;
constexpr int kOnesBytes = 2 * kTileRows * 16;              // one K=16 step of the A operand: [2 chunks][128 rows][8 bf16]
constexpr int kCandSlotBytes = 2 * kTileRows * 8;           // (max, column) per half, row of one slot
constexpr int kLossSlotBytes = kTileRows * 4;               // running loss per row of one slot
constexpr int kBarBytes = 1024;
constexpr int kSmemLimit = 227 * 1024;

// slots per CTA: two while four A operands (two row tiles per slot) fit beside the operand images, else one
template <int D>
struct Cfg {
  static constexpr int NT = D <= 32 ? 2 : 1;
};
inline int slots_for(int d) { return d <= 32 ? 2 : 1; }

struct TcPlan {
  int n_ktiles;    // operand images per level
  int tile_bytes;  // one packed image
  int stages;      // shared-memory stages for images
  int resident;    // 1: every (level, image) has its own stage and is loaded once
  int smem_bytes;
  int a_bytes;     // one row tile's A operand (hi + lo) == one fp32 row tile
};

bool make_plan(int d, int k, int n_levels, TcPlan* p) {
  if ((d != 16 && d != 32 && d != 64) || k < 1 || n_levels < 1) return false;
  p->n_ktiles = (k + kNTile - 1) / kNTile;
  p->tile_bytes = kNTile * (4 * d + 32);
  p->a_bytes = kTileRows * d * 4;
  const int fixed = slots_for(d) * 2 * (p->a_bytes + kCandSlotBytes + kLossSlotBytes) + kOnesBytes + kBarBytes;
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
template <int D>
__global__ void rq_pack_codebooks_kernel(const float* __restrict__ codebooks, int n_levels, int k, int n_ktiles,
                                         int tile_bytes, uint8_t* __restrict__ packed) {
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
// The kernel.  A persistent CTA runs NSLOT independent "slots" of 8 warps; a slot owns one accumulator set in TMEM
// and works on TWO row tiles at a time (A and B), alternating between them step by step:
//
//   step(cur, oth):   scan(cur)            2-D fold argmax of cur's accumulator (waits for cur's MMAs, issued one
//                                          step ago -- normally long finished)
//                     slot barrier         cur's candidates visible, its accumulator drained
//                     issue MMA(oth)       one elected thread; oth's A operand was staged in the previous step.  The
//                                          tensor pipe works on oth while ...
//                     row work(cur)        ... the slot merges cur's candidates, writes ids, gathers the chosen code
//                                          rows, forms value / loss / next residual and stages cur's next A operand
//                     swap(cur, oth)
//
//   so a slot never waits for its own MMAs, the only synchronisation is one named barrier per step, and the second
//   slot (an independent instruction stream with its own accumulator) fills the issue slots the first leaves idle in
//   its global-load latencies.  NACC = 2 (one slot, one accumulator per row tile; D = 64) issues MMA(oth) before
//   scan(cur) instead, which keeps the tensor pipe busy back to back when the MMAs are the longer part.
// ---------------------------------------------------------------------------------------------------------
template <int D, bool ROT, int NSLOT, int NACC>
__global__ void __launch_bounds__(NSLOT* kSlotThreads, 1) rq_fwd_tc_kernel(RqFwdArgs a, TcParams p) {
  using L = Rc<D>;
  constexpr int RPT = L::RPT;
  static_assert(NSLOT * NACC <= 2, "two 256-column accumulators in TMEM");
  extern __shared__ __align__(1024) uint8_t smem[];
  // [A (slot, tile) x (hi | lo) ... | ones | candidates | loss | barriers | B stages ...]  (sized for slots_for(D))
  constexpr int n_slots_smem = Cfg<D>::NT;
  uint8_t* s_a = smem;
  uint8_t* s_ones = smem + n_slots_smem * 2 * p.a_bytes;
  uint8_t* s_cand = s_ones + kOnesBytes;
  float* s_loss = reinterpret_cast<float*>(s_cand + n_slots_smem * 2 * kCandSlotBytes);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_cand + n_slots_smem * 2 * (kCandSlotBytes + kLossSlotBytes));
  uint8_t* s_b = reinterpret_cast<uint8_t*>(s_bar) + kBarBytes;

  uint64_t* bar_b_full = s_bar;                     // [kMaxStages]  TMA -> MMA issuers
  uint64_t* bar_b_empty = bar_b_full + kMaxStages;  // [kMaxStages]  MMA completion (one commit per slot) -> TMA
  uint64_t* bar_unit = bar_b_empty + kMaxStages;    // [2 accumulators][2 units]  MMA completion -> scan
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_unit + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bar_b_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_b_empty[s]), NSLOT);
    }
    for (int i = 0; i < 4; ++i) ptx::mbar_init(ptx::smem_u32(&bar_unit[i]), 1);
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
  const int total_tiles = n_levels * n_k;
  const int64_t n_row_tiles = (a.n + kTileRows - 1) / kTileRows;
  const int64_t n_groups = (n_row_tiles + NSLOT - 1) / NSLOT;  // a group = NSLOT consecutive row tiles, one per slot
  const int G = static_cast<int>((n_groups - 1 - blockIdx.x) / gridDim.x + 1);  // groups of this CTA (grid <= n_groups)
  // every slot walks the CTA's groups: tile A takes the even ones, tile B the odd ones; pair round P = images both see
  const uint32_t n_images = static_cast<uint32_t>((G + 1) / 2) * total_tiles;

  const int slot = warp / kSlotWarps;
  const int wt = warp % kSlotWarps;
  const int q = wt & 3;   // TMEM lane quarter of the scan
  const int h = wt >> 2;  // which 64 columns of each unit this thread scans
  const int scan_row = q * 32 + lane;
  const int j = lane % L::RPI, kc = lane / L::RPI;  // row-cooperative coordinates
  const int rc_row0 = 16 * wt + j;                  // + g * RPI
  const uint32_t ones = ptx::smem_u32(s_ones);
  const uint32_t a_base = ptx::smem_u32(s_a + slot * 2 * p.a_bytes);              // + x * a_bytes
  const uint32_t cand_base = ptx::smem_u32(s_cand + slot * 2 * kCandSlotBytes);   // + x * kCandSlotBytes
  float* loss_base = s_loss + slot * 2 * kTileRows;                              // + x * kTileRows
  const uint32_t bar_slot = 1 + slot;  // named barrier of the slot
  const bool want_loss = a.loss != nullptr || a.level_loss != nullptr;
  // the last level's code row is only needed when something other than ids is asked for
  const bool tail_last = a.emb_out != nullptr || want_loss || a.final_residual != nullptr;
  // tile identity x (0 = A, 1 = B) -> accumulator columns / completion barriers
  auto acc_of = [&](int x) -> uint32_t { return tmem_base + (NACC == 2 ? x : slot) * kNTile; };
  auto unit_bar = [&](int x) -> uint32_t { return ptx::smem_u32(&bar_unit[2 * (NACC == 2 ? x : slot)]); };
  auto tile_row0 = [&](int k) -> int64_t {  // first row of this slot's tile in the CTA's k-th group
    return ((static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(k) * gridDim.x) * NSLOT + slot) * kTileRows;
  };

  auto load_image = [&](uint32_t i) {  // image i of this CTA's sequence -> its stage (one thread)
    const int s = p.resident ? static_cast<int>(i) : static_cast<int>(i % p.stages);
    const uint32_t bar = ptx::smem_u32(&bar_b_full[s]);
    ptx::mbar_arrive_expect_tx(bar, p.tile_bytes);
    ptx::bulk_g2s(ptx::smem_u32(s_b + static_cast<size_t>(s) * p.tile_bytes),
                  p.packed + static_cast<size_t>(i % total_tiles) * p.tile_bytes, p.tile_bytes, bar);
  };
  if (threadIdx.x == 0) {
    // resident: every image once; streamed: fill all stages but one (slot 0's issuer adds one image per pair round)
    const uint32_t first = p.resident ? static_cast<uint32_t>(total_tiles)
                                      : (n_images < static_cast<uint32_t>(p.stages - 1) ? n_images : p.stages - 1);
    for (uint32_t i = 0; i < first; ++i) load_image(i);
  }

  // ---- row-tile helpers (row-cooperative layout) ----
  auto load_x = [&](float(&r)[RPT][8], int k, int x) {
    const int64_t row0 = tile_row0(k);
    if (wt == 0 && lane == 0 && k + 2 < G) {  // pull the tile after this one into L2
      const int64_t next0 = tile_row0(k + 2);
      if (next0 < a.n) {
        const int64_t rows = a.n - next0 < kTileRows ? a.n - next0 : kTileRows;
        ptx::bulk_prefetch_l2(a.x + next0 * D, static_cast<uint32_t>(rows * D * 4));
      }
    }
#pragma unroll
    for (int g = 0; g < RPT; ++g) {
      const int row = rc_row0 + g * L::RPI;
      const int64_t grow = row0 + row;
      if (grow < a.n) {
        const float4* src = reinterpret_cast<const float4*>(a.x + grow * D + kc * 8);
        const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
        r[g][0] = v0.x, r[g][1] = v0.y, r[g][2] = v0.z, r[g][3] = v0.w;
        r[g][4] = v1.x, r[g][5] = v1.y, r[g][6] = v1.z, r[g][7] = v1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[g][i] = 0.f;
      }
      if (want_loss && kc == 0) loss_base[x * kTileRows + row] = 0.f;
    }
  };
  // residual -> A operand (bf16 hi | lo) of tile x, optionally stored as residuals[l]
  auto stage = [&](const float(&r)[RPT][8], int k, int l, int x) {
    const int64_t row0 = tile_row0(k);
    const uint32_t a_hi = a_base + x * p.a_bytes, a_lo = a_hi + p.a_bytes / 2;
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
  };
  // issue the MMAs of tile x at (group k, level l, image t); called by the slot's warp 0 after a slot barrier
  // (b_follows: tile B has a group in this pair round and will use the image after tile A)
  auto issue = [&](int k, int l, int t, int x, bool b_follows) {
    const uint32_t P = static_cast<uint32_t>((k >> 1) * total_tiles + l * n_k + t);  // pair round = image index
    const int s = p.resident ? l * n_k + t : static_cast<int>(P % p.stages);
    ptx::mbar_wait(ptx::smem_u32(&bar_b_full[s]), p.resident ? 0u : (P / p.stages) & 1u);
    ptx::tc_fence_after_sync();
    if (ptx::elect_one()) {
      const uint32_t a_hi = a_base + x * p.a_bytes;
      const uint32_t b_tile = ptx::smem_u32(s_b + static_cast<size_t>(s) * p.tile_bytes);
      const uint32_t acc_col = acc_of(x);
      const uint32_t bar0 = unit_bar(x);
      issue_unit<D>(acc_col, a_hi, a_hi + p.a_bytes / 2, ones, b_tile, 0, bar0);
      issue_unit<D>(acc_col + kUnitCols, a_hi, a_hi + p.a_bytes / 2, ones, b_tile, kUnitCols, bar0 + 8);
      // the image is released once both tiles of the slot have used it: the commit after B's (or a lone A's) MMAs
      if (!p.resident && (x == 1 || !b_follows)) {
        ptx::umma_commit(ptx::smem_u32(&bar_b_empty[s]));
        if (slot == 0) {
          // refill: image P + stages - 1 goes where image P - 1 was, once every slot's MMAs on it are done
          const uint32_t nxt = P + p.stages - 1;
          if (nxt < n_images) {
            if (P > 0) {
              const uint32_t prev = P - 1;
              ptx::mbar_wait(ptx::smem_u32(&bar_b_empty[prev % p.stages]), (prev / p.stages) & 1u);
            }
            load_image(nxt);
          }
        }
      }
    }
    __syncwarp();
  };

  // ---- state of the two row tiles: c = current, o = other; x = identity of c (0 = A, 1 = B) ----
  float rc[RPT][8], ro[RPT][8];
  int k_c = 0, l_c = 0, t_c = 0, k_o = 1, l_o = 0, t_o = 0;
  float best_c = -INFINITY, best_o = -INFINITY;
  int bcol_c = 0, bcol_o = 0;
  uint32_t ph_c = 0, ph_o = 0;  // completion-barrier phases (NACC == 1: ph_c is the slot's single sequence)
  int x = 0;

  if (k_c < G) {
    load_x(rc, k_c, 0);
    stage(rc, k_c, 0, 0);
  }
  if (k_o < G) {
    load_x(ro, k_o, 1);
    stage(ro, k_o, 0, 1);
  }
  ptx::named_bar_sync(bar_slot, kSlotThreads);
  if (wt == 0 && k_c < G) issue(k_c, 0, 0, 0, k_c + 1 < G);

  while (k_c < G || k_o < G) {
    if (NACC == 2) {
      // oth's A operand is staged and its accumulator drained (previous step): queue its MMAs behind cur's
      ptx::tc_fence_before_sync();
      ptx::named_bar_sync(bar_slot, kSlotThreads);
      if (wt == 0 && k_o < G) issue(k_o, l_o, t_o, x ^ 1, k_o + 1 < G);
    }
    if (k_c < G) {
      // ---- scan ----
      float m;
      int col;
      const uint32_t acc_addr = acc_of(x) + (static_cast<uint32_t>(q * 32) << 16);
      const uint32_t bar0 = unit_bar(x);
      scan_accumulator(acc_addr, h, bar0, bar0 + 8, ph_c, tile_row0(k_c) + scan_row < a.n, m, col);
      if (NACC == 2) ph_c ^= 1;
      if (m > best_c) {  // strict: an earlier image keeps exact ties
        best_c = m;
        bcol_c = t_c * kNTile + col;
      }
      if (t_c == n_k - 1) {
        const uint32_t caddr = cand_base + x * kCandSlotBytes + (h * kTileRows + scan_row) * 8;
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(caddr), "r"(__float_as_uint(best_c)), "r"(bcol_c) : "memory");
      }
    }
    if (NACC == 1) {
      if (k_c < G) ph_c ^= 1;  // one completion per issued tile, in scan order
      ptx::tc_fence_before_sync();
      ptx::named_bar_sync(bar_slot, kSlotThreads);
      if (wt == 0 && k_o < G) issue(k_o, l_o, t_o, x ^ 1, k_o + 1 < G);
    } else {
      ptx::named_bar_sync(bar_slot, kSlotThreads);
    }
    if (k_c < G) {
      if (t_c < n_k - 1) {
        ++t_c;
      } else {
        // ---- row work: merge the two halves of every row, write ids, gather, value / loss / next residual ----
        const int64_t row0 = tile_row0(k_c);
        const float* cb = a.codebooks + static_cast<int64_t>(l_c) * a.k * D;
        const bool last = l_c + 1 == n_levels;
        const bool tail = !last || tail_last;
        const uint32_t cand = cand_base + x * kCandSlotBytes;
        float* my_loss = loss_base + x * kTileRows;
        float e[RPT][8];
#pragma unroll
        for (int g = 0; g < RPT; ++g) {
          const int row = rc_row0 + g * L::RPI;
          uint32_t m0b, c0, m1b, c1;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(m0b), "=r"(c0) : "r"(cand + row * 8) : "memory");
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(m1b), "=r"(c1) : "r"(cand + (kTileRows + row) * 8) : "memory");
          const float m0 = __uint_as_float(m0b), m1 = __uint_as_float(m1b);
          const bool first = m0 > m1 || (m0 == m1 && c0 < c1);  // lowest column wins exact ties
          uint32_t k_sel = first ? c0 : c1;
          k_sel = k_sel < static_cast<uint32_t>(a.k) ? k_sel : static_cast<uint32_t>(a.k - 1);
          const int64_t grow = row0 + row;
          if (kc == 0 && grow < a.n) a.ids[grow * a.ids_row_stride + l_c * a.ids_level_stride] = k_sel;
          if (tail) {
            const float4* src = reinterpret_cast<const float4*>(cb + static_cast<int64_t>(k_sel) * D + kc * 8);
            const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
            e[g][0] = v0.x, e[g][1] = v0.y, e[g][2] = v0.z, e[g][3] = v0.w;
            e[g][4] = v1.x, e[g][5] = v1.y, e[g][6] = v1.z, e[g][7] = v1.w;
          }
        }
        if (tail) {
#pragma unroll
          for (int g = 0; g < RPT; ++g) {
            const int row = rc_row0 + g * L::RPI;
            const int64_t grow = row0 + row;
            const bool valid = grow < a.n;
            float o[8];
            float ll = 0.f;
            if (want_loss) {
              float sq = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float df = rc[g][i] - e[g][i];
                sq = fmaf(df, df, sq);
              }
              sq = row_sum<D>(sq);
              ll = sq + a.beta * sq;  // (modules/loss.py:41-44)
            }
            if constexpr (!ROT) {
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = e[g][i];
            } else {
              // modules/quantize.py:34-45,134-140:  o = r - 2 (r.w) w + 2 (r.u) q
              float rr = 0.f, ee = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                rr = fmaf(rc[g][i], rc[g][i], rr);
                ee = fmaf(e[g][i], e[g][i], ee);
              }
              rr = row_sum<D>(rr);
              ee = row_sum<D>(ee);
              const float inv_r = 1.0f / (sqrtf(rr) + 1e-8f);  // u = r / (|r| + 1e-8)
              const float inv_e = 1.0f / (sqrtf(ee) + 1e-8f);  // q = e / (|e| + 1e-8)
              float ss = 0.f, ru = 0.f, rs = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float u = rc[g][i] * inv_r;
                const float qv = e[g][i] * inv_e;
                const float sv = u + qv;
                ss = fmaf(sv, sv, ss);
                ru = fmaf(rc[g][i], u, ru);
                rs = fmaf(rc[g][i], sv, rs);
              }
              ss = row_sum<D>(ss);
              ru = row_sum<D>(ru);
              rs = row_sum<D>(rs);
              const float inv_s = 1.0f / fmaxf(sqrtf(ss), 1e-6f);  // w = (u+q) / max(|u+q|, 1e-6)
              const float rw2 = 2.0f * (rs * inv_s);               // 2 (r.w)
              const float ru2 = 2.0f * ru;                         // 2 (r.u)
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float u = rc[g][i] * inv_r;
                const float qv = e[g][i] * inv_e;
                const float w = (u + qv) * inv_s;
                o[i] = rc[g][i] - rw2 * w + ru2 * qv;
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) rc[g][i] = rc[g][i] - o[i];
            if (valid && a.emb_out != nullptr) {
              float4* dst = reinterpret_cast<float4*>(a.emb_out + (static_cast<int64_t>(l_c) * a.n + grow) * D + kc * 8);
              dst[0] = make_float4(o[0], o[1], o[2], o[3]);
              dst[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
            if (want_loss && kc == 0) {
              const float tot = my_loss[row] + ll;  // private to this thread
              my_loss[row] = tot;
              if (valid) {
                if (a.level_loss != nullptr) a.level_loss[static_cast<int64_t>(l_c) * a.n + grow] = ll;
                if (last && a.loss != nullptr) a.loss[grow] = tot;
              }
            }
            if (last && valid && a.final_residual != nullptr) {
              float4* dst = reinterpret_cast<float4*>(a.final_residual + grow * D + kc * 8);
              dst[0] = make_float4(rc[g][0], rc[g][1], rc[g][2], rc[g][3]);
              dst[1] = make_float4(rc[g][4], rc[g][5], rc[g][6], rc[g][7]);
            }
          }
        }
        t_c = 0;
        best_c = -INFINITY;
        bcol_c = 0;
        if (last) {  // next row tile of this kind (A: even groups, B: odd groups)
          l_c = 0;
          k_c += 2;
          if (k_c < G) load_x(rc, k_c, x);
        } else {
          ++l_c;
        }
        if (k_c < G) stage(rc, k_c, l_c, x);
      }
    }
    // ---- swap the roles of the two tiles ----
#pragma unroll
    for (int g = 0; g < RPT; ++g) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float tmp = rc[g][i];
        rc[g][i] = ro[g][i];
        ro[g][i] = tmp;
      }
    }
    { const int tk = k_c; k_c = k_o; k_o = tk; }
    { const int tl = l_c; l_c = l_o; l_o = tl; }
    { const int tt = t_c; t_c = t_o; t_o = tt; }
    { const float tb = best_c; best_c = best_o; best_o = tb; }
    { const int tc = bcol_c; bcol_c = bcol_o; bcol_o = tc; }
    if (NACC == 2) { const uint32_t tp = ph_c; ph_c = ph_o; ph_o = tp; }
    x ^= 1;
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int D>
int pack_d(const float* codebooks, int n_levels, int k, const TcPlan& plan, uint8_t* packed, cudaStream_t stream) {
  const int total_codes = n_levels * plan.n_ktiles * kNTile;
  rq_pack_codebooks_kernel<D><<<(total_codes + 127) / 128, 128, 0, stream>>>(codebooks, n_levels, k, plan.n_ktiles,
                                                                           plan.tile_bytes, packed);
  HV_CUDA_CHECK(cudaGetLastError());
  return HV_OK;
}

template <int D, int NSLOT, int NACC>
int launch_slots(const RqFwdArgs& a, bool rot, const TcPlan& plan, uint8_t* packed, const DeviceProps& props, cudaStream_t stream) {
  const int64_t n_row_tiles = (a.n + kTileRows - 1) / kTileRows;
  const int64_t n_groups = (n_row_tiles + NSLOT - 1) / NSLOT;
  const unsigned grid = static_cast<unsigned>(n_groups < props.sm_count ? n_groups : props.sm_count);
  TcParams p{packed, plan.n_ktiles, plan.tile_bytes, plan.stages, plan.resident, plan.a_bytes};
  auto go = [&](auto kernel) -> int {
    HV_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes));
    kernel<<<grid, NSLOT * kSlotThreads, plan.smem_bytes, stream>>>(a, p);
    HV_CUDA_CHECK(cudaGetLastError());
    return HV_OK;
  };
  return rot ? go(rq_fwd_tc_kernel<D, true, NSLOT, NACC>) : go(rq_fwd_tc_kernel<D, false, NSLOT, NACC>);
}

template <int D>
int launch_d(const RqFwdArgs& a, bool rot, const TcPlan& plan, uint8_t* packed, bool prepacked, cudaStream_t stream) {
  if (!prepacked)
    if (int st = pack_d<D>(a.codebooks, a.n_levels, a.k, plan, packed, stream)) return st;
  DeviceProps props;
  if (int st = device_props(&props)) return st;
  if constexpr (Cfg<D>::NT == 2) {
    // one slot per CTA while one row tile per SM covers the rows (more SMs at work), else two
    const int64_t n_row_tiles = (a.n + kTileRows - 1) / kTileRows;
    static const int forced = [] {
      const char* e = getenv("HIDVAE_TC_SLOTS");
      return e != nullptr ? atoi(e) : 0;
    }();
    const bool one = forced == 1 || (forced != 2 && n_row_tiles <= props.sm_count);  // HIDVAE_TC_SLOTS: tuning only
    return one ? launch_slots<D, 1, 1>(a, rot, plan, packed, props, stream)
               : launch_slots<D, 2, 1>(a, rot, plan, packed, props, stream);
  } else {
    return launch_slots<D, 1, 2>(a, rot, plan, packed, props, stream);
  }
}

bool use_v4() {
  static const bool v4 = [] {
    const char* e = getenv("HIDVAE_TC_IMPL");
    return e != nullptr && e[0] == 'v' && e[1] == '4';
  }();
  return v4;
}

}  // namespace

bool rq_fwd_tc_supported(int d, int k, int n_levels) {
  TcPlan plan;
  return make_plan(d, k, n_levels, &plan);
}

size_t rq_fwd_tc_workspace_bytes(int d, int k, int n_levels) {
  TcPlan plan;
  if (!make_plan(d, k, n_levels, &plan)) return 0;
  return static_cast<size_t>(n_levels) * plan.n_ktiles * plan.tile_bytes;
}

int launch_rq_pack(const float* codebooks, int n_levels, int k, int d, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream) {
  TcPlan plan;
  if (!make_plan(d, k, n_levels, &plan)) {
    set_error("hv_rq_pack_codebooks: no tcgen05 instantiation for D=%d K=%d L=%d", d, k, n_levels);
    return HV_ERR_UNSUPPORTED;
  }
  const size_t need = static_cast<size_t>(n_levels) * plan.n_ktiles * plan.tile_bytes;
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
    case 32: return pack_d<32>(codebooks, n_levels, k, plan, packed, stream);
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
  const size_t need = static_cast<size_t>(a.n_levels) * plan.n_ktiles * plan.tile_bytes;
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
  if (use_v4()) {
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
