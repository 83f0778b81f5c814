// tcgen05 kernel of the fused L-level residual quantiser for sm_100a -- generation 10 ("ticket ring").
// Shapes whose operand images stay resident in shared memory beside four A buffers (K <= 256, D = 16 / 32).
//
// What the timelines of the earlier generations showed (profiles/README.md): the tensor pipe needs 896 cycles per
// (128-row tile, level), the argmax scan ~600 ALU-pipe cycles per SM sub-partition, the per-row work ~200 issue
// slots -- but a tile-level took 3100 cycles, because every stage was a long, thin chain (2 warps with 8 rows per
// thread) and the slots were tied into a fixed round-robin, so any jitter stalled the ring.  This generation keeps the
// arithmetic (bf16 hi/lo 3-way split, fp32 TMEM accumulation, 2-D fold argmax) and rebuilds the pipeline:
//
//   ROW GROUPS  four slots x four warps (warps 8..23).  A slot owns one 128-row tile for its L levels in the
//               row-cooperative layout (a thread holds one 8-float K chunk of D/8 rows: coalesced 32-byte global
//               accesses, conflict-free 128-bit stores into the UMMA core-matrix layout).  Per level: stage the
//               residual as the bf16 hi | lo A operand -> slot barrier -> the slot's first warp takes a TICKET
//               (shared-memory atomic), waits for the ticket's accumulator (ticket & 1) to be free and issues the
//               level's 3*D/16 + 1 tcgen05.mma of N = 256 with ONE commit -> all four warps wait for the slot's scan ->
//               merge the two half-row candidates, write the id, gather the fp32 code row, next residual.
//               Tickets are taken when a slot is ready, so the order in which slots use the tensor pipe is
//               first-come-first-served, not fixed.  In encode mode the next tile's rows are loaded while the last
//               level's MMAs and scan run.
//   SCAN GROUP  eight warps (0..7: TMEM lane quarter = warp & 3, column half = warp >> 2) serve the tickets in order.
//               A thread owns one accumulator row and 128 columns: four tcgen05.ld 32x32b, each in flight while
//               the previous one is folded.  2-D fold with 3-input maxima: the eight 16-column chunks are folded into
//               16 running column classes (g_j = max3(g_j, v_j, v_16+j), one FMNMX3 per two scores) while a 3-ary tree
//               keeps the eight chunk maxima (8 FMNMX3 per 16 scores): 1.0 ALU-pipe instruction per score (was 1.5).
//               Row maximum = largest chunk maximum; chunk and class are recovered on the FMA pipe
//               (sum_j sat((g_j - m) * 2^120 + 1) * (32 + j) is 32 + j* when exactly one class attains m).  More than
//               one maximiser (duplicate code rows, all-zero rows) sends the warp through the exact first-index path,
//               so exact ties resolve to the lowest index like torch.min (modules/quantize.py:122).
//   The [N, K] score table never leaves the SM; only ids (and emb_out / loss in training) go to HBM.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace hv {
namespace {

constexpr int kTileRows = 128;
constexpr int kNTile = 256;  // codes per operand image = accumulator columns
#ifndef HV_V10_SLOTS
#define HV_V10_SLOTS 3  // measured: 3 slots x 96 registers beat 4 slots x 72 (spills), profiles/README.md
#endif
constexpr int kSlots = HV_V10_SLOTS;  // row tiles in flight per CTA
#ifndef HV_V10_SLOT_WARPS
#define HV_V10_SLOT_WARPS 4
#endif
constexpr int kSlotWarps = HV_V10_SLOT_WARPS;  // warps per slot: 4 (four rows per thread at D = 32) or 8 (two)
constexpr int kScanWarps = 8;
constexpr int kRowWarps = kSlots * kSlotWarps;
constexpr int kThreads = (kScanWarps + kRowWarps) * 32;  // 768
constexpr int kSlotThreads = kSlotWarps * 32;
constexpr int kLaunchRegs = (65536 / kThreads) / 8 * 8;  // launch_bounds(kThreads, 1): 80 at 768 threads, 96 at 640
// four slots: the scan group takes 96 registers, the row groups give theirs up (72); three slots: 96 for everybody
#ifdef HV_V10_SCAN16  // scan with 16-column loads (no second load buffer): 72 registers are enough
constexpr bool kScan16 = true;
#else
constexpr bool kScan16 = false;
#endif
#if defined(HV_V10_ROW_REGS) && defined(HV_V10_SCAN_REGS)
constexpr int kRowRegs = HV_V10_ROW_REGS, kScanRegs = HV_V10_SCAN_REGS;
#else
// four slots: the scan group takes 96 registers, the row groups give theirs up (72); three slots: 96 for everybody
constexpr int kRowRegs = kSlots == 4 ? 72 : kLaunchRegs, kScanRegs = kSlots == 4 ? 96 : kLaunchRegs;
#endif
static_assert(kRowRegs * kRowWarps + kScanRegs * kScanWarps <= kLaunchRegs * (kRowWarps + kScanWarps),
              "register hand-over must stay inside the launch allocation");
constexpr int kTmemCols = 512;
constexpr int kMaxLevels = 8;
constexpr int kOnesBytes = 2 * kTileRows * 16;     // one K=16 step of the A operand: [2 chunks][128 rows][8 bf16]
constexpr int kCandSlotBytes = 2 * kTileRows * 8;  // (max, column) per column half and row of one slot
constexpr int kLossSlotBytes = kTileRows * 4;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kQueue = 8;  // ticket -> slot ring
#ifdef HV_TC_INSTRUMENT
constexpr bool kInstr = true;
#else
constexpr bool kInstr = false;
#endif
constexpr int kBarBytes = kInstr ? 2048 : 1024;
constexpr bool kAblate = kInstr;  // HIDVAE_TC_DEBUG bit 2 replaces the code gather by arithmetic (timing experiments only)
#ifdef HV_V10_POLL  // the single waiters poll (mbarrier.test_wait) instead of parking on mbarrier.try_wait
constexpr bool kPoll = true;
#else
constexpr bool kPoll = false;
#endif
// wait of ONE thread on an mbarrier phase
__device__ __forceinline__ void wait_one(uint32_t bar, uint32_t parity) {
  if constexpr (kPoll) {
    if (ptx::mbar_test_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!ptx::mbar_test_wait(bar, parity))
      if (clock64() - t0 > 4000000000LL) __trap();
  } else {
    ptx::mbar_wait(bar, parity);
  }
}
#ifdef HV_V10_SCAN_LOW
constexpr bool kScanHigh = false;
#else  // scan group on the highest warp ids: the sub-partition arbiter favours them (measured 5 % faster)
constexpr bool kScanHigh = true;
#endif


struct Plan {
  int tile_bytes;  // one packed image (same format as rq_fwd_tc.cu)
  int a_bytes;     // one slot's A operand (hi + lo) == one fp32 row tile
  int smem_bytes;
};

bool make_plan(int d, int k, int n_levels, Plan* p) {
  if ((d != 16 && d != 32) || k < 1 || k > kNTile || n_levels < 1 || n_levels > kMaxLevels) return false;
  p->tile_bytes = kNTile * (4 * d + 32);
  p->a_bytes = kTileRows * d * 4;
  const int fixed = kSlots * (p->a_bytes + kCandSlotBytes + kLossSlotBytes) + kOnesBytes + kBarBytes;
  p->smem_bytes = fixed + n_levels * p->tile_bytes;
  return p->smem_bytes <= kSmemLimit;
}

// ---------------------------------------------------------------------------------------------------------
// Row-cooperative layout of a slot (128 threads, 128 x D fp32 rows as 8-float K chunks):
//   lane -> (j = lane % RPI, kc = lane / RPI);  warp ws owns rows [32 ws, 32 ws + 32);  the thread's g-th row is
//   32 ws + g * RPI + j.  A quarter warp shares kc and covers 8 consecutive rows, so the 128-bit stores into the
//   core-matrix layout (chunk kc of row `row` at kc * 2048 + row * 16) are conflict free.
// ---------------------------------------------------------------------------------------------------------
template <int D>
struct Rc {
  static constexpr int KC = D / 8;     // 16-byte (8 x bf16) K chunks per row == lanes per row
  static constexpr int RPI = 32 / KC;  // rows per warp instruction
  static constexpr int RPT = kTileRows * KC / (kSlotWarps * 32);  // rows per thread
  static constexpr int RPW = RPT * RPI;                           // rows per warp
  static_assert(KC >= 2 && KC <= 4 && RPT >= 1, "row-cooperative layout of generation 10: D = 16 or 32");
};

template <int D>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int m = Rc<D>::RPI; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// bf16 hi/lo split of 8 consecutive floats -> one 16-byte chunk entry each of the hi and the lo operand
#ifdef HV_V10_SPLIT_TRUNC  // timing experiment only (truncation instead of round-to-nearest: no F2FP conversions)
__device__ __forceinline__ void split8(const float (&r)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t b0 = __float_as_uint(r[2 * j]), b1 = __float_as_uint(r[2 * j + 1]);
    const float l0 = r[2 * j] - __uint_as_float(b0 & 0xFFFF0000u), l1 = r[2 * j + 1] - __uint_as_float(b1 & 0xFFFF0000u);
    h[j] = __byte_perm(b0, b1, 0x7632);
    l[j] = __byte_perm(__float_as_uint(l0), __float_as_uint(l1), 0x7632);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
#else
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
#endif

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));  // FMNMX3
  return d;
}

__device__ __forceinline__ float f(uint32_t v) { return __uint_as_float(v); }

// max of 16 floats v[o .. o+16): 8 FMNMX3 / FMNMX
__device__ __forceinline__ float max16(const uint32_t (&v)[32], int o) {
  const float a0 = max3(f(v[o]), f(v[o + 1]), f(v[o + 2])), a1 = max3(f(v[o + 3]), f(v[o + 4]), f(v[o + 5]));
  const float a2 = max3(f(v[o + 6]), f(v[o + 7]), f(v[o + 8])), a3 = max3(f(v[o + 9]), f(v[o + 10]), f(v[o + 11]));
  const float a4 = max3(f(v[o + 12]), f(v[o + 13]), f(v[o + 14]));
  return fmaxf(max3(a0, a1, a2), max3(a3, a4, f(v[o + 15])));
}

constexpr float kBig = 1.329227995784916e36f;  // 2^120

// max of 32 floats as a 3-ary tree
__device__ __forceinline__ float max32(const uint32_t (&v)[32]) { return fmaxf(max16(v, 0), max16(v, 16)); }

// Exact first-index (max, argmax) of one 32-column chunk against the running pair (slow path).
__device__ __forceinline__ void scan_chunk_exact(const uint32_t (&v)[32], int base, float& best, int& best_col) {
  const float m = max32(v);
  if (m > best) {  // strict: an earlier chunk keeps exact ties
    float t = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) t = fmaxf(t, fmaf(f(v[j]) - m, kBig, static_cast<float>(32 - j)));
    best = m;
    best_col = base + 32 - static_cast<int>(t);
  }
}

// The level's 3*D/16 + 1 MMAs (M128 N256 K16) into accumulator `acc` and ONE commit.  Called by ONE elected thread.
template <int D>
__device__ __forceinline__ void issue_level(uint32_t acc, uint32_t a_hi, uint32_t a_lo, uint32_t ones, uint32_t b_tile,
                                            uint32_t bar_done) {
  constexpr uint32_t idesc = ptx::umma_idesc_bf16(kTileRows, kNTile);
  const uint32_t hi = ptx::umma_desc_hi(128);
  constexpr uint32_t chunk_b = kNTile * 16;                        // bytes between K chunks of the B image
  constexpr uint32_t a_step = 2 * kTileRows, b_step = 2 * kNTile;  // one K=16 step = two chunks, in 16-byte units
  const uint32_t d_ahi = ptx::umma_desc_lo(a_hi, kTileRows * 16), d_alo = ptx::umma_desc_lo(a_lo, kTileRows * 16);
  const uint32_t d_one = ptx::umma_desc_lo(ones, kTileRows * 16);
  const uint32_t d_bhi = ptx::umma_desc_lo(b_tile, chunk_b);
  const uint32_t d_blo = ptx::umma_desc_lo(b_tile + (D / 8) * chunk_b, chunk_b);
  const uint32_t d_bnrm = ptx::umma_desc_lo(b_tile + 2 * (D / 8) * chunk_b, chunk_b);
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
  ptx::umma_commit(bar_done);
}

// fold one loaded 32-column group into the 16 classes and record its two 16-column chunk maxima
template <bool FIRST>
__device__ __forceinline__ void fold32(const uint32_t (&v)[32], float (&g)[16], float& cm_a, float& cm_b) {
  cm_a = max16(v, 0);
  cm_b = max16(v, 16);
#pragma unroll
  for (int j = 0; j < 16; ++j) g[j] = FIRST ? fmaxf(f(v[j]), f(v[16 + j])) : max3(g[j], f(v[j]), f(v[16 + j]));
}

__device__ __forceinline__ float max16(const uint32_t (&v)[16]) {
  const float a0 = max3(f(v[0]), f(v[1]), f(v[2])), a1 = max3(f(v[3]), f(v[4]), f(v[5]));
  const float a2 = max3(f(v[6]), f(v[7]), f(v[8])), a3 = max3(f(v[9]), f(v[10]), f(v[11]));
  const float a4 = max3(f(v[12]), f(v[13]), f(v[14]));
  return fmaxf(max3(a0, a1, a2), max3(a3, a4, f(v[15])));
}

// Scan of this thread's row over 128 accumulator columns starting at TMEM address `t0` (see the file header).
// Returns the maximum and its first column (0..127).
__device__ __forceinline__ void scan_half(uint32_t t0, float& m_out, int& col_out) {
  float g[16], cm[8];
  if constexpr (kScan16) {
    uint32_t a[16], b[16];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      ptx::tmem_ld_32x16(t0 + 32 * p, a);
      ptx::tmem_ld_32x16(t0 + 32 * p + 16, b);
      ptx::tmem_wait_ld(a, b);
      cm[2 * p] = max16(a);
      cm[2 * p + 1] = max16(b);
#pragma unroll
      for (int j = 0; j < 16; ++j) g[j] = p == 0 ? fmaxf(f(a[j]), f(b[j])) : max3(g[j], f(a[j]), f(b[j]));
    }
  } else {
  uint32_t v[32], w[32];
  ptx::tmem_ld_32x32(t0, v);
  ptx::tmem_wait_ld(v);
  ptx::tmem_ld_32x32(t0 + 32, w);  // in flight while v is folded
  fold32<true>(v, g, cm[0], cm[1]);
  ptx::tmem_wait_ld(w);
  ptx::tmem_ld_32x32(t0 + 64, v);
  fold32<false>(w, g, cm[2], cm[3]);
  ptx::tmem_wait_ld(v);
  ptx::tmem_ld_32x32(t0 + 96, w);
  fold32<false>(v, g, cm[4], cm[5]);
  ptx::tmem_wait_ld(w);
  fold32<false>(w, g, cm[6], cm[7]);
  }

  const float m = max3(max3(cm[0], cm[1], cm[2]), max3(cm[3], cm[4], cm[5]), fmaxf(cm[6], cm[7]));
  // class of the maximiser: sum_j [g_j == m] * (32 + j);  chunk: sum_c [cm_c == m] * (16 + c)   (FMA pipe)
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float e = __saturatef(fmaf(g[j] - m, kBig, 1.0f));
    s4[j & 3] = fmaf(e, static_cast<float>(32 + j), s4[j & 3]);
  }
  float c2[2] = {0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float e = __saturatef(fmaf(cm[c] - m, kBig, 1.0f));
    c2[c & 1] = fmaf(e, static_cast<float>(16 + c), c2[c & 1]);
  }
  const float cls = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  const float chk = c2[0] + c2[1];
  // exactly one class and one chunk attain m (a thread whose columns are all padding, m = -1e30, cannot win anyway;
  // NaN scores fail the test and take the exact path)
  const bool unique = (cls < 64.f && chk < 32.f) || m < -1e29f;
  float best = m;
  int col = 16 * (static_cast<int>(chk) - 16) + static_cast<int>(cls) - 32;
  if (__any_sync(0xffffffffu, !unique)) {
    uint32_t v[32];
    best = -INFINITY;
    col = 0;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      ptx::tmem_ld_32x32(t0 + 32 * c, v);
      ptx::tmem_wait_ld(v);
      scan_chunk_exact(v, 32 * c, best, col);
    }
  }
  m_out = best;
  col_out = col;
}

struct V10Params {
  const uint8_t* packed;
  int tile_bytes;
  int a_bytes;
  int debug;
};

template <int D, bool ROT>
__global__ void __launch_bounds__(kThreads, 1) rq_fwd_tc_v10_kernel(RqFwdArgs a, V10Params p) {
  using L = Rc<D>;
  constexpr int RPT = L::RPT;
  extern __shared__ __align__(1024) uint8_t smem[];
  // [A slot 0 (hi | lo) | ... | ones | candidates | loss | barriers + counters | operand images]
  uint8_t* s_a = smem;
  uint8_t* s_ones = smem + kSlots * p.a_bytes;
  uint8_t* s_cand = s_ones + kOnesBytes;
  float* s_loss = reinterpret_cast<float*>(s_cand + kSlots * kCandSlotBytes);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_cand + kSlots * (kCandSlotBytes + kLossSlotBytes));
  uint8_t* s_b = reinterpret_cast<uint8_t*>(s_bar) + kBarBytes;

  uint64_t* bar_b_full = s_bar;                      // [kMaxLevels]  TMA -> MMA issuers (images are loaded once)
  uint64_t* bar_mma_done = bar_b_full + kMaxLevels;  // [2]           MMA completion -> scan group
  uint64_t* bar_acc_free = bar_mma_done + 2;         // [2]           scan group (8 warps) -> MMA issuers
  uint64_t* bar_scan_done = bar_acc_free + 2;        // [kSlots]      scan group (8 warps) -> the slot's warps
  uint32_t* s_ticket = reinterpret_cast<uint32_t*>(bar_scan_done + kSlots);
  uint32_t* s_queue = s_ticket + 1;                  // [kQueue]  (ticket + 1) << 8 | slot
  uint32_t* s_tmem = s_queue + kQueue;
  uint32_t* s_ts = s_tmem + 1;                       // [2][192] timestamps of block 0 (instrumented builds)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int ts_n = 0;
  const bool ts_on = kInstr && (p.debug & 64) && blockIdx.x == 0 && lane == 0;
  auto stamp = [&](int who, int code) {
    if (ts_on && ts_n < 94) {
      s_ts[who * 192 + 2 * ts_n] = code;
      s_ts[who * 192 + 2 * ts_n + 1] = static_cast<uint32_t>(clock64());
      ++ts_n;
    }
  };

  const int n_levels = a.n_levels;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxLevels; ++i) ptx::mbar_init(ptx::smem_u32(&bar_b_full[i]), 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&bar_mma_done[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&bar_acc_free[i]), kScanWarps);
    }
    for (int i = 0; i < kSlots; ++i) ptx::mbar_init(ptx::smem_u32(&bar_scan_done[i]), kScanWarps);
    *s_ticket = 0;
    for (int i = 0; i < kQueue; ++i) s_queue[i] = 0;
    ptx::fence_mbar_init();
    for (int l = 0; l < n_levels; ++l) {  // operand images: resident for the whole kernel
      const uint32_t bar = ptx::smem_u32(&bar_b_full[l]);
      ptx::mbar_arrive_expect_tx(bar, p.tile_bytes);
      ptx::bulk_g2s(ptx::smem_u32(s_b + static_cast<size_t>(l) * p.tile_bytes), p.packed + static_cast<size_t>(l) * p.tile_bytes,
                    p.tile_bytes, bar);
    }
  }
  if (warp == 0) {
    __syncwarp();
    ptx::tmem_alloc(ptx::smem_u32(s_tmem), kTmemCols);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + kTileRows) {
    // constant A block that multiplies the norm pieces: row -> [1, 1, 1, 0, 0, 0, 0, 0 | 0 x 8]
    const int t = threadIdx.x - 32;
    *reinterpret_cast<uint4*>(s_ones + t * 16) = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);
    *reinterpret_cast<uint4*>(s_ones + kTileRows * 16 + t * 16) = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  const int64_t n_row_tiles = (a.n + kTileRows - 1) / kTileRows;
  const int my_tiles = static_cast<int>((n_row_tiles - 1 - blockIdx.x) / gridDim.x + 1);  // grid <= n_row_tiles
  auto tile_row0 = [&](int i) -> int64_t { return (static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x) * kTileRows; };

  const bool is_row = kScanHigh ? warp < kRowWarps : warp >= kScanWarps;
  if (is_row) {
    // =========================================== row groups ===================================================
    // Addressing is tile-relative and 32-bit: one 64-bit base per tile / level, the thread's own offsets are constants.
    if constexpr (kRowRegs < kLaunchRegs) ptx::setmaxnreg_dec<kRowRegs>();
    constexpr int kABytes = kTileRows * D * 4;
    const int rwarp = kScanHigh ? warp : warp - kScanWarps;
    const int slot = rwarp / kSlotWarps;
    const int ws = rwarp % kSlotWarps;
    const int kc = lane / L::RPI;
    const int trow = ws * L::RPW + lane % L::RPI;  // the thread's g-th row of the tile is trow + g * RPI
    const int xoff = trow * D + kc * 8;        // float offset of its piece of that row inside the tile (+ g * RPI * D)
    const uint32_t ones = ptx::smem_u32(s_ones);
    const uint32_t a_slot = ptx::smem_u32(s_a) + slot * kABytes;              // the slot's A operand: hi | lo
    const uint32_t a_mine = a_slot + kc * (kTileRows * 16) + trow * 16;       // the thread's hi entry (+ g * RPI * 16)
    const uint32_t bar_ready = 1 + slot, bar_go = 1 + kSlots + slot;          // named barriers of the slot
    const uint32_t bar_scan = ptx::smem_u32(&bar_scan_done[slot]);
    const uint2* cand = reinterpret_cast<const uint2*>(s_cand + slot * kCandSlotBytes) + trow;  // [half][row]
    float* my_loss = s_loss + slot * kTileRows + trow;
    const bool want_loss = a.loss != nullptr || a.level_loss != nullptr;
    // the last level's code row is only needed when something other than ids is asked for
    const bool tail_last = a.emb_out != nullptr || want_loss || a.final_residual != nullptr;
    const int id_rs = static_cast<int>(a.ids_row_stride);  // (launch_d checks that a tile's id offsets fit 32 bits)
    const int lvl_floats = a.k * D;
    uint32_t scan_phase = 0;

    float r[RPT][8];
    auto load_x = [&](int i) {  // rows of this CTA's i-th tile -> registers (rows beyond n read as zero)
      const int64_t row0 = tile_row0(i);
      const int rows_here = static_cast<int>(a.n - row0 < kTileRows ? a.n - row0 : kTileRows);
      const float* xt = a.x + row0 * D + xoff;
#pragma unroll
      for (int g = 0; g < RPT; ++g) {
        if (trow + g * L::RPI < rows_here) {
          const float4* src = reinterpret_cast<const float4*>(xt + g * (L::RPI * D));
          const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
          r[g][0] = v0.x, r[g][1] = v0.y, r[g][2] = v0.z, r[g][3] = v0.w;
          r[g][4] = v1.x, r[g][5] = v1.y, r[g][6] = v1.z, r[g][7] = v1.w;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) r[g][q] = 0.f;
        }
      }
    };
    auto prefetch_tile = [&](int i) {  // pull a later row tile of this slot into L2
      if (ws == 0 && lane == 0 && i < my_tiles) {
        const int64_t next0 = tile_row0(i);
        const int64_t rows = a.n - next0 < kTileRows ? a.n - next0 : kTileRows;
        ptx::bulk_prefetch_l2(a.x + next0 * D, static_cast<uint32_t>(rows * D * 4));
      }
    };

    bool have_x = false;
    for (int i = slot; i < my_tiles; i += kSlots) {
      const int64_t row0 = tile_row0(i);
      const int rows_here = static_cast<int>(a.n - row0 < kTileRows ? a.n - row0 : kTileRows);
      prefetch_tile(i + kSlots);
      if (slot == 0 && ws == 0) stamp(0, 1);
      if (!have_x) load_x(i);
      have_x = false;
      if (want_loss && kc == 0) {
#pragma unroll
        for (int g = 0; g < RPT; ++g) my_loss[g * L::RPI] = 0.f;
      }
      int64_t* id_lvl = a.ids + row0 * a.ids_row_stride;               // + l * level stride
      const float* cb = a.codebooks + kc * 8;                           // + l * K * D
      const int64_t out_lvl0 = row0 * D + xoff;                         // float offset in a [L, N, D] output (+ l * N * D)

      for (int l = 0; l < n_levels; ++l, id_lvl += a.ids_level_stride, cb += lvl_floats) {
        const bool last = l + 1 == n_levels;
        const bool tail = !last || tail_last;
        const int64_t out_lvl = out_lvl0 + static_cast<int64_t>(l) * a.n * D;
        if (slot == 0 && ws == 0) stamp(0, 2);
        // ---- stage the residual as the A operand (bf16 hi | lo), optionally store it ----
#pragma unroll
        for (int g = 0; g < RPT; ++g) {
          uint4 hi, lo;
          split8(r[g], hi, lo);
          sts128(a_mine + g * (L::RPI * 16), hi);
          sts128(a_mine + kABytes / 2 + g * (L::RPI * 16), lo);
          if (a.residuals != nullptr && trow + g * L::RPI < rows_here) {
            float4* dst = reinterpret_cast<float4*>(a.residuals + out_lvl + g * (L::RPI * D));
            dst[0] = make_float4(r[g][0], r[g][1], r[g][2], r[g][3]);
            dst[1] = make_float4(r[g][4], r[g][5], r[g][6], r[g][7]);
          }
        }
        ptx::fence_proxy_async_smem();
        if (slot == 0 && ws == 0) stamp(0, 25);
        const bool early_x = last && !tail && i + kSlots < my_tiles;  // encode: the rows are dead once staged
        if (ws == 0) {
          ptx::named_bar_sync(bar_ready, kSlotThreads);
          if (slot == 0) stamp(0, 3);
          // ---- take a ticket, wait for its accumulator, issue the level's MMAs; then wait for the slot's scan ----
          if (ptx::elect_one()) {
            const uint32_t t = atomicAdd(s_ticket, 1u);
            const uint32_t acc = t & 1u;
            wait_one(ptx::smem_u32(&bar_acc_free[acc]), ((t >> 1) & 1u) ^ 1u);
            if (i < kSlots) ptx::mbar_wait(ptx::smem_u32(&bar_b_full[l]), 0u);  // images are loaded once: only the slot's first tile can be early
            asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(ptx::smem_u32(&s_queue[t % kQueue])),
                         "r"(((t + 1u) << 8) | static_cast<uint32_t>(slot))
                         : "memory");
            ptx::tc_fence_after_sync();
            if (kInstr && slot == 0) stamp(0, 35);
            issue_level<D>(tmem_base + acc * kNTile, a_slot, a_slot + kABytes / 2, ones,
                           ptx::smem_u32(s_b + static_cast<size_t>(l) * p.tile_bytes), ptx::smem_u32(&bar_mma_done[acc]));
          }
          __syncwarp();
          if (slot == 0) stamp(0, 4);
          if (early_x) load_x(i + kSlots);  // the next tile's rows travel while the MMAs and the scan run
          if (ptx::elect_one()) wait_one(bar_scan, scan_phase);  // ONE waiter per slot; the rest park on the
          __syncwarp();                                                 // named barrier (no polling instructions)
        } else {
          ptx::named_bar_arrive(bar_ready, kSlotThreads);
          if (early_x) load_x(i + kSlots);
        }
        have_x = early_x;
        ptx::named_bar_sync(bar_go, kSlotThreads);
        scan_phase ^= 1;
        if (slot == 0 && ws == 0) stamp(0, 6);

        // ---- merge the two halves of every row, gather the code rows, write the ids ----
        uint32_t k_sel[RPT];
        float e[RPT][8];
#pragma unroll
        for (int g = 0; g < RPT; ++g) {
          const uint2 c0 = cand[g * L::RPI], c1 = cand[kTileRows + g * L::RPI];
          const float m0 = __uint_as_float(c0.x), m1 = __uint_as_float(c1.x);
          const uint32_t ks = m1 > m0 ? c1.y : c0.y;  // the lower half keeps exact ties (lowest column wins)
          k_sel[g] = ks < static_cast<uint32_t>(a.k) ? ks : static_cast<uint32_t>(a.k - 1);
          if (kAblate && (p.debug & 2)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) e[g][q] = 0.25f * r[g][q];
          } else if (tail) {
            const float4* src = reinterpret_cast<const float4*>(cb + k_sel[g] * D);
            const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
            e[g][0] = v0.x, e[g][1] = v0.y, e[g][2] = v0.z, e[g][3] = v0.w;
            e[g][4] = v1.x, e[g][5] = v1.y, e[g][6] = v1.z, e[g][7] = v1.w;
          }
        }
        if (kInstr && slot == 0 && ws == 0) stamp(0, 65);
        if (kc == 0) {
#pragma unroll
          for (int g = 0; g < RPT; ++g)
            if (trow + g * L::RPI < rows_here) id_lvl[(trow + g * L::RPI) * id_rs] = k_sel[g];
        }
        if (slot == 0 && ws == 0) stamp(0, 7);
        if (tail) {
#pragma unroll
          for (int g = 0; g < RPT; ++g) {
            const int row = trow + g * L::RPI;
            const bool valid = row < rows_here;
            float o[8];
            float ll = 0.f;
            if (want_loss) {
              float sq = 0.f;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float df = r[g][q] - e[g][q];
                sq = fmaf(df, df, sq);
              }
              sq = row_sum<D>(sq);
              ll = sq + a.beta * sq;  // (modules/loss.py:41-44)
            }
            if constexpr (!ROT) {
#pragma unroll
              for (int q = 0; q < 8; ++q) o[q] = e[g][q];
            } else {
              // modules/quantize.py:34-45,134-140:  o = r - 2 (r.w) w + 2 (r.u) q
              float rr = 0.f, ee = 0.f;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                rr = fmaf(r[g][q], r[g][q], rr);
                ee = fmaf(e[g][q], e[g][q], ee);
              }
              rr = row_sum<D>(rr);
              ee = row_sum<D>(ee);
              const float inv_r = 1.0f / (sqrtf(rr) + 1e-8f);  // u = r / (|r| + 1e-8)
              const float inv_e = 1.0f / (sqrtf(ee) + 1e-8f);  // q = e / (|e| + 1e-8)
              float ss = 0.f, ru = 0.f, rs = 0.f;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float u = r[g][q] * inv_r;
                const float qv = e[g][q] * inv_e;
                const float sv = u + qv;
                ss = fmaf(sv, sv, ss);
                ru = fmaf(r[g][q], u, ru);
                rs = fmaf(r[g][q], sv, rs);
              }
              ss = row_sum<D>(ss);
              ru = row_sum<D>(ru);
              rs = row_sum<D>(rs);
              const float inv_s = 1.0f / fmaxf(sqrtf(ss), 1e-6f);  // w = (u+q) / max(|u+q|, 1e-6)
              const float rw2 = 2.0f * (rs * inv_s);               // 2 (r.w)
              const float ru2 = 2.0f * ru;                         // 2 (r.u)
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float u = r[g][q] * inv_r;
                const float qv = e[g][q] * inv_e;
                const float w = (u + qv) * inv_s;
                o[q] = r[g][q] - rw2 * w + ru2 * qv;
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) r[g][q] = r[g][q] - o[q];
            if (valid && a.emb_out != nullptr) {
              float4* dst = reinterpret_cast<float4*>(a.emb_out + out_lvl + g * (L::RPI * D));
              dst[0] = make_float4(o[0], o[1], o[2], o[3]);
              dst[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
            if (want_loss && kc == 0) {
              const float tot = my_loss[g * L::RPI] + ll;  // private to this thread
              my_loss[g * L::RPI] = tot;
              if (valid) {
                if (a.level_loss != nullptr) a.level_loss[static_cast<int64_t>(l) * a.n + row0 + row] = ll;
                if (last && a.loss != nullptr) a.loss[row0 + row] = tot;
              }
            }
            if (last && valid && a.final_residual != nullptr) {
              float4* dst = reinterpret_cast<float4*>(a.final_residual + out_lvl0 + g * (L::RPI * D));
              dst[0] = make_float4(r[g][0], r[g][1], r[g][2], r[g][3]);
              dst[1] = make_float4(r[g][4], r[g][5], r[g][6], r[g][7]);
            }
          }
          if (kInstr && slot == 0 && ws == 0) { if (r[0][0] == 12345.f) stamp(0, 99); stamp(0, 8); }
        }
      }
    }
  } else {
    // =========================================== scan group ===================================================
    if constexpr (kScanRegs > kLaunchRegs) ptx::setmaxnreg_inc<kScanRegs>();
    const int swarp = kScanHigh ? warp - kRowWarps : warp;
    const int q = warp & 3;    // TMEM lane quarter (fixed by the hardware: warp id % 4)
    const int h = swarp >> 2;  // column half
    const int scan_row = q * 32 + lane;
    const uint32_t n_tickets = static_cast<uint32_t>(my_tiles) * static_cast<uint32_t>(n_levels);
    for (uint32_t t = 0; t < n_tickets; ++t) {
      const uint32_t acc = t & 1u, par = (t >> 1) & 1u;
      if (swarp == 0) stamp(1, 10);
      // ONE thread waits for the MMA completion (mbarrier.try_wait loops cost issue slots and load/store-pipe probes),
      // the rest of the scan group parks on a named barrier.  The slot that holds ticket t was published by the
      // issuer before its MMAs.
      if (swarp == 0) {
        if (ptx::elect_one()) wait_one(ptx::smem_u32(&bar_mma_done[acc]), par);
        __syncwarp();
      }
      ptx::named_bar_sync(1 + 2 * kSlots, kScanWarps * 32);
      const uint32_t q_addr = ptx::smem_u32(&s_queue[t % kQueue]);
      uint32_t qv = ptx::counter_ld_acquire(q_addr);
      if ((qv >> 8) != t + 1u) {
        const long long t_start = clock64();
        while (((qv = ptx::counter_ld_acquire(q_addr)) >> 8) != t + 1u) {
          if (clock64() - t_start > 4000000000LL) {
            printf("hidvae_b200: ticket wait timed out (block %d warp %d ticket %u)\n", blockIdx.x, warp, t);
            __trap();
          }
        }
      }
      const int slot = static_cast<int>(qv & 0xFFu);
      ptx::tc_fence_after_sync();
      if (swarp == 0) stamp(1, 11);
      float m;
      int col;
      scan_half(tmem_base + acc * kNTile + 128 * h + (static_cast<uint32_t>(q * 32) << 16), m, col);
      ptx::tc_fence_before_sync();
      uint2* c = reinterpret_cast<uint2*>(s_cand + slot * kCandSlotBytes) + h * kTileRows + scan_row;
      *c = make_uint2(__float_as_uint(m), static_cast<uint32_t>(col + 128 * h));
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(ptx::smem_u32(&bar_acc_free[acc]));
        ptx::mbar_arrive(ptx::smem_u32(&bar_scan_done[slot]));
      }
      if (swarp == 0) stamp(1, 12);
    }
  }

  if (ts_on && (warp == 0 || warp == (kScanHigh ? kRowWarps : kScanWarps))) s_ts[((warp == 0) != kScanHigh ? 1 : 0) * 192 + 190] = ts_n;
  ptx::tc_fence_before_sync();
  __syncthreads();
#ifdef HV_TC_INSTRUMENT
  if ((p.debug & 64) && blockIdx.x == 0 && threadIdx.x == 0) {
    for (int who = 0; who < 2; ++who) {
      const int n = s_ts[who * 192 + 94];
      for (int i = 0; i < n; ++i) printf("TS %d %u %u\n", who, s_ts[who * 192 + 2 * i], s_ts[who * 192 + 2 * i + 1]);
    }
  }
#endif
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int D>
int launch_d(const RqFwdArgs& a, bool rot, const Plan& plan, const uint8_t* packed, cudaStream_t stream) {
  DeviceProps props;
  if (int st = device_props(&props)) return st;
  const int64_t n_row_tiles = (a.n + kTileRows - 1) / kTileRows;
  const unsigned grid = static_cast<unsigned>(n_row_tiles < props.sm_count ? n_row_tiles : props.sm_count);
  static const int debug = [] {
    const char* e = getenv("HIDVAE_TC_DEBUG");
    return e != nullptr ? atoi(e) : 0;
  }();
  V10Params p{packed, plan.tile_bytes, plan.a_bytes, debug};
  auto go = [&](auto kernel) -> int {
    cudaFuncAttributes attr;
    HV_CUDA_CHECK(cudaFuncGetAttributes(&attr, kernel));
    if (attr.numRegs < kLaunchRegs) {  // setmaxnreg.inc would wait forever: refuse loudly instead
      set_error("hv_rq_forward: tcgen05 kernel (generation 10) was built with %d registers/thread, the register hand-over needs %d",
                attr.numRegs, kLaunchRegs);
      return HV_ERR_UNSUPPORTED;
    }
    HV_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes));
    kernel<<<grid, kThreads, plan.smem_bytes, stream>>>(a, p);
    HV_CUDA_CHECK(cudaGetLastError());
    return HV_OK;
  };
  return rot ? go(rq_fwd_tc_v10_kernel<D, true>) : go(rq_fwd_tc_v10_kernel<D, false>);
}

}  // namespace

bool rq_fwd_tc_v10_supported(int d, int k, int n_levels) {
  Plan plan;
  return make_plan(d, k, n_levels, &plan);
}

// `packed` = operand image written by launch_rq_pack (one 256-code image per level)
int launch_rq_fwd_tc_v10(const RqFwdArgs& a, int d, bool rot, const void* packed, cudaStream_t stream) {
  Plan plan;
  if (!make_plan(d, a.k, a.n_levels, &plan)) {
    set_error("hv_rq_forward: no generation-10 tcgen05 instantiation for D=%d K=%d L=%d", d, a.k, a.n_levels);
    return HV_ERR_UNSUPPORTED;
  }
  if (a.n == 0) return HV_OK;
  const uint8_t* img = static_cast<const uint8_t*>(packed);
  switch (d) {
    case 16: return launch_d<16>(a, rot, plan, img, stream);
    case 32: return launch_d<32>(a, rot, plan, img, stream);
    default: return HV_ERR_UNSUPPORTED;
  }
}

}  // namespace hv
