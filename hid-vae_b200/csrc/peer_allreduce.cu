// One-shot all-reduce (sum) of a small fp32 buffer over NVLink / NVSwitch peer memory -- the data-parallel exchange of
// the codebook gradient (98 KB per step).  At this size a collective is pure latency: NCCL's
// all-reduce measured 16 us at 2 GPUs and 33 us at 8 on top of a 41 us step.  Here every rank PUSHES its contribution
// straight into an inbox slot on every peer (posted 16-byte stores through the peer mappings of a symmetric allocation),
// publishes one flag per (peer, chunk) with release.sys semantics, waits for the flags of its own chunk and sums the
// `world` slots in rank order -- so every rank ends with bit-identical sums.  One launch, no intermediate hop, no
// dependency between chunks: a CTA owns one 1024-float chunk end to end, so there is no grid-wide barrier.
//
//   inbox (per rank, symmetric):  [2 parities][world source ranks][n floats]
//   flags (per rank, symmetric):  [2 parities][world source ranks][n_chunks] uint32, zero before the first call
//   seq   (per rank, private):    [n_chunks] uint32 call counters, zero before the first call
// The call number lives in device memory (each CTA bumps its own counter), so a CUDA graph can replay the launch.
// Two parities are enough: a rank reaches call s + 2 only after every peer has flagged call s + 1, i.e. has left call s.
#include "common.cuh"

namespace hv {
namespace {

constexpr int kPeerThreads = 256;
constexpr int kChunkFloats = kPeerThreads * 4;
constexpr int kMaxWorld = 16;

struct PeerArgs {
  const float* src;
  float* dst;
  int64_t n;
  const uint64_t* inbox_ptrs;  // device array [world]: base of every rank's inbox in this rank's address space
  const uint64_t* flag_ptrs;   // device array [world]: base of every rank's flags
  uint32_t* seq;
  int rank;
  int world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(PeerArgs a) {
  __shared__ uint32_t s_seq;
  const int chunk = blockIdx.x, n_chunks = gridDim.x;
  if (threadIdx.x == 0) s_seq = ++a.seq[chunk];  // this CTA's call number (only this CTA touches the counter)
  __syncthreads();
  const uint32_t seq = s_seq;
  const int64_t par = seq & 1u;
  const int64_t i = static_cast<int64_t>(chunk) * kChunkFloats + threadIdx.x * 4;
  const bool live = i < a.n;  // (n is a multiple of 4)

  // ---- push: my piece of the chunk into slot [par][rank] of every rank's inbox (my own included) ----
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) v = *reinterpret_cast<const float4*>(a.src + i);
  const int64_t slot = (par * a.world + a.rank) * a.n + i;
  for (int r = 0; r < a.world; ++r) {
    const int peer = (a.rank + r) % a.world;  // start with myself, spread the first stores over the links
    if (live) *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.inbox_ptrs[peer]) + slot) = v;
  }
  // One release per (CTA, peer): the CTA barrier orders every thread's pushes before the flag thread, and st.release.sys is
  // cumulative over them (a __threadfence_system() in every thread in front of the barrier cost a second fence round trip).
  __syncthreads();
  if (threadIdx.x < a.world) {
    const int peer = threadIdx.x;
    uint32_t* flag = reinterpret_cast<uint32_t*>(a.flag_ptrs[peer]) + (par * a.world + a.rank) * n_chunks + chunk;
    st_release_sys(flag, seq);
  }

  // ---- wait for this chunk from every rank, then sum the slots in rank order ----
  if (threadIdx.x < a.world) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(a.flag_ptrs[a.rank]) + (par * a.world + threadIdx.x) * n_chunks + chunk;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) != seq) {
      if (clock64() - t0 > 20000000000LL) {  // ~10 s: a rank that never shows up must not hang the box
        printf("hidvae_b200: peer all-reduce timed out (rank %d waits for rank %d, chunk %d, call %u)\n", a.rank, threadIdx.x, chunk, seq);
        __trap();
      }
    }
  }
  __syncthreads();
  if (live) {
    const float* mine = reinterpret_cast<const float*>(a.inbox_ptrs[a.rank]) + par * a.world * a.n + i;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < a.world; ++r) {
      const float4 t = __ldcv(reinterpret_cast<const float4*>(mine + static_cast<int64_t>(r) * a.n));  // (written by peers: no stale cache line)
      s.x += t.x, s.y += t.y, s.z += t.z, s.w += t.w;
    }
    *reinterpret_cast<float4*>(a.dst + i) = s;
  }
}

}  // namespace
}  // namespace hv

extern "C" {

int hv_peer_allreduce_chunks(int64_t n) { return n <= 0 ? 0 : static_cast<int>((n + hv::kChunkFloats - 1) / hv::kChunkFloats); }

int hv_peer_allreduce(const float* src, float* dst, int64_t n, const uint64_t* inbox_ptrs, const uint64_t* flag_ptrs,
                      uint32_t* seq, int rank, int world, void* stream) {
  using namespace hv;
  if (n <= 0 || n % 4 != 0 || world < 1 || world > kMaxWorld || rank < 0 || rank >= world) {
    set_error("hv_peer_allreduce: bad arguments n=%lld (multiple of 4) world=%d (<= %d) rank=%d", (long long)n, world, kMaxWorld, rank);
    return HV_ERR_BAD_SHAPE;
  }
  if (!src || !dst || !inbox_ptrs || !flag_ptrs || !seq) {
    set_error("hv_peer_allreduce: null pointer");
    return HV_ERR_NULL;
  }
  if (!aligned16(src) || !aligned16(dst)) {
    set_error("hv_peer_allreduce: src and dst must be 16-byte aligned");
    return HV_ERR_MISALIGNED;
  }
  PeerArgs a{src, dst, n, inbox_ptrs, flag_ptrs, seq, rank, world};
  peer_allreduce_kernel<<<hv_peer_allreduce_chunks(n), kPeerThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  HV_CUDA_CHECK(cudaGetLastError());
  return HV_OK;
}

}  // extern "C"
