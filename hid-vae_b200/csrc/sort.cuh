// Stable key sort of row indices (cub::DeviceRadixSort on 64-bit keys) shared by the k-means update (sort rows by
// cluster -> segmented sums) and the uniqueness loss (sort rows by id tuple -> runs of identical tuples).
// The library never allocates: the caller provides hv_sort_workspace_bytes(n) bytes.
#pragma once

#include "common.cuh"

namespace hv {

struct SortBuffers {
  uint64_t* keys_in;
  uint64_t* keys_out;
  uint32_t* vals_in;
  uint32_t* vals_out;
  void* temp;
  size_t temp_bytes;
};

size_t sort_workspace_bytes(int64_t n);
// false when the workspace is missing, misaligned or too small for n rows
bool sort_carve(void* workspace, size_t workspace_bytes, int64_t n, SortBuffers* out);
// keys_in / vals_in -> keys_out / vals_out, ascending, stable, bits [0, end_bit) of the key
int sort_pairs(const SortBuffers& b, int64_t n, int end_bit, cudaStream_t stream);

}  // namespace hv
