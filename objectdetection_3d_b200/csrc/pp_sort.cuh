// Internal interface of the radix sort (pp_sort.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pp {
size_t sort_workspace_bytes(int64_t n);
// leading bytes of the workspace that hold the sort's counters (histograms, tickets, look-back state)
size_t sort_zero_bytes(int64_t n);
// vals_in == nullptr means "iota": the value of element i is i.  ws_zeroed: the caller has already zeroed the first
// sort_zero_bytes(n) bytes of ws on this stream (as part of a memset of its own), so the sort issues none.
int sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                   int64_t n, void *ws, size_t ws_bytes, cudaStream_t st, bool ws_zeroed = false);
}  // namespace pp
