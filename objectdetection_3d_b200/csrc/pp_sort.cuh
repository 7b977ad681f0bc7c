// Internal interface of the radix sort (pp_sort.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pp {
size_t sort_workspace_bytes(int64_t n);
// vals_in == nullptr means "iota": the value of element i is i.
int sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                   int64_t n, void *ws, size_t ws_bytes, cudaStream_t st);
}  // namespace pp
