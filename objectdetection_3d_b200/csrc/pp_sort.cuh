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
// have_hist (needs ws_zeroed): the kernel that produced keys_in has also accumulated the four 8-bit digit histograms of
// ALL n keys into sort_hist(ws, n) (sort_hist_add below), so the sort skips its histogram launch; sort_hist returns
// nullptr for sizes the one-CTA path sorts (no histogram needed).
// tail (multi-pass sizes only, see sort_runs_tail): work the LAST digit pass does on the side, saving the caller a launch:
// out[a][rank] = in[a][value] for every element and each of `arrays` float4 arrays (the payload moves with its key), and
// *count_out = min(*count_in, count_cap).
struct SortTail {
    const float4 *in[4];
    float4 *out[4];
    int arrays;
    const int32_t *count_in;
    int32_t *count_out;
    int count_cap;
};
bool sort_runs_tail(int64_t n);
int sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                   int64_t n, void *ws, size_t ws_bytes, cudaStream_t st, bool ws_zeroed = false, bool have_hist = false,
                   const SortTail *tail = nullptr);
uint32_t *sort_hist(void *ws, int64_t n);

#ifdef __CUDACC__
// block-wide: every thread passes its key (valid = false for none); sh is 1024 words of shared memory
__device__ __forceinline__ void sort_hist_add(uint32_t *sh, uint32_t *hist, uint32_t key, bool valid)
{
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    if (valid) {
#pragma unroll
        for (int p = 0; p < 4; ++p) atomicAdd(&sh[p * 256 + ((key >> (8 * p)) & 255u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
#endif
}  // namespace pp
