// Top-k of the per-anchor scores in Anchor3DHead.get_bboxes_single (model/PointPillars.py:1056-1065:
// `_, topk_inds = max_scores.topk(nms_pre)`), as a radix select + ordered compaction + stable sort of the k survivors.
// Output: the indices of the k largest scores in descending score order; equal scores: lower index first (a
// deterministic refinement of torch.topk, whose order among equal values is unspecified).
//
//   3 x topk_hist_kernel   histogram of the next 11 / 11 / 10 key bits over the elements that match the bits decided so
//                          far (key = ~ordered(score): ascending key = descending score); every CTA re-derives the
//                          decided prefix from the previous histograms (a 2048-bin scan), so there is no extra launch
//   topk_compact_kernel    threshold key T and `need` (how many elements equal to T are taken) from the last histogram;
//                          ordered compaction (tiles in ticket order, decoupled look-back) of every element with
//                          key < T plus the first `need` elements with key == T, in index order
//   sort_pairs_u32         stable radix sort of the k (key, index) pairs -> (score desc, index asc)
//   topk_rows_kernel       int64 indices
// Each pass streams the scores once (7.7 MB at 400 x 400 x 12 anchors): HBM / L2 bandwidth work.
#include "pp_common.cuh"
#include "pp_sort.cuh"

namespace pp {
namespace {

constexpr int TK_THREADS = 256, TK_ITEMS = 16, TK_TILE = TK_THREADS * TK_ITEMS;
constexpr int TK_BINS = 2048;
__device__ __constant__ int TK_SHIFT[3] = {21, 10, 0};
__device__ __constant__ int TK_WIDTH[3] = {11, 11, 10};

struct TopkWs {
    uint32_t *hist;        // [3][TK_BINS]
    uint32_t *ticket;      // tile tickets of the compaction
    unsigned long long *status;   // [tiles] look-back state: flag (2 bits) | equals (31 bits) | greater (31 bits)
    uint32_t *ckey, *cidx, *skey, *sidx;   // [k]
    void *sort_ws;
    size_t sort_ws_bytes, zero_bytes;
    int64_t tiles;
};

TopkWs carve(void *ws, int64_t n, int64_t k, size_t *total)
{
    TopkWs w;
    w.tiles = ceil_div(n > 0 ? n : 1, TK_TILE);
    Arena a(ws, (size_t)-1);
    w.hist = a.take<uint32_t>(3 * TK_BINS);
    w.ticket = a.take<uint32_t>(4);
    w.status = a.take<unsigned long long>((size_t)w.tiles);
    w.zero_bytes = a.off;
    const size_t kk = (size_t)(k > 0 ? k : 1);
    w.ckey = a.take<uint32_t>(kk);
    w.cidx = a.take<uint32_t>(kk);
    w.skey = a.take<uint32_t>(kk);
    w.sidx = a.take<uint32_t>(kk);
    w.sort_ws_bytes = sort_workspace_bytes((int64_t)kk);
    w.sort_ws = a.take<char>(w.sort_ws_bytes);
    *total = align_up(a.off);
    return w;
}

__device__ __forceinline__ uint32_t topk_key(float s) { return ~ordered_bits(s); }

// Block-cooperative: smallest bin b with hist[0] + ... + hist[b] >= krem; krem_out = krem - (sum below b).
// All threads of the CTA call it; the result is returned to all of them.
__device__ void tk_find(const uint32_t *__restrict__ hist, int nbins, uint32_t krem, uint32_t &bin, uint32_t &krem_out)
{
    __shared__ uint32_t s_part[TK_THREADS / 32];
    __shared__ uint32_t s_res[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = nbins / TK_THREADS;                    // 8 or 4 consecutive bins per thread
    uint32_t v[TK_BINS / TK_THREADS], sum = 0;
#pragma unroll
    for (int j = 0; j < TK_BINS / TK_THREADS; ++j) {
        v[j] = j < per ? hist[tid * per + j] : 0u;
        sum += v[j];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    uint32_t before = incl - sum;
    for (int wq = 0; wq < warp; ++wq) before += s_part[wq];
    if (before < krem && krem <= before + sum) {
        uint32_t run = before;
#pragma unroll
        for (int j = 0; j < TK_BINS / TK_THREADS; ++j) {
            if (j < per && run < krem && krem <= run + v[j]) { s_res[0] = (uint32_t)(tid * per + j); s_res[1] = krem - run; }
            run += v[j];
        }
    }
    __syncthreads();
    bin = s_res[0];
    krem_out = s_res[1];
    __syncthreads();
}

// the key bits decided by passes 0 .. upto-1, and how many elements are still to be taken inside that prefix
__device__ void tk_prefix(const uint32_t *__restrict__ hist, int upto, uint32_t k, uint32_t &prefix, uint32_t &krem)
{
    prefix = 0;
    krem = k;
    for (int p = 0; p < upto; ++p) {
        uint32_t bin, kr;
        tk_find(hist + p * TK_BINS, 1 << TK_WIDTH[p], krem, bin, kr);
        prefix |= bin << TK_SHIFT[p];
        krem = kr;
    }
}

template <int PASS>
__global__ void __launch_bounds__(TK_THREADS)
topk_hist_kernel(const float *__restrict__ scores, int64_t n, uint32_t k, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_hist[TK_BINS];
    for (int i = threadIdx.x; i < TK_BINS; i += TK_THREADS) s_hist[i] = 0;
    uint32_t prefix, krem;
    tk_prefix(hist, PASS, k, prefix, krem);               // (ends with a barrier: s_hist is zeroed for everybody)
    const uint32_t hi_mask = PASS == 0 ? 0u : (0xFFFFFFFFu << (TK_SHIFT[PASS] + TK_WIDTH[PASS]));
    const uint32_t dmask = (1u << TK_WIDTH[PASS]) - 1u;
    if (PASS == 0) __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * TK_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * TK_THREADS) {
        const uint32_t key = topk_key(scores[i]);
        if ((key & hi_mask) == prefix) atomicAdd(&s_hist[(key >> TK_SHIFT[PASS]) & dmask], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TK_BINS; i += TK_THREADS)
        if (s_hist[i]) atomicAdd(hist + PASS * TK_BINS + i, s_hist[i]);
}

constexpr unsigned long long TKF_AGG = 1ull << 62, TKF_PREFIX = 2ull << 62, TKF_MASK = 3ull << 62;

__global__ void __launch_bounds__(TK_THREADS)
topk_compact_kernel(const float *__restrict__ scores, int64_t n, uint32_t k, const uint32_t *__restrict__ hist,
                    uint32_t *__restrict__ ticket, unsigned long long *status, uint32_t *__restrict__ ckey,
                    uint32_t *__restrict__ cidx)
{
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_warp[TK_THREADS / 32], s_excl;
    uint32_t T, need;
    tk_prefix(hist, 3, k, T, need);                        // the k-th smallest key, and how many elements equal to it are taken
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // blocked arrangement: a thread owns TK_ITEMS consecutive elements, so that scans preserve the index order
    const int64_t i0 = tile * TK_TILE + (int64_t)tid * TK_ITEMS;
    uint32_t key[TK_ITEMS];
    uint32_t g = 0, e = 0;
#pragma unroll
    for (int j = 0; j < TK_ITEMS; ++j) {
        key[j] = i0 + j < n ? topk_key(scores[i0 + j]) : 0xFFFFFFFFu;
        if (i0 + j < n) {
            g += key[j] < T ? 1u : 0u;
            e += key[j] == T ? 1u : 0u;
        }
    }
    // CTA-wide exclusive scan of (equals, greater) packed as (e << 31 | g): both < 2^31
    const unsigned long long mine = ((unsigned long long)e << 31) | g;
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned lo = __shfl_up_sync(0xFFFFFFFFu, (unsigned)incl, o), hi = __shfl_up_sync(0xFFFFFFFFu, (unsigned)(incl >> 32), o);
        if (lane >= o) incl += ((unsigned long long)hi << 32) | lo;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long before = incl - mine, total = 0;
#pragma unroll
    for (int wq = 0; wq < TK_THREADS / 32; ++wq) {
        const unsigned long long c = s_warp[wq];
        if (wq < warp) before += c;
        total += c;
    }
    if (tid == 0) {
        unsigned long long excl = 0;
        if (tile == 0) {
            atomicExch(status, TKF_PREFIX | total);
        } else {
            atomicExch(status + tile, TKF_AGG | total);
            int64_t t = tile - 1;
            while (true) {
                const unsigned long long v = *((volatile unsigned long long *)(status + t));
                if ((v & TKF_MASK) == 0) continue;
                excl += v & ~TKF_MASK;
                if ((v & TKF_MASK) == TKF_PREFIX) break;
                --t;
            }
            atomicExch(status + tile, TKF_PREFIX | (excl + total));
        }
        s_excl = excl;
    }
    __syncthreads();
    const unsigned long long ex = s_excl + before;         // (equals, greater) before this thread's first element
    uint32_t eq_before = (uint32_t)(ex >> 31), gr_before = (uint32_t)(ex & 0x7FFFFFFFull);
#pragma unroll
    for (int j = 0; j < TK_ITEMS; ++j) {
        if (i0 + j >= n) break;
        const bool gr = key[j] < T, eq = key[j] == T;
        if (gr || (eq && eq_before < need)) {
            // selected elements in index order: all greater ones before it + the taken equal ones before it
            const uint32_t pos = gr_before + (eq_before < need ? eq_before : need);
            ckey[pos] = key[j];
            cidx[pos] = (uint32_t)(i0 + j);
        }
        gr_before += gr ? 1u : 0u;
        eq_before += eq ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(256) topk_rows_kernel(const uint32_t *__restrict__ sidx, int64_t k, int64_t *__restrict__ rows)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < k) rows[i] = (int64_t)sidx[i];
}

}  // namespace
}  // namespace pp

using namespace pp;

extern "C" size_t pp_head_topk_workspace_bytes(int64_t n, int64_t k)
{
    size_t total;
    carve(nullptr, n, k < n ? k : n, &total);
    return total;
}

extern "C" int pp_head_topk(const float *scores, int64_t n, int64_t k, int64_t *rows, void *workspace,
                            size_t workspace_bytes, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    PP_REQUIRE(n >= 0 && n < (1ll << 31) && k >= 0, "bad n / k");
    if (k > n) k = n;
    if (k == 0) return PP_OK;
    PP_REQUIRE(scores && rows && workspace, "null pointer");
    size_t total;
    TopkWs w = carve(workspace, n, k, &total);
    if (workspace_bytes < total) {
        set_error("top-k workspace too small: %zu < %zu", workspace_bytes, total);
        return PP_ERR_WORKSPACE;
    }
    PP_CUDA_TRY(cudaMemsetAsync(workspace, 0, w.zero_bytes, st));
    prof_mark("memset");
    const unsigned hb = (unsigned)(w.tiles < 148 * 8 ? w.tiles : 148 * 8);
    topk_hist_kernel<0><<<hb, TK_THREADS, 0, st>>>(scores, n, (uint32_t)k, w.hist);
    if (int rc = check_launch("topk_hist_kernel")) return rc;
    topk_hist_kernel<1><<<hb, TK_THREADS, 0, st>>>(scores, n, (uint32_t)k, w.hist);
    if (int rc = check_launch("topk_hist_kernel")) return rc;
    topk_hist_kernel<2><<<hb, TK_THREADS, 0, st>>>(scores, n, (uint32_t)k, w.hist);
    if (int rc = check_launch("topk_hist_kernel")) return rc;
    topk_compact_kernel<<<(unsigned)w.tiles, TK_THREADS, 0, st>>>(scores, n, (uint32_t)k, w.hist, w.ticket, w.status, w.ckey, w.cidx);
    if (int rc = check_launch("topk_compact_kernel")) return rc;
    if (int rc = sort_pairs_u32(w.ckey, w.cidx, w.skey, w.sidx, k, w.sort_ws, w.sort_ws_bytes, st)) return rc;
    topk_rows_kernel<<<(unsigned)ceil_div(k, 256), 256, 0, st>>>(w.sidx, k, rows);
    return check_launch("topk_rows_kernel");
}
