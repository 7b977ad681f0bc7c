// One class of multiclass_nms (model/utils.py:376-424, nms_dim == 2):
//   prepare : AABB of the rotated box (ops/ops_torch.py:13-114), score filter (strict >, :381), sort key
//   sort    : descending score, stable (ties: lower index first)  [pp_sort.cu]
//   mask    : 64x64 tiles of the upper triangle, bit j of mask[i][cb] = iou(box_j, box_i) > thr (strict, :413)
//   sweep   : one CTA walks the 64-box blocks in order; a warp resolves each diagonal block with a
//             64-step register chain, then all threads OR the kept rows into the running removed mask.
// The IoU is evaluated exactly like bbox_iou2D (pp_boxes.cuh::rect_iou), so the keep set is identical
// to the reference's greedy loop on the same rectangles.
#include "pp_boxes.cuh"
#include "pp_common.cuh"
#include "pp_sort.cuh"

namespace pp {
namespace {

typedef unsigned long long u64;
constexpr int NMS_THREADS = 256;
constexpr int SWEEP_THREADS = 1024;

__global__ void __launch_bounds__(NMS_THREADS)
nms_prepare_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, int64_t stride, int64_t N,
                   float score_thr, float4 *__restrict__ rect, uint32_t *__restrict__ keys, int32_t *__restrict__ n_cand)
{
    int64_t i = (int64_t)blockIdx.x * NMS_THREADS + threadIdx.x;
    bool cand = false;
    if (i < N) {
        float b[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) b[k] = boxes[i * 9 + k];
        float c[8][3];
        box_corners(b, c);
        rect[i] = corners_to_rect(c);
        float s = scores[i * stride];
        cand = s > score_thr;
        keys[i] = cand ? ~ordered_bits(s) : 0xFFFFFFFFu;
    }
    unsigned m = __ballot_sync(0xFFFFFFFFu, cand);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_cand, __popc(m));
}

__global__ void __launch_bounds__(NMS_THREADS)
nms_gather_kernel(const float4 *__restrict__ rect, const uint32_t *__restrict__ order, const int32_t *__restrict__ n_cand,
                  float4 *__restrict__ srect)
{
    int64_t r = (int64_t)blockIdx.x * NMS_THREADS + threadIdx.x;
    if (r < *n_cand) srect[r] = rect[order[r]];
}

constexpr int MT_ROWS = 128;   // rows (selected boxes) per CTA, one per thread
constexpr int MT_COLS = 256;   // columns (remaining boxes) per CTA = 4 mask words

// bit j of mask[i][w] = (64w + j > i) && iou(box_{64w+j}, box_i) > thr.  Words entirely on or below the
// diagonal are never read by the sweep and are not written.
__global__ void __launch_bounds__(MT_ROWS)
nms_mask_kernel(const float4 *__restrict__ srect, const int32_t *__restrict__ n_cand, float thr, int nw_stride,
                u64 *__restrict__ mask)
{
    const int n = *n_cand;
    const int row0 = blockIdx.y * MT_ROWS, col0 = blockIdx.x * MT_COLS;
    if (row0 >= n || col0 >= n || col0 + MT_COLS <= row0) return;
    __shared__ float4 s_col[MT_COLS];
    const int t = threadIdx.x;
#pragma unroll
    for (int k = 0; k < MT_COLS / MT_ROWS; ++k) {
        const int c = col0 + t + k * MT_ROWS;
        // out-of-range columns get an empty rectangle: never intersects
        s_col[t + k * MT_ROWS] = c < n ? srect[c] : make_float4(3e38f, 3e38f, -3e38f, -3e38f);
    }
    __syncthreads();
    const int i = row0 + t;
    if (i >= n) return;
    const float4 a = srect[i];
    const bool zero_hits = 0.f > thr;        // iou == 0 still "exceeds" a negative threshold
#pragma unroll 1
    for (int wd = 0; wd < MT_COLS / 64; ++wd) {
        const int c_start = col0 + wd * 64;
        if (c_start >= n) break;
        if (c_start + 63 < i) continue;       // entirely below the diagonal: never read
        u64 bits = 0;
#pragma unroll 8
        for (int j = 0; j < 64; ++j) {
            const float4 q = s_col[wd * 64 + j];
            // fast reject: empty intersection -> overlap == 0 -> iou == 0 exactly
            const float w = __fsub_rn(fminf(q.z, a.z), fmaxf(q.x, a.x));
            const float h = __fsub_rn(fminf(q.w, a.w), fmaxf(q.y, a.y));
            bool hit = zero_hits;
            if (w > 0.f && h > 0.f) hit = rect_iou(q, a, 0, 1e-6f) > thr;   // (remaining, selected) order of :412
            bits |= (u64)hit << j;
        }
        if (c_start <= i) bits &= ~((2ull << (i - c_start)) - 1ull);      // keep only columns > i
        if (c_start + 64 > n) bits &= (1ull << (n - c_start)) - 1ull;     // and columns < n
        mask[(size_t)i * nw_stride + (c_start >> 6)] = bits;
    }
}

// Sweep: warp 0 ("resolver") walks the 64-box blocks in rank order.  For block cb it takes the final
// removed word, keeps the surviving boxes with a find-first-set chain over the prefetched diagonal word
// (one iteration per KEPT box), and ORs the kept rows' next word (cb+1) itself.  The other 31 warps
// ("spreaders") run one block behind and OR the kept rows of block cb-1 into the words >= cb+1, so their
// L2 latency overlaps the resolver's chain.  One CTA barrier per block.
__global__ void __launch_bounds__(SWEEP_THREADS)
nms_sweep_kernel(const u64 *__restrict__ mask, int nw_stride, const int32_t *__restrict__ n_cand,
                 const uint32_t *__restrict__ order, int64_t *__restrict__ keep, int32_t *__restrict__ keep_count)
{
    extern __shared__ u64 removed[];
    __shared__ u64 s_kept[2];
    const int n = *n_cand;
    const int nw = (n + 63) >> 6;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int w = tid; w < nw; w += SWEEP_THREADS) removed[w] = 0;
    if (tid < 2) s_kept[tid] = 0;
    __syncthreads();
    int kept_total = 0;
    // band[k] = {diag lo, diag hi, next lo, next hi} of block cb + k, prefetched two blocks ahead
    u64 band[3][4];
    auto load_band = [&](int b, u64 (&d)[4]) {
        d[0] = d[1] = d[2] = d[3] = 0;
        if (b < nw) {
            const int r0 = b * 64 + lane, r1 = r0 + 32;
            if (r0 < n) d[0] = mask[(size_t)r0 * nw_stride + b];
            if (r1 < n) d[1] = mask[(size_t)r1 * nw_stride + b];
            if (b + 1 < nw) {
                if (r0 < n) d[2] = mask[(size_t)r0 * nw_stride + b + 1];
                if (r1 < n) d[3] = mask[(size_t)r1 * nw_stride + b + 1];
            }
        }
    };
    if (tid < 32) {
        load_band(0, band[0]);
        load_band(1, band[1]);
    }
    for (int cb = 0; cb <= nw; ++cb) {
        if (tid < 32) {
            if (cb < nw) {
                load_band(cb + 2, band[2]);
                const int valid = min(64, n - cb * 64);
                u64 alive = ~removed[cb];
                if (valid < 64) alive &= (1ull << valid) - 1;
                u64 kept = 0;
                while (alive) {
                    const int i = __ffsll((long long)alive) - 1;
                    kept |= 1ull << i;
                    const u64 di = __shfl_sync(0xFFFFFFFFu, i < 32 ? band[0][0] : band[0][1], i & 31);
                    alive &= ~(di | (1ull << i));
                }
                if (lane == 0) s_kept[cb & 1] = kept;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int bit = lane + 32 * h;
                    if ((kept >> bit) & 1ull)
                        keep[kept_total + __popcll(kept & ((1ull << bit) - 1))] = (int64_t)order[cb * 64 + bit];
                }
                kept_total += __popcll(kept);
                // the kept rows' word cb+1 must be final before the next block is resolved
                u64 nx = (((kept >> lane) & 1ull) ? band[0][2] : 0ull) | (((kept >> (lane + 32)) & 1ull) ? band[0][3] : 0ull);
                const unsigned lo = __reduce_or_sync(0xFFFFFFFFu, (unsigned)nx);
                const unsigned hi = __reduce_or_sync(0xFFFFFFFFu, (unsigned)(nx >> 32));
                if (lane == 0 && cb + 1 < nw && (lo | hi)) atomicOr(&removed[cb + 1], ((u64)hi << 32) | lo);
#pragma unroll
                for (int k = 0; k < 4; ++k) { band[0][k] = band[1][k]; band[1][k] = band[2][k]; }
            }
        } else if (cb >= 1) {
            // spreaders: block cb-1 -> words >= cb+1
            const int pb = cb - 1;
            const u64 kept = s_kept[pb & 1];
            if (kept) {
                for (int w = cb + 1 + (tid - 32); w < nw; w += SWEEP_THREADS - 32) {
                    u64 acc = 0, k = kept;
                    while (k) {
                        const int i = __ffsll((long long)k) - 1;
                        k &= k - 1;
                        acc |= mask[(size_t)(pb * 64 + i) * nw_stride + w];
                    }
                    if (acc) atomicOr(&removed[w], acc);
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) *keep_count = kept_total;
}

struct NmsWs {
    int32_t *n_cand;
    float4 *rect, *srect;
    uint32_t *keys, *keys_sorted, *order;
    u64 *mask;
    void *sort_ws;
    size_t sort_ws_bytes;
    int nw;
};

NmsWs carve(void *ws, int64_t N, size_t *total)
{
    NmsWs w;
    int64_t n1 = N > 0 ? N : 1;
    w.nw = (int)ceil_div(n1, 64);
    Arena a(ws, (size_t)-1);
    w.n_cand = a.take<int32_t>(64);
    w.rect = a.take<float4>((size_t)n1);
    w.srect = a.take<float4>((size_t)n1);
    w.keys = a.take<uint32_t>((size_t)n1);
    w.keys_sorted = a.take<uint32_t>((size_t)n1);
    w.order = a.take<uint32_t>((size_t)n1);
    w.mask = a.take<u64>((size_t)n1 * w.nw);
    w.sort_ws_bytes = sort_workspace_bytes(n1);
    w.sort_ws = a.take<char>(w.sort_ws_bytes);
    *total = align_up(a.off);
    return w;
}

}  // namespace
}  // namespace pp

using namespace pp;

extern "C" size_t pp_nms_workspace_bytes(int64_t N)
{
    size_t total;
    carve(nullptr, N, &total);
    return total;
}

extern "C" int pp_nms(const float *boxes9, const float *scores, int64_t score_stride, int64_t N, float score_thr,
                      float iou_thr, int64_t *keep, int32_t *keep_count, void *workspace, size_t workspace_bytes,
                      pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    PP_REQUIRE(keep_count, "null keep_count");
    PP_REQUIRE(N >= 0 && N <= 131072, "N must be in [0, 131072]");
    if (N == 0) {
        PP_CUDA_TRY(cudaMemsetAsync(keep_count, 0, sizeof(int32_t), st));
        return PP_OK;
    }
    PP_REQUIRE(boxes9 && scores && keep && workspace, "null pointer");
    PP_REQUIRE(score_stride >= 1, "bad score stride");
    size_t total;
    NmsWs w = carve(workspace, N, &total);
    if (workspace_bytes < total) {
        set_error("nms workspace too small: %zu < %zu", workspace_bytes, total);
        return PP_ERR_WORKSPACE;
    }
    PP_CUDA_TRY(cudaMemsetAsync(w.n_cand, 0, sizeof(int32_t), st));
    prof_mark("memset");
    const unsigned nb = (unsigned)ceil_div(N, NMS_THREADS);
    nms_prepare_kernel<<<nb, NMS_THREADS, 0, st>>>(boxes9, scores, score_stride, N, score_thr, w.rect, w.keys, w.n_cand);
    if (int rc = check_launch("nms_prepare_kernel")) return rc;
    if (int rc = sort_pairs_u32(w.keys, nullptr, w.keys_sorted, w.order, N, w.sort_ws, w.sort_ws_bytes, st)) return rc;
    nms_gather_kernel<<<nb, NMS_THREADS, 0, st>>>(w.rect, w.order, w.n_cand, w.srect);
    if (int rc = check_launch("nms_gather_kernel")) return rc;
    dim3 grid((unsigned)ceil_div(N, MT_COLS), (unsigned)ceil_div(N, MT_ROWS));
    nms_mask_kernel<<<grid, MT_ROWS, 0, st>>>(w.srect, w.n_cand, iou_thr, w.nw, w.mask);
    if (int rc = check_launch("nms_mask_kernel")) return rc;
    size_t smem = (size_t)w.nw * sizeof(u64);
    nms_sweep_kernel<<<1, SWEEP_THREADS, smem, st>>>(w.mask, w.nw, w.n_cand, w.order, keep, keep_count);
    return check_launch("nms_sweep_kernel");
}
