// The steps either side of the hot path, on the device so that a raw tile is uploaded once:
//   before voxelization : PointPillars.preprocess (model/PointPillars.py:241-266) = global_outlier_check
//                         (ops/ops_numpy.py:111-115) + range filter + feature selection, and the centroid step of
//                         CustomVoxelizer.voxelize (model/utils.py:15-43)
//   after the scatter   : the dense -> sparse glue of SubmanifoldSparseRPN.forward (model/PointPillars.py:766-789)
// All three end in an order-preserving compaction (numpy boolean indexing / torch.where order), built from one
// primitive: per-CTA counts -> one-CTA scan -> scatter.
#include "pp_common.cuh"

namespace pp {
namespace {

constexpr int PT_THREADS = 256;

// ---- order-preserving compaction of the set bits of a flag array -------------------------------------------------
__global__ void __launch_bounds__(PT_THREADS)
flag_count_kernel(const uint8_t *__restrict__ flags, int64_t n, int32_t *__restrict__ block_count)
{
    const int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x;
    const int c = __syncthreads_count(i < n && flags[i]);
    if (threadIdx.x == 0) block_count[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024) block_scan_kernel(int32_t *__restrict__ block_count, int64_t nb, int32_t *__restrict__ total)
{
    // exclusive scan of nb counts by one CTA, 1024 at a time with a running carry
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < nb; b0 += 1024) {
        const int64_t i = b0 + tid;
        const int v = i < nb ? block_count[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int base = s_carry + incl - v;
        for (int k = 0; k < warp; ++k) base += s_warp[k];
        if (i < nb) block_count[i] = base;
        __syncthreads();
        if (tid == 1023) s_carry = base + v;
        __syncthreads();
    }
    if (tid == 0) *total = s_carry;
}

// position of element i among the set flags, or -1
__device__ __forceinline__ int compact_slot(const uint8_t *flags, int64_t n, const int32_t *block_base, int64_t i)
{
    __shared__ int s_w[PT_THREADS / 32];
    const bool f = i < n && flags[i];
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    int base = block_base[blockIdx.x];
    for (int k = 0; k < warp; ++k) base += s_w[k];
    return f ? base + __popc(bal & lanemask_lt()) : -1;
}

// ---- preprocess: statistics -------------------------------------------------------------------------------------
// sums[0..2] += xyz (float64);  second pass: sums[3] += norm, sums[4] += norm^2 with norm = |p - mean| (float32 like
// the reference's array math, accumulated in float64)
__global__ void __launch_bounds__(PT_THREADS)
sum_xyz_kernel(const float *__restrict__ pts, int64_t n, int C, double *sums)
{
    double s[3] = {0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT_THREADS)
        for (int k = 0; k < 3; ++k) s[k] += (double)pts[i * C + k];
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xFFFFFFFFu, s[k], o);
        if ((threadIdx.x & 31) == 0) atomicAdd(sums + k, s[k]);
    }
}

__device__ __forceinline__ float point_norm(const float *p, const float mean[3])
{
    // ((p - mean) ** 2).sum(axis=1) ** 0.5 in float32, ops/ops_numpy.py:113
    const float dx = __fsub_rn(p[0], mean[0]), dy = __fsub_rn(p[1], mean[1]), dz = __fsub_rn(p[2], mean[2]);
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

__global__ void __launch_bounds__(PT_THREADS)
norm_stats_kernel(const float *__restrict__ pts, int64_t n, int C, double *sums)
{
    const float mean[3] = {(float)(sums[0] / (double)n), (float)(sums[1] / (double)n), (float)(sums[2] / (double)n)};
    double s1 = 0, s2 = 0;
    for (int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT_THREADS) {
        const double v = (double)point_norm(pts + i * C, mean);
        s1 += v;
        s2 += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o);
        s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(sums + 3, s1); atomicAdd(sums + 4, s2); }
}

// ---- preprocess: statistics in numpy's own order of operations (exact mode) ------------------------------------------
// global_outlier_check is float32 numpy code (ops/ops_numpy.py:111-115) and which rows it keeps depends on how numpy
// rounds its reductions (probed against numpy 2.3, the version the oracle is pinned with):
//   np.mean(a[:, :3], axis=0)   the reduced axis is not the contiguous one -> numpy adds row after row into one float32
//                               accumulator per column: a SEQUENTIAL sum (it differs from the pairwise and from the
//                               float64 result in the 5th digit at 2e5 points)
//   np.mean / np.std of the 1-D norm array: pairwise summation -- halves (left half rounded down to a multiple of 8)
//                               down to blocks of <= 128 elements, each summed with 8 interleaved accumulators
// Both orders are reproduced literally.  The sequential sum is a dependent chain of float adds (three threads, the
// other threads of the CTA stage the next chunk of points in shared memory); the pairwise tree is walked by one thread
// while all threads of the CTA sum the leaves.
constexpr int NPS_THREADS = 1024, NPS_CHUNK = 1024, PW_BLOCK = 128;

__global__ void __launch_bounds__(NPS_THREADS)
np_colmean3_kernel(const float *__restrict__ pts, int64_t n, int C, float *__restrict__ stats)
{
    __shared__ float s_buf[2][3][NPS_CHUNK];
    const int tid = threadIdx.x;
    float acc = 0.f;
    const int64_t chunks = ceil_div(n, NPS_CHUNK);
    auto stage = [&](int64_t c) {
        const int64_t i = c * NPS_CHUNK + tid;
        if (c < chunks && i < n) {
#pragma unroll
            for (int k = 0; k < 3; ++k) s_buf[c & 1][k][tid] = pts[i * C + k];
        }
    };
    stage(0);
    __syncthreads();
    for (int64_t c = 0; c < chunks; ++c) {
        stage(c + 1);
        if (tid < 3) {
            const float *col = s_buf[c & 1][tid];
            const int m = (int)(n - c * NPS_CHUNK < NPS_CHUNK ? n - c * NPS_CHUNK : NPS_CHUNK);
            int j = 0;
            for (; j + 8 <= m; j += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = col[j + u];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, v[u]);
            }
            for (; j < m; ++j) acc = __fadd_rn(acc, col[j]);
        }
        __syncthreads();
    }
    if (tid < 3) stats[tid] = __fdiv_rn(acc, (float)n);
}

__device__ __forceinline__ float np_leaf_sum(const float *__restrict__ a, int n)
{
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
}

struct PwFrame { int64_t n; int st; float l; };

// out[0] = numpy's pairwise float32 sum of a[0, n).  One CTA; leaf_start / leaf_val: ceil(n / 64) + 2 entries.
__global__ void __launch_bounds__(NPS_THREADS)
np_pairwise_sum_kernel(const float *__restrict__ a, int64_t n, int64_t *__restrict__ leaf_start, float *__restrict__ leaf_val,
                       float *__restrict__ out)
{
    __shared__ int s_leaves;
    if (threadIdx.x == 0) {
        // the leaves of the recursion, left to right
        int64_t stk_s[64], stk_n[64];
        int sp = 0, L = 0;
        stk_s[0] = 0; stk_n[0] = n; sp = 1;
        while (sp) {
            const int64_t s0 = stk_s[sp - 1], m = stk_n[sp - 1];
            --sp;
            if (m <= PW_BLOCK) { leaf_start[L++] = s0; continue; }
            int64_t n2 = m / 2;
            n2 -= n2 % 8;
            stk_s[sp] = s0 + n2; stk_n[sp] = m - n2; ++sp;      // right half: popped after the left one
            stk_s[sp] = s0; stk_n[sp] = n2; ++sp;
        }
        leaf_start[L] = n;
        s_leaves = L;
    }
    __syncthreads();
    const int L = s_leaves;
    for (int k = threadIdx.x; k < L; k += NPS_THREADS)
        leaf_val[k] = np_leaf_sum(a + leaf_start[k], (int)(leaf_start[k + 1] - leaf_start[k]));
    __syncthreads();
    if (threadIdx.x == 0) {
        // the same recursion again, consuming the leaf sums in order: sum(left) + sum(right)
        PwFrame stk[64];
        int sp = 0, next = 0;
        float ret = 0.f;
        stk[sp++] = {n, 0, 0.f};
        while (sp) {
            PwFrame &f = stk[sp - 1];
            if (f.st == 0) {
                if (f.n <= PW_BLOCK) { ret = leaf_val[next++]; --sp; continue; }
                f.st = 1;
                int64_t n2 = f.n / 2;
                n2 -= n2 % 8;
                stk[sp++] = {n2, 0, 0.f};
            } else if (f.st == 1) {
                f.l = ret;
                f.st = 2;
                int64_t n2 = f.n / 2;
                n2 -= n2 % 8;
                const int64_t rn = f.n - n2;
                stk[sp++] = {rn, 0, 0.f};
            } else {
                ret = __fadd_rn(f.l, ret);
                --sp;
            }
        }
        out[0] = ret;
    }
}

__global__ void __launch_bounds__(PT_THREADS)
np_norm_kernel(const float *__restrict__ pts, int64_t n, int C, const float *__restrict__ stats, float *__restrict__ norm)
{
    const float mean[3] = {stats[0], stats[1], stats[2]};
    for (int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT_THREADS)
        norm[i] = point_norm(pts + i * C, mean);
}

// np.std's centred squares: x = norm - mean(norm); x * x   (mean(norm) = pairwise sum / n, float32)
__global__ void __launch_bounds__(PT_THREADS)
np_sqdev_kernel(const float *__restrict__ norm, int64_t n, float *__restrict__ stats, float *__restrict__ dev)
{
    const float m = __fdiv_rn(stats[3], (float)n);
    if (blockIdx.x == 0 && threadIdx.x == 0) stats[5] = m;
    for (int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT_THREADS) {
        const float x = __fsub_rn(norm[i], m);
        dev[i] = __fmul_rn(x, x);
    }
}

struct Range6 { float lo[3], hi[3]; };

// keep = norm < mean(norm) + 5 std(norm)  (:115)  and  lo <= xyz < hi  (model/PointPillars.py:251-252)
__global__ void __launch_bounds__(PT_THREADS)
preprocess_flag_kernel(const float *__restrict__ pts, int64_t n, int C, const double *__restrict__ sums, int outlier,
                       const Range6 rg, uint8_t *__restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x;
    if (i >= n) return;
    const float *p = pts + i * C;
    bool keep = true;
    if (outlier) {
        const float mean[3] = {(float)(sums[0] / (double)n), (float)(sums[1] / (double)n), (float)(sums[2] / (double)n)};
        const double m = sums[3] / (double)n;
        double var = sums[4] / (double)n - m * m;
        var = var > 0 ? var : 0;
        const float thr = (float)(m + 5.0 * sqrt(var));
        keep = point_norm(p, mean) < thr;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) keep = keep && p[k] >= rg.lo[k] && p[k] < rg.hi[k];
    flags[i] = keep ? 1 : 0;
}

// exact mode: threshold = np.mean(norm) + 5 * np.std(norm) in float32 (numpy 2 scalar rules), norm from the array
__global__ void __launch_bounds__(PT_THREADS)
preprocess_flag_exact_kernel(const float *__restrict__ pts, int64_t n, int C, const float *__restrict__ stats,
                             const float *__restrict__ norm, const Range6 rg, uint8_t *__restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x;
    if (i >= n) return;
    const float *p = pts + i * C;
    const float std_ = __fsqrt_rn(__fdiv_rn(stats[4], (float)n));
    const float thr = __fadd_rn(stats[5], __fmul_rn(5.f, std_));
    bool keep = norm[i] < thr;
#pragma unroll
    for (int k = 0; k < 3; ++k) keep = keep && p[k] >= rg.lo[k] && p[k] < rg.hi[k];
    flags[i] = keep ? 1 : 0;
}

struct FeatSel { int idx[16]; int n; };

__global__ void __launch_bounds__(PT_THREADS)
preprocess_gather_kernel(const float *__restrict__ pts, int64_t n, int C, const uint8_t *__restrict__ flags,
                         const int32_t *__restrict__ block_base, const FeatSel fs, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x;
    const int slot = compact_slot(flags, n, block_base, i);
    if (slot < 0) return;
    for (int k = 0; k < fs.n; ++k) out[(int64_t)slot * fs.n + k] = pts[i * C + fs.idx[k]];
}

// ---- CustomVoxelizer centroids: np.sum(vox, axis=1) / vp  ++  vp  (model/utils.py:34-43) ------------------------
__global__ void __launch_bounds__(PT_THREADS)
centroid_kernel(const float *__restrict__ voxels, const int32_t *__restrict__ num, int64_t M, int P, int C,
                float *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x;      // one thread per (pillar, feature)
    if (t >= M * (C + 1)) return;
    const int64_t m = t / (C + 1);
    const int c = (int)(t - m * (C + 1));
    const float nf = (float)num[m];
    if (c == C) { out[t] = nf; return; }
    float s = 0.f;
    for (int p = 0; p < P; ++p) s = __fadd_rn(s, voxels[(m * P + p) * C + c]);     // sequential over the points
    out[t] = __fdiv_rn(s, nf);
}

__global__ void __launch_bounds__(PT_THREADS)
minmax_kernel(const float *__restrict__ pts, int64_t n, int C, uint32_t *mm /* [6] ordered bits: min xyz, max xyz */)
{
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT_THREADS)
        for (int k = 0; k < 3; ++k) { const float v = pts[i * C + k]; lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
    for (int k = 0; k < 3; ++k) {
        const unsigned l = __reduce_min_sync(0xFFFFFFFFu, ordered_bits(lo[k])), h = __reduce_max_sync(0xFFFFFFFFu, ordered_bits(hi[k]));
        if ((threadIdx.x & 31) == 0) { atomicMin(mm + k, l); atomicMax(mm + 3 + k, h); }
    }
}
__global__ void minmax_finish_kernel(const uint32_t *mm, float *out6)
{
    const int k = threadIdx.x;
    if (k >= 6) return;
    const uint32_t u = mm[k];
    out6[k] = __uint_as_float(u ^ ((u & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu));
}

// ---- dense -> sparse: cells with any non-zero channel, row-major order (torch.where), + their feature rows ------
__global__ void __launch_bounds__(PT_THREADS)
dense_flag_kernel(const float *__restrict__ x, int C, int64_t HW, int64_t cells /* B * HW */, uint8_t *__restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x;      // b * HW + y * W + x
    if (i >= cells) return;
    const int64_t b = i / HW, pos = i - b * HW;
    const float *p = x + b * C * HW + pos;
    bool any = false;
    for (int c = 0; c < C; ++c) any = any || (p[(int64_t)c * HW] != 0.f);
    flags[i] = any ? 1 : 0;
}

__global__ void __launch_bounds__(PT_THREADS)
dense_gather_kernel(const float *__restrict__ x, int C, int W, int64_t HW, int64_t cells, const uint8_t *__restrict__ flags,
                    const int32_t *__restrict__ block_base, int32_t *__restrict__ coords, float *__restrict__ values)
{
    const int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x;
    const int slot = compact_slot(flags, cells, block_base, i);
    if (slot < 0) return;
    const int64_t b = i / HW, pos = i - b * HW;
    coords[(int64_t)slot * 3 + 0] = (int32_t)b;
    coords[(int64_t)slot * 3 + 1] = (int32_t)(pos / W);
    coords[(int64_t)slot * 3 + 2] = (int32_t)(pos % W);
    const float *p = x + b * C * HW + pos;
    for (int c = 0; c < C; ++c) values[(int64_t)slot * C + c] = p[(int64_t)c * HW];
}

int run_compaction_counts(const uint8_t *flags, int64_t n, int32_t *block_base, int32_t *total, cudaStream_t st)
{
    const int64_t nb = ceil_div(n, PT_THREADS);
    flag_count_kernel<<<(unsigned)nb, PT_THREADS, 0, st>>>(flags, n, block_base);
    if (int rc = check_launch("flag_count_kernel")) return rc;
    block_scan_kernel<<<1, 1024, 0, st>>>(block_base, nb, total);
    return check_launch("block_scan_kernel");
}

unsigned stream_grid(int64_t n) { int64_t g = ceil_div(n, PT_THREADS); return (unsigned)(g < 148 * 8 ? (g > 0 ? g : 1) : 148 * 8); }

}  // namespace
}  // namespace pp

using namespace pp;

extern "C" size_t pp_compact_workspace_bytes(int64_t n)
{
    if (n < 0) return 0;
    // flags + per-CTA counts + float64 statistics / ordered min-max words
    return align_up((size_t)(n > 0 ? n : 1)) + align_up((size_t)(ceil_div(n > 0 ? n : 1, PT_THREADS) + 1) * 4) + 256;
}

extern "C" size_t pp_preprocess_workspace_bytes(int64_t n)
{
    if (n < 0) return 0;
    const size_t n1 = (size_t)(n > 0 ? n : 1), leaves = n1 / 64 + 4;
    // compaction + (exact statistics) the norm array, the centred squares, the leaves of the pairwise tree
    return pp_compact_workspace_bytes(n) + 2 * align_up(n1 * sizeof(float)) + align_up(leaves * sizeof(int64_t)) +
           align_up(leaves * sizeof(float));
}

namespace {
struct CompactWs { uint8_t *flags; int32_t *block_base; double *stats; };
CompactWs carve_compact(void *ws, int64_t n)
{
    CompactWs w;
    char *p = (char *)ws;
    w.stats = (double *)p;                         // 256 bytes: 5 doubles, or 6 ordered words
    w.flags = (uint8_t *)(p + 256);
    w.block_base = (int32_t *)(p + 256 + align_up((size_t)(n > 0 ? n : 1)));
    return w;
}
}  // namespace

extern "C" int pp_preprocess_points(const float *points, int64_t n, int C, int outlier_check, const float *range6_host,
                                    const int32_t *features_host, int n_features, float *out, int32_t *out_count,
                                    void *workspace, size_t workspace_bytes, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    PP_REQUIRE(n >= 0 && C >= 3 && out_count, "bad arguments");
    PP_REQUIRE(range6_host && features_host && n_features > 0 && n_features <= 16, "1..16 selected features");
    for (int k = 0; k < n_features; ++k) PP_REQUIRE(features_host[k] >= 0 && features_host[k] < C, "feature index out of range");
    if (n == 0) {
        PP_CUDA_TRY(cudaMemsetAsync(out_count, 0, sizeof(int32_t), st));
        return PP_OK;
    }
    PP_REQUIRE(points && out && workspace, "null pointer");
    if (workspace_bytes < pp_compact_workspace_bytes(n)) {
        set_error("preprocess workspace too small");
        return PP_ERR_WORKSPACE;
    }
    CompactWs w = carve_compact(workspace, n);
    Range6 rg;
    for (int k = 0; k < 3; ++k) { rg.lo[k] = range6_host[k]; rg.hi[k] = range6_host[3 + k]; }
    const unsigned nb = (unsigned)ceil_div(n, PT_THREADS);
    if (outlier_check == PP_OUTLIER_EXACT) {
        // numpy's own order of operations: the kept rows are bit-exact with the reference's (T0)
        if (workspace_bytes < pp_preprocess_workspace_bytes(n)) {
            set_error("preprocess workspace too small for the exact statistics (pp_preprocess_workspace_bytes)");
            return PP_ERR_WORKSPACE;
        }
        const size_t n1 = (size_t)n, leaves = n1 / 64 + 4;
        char *q = (char *)workspace + pp_compact_workspace_bytes(n);
        float *norm = (float *)q;                       q += align_up(n1 * sizeof(float));
        float *dev = (float *)q;                        q += align_up(n1 * sizeof(float));
        int64_t *leaf_start = (int64_t *)q;             q += align_up(leaves * sizeof(int64_t));
        float *leaf_val = (float *)q;
        float *fs = (float *)w.stats;                   // [0..2] mean xyz, [3] sum(norm), [4] sum(dev), [5] mean(norm)
        np_colmean3_kernel<<<1, NPS_THREADS, 0, st>>>(points, n, C, fs);
        if (int rc = check_launch("np_colmean3_kernel")) return rc;
        np_norm_kernel<<<stream_grid(n), PT_THREADS, 0, st>>>(points, n, C, fs, norm);
        if (int rc = check_launch("np_norm_kernel")) return rc;
        np_pairwise_sum_kernel<<<1, NPS_THREADS, 0, st>>>(norm, n, leaf_start, leaf_val, fs + 3);
        if (int rc = check_launch("np_pairwise_sum_kernel")) return rc;
        np_sqdev_kernel<<<stream_grid(n), PT_THREADS, 0, st>>>(norm, n, fs, dev);
        if (int rc = check_launch("np_sqdev_kernel")) return rc;
        np_pairwise_sum_kernel<<<1, NPS_THREADS, 0, st>>>(dev, n, leaf_start, leaf_val, fs + 4);
        if (int rc = check_launch("np_pairwise_sum_kernel")) return rc;
        preprocess_flag_exact_kernel<<<nb, PT_THREADS, 0, st>>>(points, n, C, fs, norm, rg, w.flags);
        if (int rc = check_launch("preprocess_flag_exact_kernel")) return rc;
    } else {
        if (outlier_check) {
            // fast statistics: float64 accumulation, fully parallel (kept rows may differ from numpy's by a row or two
            // per 1e5 points: those within rounding of the 5-sigma threshold)
            PP_CUDA_TRY(cudaMemsetAsync(w.stats, 0, 5 * sizeof(double), st));
            sum_xyz_kernel<<<stream_grid(n), PT_THREADS, 0, st>>>(points, n, C, w.stats);
            if (int rc = check_launch("sum_xyz_kernel")) return rc;
            norm_stats_kernel<<<stream_grid(n), PT_THREADS, 0, st>>>(points, n, C, w.stats);
            if (int rc = check_launch("norm_stats_kernel")) return rc;
        }
        preprocess_flag_kernel<<<nb, PT_THREADS, 0, st>>>(points, n, C, w.stats, outlier_check, rg, w.flags);
        if (int rc = check_launch("preprocess_flag_kernel")) return rc;
    }
    if (int rc = run_compaction_counts(w.flags, n, w.block_base, out_count, st)) return rc;
    FeatSel fs;
    fs.n = n_features;
    for (int k = 0; k < n_features; ++k) fs.idx[k] = features_host[k];
    preprocess_gather_kernel<<<nb, PT_THREADS, 0, st>>>(points, n, C, w.flags, w.block_base, fs, out);
    return check_launch("preprocess_gather_kernel");
}

extern "C" int pp_points_minmax(const float *points, int64_t n, int C, float *out6, void *workspace, size_t workspace_bytes,
                                pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    PP_REQUIRE(n > 0 && C >= 3 && points && out6 && workspace && workspace_bytes >= 256, "bad arguments");
    uint32_t *mm = (uint32_t *)workspace;
    PP_CUDA_TRY(cudaMemsetAsync(mm, 0xFF, 3 * sizeof(uint32_t), st));
    PP_CUDA_TRY(cudaMemsetAsync(mm + 3, 0, 3 * sizeof(uint32_t), st));
    minmax_kernel<<<stream_grid(n), PT_THREADS, 0, st>>>(points, n, C, mm);
    if (int rc = check_launch("minmax_kernel")) return rc;
    minmax_finish_kernel<<<1, 32, 0, st>>>(mm, out6);
    return check_launch("minmax_finish_kernel");
}

extern "C" int pp_voxel_centroids(const float *voxels, const int32_t *num_points, int64_t M, int P, int C, float *out,
                                  pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(M >= 0 && P > 0 && C > 0, "bad shape");
    if (M == 0) return PP_OK;
    PP_REQUIRE(voxels && num_points && out, "null pointer");
    centroid_kernel<<<(unsigned)ceil_div(M * (C + 1), PT_THREADS), PT_THREADS, 0, (cudaStream_t)stream>>>(voxels, num_points, M, P, C, out);
    return check_launch("centroid_kernel");
}

extern "C" int pp_dense_to_sparse(const float *x, int B, int C, int H, int W, int32_t *coords, float *values, int32_t *nnz,
                                  void *workspace, size_t workspace_bytes, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    PP_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && x && coords && values && nnz && workspace, "bad arguments");
    const int64_t HW = (int64_t)H * W, cells = (int64_t)B * HW;
    if (workspace_bytes < pp_compact_workspace_bytes(cells)) {
        set_error("dense_to_sparse workspace too small");
        return PP_ERR_WORKSPACE;
    }
    CompactWs w = carve_compact(workspace, cells);
    const unsigned nb = (unsigned)ceil_div(cells, PT_THREADS);
    dense_flag_kernel<<<nb, PT_THREADS, 0, st>>>(x, C, HW, cells, w.flags);
    if (int rc = check_launch("dense_flag_kernel")) return rc;
    if (int rc = run_compaction_counts(w.flags, cells, w.block_base, nnz, st)) return rc;
    dense_gather_kernel<<<nb, PT_THREADS, 0, st>>>(x, C, W, HW, cells, w.flags, w.block_base, coords, values);
    return check_launch("dense_gather_kernel");
}
