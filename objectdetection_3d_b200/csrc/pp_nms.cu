// One class of multiclass_nms (model/utils.py:376-424; nms_dim == 2 and, through the pair-test template, nms_dim == 3
// and the rotated-BEV extension):
//   prepare : AABB of the rotated box (ops/ops_torch.py:13-114), pair-test geometry, score filter (strict >, :381), sort
//             key, the sort's digit histograms
//   sort    : descending score, stable (ties: lower index first); the last pass moves the rectangles and the geometry
//             with their keys  [pp_sort.cu]
//   level 1 : greedy NMS of the NMS_LEVEL1 best candidates = mask + sweep
//   filter  : drops every lower candidate the level-1 keep set suppresses, compacts the survivors in rank order
//   level 2 : greedy NMS of the survivors = mask + sweep
//   mask    : tiles of the upper triangle, bit j of mask[i][cb] = iou(box_j, box_i) > thr (strict, :413); the clipped
//             pair tests are queued per tile and evaluated with full warps
//   sweep   : one CTA; a warp decides each 64-box block's keep set with warp OR-reductions and ORs the kept rows into
//             the next 31 words from a shared-memory band, the other warps OR them into the words beyond
// The rectangle IoU is evaluated exactly like bbox_iou2D (pp_boxes.cuh::rect_iou), so the keep set is identical to the
// reference's greedy loop on the same rectangles.
#include <cuda_fp16.h>

#include "pp_boxes.cuh"
#include "pp_common.cuh"
#include "pp_sort.cuh"

namespace pp {
namespace {

typedef unsigned long long u64;
constexpr int NMS_THREADS = 256;
// device scalars of one pp_nms call
enum { SC_N = 0, SC_N1 = 1, SC_K1 = 2, SC_N2 = 3, SC_TICKET = 4, SC_COUNT = 8 };
constexpr int NMS_LEVEL1 = 2048;   // boxes resolved first; the rest is filtered against their keep set before its own NMS
constexpr int SWEEP_THREADS = 1024;

// Per-box data of the pair tests that need more than the bounding rectangle: PP_NMS_ROT_BEV uses a0 and a1.xy (RRect),
// PP_NMS_BOX3D a0..a2 (Box3).
struct Aux {
    float4 *a0, *a1, *a2;
};
template <int MODE> struct PairGeom;
template <> struct PairGeom<PP_NMS_AABB2D> {
    struct T {};
    static __device__ __forceinline__ T load(const float4 *, const float4 *, const float4 *, int) { return T(); }
    static __device__ __forceinline__ float iou(const T &, const T &) { return 0.f; }
    static __device__ __forceinline__ bool exceeds(const T &, const T &, float) { return false; }
};
template <> struct PairGeom<PP_NMS_ROT_BEV> {
    typedef RRect T;
    static __device__ __forceinline__ T load(const float4 *a0, const float4 *a1, const float4 *, int i)
    {
        const float4 h = a1[i];
        return rrect_pack(a0[i], make_float2(h.x, h.y));
    }
    static __device__ __forceinline__ bool maybe(const T &, const T &, float) { return true; }
    static __device__ __forceinline__ bool exact(const T &a, const T &b, float thr) { return rrect_iou(a, b) > thr; }
    static __device__ __forceinline__ bool exceeds(const T &a, const T &b, float thr) { return rrect_iou(a, b) > thr; }
};
template <> struct PairGeom<PP_NMS_BOX3D> {
    typedef Box3 T;
    static __device__ __forceinline__ T load(const float4 *a0, const float4 *a1, const float4 *a2, int i)
    {
        return box3_load(a0[i], a1[i], a2[i]);
    }
    // Can iou exceed thr at all?  Most pairs are decided here, in fp32: an axis of either box separates them (the
    // intersection is empty and 0 > thr is false), or the intersection, which lies inside box b and inside the
    // b-aligned bounding box of a (and vice versa), is too small to reach the threshold (iou is increasing in the
    // volume; the bound is inflated by 1e-3 relative, far above its fp32 error).
    static __device__ __forceinline__ bool maybe(const T &a, const T &b, float thr)
    {
        if (thr < 0.f) return true;
        float ub1, ub2;
        if (!box3_proj_bound(a, b, ub1) || !box3_proj_bound(b, a, ub2)) return false;
        const float ub = fminf(ub1, ub2) * 1.001f;
        const float va = fabsf(det3(a.e[0], a.e[1], a.e[2])), vb = fabsf(det3(b.e[0], b.e[1], b.e[2]));
        // iou(ub) = ub / (va + vb - ub) <= thr  <=>  ub * (1 + thr) <= thr * (va + vb)
        return !(ub * (1.f + thr) <= thr * (va + vb));
    }
    static __device__ __forceinline__ bool exact(const T &a, const T &b, float thr) { return box3_iou(a, b, nullptr) > thr; }
    static __device__ __forceinline__ bool exceeds(const T &a, const T &b, float thr) { return maybe(a, b, thr) && exact(a, b, thr); }
};

constexpr int PREP_THREADS = 1024;          // large CTAs: each adds its keys' digit histograms to the sort's with few atomics
__global__ void __launch_bounds__(PREP_THREADS)
nms_prepare_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, int64_t stride, int64_t N,
                   float score_thr, float4 *__restrict__ rect, uint32_t *__restrict__ keys, int32_t *__restrict__ n_cand,
                   int mode, const Aux aux, uint32_t *__restrict__ sort_hist_out)
{
    __shared__ uint32_t s_hist[1024];
    int64_t i = (int64_t)blockIdx.x * PREP_THREADS + threadIdx.x;
    bool cand = false;
    uint32_t key = 0xFFFFFFFFu;
    if (i < N) {
        float b[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) b[k] = boxes[i * 9 + k];
        if (mode == PP_NMS_ROT_BEV) {
            // rotated footprint (x, y, dx, dy, rz); its bounding rectangle drives the prefilter
            const RRect r = rrect_from_box9(b);
            rect[i] = rrect_aabb(r);
            aux.a0[i] = rrect_lo(r);
            aux.a1[i] = make_float4(r.c, r.s, 0.f, 0.f);
        } else {
            float c[8][3];
            box_corners(b, c);
            rect[i] = corners_to_rect(c);
            if (mode == PP_NMS_BOX3D) {
                float4 q0, q1, q2;
                box3_store(box3_from_corner_array(c), q0, q1, q2);
                aux.a0[i] = q0; aux.a1[i] = q1; aux.a2[i] = q2;
            }
        }
        float s = scores[i * stride];
        cand = s > score_thr;
        key = cand ? ~ordered_bits(s) : 0xFFFFFFFFu;
        keys[i] = key;
    }
    unsigned m = __ballot_sync(0xFFFFFFFFu, cand);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_cand, __popc(m));
    if (sort_hist_out) sort_hist_add(s_hist, sort_hist_out, key, i < N);      // (the sort orders all N keys)
}

__global__ void __launch_bounds__(NMS_THREADS)
nms_gather_kernel(const float4 *__restrict__ rect, const uint32_t *__restrict__ order, int32_t *__restrict__ sc,
                  float4 *__restrict__ srect, int level1, const Aux aux, const Aux saux)
{
    int64_t r = (int64_t)blockIdx.x * NMS_THREADS + threadIdx.x;
    const int n = sc[SC_N];
    if (r == 0) sc[SC_N1] = n < level1 ? n : level1;
    if (r < n) {
        const uint32_t o = order[r];
        srect[r] = rect[o];
        if (aux.a0) { saux.a0[r] = aux.a0[o]; saux.a1[r] = aux.a1[o]; }
        if (aux.a2) saux.a2[r] = aux.a2[o];
    }
}

constexpr int SW_L = 31;                    // words after the diagonal handled by the sweep's resolver warp (one per lane)
constexpr int SW_RING = 6;                  // band blocks in flight in the sweep
constexpr int SW_BAND = (SW_L + 1) * 64;    // u64 per band block (64 rows x the diagonal word and the SW_L words after it)
// Layout of one band block: the 64 diagonal words first (the resolver's lanes read rows l and l + 32 of them), then the
// rows' next SW_L words row by row (a kept row's words are read by consecutive lanes: no bank conflicts).
__device__ __forceinline__ size_t band_index(int row, int kb)
{
    const size_t blk = (size_t)(row >> 6) * SW_BAND;
    const int r = row & 63;
    return kb == 0 ? blk + r : blk + 64 + (size_t)r * SW_L + (kb - 1);
}
constexpr size_t SW_SMEM_MAX = 200 * 1024;  // dynamic shared memory the sweep may ask for

constexpr int MT_ROWS = 128;   // rows (selected boxes) per CTA, one per thread
constexpr int MT_COLS = 256;   // columns (remaining boxes) per CTA = 4 mask words

// bit j of mask[i][w] = (64w + j > i) && iou(box_{64w+j}, box_i) > thr.  Words entirely below the
// diagonal are never read by the sweep and are not written.
// Two phases per 64-column word: (1) a 4-compare interval test marks the columns whose rectangle
// intersects the row's (a superset of the hits when thr >= 0); (2) only those columns get the exact
// bbox_iou2D evaluation (rect_iou, IEEE division), so decisions equal the reference's `iou > thr`.
// (x1, y1) rounded down and (x2, y2) rounded up to half precision, clamped to the finite half range
__device__ __forceinline__ uint2 rect_to_half(const float4 r)
{
    const float L = 60000.f;
    const __half2 lo = __halves2half2(__float2half_rd(fminf(fmaxf(r.x, -L), L)), __float2half_rd(fminf(fmaxf(r.y, -L), L)));
    const __half2 hi = __halves2half2(__float2half_ru(fminf(fmaxf(r.z, -L), L)), __float2half_ru(fminf(fmaxf(r.w, -L), L)));
    uint2 o;
    o.x = *reinterpret_cast<const unsigned *>(&lo);
    o.y = *reinterpret_cast<const unsigned *>(&hi);
    return o;
}

template <bool PREFILTER, int MODE>
__global__ void __launch_bounds__(MT_ROWS)
nms_mask_kernel(const float4 *__restrict__ srect, const int32_t *__restrict__ n_cand, float thr, int nw_stride,
                u64 *__restrict__ mask, u64 *__restrict__ band, const Aux saux)
{
    typedef PairGeom<MODE> G;
    const int n = *n_cand;
    const int row0 = blockIdx.y * MT_ROWS, col0 = blockIdx.x * MT_COLS;
    if (row0 >= n || col0 >= n || col0 + MT_COLS <= row0) return;
    __shared__ float4 s_col[MT_COLS];
    __shared__ uint2 s_colh[MT_COLS];     // the same rectangles as conservative half2 pairs: (x1,y1) down, (x2,y2) up
    const int t = threadIdx.x;
#pragma unroll
    for (int k = 0; k < MT_COLS / MT_ROWS; ++k) {
        const int c = col0 + t + k * MT_ROWS;
        // out-of-range columns get an empty rectangle: never intersects
        const float4 q = c < n ? srect[c] : make_float4(3e38f, 3e38f, -3e38f, -3e38f);
        s_col[t + k * MT_ROWS] = q;
        s_colh[t + k * MT_ROWS] = rect_to_half(q);
    }
    __syncthreads();
    const int i = row0 + t;
    if (i >= n) return;
    const float4 a = srect[i];
    const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const uint2 ah = rect_to_half(a);
    const __half2 a_lo = *reinterpret_cast<const __half2 *>(&ah.x), a_hi = *reinterpret_cast<const __half2 *>(&ah.y);
#pragma unroll 1
    for (int wd = 0; wd < MT_COLS / 64; ++wd) {
        const int c_start = col0 + wd * 64;
        if (c_start >= n) break;
        if (c_start + 63 < i) continue;       // entirely below the diagonal: never read
        unsigned lo = 0xFFFFFFFFu, hi = 0xFFFFFFFFu;
        if (PREFILTER) {
            // Phase 1 on packed halves: both components of (q_hi - a_lo) and (a_hi - q_lo) must be >= 0.  The
            // halves are rounded outwards, so this marks a superset of the intersecting pairs with 2 HADD2 + 3
            // integer ops per pair; phase 2 decides every marked pair exactly in fp32.
            lo = hi = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint2 qh = s_colh[wd * 64 + j];
                const __half2 d1 = __hsub2(*reinterpret_cast<const __half2 *>(&qh.y), a_lo);
                const __half2 d2 = __hsub2(a_hi, *reinterpret_cast<const __half2 *>(&qh.x));
                const unsigned sg = (*reinterpret_cast<const unsigned *>(&d1) | *reinterpret_cast<const unsigned *>(&d2)) & 0x80008000u;
                if (sg == 0) lo |= 1u << j;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint2 qh = s_colh[wd * 64 + 32 + j];
                const __half2 d1 = __hsub2(*reinterpret_cast<const __half2 *>(&qh.y), a_lo);
                const __half2 d2 = __hsub2(a_hi, *reinterpret_cast<const __half2 *>(&qh.x));
                const unsigned sg = (*reinterpret_cast<const unsigned *>(&d1) | *reinterpret_cast<const unsigned *>(&d2)) & 0x80008000u;
                if (sg == 0) hi |= 1u << j;
            }
        }
        u64 cand = ((u64)hi << 32) | lo;
        if (c_start <= i) cand &= ~((2ull << (i - c_start)) - 1ull);      // keep only columns > i
        if (c_start + 64 > n) cand &= (1ull << (n - c_start)) - 1ull;     // and columns < n
        u64 bits = 0;
        while (cand) {
            const int j = __ffsll((long long)cand) - 1;
            cand &= cand - 1;
            const float4 q = s_col[wd * 64 + j];
            bool hit;
            if (PREFILTER) {
                // iou > thr  <=>  overlap > thr * union, decided without the division unless the two sides are
                // within 1e-6 relative of each other (then the reference's exact IEEE quotient is evaluated)
                float w = __fsub_rn(fminf(q.z, a.z), fmaxf(q.x, a.x));
                float h = __fsub_rn(fminf(q.w, a.w), fmaxf(q.y, a.y));
                w = w < 0.f ? 0.f : w;
                h = h < 0.f ? 0.f : h;
                const float ov = __fmul_rn(w, h);
                // (remaining, selected) argument order of model/utils.py:412 -> area1 = column box
                const float area_q = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
                const float uni = fmaxf(__fsub_rn(__fadd_rn(area_q, area_a), ov), 1e-6f);
                const float tu = __fmul_rn(thr, uni);
                hit = ov > __fmul_rn(tu, 1.000001f) && ov > 1e-30f;
                if (!hit && ov >= __fmul_rn(tu, 0.999999f)) hit = __fdiv_rn(ov, uni) > thr;
            } else {
                hit = rect_iou(q, a, 0, 1e-6f) > thr;
            }
            if (hit) bits |= 1ull << j;
        }
        const int cw = c_start >> 6, kb = cw - (i >> 6);
        mask[(size_t)i * nw_stride + cw] = bits;
        // block-major copy of the diagonal band, streamed by the sweep with one bulk copy per block
        if (kb <= SW_L) band[band_index(i, kb)] = bits;
    }
}

// The clipped pair tests (rotated BEV footprint, oriented 3-D box) cost hundreds to thousands of instructions per pair
// and only a few percent of the pairs that survive the interval prefilter need them.  Evaluated by the row's own thread
// (as the rectangle test above is) nearly every exact test would run with one active lane per warp -- and even a queue
// per 64-column word fills three lanes of a warp on average (ncu, 20k boxes).  So the CTA keeps two queues in shared
// memory for its whole 128 x 256 tile: (1) every row thread pushes the pairs whose rectangles overlap (interval
// prefilter, then the fp32 rectangle overlap the IoU definition requires); (2) all threads pop those, run the cheap
// conservative test of the mode (PairGeom::maybe) and push the pairs it cannot decide onto the second queue; (3) when
// that one is nearly full, and at the end of the tile, all threads pop it and run the exact test with full warps,
// setting the mask bits with shared-memory atomics.
constexpr int MC_Q1 = 12288;       // entries of the first queue; a 64-column word adds at most 128 x 64
constexpr int MC_Q2 = 2048;        // entries of the second
template <int MODE> struct ClipSmem {
    float4 col[MT_COLS];
    uint2 colh[MT_COLS];
    float4 a0[MT_COLS], a1[MT_COLS], a2[MODE == PP_NMS_BOX3D ? MT_COLS : 1];       // geometry of the column boxes
    float4 r0[MT_ROWS], r1[MT_ROWS], r2[MODE == PP_NMS_BOX3D ? MT_ROWS : 1];       // and of the row boxes
    u64 bits[MT_COLS / 64][MT_ROWS];
    uint16_t q1[MC_Q1], q2[MC_Q2];
    int n1, n2;
};

template <bool PREFILTER, int MODE>
__global__ void __launch_bounds__(MT_ROWS, 3)
nms_mask_clip_kernel(const float4 *__restrict__ srect, const int32_t *__restrict__ n_cand, float thr, int nw_stride,
                     u64 *__restrict__ mask, u64 *__restrict__ band, const Aux saux)
{
    typedef PairGeom<MODE> G;
    const int n = *n_cand;
    const int row0 = blockIdx.y * MT_ROWS, col0 = blockIdx.x * MT_COLS;
    if (row0 >= n || col0 >= n || col0 + MT_COLS <= row0) return;
    extern __shared__ __align__(16) unsigned char clip_smem[];
    ClipSmem<MODE> &S = *reinterpret_cast<ClipSmem<MODE> *>(clip_smem);
    const int t = threadIdx.x;
#pragma unroll
    for (int k = 0; k < MT_COLS / MT_ROWS; ++k) {
        const int c = col0 + t + k * MT_ROWS;
        const float4 q = c < n ? srect[c] : make_float4(3e38f, 3e38f, -3e38f, -3e38f);
        S.col[t + k * MT_ROWS] = q;
        S.colh[t + k * MT_ROWS] = rect_to_half(q);
        if (c < n) {
            S.a0[t + k * MT_ROWS] = saux.a0[c];
            S.a1[t + k * MT_ROWS] = saux.a1[c];
            if (MODE == PP_NMS_BOX3D) S.a2[t + k * MT_ROWS] = saux.a2[c];
        }
    }
    const int i = row0 + t;
    const bool row_ok = i < n;
    float4 a = make_float4(3e38f, 3e38f, -3e38f, -3e38f);
    if (row_ok) {
        a = srect[i];
        S.r0[t] = saux.a0[i];
        S.r1[t] = saux.a1[i];
        if (MODE == PP_NMS_BOX3D) S.r2[t] = saux.a2[i];
    }
    if (t == 0) { S.n1 = 0; S.n2 = 0; }
    __syncthreads();
    const uint2 ah = rect_to_half(a);
    const __half2 a_lo = *reinterpret_cast<const __half2 *>(&ah.x), a_hi = *reinterpret_cast<const __half2 *>(&ah.y);
    const bool zero_hits = 0.f > thr;
    int wd = 0;
    bool filling = true;
    // one loop, one copy of each test in the code: fill the first queue word by word; drain it when the next word
    // might not fit, and after the last word
#pragma unroll 1
    while (true) {
        if (filling) {
            const int c_start = col0 + wd * 64;
            u64 bits = 0;
            if (row_ok && c_start < n && c_start + 63 >= i) {         // words entirely below the diagonal are never read
                unsigned lo = 0xFFFFFFFFu, hi = 0xFFFFFFFFu;
                if (PREFILTER) {
                    lo = hi = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint2 qh = S.colh[wd * 64 + j];
                        const __half2 d1 = __hsub2(*reinterpret_cast<const __half2 *>(&qh.y), a_lo);
                        const __half2 d2 = __hsub2(a_hi, *reinterpret_cast<const __half2 *>(&qh.x));
                        const unsigned sg = (*reinterpret_cast<const unsigned *>(&d1) | *reinterpret_cast<const unsigned *>(&d2)) & 0x80008000u;
                        if (sg == 0) lo |= 1u << j;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint2 qh = S.colh[wd * 64 + 32 + j];
                        const __half2 d1 = __hsub2(*reinterpret_cast<const __half2 *>(&qh.y), a_lo);
                        const __half2 d2 = __hsub2(a_hi, *reinterpret_cast<const __half2 *>(&qh.x));
                        const unsigned sg = (*reinterpret_cast<const unsigned *>(&d1) | *reinterpret_cast<const unsigned *>(&d2)) & 0x80008000u;
                        if (sg == 0) hi |= 1u << j;
                    }
                }
                u64 cand = ((u64)hi << 32) | lo;
                if (c_start <= i) cand &= ~((2ull << (i - c_start)) - 1ull);      // keep only columns > i
                if (c_start + 64 > n) cand &= (1ull << (n - c_start)) - 1ull;     // and columns < n
                while (cand) {
                    const int j = __ffsll((long long)cand) - 1;
                    cand &= cand - 1;
                    const float4 q = S.col[wd * 64 + j];
                    // same definition as pp_iou_rotated_bev / pp_box3d_overlap: 0 unless the fp32 xy bounding rectangles
                    // overlap, else the clipped-polygon (polyhedron) IoU, symmetric in its arguments
                    if (fminf(a.z, q.z) > fmaxf(a.x, q.x) && fminf(a.w, q.w) > fmaxf(a.y, q.y))
                        S.q1[atomicAdd(&S.n1, 1)] = (uint16_t)((t << 8) | (wd * 64 + j));
                    else if (zero_hits)
                        bits |= 1ull << j;
                }
            }
            S.bits[wd][t] = bits;
            ++wd;
            __syncthreads();
            const int n1_now = S.n1;
            __syncthreads();                                           // (the next word's pushes change n1)
            const bool last = wd == MT_COLS / 64 || col0 + wd * 64 >= n;
            if (!last && n1_now <= MC_Q1 - MT_ROWS * 64) continue;     // (uniform over the CTA)
            filling = false;
        }
        // drain: first queue -> cheap test -> second queue -> exact test
        const int n1 = S.n1;
        int base = 0, n2 = 0;                                          // n2: entries of the second queue (uniform)
#pragma unroll 1
        while (true) {
#pragma unroll 1
            for (; base < n1 && n2 <= MC_Q2 - MT_ROWS; base += MT_ROWS) {
                bool pass = false;
                if (base + t < n1) {
                    const int pr = S.q1[base + t], r = pr >> 8, c = pr & 255;
                    pass = G::maybe(G::load(S.a0, S.a1, S.a2, c), G::load(S.r0, S.r1, S.r2, r), thr);
                    if (pass) S.q2[atomicAdd(&S.n2, 1)] = (uint16_t)pr;
                }
                n2 += __syncthreads_count(pass);
            }
            for (int k = t; k < n2; k += MT_ROWS) {
                const int pr = S.q2[k], r = pr >> 8, c = pr & 255;
                if (G::exact(G::load(S.a0, S.a1, S.a2, c), G::load(S.r0, S.r1, S.r2, r), thr))
                    atomicOr(&S.bits[c >> 6][r], 1ull << (c & 63));
            }
            __syncthreads();
            if (t == 0) S.n2 = 0;
            n2 = 0;
            __syncthreads();
            if (base >= n1) break;
        }
        if (t == 0) S.n1 = 0;
        __syncthreads();
        if (wd == MT_COLS / 64 || col0 + wd * 64 >= n) break;
        filling = true;
    }
    if (row_ok) {
#pragma unroll 1
        for (int w = 0; w < wd; ++w) {
            const int c_start = col0 + w * 64;
            if (c_start + 63 < i) continue;
            const u64 out = S.bits[w][t];
            const int cw = c_start >> 6, kb = cw - (i >> 6);
            mask[(size_t)i * nw_stride + cw] = out;
            if (kb <= SW_L) band[band_index(i, kb)] = out;
        }
    }
}

// ---- sweep -------------------------------------------------------------------------------------
// One CTA.  Warp 31 ("resolver") walks the 64-box blocks in rank order.  Per block it takes the final removed word,
// decides the block's keep set with the whole warp -- lane l holds rows l and l + 32 of the 64 x 64 diagonal tile; per
// round the candidates that no other candidate suppresses are kept and what they suppress leaves the candidate set
// (two OR-reductions over the warp per round, as many rounds as the longest suppression chain inside the block, not
// one step per kept box) -- and ORs the kept rows into the next SW_L words itself, again as warp OR-reductions, from a
// band of the mask that it streams into a shared-memory ring with bulk copies SW_RING - 1 blocks ahead.  The other 31
// warps ("owners") take the resolved blocks round-robin and OR their kept rows into all words beyond the band, lanes
// over consecutive words (coalesced 256-byte reads of the mask, 8 rows in flight); the resolver needs block b - SW_L - 1
// finished when it reaches block b, so the owners' L2 latency hides behind SW_L resolver steps.  Synchronisation is
// through shared-memory flags only.

// acquire/release fence at CTA scope (MEMBAR.ALL.CTA): cheaper than __threadfence_block()'s fence.sc
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ u64 warp_or64(u64 v)
{
    const uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)v), hi = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)(v >> 32));
    return ((u64)hi << 32) | lo;
}
constexpr int SW_OWN_ROWS = 8;              // mask rows an owner lane has in flight (64 registers per thread at 1024 threads)

__global__ void __launch_bounds__(SWEEP_THREADS)
nms_sweep_kernel(const u64 *__restrict__ mask, const u64 *__restrict__ band, int nw_stride,
                 const int32_t *__restrict__ n_cand, const uint32_t *__restrict__ order, int64_t *__restrict__ keep,
                 const int32_t *__restrict__ keep_base, int32_t *__restrict__ keep_count,
                 int32_t *__restrict__ kept_rank, int32_t *__restrict__ kept_n, int nw_cap)
{
    extern __shared__ __align__(16) u64 sw_smem[];
    u64 *ring = sw_smem;                                   // [SW_RING][SW_L+1][64]
    u64 *removed = ring + SW_RING * SW_BAND;               // [nw_cap] owners' contributions (atomics)
    u64 *kept_arr = removed + nw_cap;                      // [nw_cap]
    u64 *removed_r = kept_arr + nw_cap;                    // [nw_cap] resolver-side contributions (no atomics)
    volatile int *done = (volatile int *)(removed_r + nw_cap);   // [nw_cap] block b's rows are in every word beyond its band
    __shared__ volatile int s_resolved;
    __shared__ __align__(8) u64 s_mbar[SW_RING];
    const int n = *n_cand;
    const int nw = (n + 63) >> 6;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef PP_TIMING
    long long t_start = clock64(), t_loop0 = 0, t_loop1 = 0, t_wait = 0, t_chain = 0;
    int rounds = 0;
#endif
    for (int w = tid; w < nw; w += SWEEP_THREADS) { removed[w] = 0; removed_r[w] = 0; done[w] = 0; }
    if (tid == 0) {
        s_resolved = 0;
        for (int r = 0; r < SW_RING; ++r)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_mbar[r])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == SWEEP_THREADS / 32 - 1) {
        // ------------------------------------------------------------------ resolver
        // band of block b (16 KB, contiguous) -> ring slot b % SW_RING with one TMA bulk copy, SW_RING-1 ahead
        auto issue_band = [&](int b) {
            if (b < nw && lane == 0) {
                const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_mbar[b % SW_RING]);
                const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + (b % SW_RING) * SW_BAND);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(SW_BAND * 8) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(band + (size_t)b * SW_BAND), "r"(SW_BAND * 8), "r"(bar) : "memory");
            }
        };
        for (int b = 0; b < SW_RING - 1; ++b) issue_band(b);
#ifdef PP_TIMING
        t_loop0 = clock64();
#endif
        for (int b = 0; b < nw; ++b) {
#ifdef PP_TIMING
            const long long tw0 = clock64();
#endif
            issue_band(b + SW_RING - 1);          // its slot was consumed in step b - 1
            {
                const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_mbar[b % SW_RING]);
                const unsigned parity = (b / SW_RING) & 1;
                unsigned ok;
                do {
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                } while (!ok);
            }
            const u64 *slot = ring + (b % SW_RING) * SW_BAND;
            if (b > SW_L) {
                while (done[b - SW_L - 1] == 0) { }      // pure spin: __nanosleep granularity (~1 us) would dominate the step
                fence_cta();
            }
#ifdef PP_TIMING
            const long long tw1 = clock64();
            t_wait += tw1 - tw0;
#endif
            const int valid = min(64, n - b * 64);
            u64 cand = ~(*(volatile u64 *)(removed + b) | *(volatile u64 *)(removed_r + b));
            if (valid < 64) cand &= (1ull << valid) - 1;
            // rows lane and lane + 32 of the diagonal tile (bit j of row i is set only for j > i); rows that are not
            // boxes hold whatever the workspace held, and are never candidates
            const u64 d0 = slot[lane], d1 = slot[32 + lane];
            u64 kept = 0;
            while (cand) {
                const u64 mine = (((cand >> lane) & 1ull) ? d0 : 0ull) | (((cand >> (32 + lane)) & 1ull) ? d1 : 0ull);
                const u64 nk = cand & ~warp_or64(mine);           // candidates no candidate suppresses: kept
                kept |= nk;
                const u64 theirs = (((nk >> lane) & 1ull) ? d0 : 0ull) | (((nk >> (32 + lane)) & 1ull) ? d1 : 0ull);
                cand &= ~(nk | warp_or64(theirs));
#ifdef PP_TIMING
                ++rounds;
#endif
            }
#ifdef PP_TIMING
            t_chain += clock64() - tw1;
#endif
            if (lane == 0) {
                kept_arr[b] = kept;
                if (nw > SW_L + 1) {           // (no owners otherwise: the band holds every word)
                    fence_cta();
                    s_resolved = b + 1;
                }
            }
            // the kept rows' next SW_L words: lane k ORs word b + k of every kept row.  All 64 rows are read (consecutive
            // lanes read consecutive words of a row: no bank conflicts) and masked by their kept bit: 64 independent
            // loads the compiler batches, instead of one dependent find-first-set + load per kept box.
            if (kept) {
                const u64 *rows = slot + 64 + (lane > 0 ? lane - 1 : 0);
                const uint32_t klo = (uint32_t)kept, khi = (uint32_t)(kept >> 32);
                uint32_t vlo = 0, vhi = 0;
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const u64 t = rows[i * SW_L];
                    if ((i < 32 ? klo >> i : khi >> (i - 32)) & 1u) { vlo |= (uint32_t)t; vhi |= (uint32_t)(t >> 32); }
                }
                const u64 v = ((u64)vhi << 32) | vlo;
                if (v && lane > 0 && b + lane < nw) *(volatile u64 *)(removed_r + b + lane) |= v;   // one lane per word per step
            }
            __syncwarp();      // orders the lanes' shared-memory traffic; the slot may be refilled next step
        }
#ifdef PP_TIMING
        t_loop1 = clock64();
        if (lane == 0)
            printf("sweep n=%d blocks=%d: init %lld | loop %lld cycles = %lld per block (wait %lld, rounds %lld, %d rounds)\n", n, nw,
                   t_loop0 - t_start, t_loop1 - t_loop0, (t_loop1 - t_loop0) / (nw > 0 ? nw : 1), t_wait / (nw > 0 ? nw : 1),
                   t_chain / (nw > 0 ? nw : 1), rounds);
#endif
    } else {
        // ------------------------------------------------------------------ owners
        const int n_owner = SWEEP_THREADS / 32 - 1;
        for (int b = warp; b + SW_L + 1 < nw; b += n_owner) {
            while (s_resolved <= b) __nanosleep(100);
            fence_cta();
            u64 k = *(volatile u64 *)(kept_arr + b);
            const int w0 = b + SW_L + 1;
            // rows of this block, SW_OWN_ROWS at a time; per group every lane ORs its words and issues one atomic each
            while (k) {
                int row[SW_OWN_ROWS];
#pragma unroll
                for (int j = 0; j < SW_OWN_ROWS; ++j) {
                    row[j] = -1;
                    if (k) {
                        const int i = __ffsll((long long)k) - 1;
                        k &= k - 1;
                        row[j] = b * 64 + i;
                    }
                }
                for (int w = w0 + lane; w < nw; w += 32) {
                    u64 v[SW_OWN_ROWS];
#pragma unroll
                    for (int j = 0; j < SW_OWN_ROWS; ++j) v[j] = row[j] >= 0 ? mask[(size_t)row[j] * nw_stride + w] : 0ull;
                    u64 acc = 0;
#pragma unroll
                    for (int j = 0; j < SW_OWN_ROWS; ++j) acc |= v[j];
                    if (acc) atomicOr(removed + w, acc);
                }
            }
            __syncwarp();
            if (lane == 0) {
                fence_cta();
                done[b] = 1;
            }
        }
    }
    // ---- expand the kept bitmaps to original indices, in rank order (all threads) -------------------
    // exclusive prefix of the words' kept counts (block scan per 1024 words), then one thread per box
    __syncthreads();
    __shared__ int s_warp_sum[SWEEP_THREADS / 32];
    __shared__ int s_base;
    int *prefix = reinterpret_cast<int *>(removed);    // (the removed words are no longer needed)
    const int base0 = keep_base ? *keep_base : 0;      // keep entries of the previous level come first
    if (tid == 0) s_base = base0;
    __syncthreads();
    for (int w0 = 0; w0 < nw; w0 += SWEEP_THREADS) {
        const int w = w0 + tid;
        const int cnt = w < nw ? __popcll(kept_arr[w]) : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp_sum[warp] = incl;
        __syncthreads();
        int pos = s_base + incl - cnt;
        for (int q = 0; q < warp; ++q) pos += s_warp_sum[q];
        if (w < nw) prefix[w] = pos;
        __syncthreads();
        if (tid == SWEEP_THREADS - 1) s_base = pos + cnt;      // last thread holds the running total
        __syncthreads();
    }
    for (int i = tid; i < n; i += SWEEP_THREADS) {
        const u64 kept = kept_arr[i >> 6];
        if ((kept >> (i & 63)) & 1ull) {
            const int pos = prefix[i >> 6] + __popcll(kept & ((1ull << (i & 63)) - 1ull));
            if (kept_rank) kept_rank[pos - base0] = i;
            keep[pos] = (int64_t)order[i];
        }
    }
    if (tid == 0) {
        *keep_count = s_base;
        if (kept_n) *kept_n = s_base - base0;
#ifdef PP_TIMING
        printf("sweep total %lld cycles\n", clock64() - t_start);
#endif
    }
}

// Boxes below the first level that the first level's keep set does not suppress, compacted in rank order: they
// only need NMS among themselves afterwards (every box above them has already been accounted for).
constexpr uint32_t FLT_AGG = 1u << 30, FLT_PREFIX = 2u << 30, FLT_MASK = 3u << 30;

constexpr int FLT_BOXES = 64;                       // boxes per CTA
constexpr int FLT_SPLIT = NMS_THREADS / FLT_BOXES;  // threads that share one box (each scans 1/FLT_SPLIT of the keep set)

template <int MODE>
__global__ void __launch_bounds__(NMS_THREADS)
nms_filter_kernel(const float4 *__restrict__ srect, const uint32_t *__restrict__ order, int32_t *__restrict__ sc,
                  const int32_t *__restrict__ kept_rank, float thr, float4 *__restrict__ srect2,
                  uint32_t *__restrict__ order2, uint32_t *status, const Aux saux, const Aux saux2)
{
    typedef PairGeom<MODE> G;
    __shared__ float4 s_k[NMS_THREADS];
    __shared__ float4 s_k0[MODE != PP_NMS_AABB2D ? NMS_THREADS : 1];
    __shared__ float4 s_k1[MODE != PP_NMS_AABB2D ? NMS_THREADS : 1];
    __shared__ float4 s_k2[MODE == PP_NMS_BOX3D ? NMS_THREADS : 1];
    __shared__ uint32_t s_tile, s_excl;
    __shared__ uint32_t s_warp[NMS_THREADS / 32];
    __shared__ unsigned char s_dead[FLT_BOXES];
    const int n = sc[SC_N], n1 = sc[SC_N1], k1 = sc[SC_K1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int num_tiles = (n - n1 + FLT_BOXES - 1) / FLT_BOXES;
    if ((int)blockIdx.x >= num_tiles) return;
    if (tid == 0) s_tile = atomicAdd((uint32_t *)&sc[SC_TICKET], 1u);
    if (tid < FLT_BOXES) s_dead[tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int bi = tid / FLT_SPLIT, part = tid % FLT_SPLIT;      // box of this thread, its share of the keep set
    const int r = n1 + (int)tile * FLT_BOXES + bi;
    const bool valid = r < n;
    const float4 box = valid ? srect[r] : make_float4(3e38f, 3e38f, -3e38f, -3e38f);
    typename G::T gbox;
    if (MODE != PP_NMS_AABB2D && valid) gbox = G::load(saux.a0, saux.a1, saux.a2, r);
    const bool zero_hits = 0.f > thr;
    bool dead = false;
    for (int k0 = 0; k0 < k1; k0 += NMS_THREADS) {
        if (k0 + tid < k1) {
            const int kr = kept_rank[k0 + tid];
            s_k[tid] = srect[kr];
            if (MODE != PP_NMS_AABB2D) { s_k0[tid] = saux.a0[kr]; s_k1[tid] = saux.a1[kr]; }
            if (MODE == PP_NMS_BOX3D) s_k2[tid] = saux.a2[kr];
        }
        __syncthreads();
        const int kn = min(NMS_THREADS, k1 - k0);
        if (valid && !dead) {
            // four kept boxes per step (independent loads and compares: the loop is latency-bound at two CTAs per SM)
            for (int j0 = part; j0 < kn && !dead; j0 += 4 * FLT_SPLIT) {
                bool any = false;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u * FLT_SPLIT;
                    if (j < kn) {
                        const float4 q = s_k[j];
                        // empty intersection -> iou == 0 exactly; otherwise bbox_iou2D with (remaining, selected) = (box, kept)
                        const bool apart = !(fminf(box.z, q.z) > fmaxf(box.x, q.x) && fminf(box.w, q.w) > fmaxf(box.y, q.y));
                        bool hit;
                        if (apart) hit = zero_hits;
                        else if (MODE != PP_NMS_AABB2D) hit = G::exceeds(gbox, G::load(s_k0, s_k1, s_k2, j), thr);
                        else hit = rect_iou(box, q, 0, 1e-6f) > thr;
                        any |= hit;
                    }
                }
                dead = any;
            }
        }
        __syncthreads();
    }
    if (dead) s_dead[bi] = 1;
    __syncthreads();
    // ordered compaction: block scan of the survivor flags + exclusive prefix over the (ticket-ordered) tiles
    const bool alive = tid < FLT_BOXES && (n1 + (int)tile * FLT_BOXES + tid) < n && !s_dead[tid];
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, alive);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NMS_THREADS / 32; ++w) {
        const uint32_t c = s_warp[w];
        if (w < warp) wbase += c;
        total += c;
    }
    // exclusive prefix over the (ticket-ordered) tiles: the tile publishes its count, then all its threads read the counts
    // of ALL predecessors with independent loads (a few hundred tiles: one L2 round trip and a block reduction instead of
    // a look-back chain of dependent loads, which was most of this kernel's time)
    if (tid == 0) atomicExch(status + tile, FLT_AGG | total);
    uint32_t before = 0;
    for (uint32_t t = tid; t < tile; t += NMS_THREADS) {
        uint32_t v;
        do { v = *((volatile uint32_t *)(status + t)); } while ((v & FLT_MASK) == 0);
        before += v & ~FLT_MASK;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, o);
    __syncthreads();                      // (s_warp is read above)
    if (lane == 0) s_warp[warp] = before;
    __syncthreads();
    if (tid == 0) {
        uint32_t excl = 0;
#pragma unroll
        for (int w = 0; w < NMS_THREADS / 32; ++w) excl += s_warp[w];
        s_excl = excl;
        if ((int)tile == num_tiles - 1) sc[SC_N2] = (int32_t)(excl + total);
    }
    __syncthreads();
    if (alive) {
        const int rr = n1 + (int)tile * FLT_BOXES + tid;
        const uint32_t pos = s_excl + wbase + __popc(bal & lanemask_lt());
        srect2[pos] = srect[rr];
        order2[pos] = order[rr];
        if (MODE != PP_NMS_AABB2D) { saux2.a0[pos] = saux.a0[rr]; saux2.a1[pos] = saux.a1[rr]; }
        if (MODE == PP_NMS_BOX3D) saux2.a2[pos] = saux.a2[rr];
    }
}

// The filter for the clipped pair tests, queued like nms_mask_clip_kernel: every thread pushes the (box, kept box)
// pairs whose rectangles overlap onto a first shared-memory queue, FC_CHUNK kept boxes per round; when a further round
// might not fit, and after the last one, all threads pop it, run the cheap test of the mode and push the undecided
// pairs onto a second queue, which is evaluated exactly with full warps when it is nearly full and at the end.  The
// ordered compaction of the survivors is the same as in nms_filter_kernel.
constexpr int FC_CHUNK = 64;
constexpr int FC_Q1 = 8192;        // entries of the first queue; a round adds at most FLT_BOXES x FC_CHUNK
constexpr int FC_Q2 = 2048;

template <int MODE>
__global__ void __launch_bounds__(NMS_THREADS, 2)
nms_filter_clip_kernel(const float4 *__restrict__ srect, const uint32_t *__restrict__ order, int32_t *__restrict__ sc,
                       const int32_t *__restrict__ kept_rank, float thr, float4 *__restrict__ srect2,
                       uint32_t *__restrict__ order2, uint32_t *status, const Aux saux, const Aux saux2)
{
    typedef PairGeom<MODE> G;
    __shared__ float4 s_k[FC_CHUNK];
    __shared__ float4 s_b[FLT_BOXES], s_b0[FLT_BOXES], s_b1[FLT_BOXES];
    __shared__ float4 s_b2[MODE == PP_NMS_BOX3D ? FLT_BOXES : 1];
    __shared__ uint32_t s_q1[FC_Q1];                       // (box << 16) | index into the keep set
    __shared__ uint32_t s_q2[FC_Q2];
    __shared__ int s_n1, s_n2;
    __shared__ uint32_t s_tile, s_excl;
    __shared__ uint32_t s_warp[NMS_THREADS / 32];
    __shared__ unsigned char s_dead[FLT_BOXES];
    const int n = sc[SC_N], n1 = sc[SC_N1], k1 = sc[SC_K1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int num_tiles = (n - n1 + FLT_BOXES - 1) / FLT_BOXES;
    if ((int)blockIdx.x >= num_tiles) return;
    if (tid == 0) { s_tile = atomicAdd((uint32_t *)&sc[SC_TICKET], 1u); s_n1 = 0; s_n2 = 0; }
    if (tid < FLT_BOXES) s_dead[tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tid < FLT_BOXES) {
        const int r = n1 + (int)tile * FLT_BOXES + tid;
        if (r < n) {
            s_b[tid] = srect[r];
            s_b0[tid] = saux.a0[r];
            s_b1[tid] = saux.a1[r];
            if (MODE == PP_NMS_BOX3D) s_b2[tid] = saux.a2[r];
        } else {
            s_b[tid] = make_float4(3e38f, 3e38f, -3e38f, -3e38f);
            s_dead[tid] = 2;                                   // not a box
        }
    }
    const int bi = tid / FLT_SPLIT, part = tid % FLT_SPLIT;
    const bool zero_hits = 0.f > thr;
    int k0 = 0;
#pragma unroll 1
    while (true) {
        // fill: rounds of FC_CHUNK kept boxes until the next round might not fit
#pragma unroll 1
        for (; k0 < k1; k0 += FC_CHUNK) {
            __syncthreads();                                   // s_k free, s_n1 settled
            if (s_n1 > FC_Q1 - FLT_BOXES * FC_CHUNK) break;    // (uniform: the next push comes after the next barrier)
            if (tid < FC_CHUNK && k0 + tid < k1) s_k[tid] = srect[kept_rank[k0 + tid]];
            __syncthreads();
            const int kn = min(FC_CHUNK, k1 - k0);
            if (!s_dead[bi]) {
                const float4 box = s_b[bi];
                for (int j = part; j < kn; j += FLT_SPLIT) {
                    const float4 q = s_k[j];
                    // empty rectangle intersection -> iou == 0 exactly
                    if (fminf(box.z, q.z) > fmaxf(box.x, q.x) && fminf(box.w, q.w) > fmaxf(box.y, q.y))
                        s_q1[atomicAdd(&s_n1, 1)] = ((uint32_t)bi << 16) | (uint32_t)(k0 + j);
                    else if (zero_hits)
                        s_dead[bi] = 1;
                }
            }
        }
        __syncthreads();
        // drain
        const int q1n = s_n1;
        int base = 0, n2 = 0;
#pragma unroll 1
        while (true) {
#pragma unroll 1
            for (; base < q1n && n2 <= FC_Q2 - NMS_THREADS; base += NMS_THREADS) {
                bool pass = false;
                if (base + tid < q1n) {
                    const uint32_t pr = s_q1[base + tid];
                    const int b = pr >> 16, kr = kept_rank[pr & 0xFFFFu];
                    pass = !s_dead[b] && G::maybe(G::load(s_b0, s_b1, s_b2, b), G::load(saux.a0, saux.a1, saux.a2, kr), thr);
                    if (pass) s_q2[atomicAdd(&s_n2, 1)] = pr;
                }
                n2 += __syncthreads_count(pass);
            }
            for (int k = tid; k < n2; k += NMS_THREADS) {
                const uint32_t pr = s_q2[k];
                const int b = pr >> 16, kr = kept_rank[pr & 0xFFFFu];
                if (!s_dead[b] && G::exact(G::load(s_b0, s_b1, s_b2, b), G::load(saux.a0, saux.a1, saux.a2, kr), thr)) s_dead[b] = 1;
            }
            __syncthreads();
            if (tid == 0) s_n2 = 0;
            n2 = 0;
            __syncthreads();
            if (base >= q1n) break;
        }
        if (tid == 0) s_n1 = 0;
        __syncthreads();
        if (k0 >= k1) break;
    }
    // ordered compaction: block scan of the survivor flags + exclusive prefix over the (ticket-ordered) tiles
    const bool alive = tid < FLT_BOXES && !s_dead[tid];
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, alive);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NMS_THREADS / 32; ++w) {
        const uint32_t c = s_warp[w];
        if (w < warp) wbase += c;
        total += c;
    }
    // exclusive prefix over the (ticket-ordered) tiles: the tile publishes its count, then all its threads read the counts
    // of ALL predecessors with independent loads (a few hundred tiles: one L2 round trip and a block reduction instead of
    // a look-back chain of dependent loads, which was most of this kernel's time)
    if (tid == 0) atomicExch(status + tile, FLT_AGG | total);
    uint32_t before = 0;
    for (uint32_t t = tid; t < tile; t += NMS_THREADS) {
        uint32_t v;
        do { v = *((volatile uint32_t *)(status + t)); } while ((v & FLT_MASK) == 0);
        before += v & ~FLT_MASK;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, o);
    __syncthreads();                      // (s_warp is read above)
    if (lane == 0) s_warp[warp] = before;
    __syncthreads();
    if (tid == 0) {
        uint32_t excl = 0;
#pragma unroll
        for (int w = 0; w < NMS_THREADS / 32; ++w) excl += s_warp[w];
        s_excl = excl;
        if ((int)tile == num_tiles - 1) sc[SC_N2] = (int32_t)(excl + total);
    }
    __syncthreads();
    if (alive) {
        const int rr = n1 + (int)tile * FLT_BOXES + tid;
        const uint32_t pos = s_excl + wbase + __popc(bal & lanemask_lt());
        srect2[pos] = srect[rr];
        order2[pos] = order[rr];
        saux2.a0[pos] = saux.a0[rr];
        saux2.a1[pos] = saux.a1[rr];
        if (MODE == PP_NMS_BOX3D) saux2.a2[pos] = saux.a2[rr];
    }
}

struct NmsWs {
    int32_t *sc;
    uint32_t *status;          // look-back state of the filter's compaction
    u64 *band1, *band2;
    size_t zero_bytes;
    float4 *rect, *srect, *srect2;
    uint32_t *keys, *keys_sorted, *order, *order2;
    int32_t *kept_rank;
    Aux aux, saux, saux2;                  // unsorted / sorted / level-2 geometry (modes other than AABB2D)
    u64 *mask1, *mask2;
    void *sort_ws;
    size_t sort_ws_bytes;
    int nw1, nw2;
};

NmsWs carve(void *ws, int64_t N, int mode, size_t *total)
{
    NmsWs w;
    const int64_t n1 = N > 0 ? N : 1;
    const int64_t l1 = n1 < NMS_LEVEL1 ? n1 : NMS_LEVEL1;
    const int64_t l2 = n1 > NMS_LEVEL1 ? n1 - NMS_LEVEL1 : 0;
    w.nw1 = (int)ceil_div(l1, 64);
    w.nw2 = (int)ceil_div(l2 > 0 ? l2 : 1, 64);
    Arena a(ws, (size_t)-1);
    w.sc = a.take<int32_t>(64);
    w.status = a.take<uint32_t>((size_t)ceil_div(l2 > 0 ? l2 : 1, FLT_BOXES));
    // the sort's counters are zeroed together with ours (one memset per call)
    w.sort_ws_bytes = sort_workspace_bytes(n1);
    w.sort_ws = a.take<char>(w.sort_ws_bytes);
    w.zero_bytes = (size_t)((char *)w.sort_ws - (char *)ws) + sort_zero_bytes(n1);
    // (the bands need no zeroing: every word the sweep reads of a box's rows is written by the mask kernel, and rows
    // that are not boxes are never candidates)
    w.band1 = a.take<u64>((size_t)w.nw1 * SW_BAND);
    w.band2 = a.take<u64>(l2 > 0 ? (size_t)w.nw2 * SW_BAND : 1);
    w.rect = a.take<float4>((size_t)n1);
    w.srect = a.take<float4>((size_t)n1);
    w.srect2 = a.take<float4>((size_t)(l2 > 0 ? l2 : 1));
    w.keys = a.take<uint32_t>((size_t)n1);
    w.keys_sorted = a.take<uint32_t>((size_t)n1);
    w.order = a.take<uint32_t>((size_t)n1);
    w.order2 = a.take<uint32_t>((size_t)(l2 > 0 ? l2 : 1));
    w.kept_rank = a.take<int32_t>((size_t)l1);
    const int naux = mode == PP_NMS_BOX3D ? 3 : (mode == PP_NMS_ROT_BEV ? 2 : 0);
    float4 **slots[3][3] = {{&w.aux.a0, &w.aux.a1, &w.aux.a2}, {&w.saux.a0, &w.saux.a1, &w.saux.a2},
                            {&w.saux2.a0, &w.saux2.a1, &w.saux2.a2}};
    for (int g = 0; g < 3; ++g)
        for (int k = 0; k < 3; ++k)
            *slots[g][k] = k < naux ? a.take<float4>((size_t)(g == 2 ? (l2 > 0 ? l2 : 1) : n1)) : nullptr;
    w.mask1 = a.take<u64>((size_t)l1 * w.nw1);
    w.mask2 = a.take<u64>(l2 > 0 ? (size_t)l2 * w.nw2 : 1);
    *total = align_up(a.off);
    return w;
}

int launch_level(const float4 *rects, const int32_t *n_ptr, int64_t n_max, float thr, int nw, u64 *mask, u64 *band,
                 const uint32_t *order, int64_t *keep, const int32_t *keep_base, int32_t *keep_count,
                 int32_t *kept_rank, int32_t *kept_n, int mode, const Aux saux, cudaStream_t st)
{
    dim3 grid((unsigned)ceil_div(n_max, MT_COLS), (unsigned)ceil_div(n_max, MT_ROWS));
    // thr >= 0: a pair whose bounding rectangles are apart has iou == 0, which only exceeds a negative threshold
#define PP_MASK(PF, MD) nms_mask_kernel<PF, MD><<<grid, MT_ROWS, 0, st>>>(rects, n_ptr, thr, nw, mask, band, saux)
#define PP_MASKC(PF, MD)                                                                                              \
    do {                                                                                                              \
        cudaFuncSetAttribute(nms_mask_clip_kernel<PF, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize,               \
                             (int)sizeof(ClipSmem<MD>));                                                              \
        nms_mask_clip_kernel<PF, MD><<<grid, MT_ROWS, sizeof(ClipSmem<MD>), st>>>(rects, n_ptr, thr, nw, mask, band, saux); \
    } while (0)
    if (mode == PP_NMS_BOX3D) { if (thr >= 0.f) PP_MASKC(true, PP_NMS_BOX3D); else PP_MASKC(false, PP_NMS_BOX3D); }
    else if (mode == PP_NMS_ROT_BEV) { if (thr >= 0.f) PP_MASKC(true, PP_NMS_ROT_BEV); else PP_MASKC(false, PP_NMS_ROT_BEV); }
    else { if (thr >= 0.f) PP_MASK(true, PP_NMS_AABB2D); else PP_MASK(false, PP_NMS_AABB2D); }
#undef PP_MASK
#undef PP_MASKC
    if (int rc = check_launch("nms_mask_kernel")) return rc;
    const size_t smem = ((size_t)SW_RING * SW_BAND + 3 * (size_t)nw) * sizeof(u64) + (size_t)nw * sizeof(int);
    PP_REQUIRE(smem <= SW_SMEM_MAX, "too many boxes for the sweep's shared memory");
    nms_sweep_kernel<<<1, SWEEP_THREADS, smem, st>>>(mask, band, nw, n_ptr, order, keep, keep_base, keep_count, kept_rank,
                                                     kept_n, nw);
    return check_launch("nms_sweep_kernel");
}

}  // namespace
}  // namespace pp

using namespace pp;

extern "C" size_t pp_nms_workspace_bytes_mode(int64_t N, int iou_mode)
{
    size_t total;
    carve(nullptr, N, iou_mode, &total);
    return total;
}

extern "C" size_t pp_nms_workspace_bytes(int64_t N) { return pp_nms_workspace_bytes_mode(N, PP_NMS_AABB2D); }

extern "C" int pp_nms_mode(const float *boxes9, const float *scores, int64_t score_stride, int64_t N, float score_thr,
                           float iou_thr, int iou_mode, int64_t *keep, int32_t *keep_count, void *workspace,
                           size_t workspace_bytes, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    PP_REQUIRE(keep_count, "null keep_count");
    PP_REQUIRE(N >= 0 && N <= 131072, "N must be in [0, 131072]");
    PP_REQUIRE(iou_mode == PP_NMS_AABB2D || iou_mode == PP_NMS_ROT_BEV || iou_mode == PP_NMS_BOX3D, "unknown iou_mode");
    if (N == 0) {
        PP_CUDA_TRY(cudaMemsetAsync(keep_count, 0, sizeof(int32_t), st));
        return PP_OK;
    }
    PP_REQUIRE(boxes9 && scores && keep && workspace, "null pointer");
    PP_REQUIRE(score_stride >= 1, "bad score stride");
    size_t total;
    NmsWs w = carve(workspace, N, iou_mode, &total);
    if (workspace_bytes < total) {
        set_error("nms workspace too small: %zu < %zu", workspace_bytes, total);
        return PP_ERR_WORKSPACE;
    }
    PP_CUDA_TRY(cudaMemsetAsync(workspace, 0, w.zero_bytes, st));     // scalars, look-back state, the sort's counters
    prof_mark("memset");
    // per device and context, cheap: set on every call (a process may drive several GPUs, from several threads)
    PP_CUDA_TRY(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM_MAX));
    const unsigned nb = (unsigned)ceil_div(N, NMS_THREADS);
    nms_prepare_kernel<<<(unsigned)ceil_div(N, PREP_THREADS), PREP_THREADS, 0, st>>>(
        boxes9, scores, score_stride, N, score_thr, w.rect, w.keys, w.sc + SC_N, iou_mode, w.aux, sort_hist(w.sort_ws, N));
    if (int rc = check_launch("nms_prepare_kernel")) return rc;
    // the sort's last pass moves the rectangles (and the geometry of the clipped modes) with their keys and sets the size
    // of level 1; sizes that the one-CTA sort handles use the gather kernel
    SortTail tail = {};
    tail.in[0] = w.rect; tail.out[0] = w.srect; tail.arrays = 1;
    const float4 *ain[3] = {w.aux.a0, w.aux.a1, w.aux.a2};
    float4 *aout[3] = {w.saux.a0, w.saux.a1, w.saux.a2};
    for (int k = 0; k < 3; ++k)
        if (ain[k]) { tail.in[tail.arrays] = ain[k]; tail.out[tail.arrays] = aout[k]; ++tail.arrays; }
    tail.count_in = w.sc + SC_N; tail.count_out = w.sc + SC_N1; tail.count_cap = NMS_LEVEL1;
    const bool fused_gather = sort_runs_tail(N);
    if (int rc = sort_pairs_u32(w.keys, nullptr, w.keys_sorted, w.order, N, w.sort_ws, w.sort_ws_bytes, st, true, true,
                                fused_gather ? &tail : nullptr))
        return rc;
    if (!fused_gather) {
        nms_gather_kernel<<<nb, NMS_THREADS, 0, st>>>(w.rect, w.order, w.sc, w.srect, NMS_LEVEL1, w.aux, w.saux);
        if (int rc = check_launch("nms_gather_kernel")) return rc;
    }
    // level 1: greedy NMS of the NMS_LEVEL1 best-scored candidates
    const int64_t l1 = N < NMS_LEVEL1 ? N : NMS_LEVEL1;
    if (int rc = launch_level(w.srect, w.sc + SC_N1, l1, iou_thr, w.nw1, w.mask1, w.band1, w.order, keep, nullptr,
                              keep_count, w.kept_rank, w.sc + SC_K1, iou_mode, w.saux, st))
        return rc;
    if (N <= NMS_LEVEL1) return PP_OK;
    // the other candidates: drop those suppressed by level 1's keep set, compact in rank order, NMS among themselves
    const int64_t l2 = N - NMS_LEVEL1;
    const unsigned fb = (unsigned)ceil_div(l2, FLT_BOXES);
#define PP_FILTER(KERNEL, MD) KERNEL<MD><<<fb, NMS_THREADS, 0, st>>>(w.srect, w.order, w.sc, w.kept_rank, iou_thr, w.srect2, \
                                                                    w.order2, w.status, w.saux, w.saux2)
    if (iou_mode == PP_NMS_BOX3D) PP_FILTER(nms_filter_clip_kernel, PP_NMS_BOX3D);
    else if (iou_mode == PP_NMS_ROT_BEV) PP_FILTER(nms_filter_clip_kernel, PP_NMS_ROT_BEV);
    else PP_FILTER(nms_filter_kernel, PP_NMS_AABB2D);
#undef PP_FILTER
    if (int rc = check_launch("nms_filter_kernel")) return rc;
    return launch_level(w.srect2, w.sc + SC_N2, l2, iou_thr, w.nw2, w.mask2, w.band2, w.order2, keep, w.sc + SC_K1,
                        keep_count, nullptr, nullptr, iou_mode, w.saux2, st);
}

extern "C" int pp_nms(const float *boxes9, const float *scores, int64_t score_stride, int64_t N, float score_thr,
                      float iou_thr, int64_t *keep, int32_t *keep_count, void *workspace, size_t workspace_bytes,
                      pp_stream_t stream)
{
    return pp_nms_mode(boxes9, scores, score_stride, N, score_thr, iou_thr, PP_NMS_AABB2D, keep, keep_count, workspace,
                       workspace_bytes, stream);
}
