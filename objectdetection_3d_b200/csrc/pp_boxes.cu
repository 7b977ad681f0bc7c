// Stage 3 element-wise box math: BBoxCoder (model/utils.py:276-337), limit_period (:339-350),
// anchor grid (:168-264), box corners / xy bounding rectangle (ops/ops_torch.py:13-256) and the
// axis-aligned IoU matrix (:538-607).  Same op order as the eager torch code with every op rounded
// separately (-fmad=false + __f*_rn), so IoU on identical rectangles is bit-identical.
#include "pp_boxes.cuh"
#include "pp_common.cuh"

namespace pp {
namespace {

constexpr int BX_THREADS = 256;

__global__ void __launch_bounds__(BX_THREADS)
box_encode_kernel(const float *__restrict__ src, const float *__restrict__ dst, int64_t K, float *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * BX_THREADS + threadIdx.x;
    if (i >= K) return;
    float a[9], g[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { a[k] = src[i * 9 + k]; g[k] = dst[i * 9 + k]; }
    float zg = __fadd_rn(g[2], __fdiv_rn(g[5], 2.0f));
    float za = __fadd_rn(a[2], __fdiv_rn(a[5], 2.0f));
    float diag = __fsqrt_rn(__fadd_rn(__fmul_rn(a[3], a[3]), __fmul_rn(a[4], a[4])));
    float *o = out + i * 9;
    o[0] = __fdiv_rn(__fsub_rn(g[0], a[0]), diag);
    o[1] = __fdiv_rn(__fsub_rn(g[1], a[1]), diag);
    o[2] = __fdiv_rn(__fsub_rn(zg, za), a[5]);
    o[3] = logf(__fdiv_rn(g[3], a[3]));
    o[4] = logf(__fdiv_rn(g[4], a[4]));
    o[5] = logf(__fdiv_rn(g[5], a[5]));
    o[6] = __fsub_rn(g[6], a[6]);
    o[7] = __fsub_rn(g[7], a[7]);
    o[8] = __fsub_rn(g[8], a[8]);
}

__global__ void __launch_bounds__(BX_THREADS)
box_decode_kernel(const float *__restrict__ anchors, const float *__restrict__ deltas, int64_t K,
                  float *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * BX_THREADS + threadIdx.x;
    if (i >= K) return;
    float a[9], t[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { a[k] = anchors[i * 9 + k]; t[k] = deltas[i * 9 + k]; }
    float o[9];
    decode_one(a, t, o);
#pragma unroll
    for (int k = 0; k < 9; ++k) out[i * 9 + k] = o[k];
}

__global__ void __launch_bounds__(BX_THREADS)
limit_period_kernel(const float *__restrict__ val, int64_t n, float offset, float period, float *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * BX_THREADS + threadIdx.x;
    if (i >= n) return;
    float v = val[i];
    out[i] = __fsub_rn(v, __fmul_rn(floorf(__fadd_rn(__fdiv_rn(v, period), offset)), period));
}

struct AnchorSpec {
    float range[6];
    float sizes[16 * 3];
    float rots[16 * 3];
    int S, R, D, H, W;
};

// anchor i of grid_anchors(...).reshape(-1, 9): i = (((z*H + y)*W + x)*S + s)*R + r
__device__ __forceinline__ void anchor_at(const AnchorSpec &a, int64_t i, float o[9])
{
    int r = (int)(i % a.R);
    int64_t t = i / a.R;
    int s = (int)(t % a.S); t /= a.S;
    int x = (int)(t % a.W); t /= a.W;
    int y = (int)(t % a.H);
    int z = (int)(t / a.H);
    o[0] = linspace_at(a.range[0], a.range[3], a.W, x);
    o[1] = linspace_at(a.range[1], a.range[4], a.H, y);
    o[2] = linspace_at(a.range[2], a.range[5], a.D, z);
    o[3] = a.sizes[s * 3]; o[4] = a.sizes[s * 3 + 1]; o[5] = a.sizes[s * 3 + 2];
    o[6] = a.rots[r * 3]; o[7] = a.rots[r * 3 + 1]; o[8] = a.rots[r * 3 + 2];
}

__device__ __forceinline__ float sigmoid_f32(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// ---- fused head post-processing, Anchor3DHead.get_bboxes_single (model/PointPillars.py:1040-1092) -----------------
// The reference materialises every anchor (69 MB at its 400 x 400 map), permutes the three head tensors, decodes all
// anchors and only then keeps nms_pre of them.  Here: (1) one pass over the class logits gives the per-anchor score
// (sigmoid + class max, :1050-1058) straight from the (channels, H, W) layout; (2) after the top-k, the nms_pre
// survivors get their anchor generated on the fly, their deltas / logits / direction logits gathered from the head
// tensors, and are decoded (decode is element-wise, so decode-after-select equals the reference's decode-all).
__global__ void __launch_bounds__(BX_THREADS)
head_max_scores_kernel(const float *__restrict__ cls, int A_per, int ncls, int64_t HW, float *__restrict__ max_scores)
{
    const int64_t pos = (int64_t)blockIdx.x * BX_THREADS + threadIdx.x;      // y * W + x: coalesced channel rows
    if (pos >= HW) return;
    for (int a = 0; a < A_per; ++a) {
        float m = -1.f;
        for (int c = 0; c < ncls; ++c) m = fmaxf(m, sigmoid_f32(cls[(int64_t)(a * ncls + c) * HW + pos]));
        max_scores[pos * A_per + a] = m;
    }
}

__global__ void __launch_bounds__(128)
head_select_decode_kernel(const float *__restrict__ cls, const float *__restrict__ reg, const float *__restrict__ dirs,
                          const int64_t *__restrict__ rows, int64_t K, const AnchorSpec spec, int ncls,
                          float *__restrict__ boxes, float *__restrict__ scores, int32_t *__restrict__ dir_bits)
{
    const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (j >= K) return;
    const int64_t r = rows ? rows[j] : j;
    const int A_per = spec.S * spec.R;
    const int64_t HW = (int64_t)spec.H * spec.W;
    const int64_t pos = r / A_per;
    const int a = (int)(r - pos * A_per);
    float an[9], t[9], o[9];
    anchor_at(spec, r, an);
#pragma unroll
    for (int k = 0; k < 9; ++k) t[k] = reg[(int64_t)(a * 9 + k) * HW + pos];
    decode_one(an, t, o);
#pragma unroll
    for (int k = 0; k < 9; ++k) boxes[j * 9 + k] = o[k];
    for (int c = 0; c < ncls; ++c) scores[j * ncls + c] = sigmoid_f32(cls[(int64_t)(a * ncls + c) * HW + pos]);
    // row r of dir_preds.permute(1, 2, 0).reshape(-1, 6) = channels 6a .. 6a+5 of the concatenated tensor (:1045-1048)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float d0 = dirs[(int64_t)(a * 6 + 2 * k) * HW + pos], d1 = dirs[(int64_t)(a * 6 + 2 * k + 1) * HW + pos];
        dir_bits[j * 3 + k] = d1 > d0 ? 1 : 0;                  // torch.max(...)[1]: first maximum
    }
}

// direction fix-up of the kept boxes, :1085-1092: rot = limit_period(rot - off, 1, pi) + off + pi * bit
__global__ void __launch_bounds__(BX_THREADS)
head_fixup_kernel(float *__restrict__ boxes, const int32_t *__restrict__ dir_bits, int64_t K, float dir_offset)
{
    const int64_t i = (int64_t)blockIdx.x * BX_THREADS + threadIdx.x;
    if (i >= K * 3) return;
    const int64_t j = i / 3;
    const int k = (int)(i - j * 3);
    const float period = 3.14159274101257324f;                  // float32(np.pi)
    const float v = __fsub_rn(boxes[j * 9 + 6 + k], dir_offset);
    const float lp = __fsub_rn(v, __fmul_rn(floorf(__fadd_rn(__fdiv_rn(v, period), 1.0f)), period));
    boxes[j * 9 + 6 + k] = __fadd_rn(__fadd_rn(lp, dir_offset), __fmul_rn(period, (float)dir_bits[j * 3 + k]));
}

__global__ void __launch_bounds__(BX_THREADS) grid_anchors_kernel(const AnchorSpec a, int64_t total, float *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * BX_THREADS + threadIdx.x;   // one thread per anchor
    if (i >= total) return;
    float *o = out + i * 9;
    float v[9];
    anchor_at(a, i, v);
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = v[k];
}

__global__ void __launch_bounds__(BX_THREADS)
box_corners_kernel(const float *__restrict__ boxes, int64_t N, float *__restrict__ corners, float *__restrict__ rect)
{
    int64_t i = (int64_t)blockIdx.x * BX_THREADS + threadIdx.x;
    if (i >= N) return;
    float b[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) b[k] = boxes[i * 9 + k];
    float c[8][3];
    box_corners(b, c);
    if (corners) {
#pragma unroll
        for (int v = 0; v < 8; ++v)
#pragma unroll
            for (int k = 0; k < 3; ++k) corners[i * 24 + v * 3 + k] = c[v][k];
    }
    if (rect) {
        float4 r = corners_to_rect(c);
        reinterpret_cast<float4 *>(rect)[i] = r;
    }
}

// (m,n) IoU matrix: one thread per output element, column boxes staged in shared memory.
__global__ void __launch_bounds__(BX_THREADS)
iou2d_kernel(const float *__restrict__ b1, int64_t m, const float *__restrict__ b2, int64_t n, int mode, float eps,
             float *__restrict__ out)
{
    __shared__ float4 s_col[BX_THREADS];
    const int64_t j0 = (int64_t)blockIdx.x * BX_THREADS;
    const int64_t j = j0 + threadIdx.x;
    if (j < n) s_col[threadIdx.x] = reinterpret_cast<const float4 *>(b2)[j];
    __syncthreads();
    if (j >= n) return;
    const float4 q = s_col[threadIdx.x];
    const int64_t i0 = (int64_t)blockIdx.y * 32;
    for (int64_t i = i0; i < i0 + 32 && i < m; ++i) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(b1) + i);
        out[i * n + j] = rect_iou(a, q, mode, eps);
    }
}

// (m,n) rotated BEV IoU: one thread per column box (staged in shared memory), 32 rows per CTA
__global__ void __launch_bounds__(BX_THREADS)
iou_rot_kernel(const float *__restrict__ b1, int64_t m, const float *__restrict__ b2, int64_t n, float *__restrict__ out)
{
    __shared__ RRect s_row[32];
    const int64_t i0 = (int64_t)blockIdx.y * 32;
    if (threadIdx.x < 32 && i0 + threadIdx.x < m) s_row[threadIdx.x] = rrect_from_box9(b1 + (i0 + threadIdx.x) * 9);
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * BX_THREADS + threadIdx.x;
    if (j >= n) return;
    const RRect q = rrect_from_box9(b2 + j * 9);
    const float4 qa = rrect_aabb(q);
    for (int r = 0; r < 32 && i0 + r < m; ++r) {
        const RRect a = s_row[r];
        const float4 aa = rrect_aabb(a);
        float v = 0.f;
        if (fminf(aa.z, qa.z) > fmaxf(aa.x, qa.x) && fminf(aa.w, qa.w) > fmaxf(aa.y, qa.y)) v = rrect_iou(a, q);
        out[(i0 + r) * n + j] = v;
    }
}

// (n,m) oriented 3-D IoU from corners: one thread per column box, 32 row boxes per CTA staged in shared memory
__global__ void __launch_bounds__(128)
iou3d_kernel(const float *__restrict__ c1, int64_t n, const float *__restrict__ c2, int64_t m, float *__restrict__ vol,
             float *__restrict__ iou)
{
    __shared__ Box3 s_row[32];
    __shared__ float4 s_rect[32];
    const int64_t i0 = (int64_t)blockIdx.y * 32;
    if (threadIdx.x < 32 && i0 + threadIdx.x < n) {
        const float *c = c1 + (i0 + threadIdx.x) * 24;
        s_row[threadIdx.x] = box3_from_corners(c);
        s_rect[threadIdx.x] = corners_xy_rect(c);
    }
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (j >= m) return;
    const Box3 q = box3_from_corners(c2 + j * 24);
    const float4 qr = corners_xy_rect(c2 + j * 24);
    for (int r = 0; r < 32 && i0 + r < n; ++r) {
        // 0 unless the xy bounding rectangles of the corners overlap (the same pre-test the NMS tiles use)
        float v = 0.f, u = 0.f;
        const float4 ar = s_rect[r];
        if (fminf(ar.z, qr.z) > fmaxf(ar.x, qr.x) && fminf(ar.w, qr.w) > fmaxf(ar.y, qr.y)) u = box3_iou(s_row[r], q, &v);
        if (vol) vol[(i0 + r) * m + j] = v;
        iou[(i0 + r) * m + j] = u;
    }
}

// check_coplanar / check_nonzero, ops/ops_torch.py:610-690: flags bit 0 = fails the coplanarity test, bit 1 = a face
// triangle with area < eps.  (The reference sums the plane residuals of all six faces before taking the absolute
// value, :637-641; kept.)
__device__ __forceinline__ void normalize3(float v[3])
{
    const float nrm = fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12f);
    v[0] /= nrm; v[1] /= nrm; v[2] /= nrm;
}
__device__ __forceinline__ void cross3(const float a[3], const float b[3], float o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

__global__ void __launch_bounds__(BX_THREADS)
box3d_check_kernel(const float *__restrict__ corners, int64_t n, float eps, int32_t *__restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * BX_THREADS + threadIdx.x;
    if (i >= n) return;
    float c[8][3];
#pragma unroll
    for (int k = 0; k < 24; ++k) c[k / 3][k % 3] = corners[i * 24 + k];
    const int planes[6][4] = {{0, 1, 2, 3}, {3, 2, 6, 7}, {0, 1, 5, 4}, {0, 3, 7, 4}, {1, 2, 6, 5}, {4, 5, 6, 7}};
    const int tris[12][3] = {{0, 1, 2}, {0, 3, 2}, {4, 5, 6}, {4, 6, 7}, {1, 5, 6}, {1, 6, 2},
                             {0, 4, 7}, {0, 7, 3}, {3, 2, 6}, {3, 6, 7}, {0, 1, 5}, {0, 4, 5}};
    float res = 0.f;
    for (int p = 0; p < 6; ++p) {
        float e0[3], e1[3], d[3], nrm[3];
        for (int k = 0; k < 3; ++k) {
            e0[k] = c[planes[p][1]][k] - c[planes[p][0]][k];
            e1[k] = c[planes[p][2]][k] - c[planes[p][0]][k];
            d[k] = c[planes[p][3]][k] - c[planes[p][0]][k];
        }
        normalize3(e0); normalize3(e1);
        cross3(e0, e1, nrm);
        normalize3(nrm);
        res += d[0] * nrm[0] + d[1] * nrm[1] + d[2] * nrm[2];
    }
    int f = (fabsf(res) < eps) ? 0 : 1;
    for (int t = 0; t < 12; ++t) {
        float a[3], b[3], x[3];
        for (int k = 0; k < 3; ++k) {
            a[k] = c[tris[t][1]][k] - c[tris[t][0]][k];
            b[k] = c[tris[t][2]][k] - c[tris[t][0]][k];
        }
        cross3(a, b, x);
        if (sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) / 2.f < eps) f |= 2;
    }
    flags[i] = f;
}

// ---- target assignment, Anchor3DHead.assign_bboxes (model/PointPillars.py:964-978) -------------------------------
// IoU(ground truths x anchors) fused with its two reductions, so the (G, A) matrix (hundreds of MB at the reference's
// map sizes) is never stored.  Pass 1: per anchor max / first argmax over the ground truths (:968) and per ground
// truth max over the anchors (:971); pass 2: low-quality matching (:976-978), the anchors that tie a ground truth's
// best IoU -- the pair IoU is recomputed, which is bit-identical, so `==` means what it means in the reference.
constexpr int AS_THREADS = 256;
constexpr int AS_GT = 64;          // ground truths staged per pass

template <int MODE> struct AssignGeom;
template <> struct AssignGeom<PP_NMS_AABB2D> {
    typedef float4 T;
    static __device__ __forceinline__ T load(const float *p, int64_t i) { return reinterpret_cast<const float4 *>(p)[i]; }
    // bbox_iou2D(bboxes1 = ground truths, bboxes2 = anchors)
    static __device__ __forceinline__ float iou(const T &gt, const T &an) { return rect_iou(gt, an, 0, 1e-6f); }
};
template <> struct AssignGeom<PP_NMS_BOX3D> {
    struct T { Box3 b; float4 r; };
    static __device__ __forceinline__ T load(const float *p, int64_t i)
    {
        T t;
        t.b = box3_from_corners(p + i * 24);
        t.r = corners_xy_rect(p + i * 24);
        return t;
    }
    static __device__ __forceinline__ float iou(const T &gt, const T &an)
    {
        if (!(fminf(gt.r.z, an.r.z) > fmaxf(gt.r.x, an.r.x) && fminf(gt.r.w, an.r.w) > fmaxf(gt.r.y, an.r.y))) return 0.f;
        return box3_iou(gt.b, an.b, nullptr);
    }
};

template <int MODE, bool LOWQ>
__global__ void __launch_bounds__(AS_THREADS)
assign_kernel(const float *__restrict__ gt, int G, const float *__restrict__ anchors, int64_t A, float lo_thr,
              float *__restrict__ max_ov, int32_t *__restrict__ argmax, float *gt_max, uint8_t *__restrict__ lowq)
{
    typedef AssignGeom<MODE> Geo;
    __shared__ typename Geo::T s_gt[AS_GT];
    __shared__ float s_gmax[AS_GT];
    const int64_t a = (int64_t)blockIdx.x * AS_THREADS + threadIdx.x;
    const bool valid = a < A;
    typename Geo::T an;
    if (valid) an = Geo::load(anchors, a);
    float best = -1.f;
    int best_g = 0;
    bool flag = false;
    for (int g0 = 0; g0 < G; g0 += AS_GT) {
        const int gn = min(AS_GT, G - g0);
        __syncthreads();
        if ((int)threadIdx.x < gn) {
            s_gt[threadIdx.x] = Geo::load(gt, g0 + threadIdx.x);
            s_gmax[threadIdx.x] = LOWQ ? gt_max[g0 + threadIdx.x] : 0.f;
        }
        __syncthreads();
        for (int g = 0; g < gn; ++g) {
            if (LOWQ) {
                const float gm = s_gmax[g];
                if (!(gm >= lo_thr)) continue;                             // :977 (CTA-uniform)
                if (valid && Geo::iou(s_gt[g], an) == gm) flag = true;      // :978
            } else {
                const float v = valid ? Geo::iou(s_gt[g], an) : 0.f;
                if (v > best) { best = v; best_g = g0 + g; }               // first maximum
                // IoU >= 0: the float order equals the order of the bit patterns
                const unsigned wmax = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(v));
                if (wmax != 0u && (threadIdx.x & 31) == 0) atomicMax((unsigned *)&s_gmax[g], wmax);
            }
        }
        if (!LOWQ) {
            __syncthreads();
            if ((int)threadIdx.x < gn && __float_as_uint(s_gmax[threadIdx.x]) != 0u)
                atomicMax((unsigned *)(gt_max + g0 + threadIdx.x), __float_as_uint(s_gmax[threadIdx.x]));
        }
    }
    if (!valid) return;
    if (LOWQ) {
        lowq[a] = flag ? 1 : 0;
    } else {
        max_ov[a] = best;
        argmax[a] = best_g;
    }
}

__global__ void __launch_bounds__(BX_THREADS)
iou_jit_kernel(const float *__restrict__ boxes, int64_t N, const float *__restrict__ query, int64_t K, double eps,
               float *__restrict__ out)
{
    int64_t t = (int64_t)blockIdx.x * BX_THREADS + threadIdx.x;
    if (t >= N * K) return;
    int64_t nI = t / K, k = t % K;
    const float4 b = __ldg(reinterpret_cast<const float4 *>(boxes) + nI);
    const float4 q = __ldg(reinterpret_cast<const float4 *>(query) + k);
    // numba: f32 - f32 stays f32, "+ eps" (float64) promotes (ops/ops_numba.py:22-35)
    double box_area = ((double)__fsub_rn(q.z, q.x) + eps) * ((double)__fsub_rn(q.w, q.y) + eps);
    double iw = (double)__fsub_rn(fminf(b.z, q.z), fmaxf(b.x, q.x)) + eps;
    float r = 0.f;
    if (iw > 0) {
        double ih = (double)__fsub_rn(fminf(b.w, q.w), fmaxf(b.y, q.y)) + eps;
        if (ih > 0) {
            double ua = __dadd_rn(__dmul_rn((double)__fsub_rn(b.z, b.x) + eps, (double)__fsub_rn(b.w, b.y) + eps), box_area);
            ua = __dsub_rn(ua, __dmul_rn(iw, ih));
            r = (float)(__dmul_rn(iw, ih) / ua);
        }
    }
    out[t] = r;
}

}  // namespace
}  // namespace pp

using namespace pp;

#define GRID1(n) (unsigned)ceil_div((n), BX_THREADS), BX_THREADS, 0, (cudaStream_t)stream

extern "C" int pp_box_encode(const float *src, const float *dst, int64_t K, float *out, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(K >= 0, "K < 0");
    if (K == 0) return PP_OK;
    PP_REQUIRE(src && dst && out, "null pointer");
    box_encode_kernel<<<GRID1(K)>>>(src, dst, K, out);
    return check_launch("box_encode_kernel");
}

extern "C" int pp_box_decode(const float *anchors, const float *deltas, int64_t K, float *out, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(K >= 0, "K < 0");
    if (K == 0) return PP_OK;
    PP_REQUIRE(anchors && deltas && out, "null pointer");
    box_decode_kernel<<<GRID1(K)>>>(anchors, deltas, K, out);
    return check_launch("box_decode_kernel");
}

extern "C" int pp_limit_period(const float *val, int64_t n, float offset, float period, float *out, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return PP_OK;
    PP_REQUIRE(val && out, "null pointer");
    limit_period_kernel<<<GRID1(n)>>>(val, n, offset, period, out);
    return check_launch("limit_period_kernel");
}

extern "C" int pp_grid_anchors(const float *range6_host, const float *sizes_host, int S, const float *rots_host, int R,
                               int D, int H, int W, float *out, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(range6_host && sizes_host && rots_host && out, "null pointer");
    PP_REQUIRE(S > 0 && S <= 16 && R > 0 && R <= 16, "1..16 sizes / rotations supported");
    PP_REQUIRE(D > 0 && H > 0 && W > 0, "bad feature map");
    AnchorSpec a;
    for (int i = 0; i < 6; ++i) a.range[i] = range6_host[i];
    for (int i = 0; i < S * 3; ++i) a.sizes[i] = sizes_host[i];
    for (int i = 0; i < R * 3; ++i) a.rots[i] = rots_host[i];
    a.S = S; a.R = R; a.D = D; a.H = H; a.W = W;
    int64_t total = (int64_t)D * H * W * S * R;
    grid_anchors_kernel<<<GRID1(total)>>>(a, total, out);
    return check_launch("grid_anchors_kernel");
}

extern "C" int pp_box_corners3d(const float *boxes, int64_t N, float *corners, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(N >= 0, "N < 0");
    if (N == 0) return PP_OK;
    PP_REQUIRE(boxes && corners, "null pointer");
    box_corners_kernel<<<GRID1(N)>>>(boxes, N, corners, nullptr);
    return check_launch("box_corners_kernel");
}

extern "C" int pp_box_aabb2d(const float *boxes, int64_t N, float *rect, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(N >= 0, "N < 0");
    if (N == 0) return PP_OK;
    PP_REQUIRE(boxes && rect, "null pointer");
    PP_REQUIRE((uintptr_t)rect % 16 == 0, "rect must be 16-byte aligned");
    box_corners_kernel<<<GRID1(N)>>>(boxes, N, nullptr, rect);
    return check_launch("box_corners_kernel");
}

extern "C" int pp_bbox_iou2d(const float *b1, int64_t m, const float *b2, int64_t n, int mode, float eps, float *out,
                             pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(m >= 0 && n >= 0, "negative size");
    PP_REQUIRE(mode == PP_IOU || mode == PP_IOF || mode == PP_GIOU, "Unsupported mode");
    if (m == 0 || n == 0) return PP_OK;
    PP_REQUIRE(b1 && b2 && out, "null pointer");
    PP_REQUIRE((uintptr_t)b1 % 16 == 0 && (uintptr_t)b2 % 16 == 0, "boxes must be 16-byte aligned");
    PP_REQUIRE(ceil_div(m, 32) < 65536, "too many rows for one launch");
    dim3 grid((unsigned)ceil_div(n, BX_THREADS), (unsigned)ceil_div(m, 32));
    iou2d_kernel<<<grid, BX_THREADS, 0, (cudaStream_t)stream>>>(b1, m, b2, n, mode, eps, out);
    return check_launch("iou2d_kernel");
}

extern "C" int pp_iou_jit(const float *boxes, int64_t N, const float *query, int64_t K, double eps, float *out,
                          pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(N >= 0 && K >= 0, "negative size");
    if (N == 0 || K == 0) return PP_OK;
    PP_REQUIRE(boxes && query && out, "null pointer");
    PP_REQUIRE((uintptr_t)boxes % 16 == 0 && (uintptr_t)query % 16 == 0, "boxes must be 16-byte aligned");
    iou_jit_kernel<<<GRID1(N * K)>>>(boxes, N, query, K, eps, out);
    return check_launch("iou_jit_kernel");
}

extern "C" int pp_iou_rotated_bev(const float *b1, int64_t m, const float *b2, int64_t n, float *out, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(m >= 0 && n >= 0, "negative size");
    if (m == 0 || n == 0) return PP_OK;
    PP_REQUIRE(b1 && b2 && out, "null pointer");
    PP_REQUIRE(ceil_div(m, 32) < 65536, "too many rows for one launch");
    dim3 grid((unsigned)ceil_div(n, BX_THREADS), (unsigned)ceil_div(m, 32));
    iou_rot_kernel<<<grid, BX_THREADS, 0, (cudaStream_t)stream>>>(b1, m, b2, n, out);
    return check_launch("iou_rot_kernel");
}

extern "C" int pp_box3d_overlap(const float *corners1, int64_t n, const float *corners2, int64_t m, float *vol, float *iou,
                                pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(n >= 0 && m >= 0, "negative size");
    if (n == 0 || m == 0) return PP_OK;
    PP_REQUIRE(corners1 && corners2 && iou, "null pointer");
    PP_REQUIRE(ceil_div(n, 32) < 65536, "too many rows for one launch");
    dim3 grid((unsigned)ceil_div(m, 128), (unsigned)ceil_div(n, 32));
    iou3d_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(corners1, n, corners2, m, vol, iou);
    return check_launch("iou3d_kernel");
}

extern "C" int pp_box3d_check(const float *corners, int64_t n, float eps, int32_t *flags, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(n >= 0, "negative size");
    if (n == 0) return PP_OK;
    PP_REQUIRE(corners && flags, "null pointer");
    box3d_check_kernel<<<(unsigned)ceil_div(n, BX_THREADS), BX_THREADS, 0, (cudaStream_t)stream>>>(corners, n, eps, flags);
    return check_launch("box3d_check_kernel");
}

extern "C" int pp_assign_overlaps(const float *gt, int64_t G, const float *anchors, int64_t A, int iou_mode, float lo_thr,
                                  float *max_ov, int32_t *argmax, float *gt_max, uint8_t *lowq, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    PP_REQUIRE(G >= 1 && G < (1 << 30) && A >= 0, "bad sizes (needs at least one ground truth)");
    PP_REQUIRE(iou_mode == PP_NMS_AABB2D || iou_mode == PP_NMS_BOX3D, "iou_mode must be PP_NMS_AABB2D or PP_NMS_BOX3D");
    PP_REQUIRE(gt && gt_max, "null pointer");
    PP_CUDA_TRY(cudaMemsetAsync(gt_max, 0, (size_t)G * sizeof(float), st));
    prof_mark("memset");
    if (A == 0) return PP_OK;
    PP_REQUIRE(anchors && max_ov && argmax && lowq, "null pointer");
    PP_REQUIRE(iou_mode != PP_NMS_AABB2D || (((uintptr_t)gt | (uintptr_t)anchors) % 16 == 0), "rectangles must be 16-byte aligned");
    const unsigned grid = (unsigned)ceil_div(A, AS_THREADS);
    if (iou_mode == PP_NMS_AABB2D) {
        assign_kernel<PP_NMS_AABB2D, false><<<grid, AS_THREADS, 0, st>>>(gt, (int)G, anchors, A, lo_thr, max_ov, argmax, gt_max, lowq);
        if (int rc = check_launch("assign_kernel")) return rc;
        assign_kernel<PP_NMS_AABB2D, true><<<grid, AS_THREADS, 0, st>>>(gt, (int)G, anchors, A, lo_thr, max_ov, argmax, gt_max, lowq);
    } else {
        assign_kernel<PP_NMS_BOX3D, false><<<grid, AS_THREADS, 0, st>>>(gt, (int)G, anchors, A, lo_thr, max_ov, argmax, gt_max, lowq);
        if (int rc = check_launch("assign_kernel")) return rc;
        assign_kernel<PP_NMS_BOX3D, true><<<grid, AS_THREADS, 0, st>>>(gt, (int)G, anchors, A, lo_thr, max_ov, argmax, gt_max, lowq);
    }
    return check_launch("assign_kernel");
}

static int fill_anchor_spec(AnchorSpec &a, const float *range6_host, const float *sizes_host, int S, const float *rots_host,
                            int R, int D, int H, int W)
{
    PP_REQUIRE(range6_host && sizes_host && rots_host, "null pointer");
    PP_REQUIRE(S > 0 && S <= 16 && R > 0 && R <= 16, "1..16 sizes / rotations supported");
    PP_REQUIRE(D > 0 && H > 0 && W > 0, "bad feature map");
    for (int i = 0; i < 6; ++i) a.range[i] = range6_host[i];
    for (int i = 0; i < S * 3; ++i) a.sizes[i] = sizes_host[i];
    for (int i = 0; i < R * 3; ++i) a.rots[i] = rots_host[i];
    a.S = S; a.R = R; a.D = D; a.H = H; a.W = W;
    return PP_OK;
}

extern "C" int pp_head_max_scores(const float *cls, int anchors_per_pos, int ncls, int H, int W, float *max_scores,
                                  pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(cls && max_scores, "null pointer");
    PP_REQUIRE(anchors_per_pos > 0 && ncls > 0 && H > 0 && W > 0, "bad shape");
    const int64_t HW = (int64_t)H * W;
    head_max_scores_kernel<<<GRID1(HW)>>>(cls, anchors_per_pos, ncls, HW, max_scores);
    return check_launch("head_max_scores_kernel");
}

extern "C" int pp_head_select_decode(const float *cls, const float *reg, const float *dirs, const int64_t *rows, int64_t K,
                                     const float *range6_host, const float *sizes_host, int S, const float *rots_host,
                                     int R, int ncls, int H, int W, float *boxes, float *scores, int32_t *dir_bits,
                                     pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(K >= 0 && ncls > 0, "bad shape");
    AnchorSpec a;
    if (int rc = fill_anchor_spec(a, range6_host, sizes_host, S, rots_host, R, 1, H, W)) return rc;
    if (K == 0) return PP_OK;
    PP_REQUIRE(cls && reg && dirs && boxes && scores && dir_bits, "null pointer");
    head_select_decode_kernel<<<(unsigned)ceil_div(K, 128), 128, 0, (cudaStream_t)stream>>>(cls, reg, dirs, rows, K, a, ncls,
                                                                                          boxes, scores, dir_bits);
    return check_launch("head_select_decode_kernel");
}

extern "C" int pp_head_direction_fixup(float *boxes, const int32_t *dir_bits, int64_t K, float dir_offset, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(K >= 0, "K < 0");
    if (K == 0) return PP_OK;
    PP_REQUIRE(boxes && dir_bits, "null pointer");
    head_fixup_kernel<<<GRID1(K * 3)>>>(boxes, dir_bits, K, dir_offset);
    return check_launch("head_fixup_kernel");
}
