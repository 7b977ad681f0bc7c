// Device-side box math shared by pp_boxes.cu and pp_nms.cu.  Every operation is rounded separately
// (__f*_rn) in the order of the reference's eager torch ops.
#pragma once
#include <cuda_runtime.h>

namespace pp {

// BBoxCoder.decode, model/utils.py:309-337
__device__ __forceinline__ void decode_one(const float a[9], const float t[9], float o[9])
{
    float za = __fadd_rn(a[2], __fdiv_rn(a[5], 2.0f));
    float diag = __fsqrt_rn(__fadd_rn(__fmul_rn(a[3], a[3]), __fmul_rn(a[4], a[4])));
    o[0] = __fadd_rn(__fmul_rn(t[0], diag), a[0]);
    o[1] = __fadd_rn(__fmul_rn(t[1], diag), a[1]);
    o[2] = __fadd_rn(__fmul_rn(t[2], a[5]), za);
    o[3] = __fmul_rn(expf(t[3]), a[3]);
    o[4] = __fmul_rn(expf(t[4]), a[4]);
    o[5] = __fmul_rn(expf(t[5]), a[5]);
    o[6] = __fadd_rn(t[6], a[6]);
    o[7] = __fadd_rn(t[7], a[7]);
    o[8] = __fadd_rn(t[8], a[8]);
}

// torch.linspace (CUDA kernel formula): symmetric evaluation from both ends, model/utils.py:227-239
__device__ __forceinline__ float linspace_at(float start, float end, int steps, int i)
{
    if (steps == 1) return start;
    float step = __fdiv_rn(__fsub_rn(end, start), (float)(steps - 1));
    if (i < steps / 2) return __fadd_rn(start, __fmul_rn(step, (float)i));
    return __fsub_rn(end, __fmul_rn(step, (float)(steps - i - 1)));
}

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2)
{
    return __fadd_rn(__fadd_rn(__fadd_rn(0.f, __fmul_rn(a0, b0)), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// 8 corners of a 9-parameter box: ops/ops_torch.py:160-256.  R = (Rz*Ry)*Rx, pivot = bottom centre.
__device__ __forceinline__ void box_corners(const float b[9], float out[8][3])
{
    const float x = b[0], y = b[1], z = b[2];
    const float hx = __fmul_rn(b[3], 0.5f), hy = __fmul_rn(b[4], 0.5f);
    const float xl = __fsub_rn(x, hx), xh = __fadd_rn(x, hx);
    const float yl = __fsub_rn(y, hy), yh = __fadd_rn(y, hy);
    const float zt = __fadd_rn(z, b[5]);
    const float vx[8] = {xl, xh, xh, xl, xl, xh, xh, xl};
    const float vy[8] = {yl, yl, yh, yh, yl, yl, yh, yh};
    const float vz[8] = {z, z, z, z, zt, zt, zt, zt};
    float sx, cx, sy, cy, sz, cz;
    sincosf(b[6], &sx, &cx);
    sincosf(b[7], &sy, &cy);
    sincosf(b[8], &sz, &cz);
    const float Rx[3][3] = {{1.f, 0.f, 0.f}, {0.f, cx, -sx}, {0.f, sx, cx}};
    const float Ry[3][3] = {{cy, 0.f, sy}, {0.f, 1.f, 0.f}, {-sy, 0.f, cy}};
    const float Rz[3][3] = {{cz, -sz, 0.f}, {sz, cz, 0.f}, {0.f, 0.f, 1.f}};
    float T[3][3], R[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) T[i][j] = dot3(Rz[i][0], Rz[i][1], Rz[i][2], Ry[0][j], Ry[1][j], Ry[2][j]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[i][j] = dot3(T[i][0], T[i][1], T[i][2], Rx[0][j], Rx[1][j], Rx[2][j]);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const float c0 = __fsub_rn(vx[v], x), c1 = __fsub_rn(vy[v], y), c2 = __fsub_rn(vz[v], z);
        out[v][0] = __fadd_rn(dot3(c0, c1, c2, R[0][0], R[0][1], R[0][2]), x);
        out[v][1] = __fadd_rn(dot3(c0, c1, c2, R[1][0], R[1][1], R[1][2]), y);
        out[v][2] = __fadd_rn(dot3(c0, c1, c2, R[2][0], R[2][1], R[2][2]), z);
    }
}

// xy bounding rectangle of the rotated corners: ops/ops_torch.py:111-114
__device__ __forceinline__ float4 corners_to_rect(const float c[8][3])
{
    float x1 = c[0][0], x2 = c[0][0], y1 = c[0][1], y2 = c[0][1];
#pragma unroll
    for (int v = 1; v < 8; ++v) {
        x1 = fminf(x1, c[v][0]); x2 = fmaxf(x2, c[v][0]);
        y1 = fminf(y1, c[v][1]); y2 = fmaxf(y2, c[v][1]);
    }
    return make_float4(x1, y1, x2, y2);
}

// bbox_iou2D for one pair (a from bboxes1, b from bboxes2): ops/ops_torch.py:572-607
__device__ __forceinline__ float rect_iou(const float4 a, const float4 b, int mode, float eps)
{
    const float area1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    w = w < 0.f ? 0.f : w;
    h = h < 0.f ? 0.f : h;
    const float overlap = __fmul_rn(w, h);
    float uni = (mode == 1) ? area1 : __fsub_rn(__fadd_rn(area1, area2), overlap);
    uni = fmaxf(uni, eps);
    const float iou = __fdiv_rn(overlap, uni);
    if (mode != 2) return iou;
    float ew = __fsub_rn(fmaxf(a.z, b.z), fminf(a.x, b.x));
    float eh = __fsub_rn(fmaxf(a.w, b.w), fminf(a.y, b.y));
    ew = ew < 0.f ? 0.f : ew;
    eh = eh < 0.f ? 0.f : eh;
    const float ea = fmaxf(__fmul_rn(ew, eh), eps);
    return __fsub_rn(iou, __fdiv_rn(__fsub_rn(ea, uni), ea));
}

// ---- rotated BEV rectangles (extension: the north star's "rotated BEV IoU"; not in the reference) -----------------
// BEV footprint of a 9-parameter box: centre (x, y), size (dx, dy), yaw rz.
struct RRect {
    float cx, cy, hx, hy, c, s;
};

__device__ __forceinline__ RRect rrect_from_box9(const float *b)
{
    RRect r;
    r.cx = b[0]; r.cy = b[1];
    r.hx = 0.5f * b[3]; r.hy = 0.5f * b[4];
    sincosf(b[8], &r.s, &r.c);
    return r;
}

// axis-aligned bounding rectangle of the footprint (conservative pre-filter for the pair tests)
__device__ __forceinline__ float4 rrect_aabb(const RRect &r)
{
    const float ex = fabsf(r.c) * r.hx + fabsf(r.s) * r.hy, ey = fabsf(r.s) * r.hx + fabsf(r.c) * r.hy;
    return make_float4(r.cx - ex, r.cy - ey, r.cx + ex, r.cy + ey);
}

// (intersection area and IoU of two footprints: rrect_inter_area / rrect_iou at the end of this file)
// RRect <-> (float4, float2) for 16/8-byte loads and stores
__device__ __forceinline__ float4 rrect_lo(const RRect &r) { return make_float4(r.cx, r.cy, r.hx, r.hy); }
__device__ __forceinline__ float2 rrect_hi(const RRect &r) { return make_float2(r.c, r.s); }
__device__ __forceinline__ RRect rrect_pack(const float4 lo, const float2 hi)
{
    RRect r;
    r.cx = lo.x; r.cy = lo.y; r.hx = lo.z; r.hy = lo.w; r.c = hi.x; r.s = hi.y;
    return r;
}

// ---- oriented 3-D boxes (ops/ops_torch.py:692-755 -> pytorch3d _C.iou_box3d; parity unpinned, see oracle) ----------
// A box is the parallelepiped o = v0, e1 = v1 - v0, e2 = v3 - v0, e3 = v4 - v0 of its 8 corners (reference order).
struct Box3 {
    float o[3], e[3][3];
};
// (tests/host/iou_host.cu defines PP_B3_FN as `__host__ __device__ inline` to check this arithmetic against the oracle
// on a machine without a GPU; the library itself is only ever built for the device)
#ifndef PP_B3_FN
#define PP_B3_FN __device__ __forceinline__
// The two IoU entry points are NOT inlined: every kernel of a translation unit then runs the same instructions, so the
// IoU matrix kernels, the assignment kernel and the NMS kernels agree bit for bit (inlined copies may contract
// multiply-adds differently).
#define PP_B3_ENTRY static __device__ __noinline__
#endif
#ifndef PP_B3_ENTRY
#define PP_B3_ENTRY PP_B3_FN
#endif

__device__ __forceinline__ Box3 box3_from_corners(const float *c /* (8,3) */)
{
    Box3 b;
    const int nb[3] = {1, 3, 4};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b.o[k] = c[k];
#pragma unroll
        for (int j = 0; j < 3; ++j) b.e[j][k] = c[nb[j] * 3 + k] - c[k];
    }
    return b;
}
__device__ __forceinline__ Box3 box3_from_corner_array(const float c[8][3])
{
    Box3 b;
    const int nb[3] = {1, 3, 4};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b.o[k] = c[0][k];
#pragma unroll
        for (int j = 0; j < 3; ++j) b.e[j][k] = c[nb[j]][k] - c[0][k];
    }
    return b;
}
// 12 floats <-> 3 float4
__device__ __forceinline__ void box3_store(const Box3 &b, float4 &q0, float4 &q1, float4 &q2)
{
    q0 = make_float4(b.o[0], b.o[1], b.o[2], b.e[0][0]);
    q1 = make_float4(b.e[0][1], b.e[0][2], b.e[1][0], b.e[1][1]);
    q2 = make_float4(b.e[1][2], b.e[2][0], b.e[2][1], b.e[2][2]);
}
__device__ __forceinline__ Box3 box3_load(const float4 q0, const float4 q1, const float4 q2)
{
    Box3 b;
    b.o[0] = q0.x; b.o[1] = q0.y; b.o[2] = q0.z; b.e[0][0] = q0.w;
    b.e[0][1] = q1.x; b.e[0][2] = q1.y; b.e[1][0] = q1.z; b.e[1][1] = q1.w;
    b.e[1][2] = q2.x; b.e[2][0] = q2.y; b.e[2][1] = q2.z; b.e[2][2] = q2.w;
    return b;
}

// xy bounding rectangle of 8 corners stored as (8,3)
__device__ __forceinline__ float4 corners_xy_rect(const float *c)
{
    float x1 = c[0], x2 = c[0], y1 = c[1], y2 = c[1];
#pragma unroll
    for (int v = 1; v < 8; ++v) {
        x1 = fminf(x1, c[v * 3]); x2 = fmaxf(x2, c[v * 3]);
        y1 = fminf(y1, c[v * 3 + 1]); y2 = fmaxf(y2, c[v * 3 + 1]);
    }
    return make_float4(x1, y1, x2, y2);
}

__device__ __forceinline__ float det3(const float a[3], const float b[3], const float c[3])
{
    return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
}
// rows of the inverse of the matrix whose columns are e[0], e[1], e[2]
__device__ __forceinline__ void dual3(const float e[3][3], float det, float g[3][3])
{
    const float r = 1.f / det;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float *a = e[(k + 1) % 3], *b = e[(k + 2) % 3];
        g[k][0] = (a[1] * b[2] - a[2] * b[1]) * r;
        g[k][1] = (a[2] * b[0] - a[0] * b[2]) * r;
        g[k][2] = (a[0] * b[1] - a[1] * b[0]) * r;
    }
}

// ---- exact intersection volume, float64, branch-free ------------------------------------------------------------
// a is mapped into b's unit-cube frame (cube centre at the origin).  By the divergence theorem the volume is
//     vol(b) / 3 * ( sum over a's faces  sigma_f det(O_f, E1_f, E2_f) A_f  +  1/2 * sum over the cube's faces A'_f )
// where A_f is the AREA, in the face's own (u, v) in [0,1]^2 parameters, of the part of the face inside the other box:
// the unit square cut by three strips  lo <= c0[k] + cu[k] u + cv[k] v <= hi  (one per axis of the other box).  That
// area is Green's integral of u dv over the boundary, edge by edge: the edge u = 1 and the six strip lines, each cut
// to the interval the other constraints leave of it.  No polygon is ever built, so there are no loops over vertices,
// no local arrays and no divergence -- but the two edges that meet in a vertex compute it separately, and the two
// results must agree far below the size of the face for the contributions to cancel; lines that cross at a shallow
// angle amplify rounding by 1/sin, so the evaluation is in float64 (half the fp32 issue rate on B200, measured:
// scripts/micro/fp64_bench.cu), which also makes the result exact to fp32 output precision.
// Faces of a are cut with closed strips (+eps), faces of the cube with open ones (-eps): coincident faces count once.
constexpr double B3_EPS = 1e-9;

// 1 / x by two Newton steps on the fp32 reciprocal; +-inf for |x| < 1e-30 (lines taken as parallel)
PP_B3_FN bool b3_tiny(double x) { return fabsf((float)x) < 1e-30f; }
PP_B3_FN double b3_rcp(double x)
{
    const float xf = (float)x;
#ifdef __CUDA_ARCH__
    const float rf = __frcp_rn(xf);
#else
    const float rf = 1.0f / xf;
#endif
    double r = (double)rf;
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    const double inf = x < 0.0 ? -(double)INFINITY : (double)INFINITY;
    return b3_tiny(x) ? inf : r;
}
// [tmin, tmax] &= { t : lo <= f0 + t / r <= hi }   (r = 1 / slope, +-inf when the slope is 0)
PP_B3_FN void b3_cut(double &tmin, double &tmax, double f0, double r, bool pos, double lo, double hi)
{
    const double t1 = (lo - f0) * r, t2 = (hi - f0) * r;
    const double a = pos ? t1 : t2, b = pos ? t2 : t1;
    tmin = a > tmin ? a : tmin;          // (a NaN -- a bound that coincides with a parallel line -- leaves the interval as it is)
    tmax = b < tmax ? b : tmax;
}
// what the three strips share between the two opposite faces of one axis
struct B3Lines {
    double cu[3], cv[3];      // strip k: c0[k] + cu[k] u + cv[k] v
    double rcu[3], rcv[3];    // 1 / cu, 1 / cv
    double rnn[3];            // 1 / (cu^2 + cv^2); 0 for a strip that does not depend on (u, v)
    double rx[3];             // 1 / (cu[j] cv[k] - cv[j] cu[k]) for (j, k) = (1, 2), (2, 0), (0, 1)
};
PP_B3_FN void b3_lines(B3Lines &L)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        L.rcu[k] = b3_rcp(L.cu[k]);
        L.rcv[k] = b3_rcp(L.cv[k]);
        const double nn = L.cu[k] * L.cu[k] + L.cv[k] * L.cv[k];
        const double r = b3_rcp(nn);
        L.rnn[k] = b3_tiny(nn) ? 0.0 : r;
        const int j = (k + 1) % 3, m = (k + 2) % 3;
        L.rx[k] = b3_rcp(L.cu[j] * L.cv[m] - L.cv[j] * L.cu[m]);
    }
}
// area of { (u, v) in [0,1]^2 : lo <= c0[k] + cu[k] u + cv[k] v <= hi, k = 0..2 }
PP_B3_FN double b3_area(const B3Lines &L, const double c0[3], const double lo[3], const double hi[3])
{
    // edge u = 1, parameter v
    double tmin = 0.0, tmax = 1.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) b3_cut(tmin, tmax, c0[k] + L.cu[k], L.rcv[k], L.rcv[k] >= 0.0, lo[k], hi[k]);
    double area = tmax > tmin ? tmax - tmin : 0.0;
    // strip lines: P + t D with D = (cv, -cu); the interior is on the left of D for the lo line, on the right for hi
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int j = (k + 1) % 3, m = (k + 2) % 3;
        const double du = L.cv[k], dv = -L.cu[k];
        // reciprocal slopes of the other constraints along D: u: 1 / du, v: 1 / dv, strip j: 1 / (cu[j] du + cv[j] dv)
        const double r_u = L.rcv[k], r_v = -L.rcu[k];
        const double r_j = -L.rx[m], r_m = L.rx[j];       // 1 / (cu[j] cv[k] - cv[j] cu[k]),  1 / (cu[m] cv[k] - cv[m] cu[k])
        const bool p_u = r_u >= 0.0, p_v = r_v >= 0.0, p_j = r_j >= 0.0, p_m = r_m >= 0.0;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const double sc = ((side ? hi[k] : lo[k]) - c0[k]) * L.rnn[k];
            const double pu = L.cu[k] * sc, pv = L.cv[k] * sc;
            double t0 = -1e300, t1 = 1e300;
            b3_cut(t0, t1, pu, r_u, p_u, 0.0, 1.0);
            b3_cut(t0, t1, pv, r_v, p_v, 0.0, 1.0);
            b3_cut(t0, t1, c0[j] + L.cu[j] * pu + L.cv[j] * pv, r_j, p_j, lo[j], hi[j]);
            b3_cut(t0, t1, c0[m] + L.cu[m] * pu + L.cv[m] * pv, r_m, p_m, lo[m], hi[m]);
            const double c = 0.5 * (2.0 * pu + (t0 + t1) * du) * ((t1 - t0) * dv);
            const bool ok = t1 > t0 && L.rnn[k] != 0.0;
            area += ok ? (side ? -c : c) : 0.0;
        }
    }
    return area;
}

PP_B3_FN double det3d(const double a[3], const double b[3], const double c[3])
{
    return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
}
// rows of the inverse of the matrix whose columns are e[0], e[1], e[2]
PP_B3_FN void dual3d(const double e[3][3], double det, double g[3][3])
{
    const double r = 1.0 / det;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double *a = e[(k + 1) % 3], *b = e[(k + 2) % 3];
        g[k][0] = (a[1] * b[2] - a[2] * b[1]) * r;
        g[k][1] = (a[2] * b[0] - a[0] * b[2]) * r;
        g[k][2] = (a[0] * b[1] - a[1] * b[0]) * r;
    }
}

PP_B3_FN double box3_inter_volume(const Box3 &a, const Box3 &b, double &va, double &vb)
{
    double ae[3][3], be[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) { ae[k][c] = (double)a.e[k][c]; be[k][c] = (double)b.e[k][c]; }
    const double d1 = det3d(ae[0], ae[1], ae[2]), d2 = det3d(be[0], be[1], be[2]);
    va = fabs(d1); vb = fabs(d2);
    if (!(va > 0.0) || !(vb > 0.0)) return 0.0;
    double g2[3][3];
    dual3d(be, d2, g2);
    // org[0] / mat[0]: a in b's frame (origin po, edges pe[k]); org[1] / mat[1]: the cube in a's frame (the image of
    // the corner (-1/2,-1/2,-1/2) and of the unit steps along b's axes) -- the same shape, so one loop body serves both
    double org[2][3], mat[2][3][3];
    const double dx = (double)a.o[0] - (double)b.o[0], dy = (double)a.o[1] - (double)b.o[1], dz = (double)a.o[2] - (double)b.o[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        org[0][r] = g2[r][0] * dx + g2[r][1] * dy + g2[r][2] * dz - 0.5;
#pragma unroll
        for (int k = 0; k < 3; ++k) mat[0][k][r] = g2[r][0] * ae[k][0] + g2[r][1] * ae[k][1] + g2[r][2] * ae[k][2];
    }
    // quick reject: the image of a lies entirely beyond one side of the cube
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double lo = org[0][r] + fmin(mat[0][0][r], 0.0) + fmin(mat[0][1][r], 0.0) + fmin(mat[0][2][r], 0.0);
        const double hi = org[0][r] + fmax(mat[0][0][r], 0.0) + fmax(mat[0][1][r], 0.0) + fmax(mat[0][2][r], 0.0);
        if (lo >= 0.5 || hi <= -0.5) return 0.0;
    }
    const double dp = det3d(mat[0][0], mat[0][1], mat[0][2]);
    double gp[3][3];
    dual3d(mat[0], dp, gp);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double off = -(gp[k][0] * org[0][0] + gp[k][1] * org[0][1] + gp[k][2] * org[0][2]);
        org[1][k] = off - 0.5 * (gp[k][0] + gp[k][1] + gp[k][2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) mat[1][c][k] = gp[k][c];
    }
    const double sg = dp < 0.0 ? -1.0 : 1.0;
    // The widening of the closed strips (b's frame) and the narrowing of the open ones (a's frame) are the SAME length
    // in space, B3_EPS times the longest edge of the pair, expressed in each axis' own unit: two nearly coplanar faces
    // -- coplanar up to the fp32 rounding of the corners, i.e. tilted against each other by 1e-8 -- are then told
    // apart by the same distance from both sides, and every point of their common plane is counted exactly once.
    double len2 = 0.0, la[3], lb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        la[k] = ae[k][0] * ae[k][0] + ae[k][1] * ae[k][1] + ae[k][2] * ae[k][2];
        lb[k] = be[k][0] * be[k][0] + be[k][1] * be[k][1] + be[k][2] * be[k][2];
        len2 = fmax(len2, fmax(la[k], lb[k]));
    }
    double lo[2][3], hi[2][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double eb = B3_EPS * sqrt(len2 / lb[k]), ea = B3_EPS * sqrt(len2 / la[k]);
        lo[0][k] = -0.5 - eb; hi[0][k] = 0.5 + eb;
        lo[1][k] = ea;        hi[1][k] = 1.0 - ea;
    }
    double sum = 0.0;
#pragma unroll 1
    for (int task = 0; task < 6; ++task) {
        const int which = task >= 3 ? 1 : 0, ax = task - 3 * which;
        const int ia = ax == 2 ? 0 : ax + 1, ib = ax == 0 ? 2 : ax - 1;
        B3Lines L;
        double c0a[3], c0b[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            c0a[k] = org[which][k];
            c0b[k] = org[which][k] + mat[which][ax][k];
            L.cu[k] = mat[which][ia][k];
            L.cv[k] = mat[which][ib][k];
        }
        b3_lines(L);
        const double a0 = b3_area(L, c0a, lo[which], hi[which]), a1 = b3_area(L, c0b, lo[which], hi[which]);
        // cone weights: a's faces sigma * det(O, E1, E2) = -h0 (low side), h0 + dp (high side), times sign(dp);
        // the cube's faces are at distance 1/2 from the apex
        const double h0 = det3d(c0a, mat[0][ia], mat[0][ib]);
        sum += which ? 0.5 * (a0 + a1) : sg * ((h0 + dp) * a1 - h0 * a0);
    }
    double v = sum * (1.0 / 3.0) * vb;
    v = fmax(v, 0.0);
    return fmin(v, fmin(va, vb));
}

// Image of a in b's unit frame, axis by axis: false when an axis of b separates the boxes; else ub = vol(b) x the
// product of the overlap lengths, an upper bound of the intersection volume (the intersection lies inside b and
// inside the b-aligned bounding box of a).
__device__ __forceinline__ bool box3_proj_bound(const Box3 &a, const Box3 &b, float &ub)
{
    const float d2 = det3(b.e[0], b.e[1], b.e[2]);
    if (d2 == 0.f) { ub = 0.f; return false; }
    float g2[3][3];
    dual3(b.e, d2, g2);
    const float dx = a.o[0] - b.o[0], dy = a.o[1] - b.o[1], dz = a.o[2] - b.o[2];
    float prod = fabsf(d2);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float po = g2[r][0] * dx + g2[r][1] * dy + g2[r][2] * dz;
        float lo = po, hi = po;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float pe = g2[r][0] * a.e[k][0] + g2[r][1] * a.e[k][1] + g2[r][2] * a.e[k][2];
            lo += fminf(pe, 0.f);
            hi += fmaxf(pe, 0.f);
        }
        const float len = fminf(hi, 1.f) - fmaxf(lo, 0.f);
        if (!(len > 0.f)) { ub = 0.f; return false; }
        prod *= len;
    }
    ub = prod;
    return true;
}

// Symmetric by construction (canonical argument order), like rrect_iou.
PP_B3_ENTRY float box3_iou(const Box3 &a, const Box3 &b, float *vol_out)
{
    bool swap = false;
#pragma unroll
    for (int k = 2; k >= 0; --k)
        if (a.o[k] != b.o[k]) swap = a.o[k] > b.o[k];
    const Box3 &p = swap ? b : a, &q = swap ? a : b;      // one instance of the volume code
    double vp, vq;
    const double v = box3_inter_volume(p, q, vp, vq);
    if (vol_out) *vol_out = (float)v;
    const double u = vp + vq - v;
    return u > 0.0 ? (float)(v / u) : 0.f;
}

// ---- rotated BEV rectangles: intersection area, float64, branch-free (r2) -----------------------------------------
// The 2-D case of the face-area routine above: in a's own (u, v) in [0,1]^2 parameters the part of a inside b is the unit
// square cut by two strips (b's two axes), so area(a n b) = area(a) * that area.  Strips are closed and widened by eps
// so that an edge of a lying on an edge of b is counted by the square's edge only.
PP_B3_FN double strips2_area(const double c0[2], const double cu[2], const double cv[2], const double lo[2], const double hi[2])
{
    double rcu[2], rcv[2], rnn[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        rcu[k] = b3_rcp(cu[k]);
        rcv[k] = b3_rcp(cv[k]);
        const double nn = cu[k] * cu[k] + cv[k] * cv[k];
        const double r = b3_rcp(nn);
        rnn[k] = b3_tiny(nn) ? 0.0 : r;
    }
    const double rx = b3_rcp(cu[0] * cv[1] - cv[0] * cu[1]);
    double tmin = 0.0, tmax = 1.0;
#pragma unroll
    for (int k = 0; k < 2; ++k) b3_cut(tmin, tmax, c0[k] + cu[k], rcv[k], rcv[k] >= 0.0, lo[k], hi[k]);
    double area = tmax > tmin ? tmax - tmin : 0.0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int j = 1 - k;
        const double du = cv[k], dv = -cu[k];
        const double r_u = rcv[k], r_v = -rcu[k], r_j = k ? rx : -rx;     // 1 / (cu[j] du + cv[j] dv)
        const bool p_u = r_u >= 0.0, p_v = r_v >= 0.0, p_j = r_j >= 0.0;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const double sc = ((side ? hi[k] : lo[k]) - c0[k]) * rnn[k];
            const double pu = cu[k] * sc, pv = cv[k] * sc;
            double t0 = -1e300, t1 = 1e300;
            b3_cut(t0, t1, pu, r_u, p_u, 0.0, 1.0);
            b3_cut(t0, t1, pv, r_v, p_v, 0.0, 1.0);
            b3_cut(t0, t1, c0[j] + cu[j] * pu + cv[j] * pv, r_j, p_j, lo[j], hi[j]);
            const double c = 0.5 * (2.0 * pu + (t0 + t1) * du) * ((t1 - t0) * dv);
            const bool ok = t1 > t0 && rnn[k] != 0.0;
            area += ok ? (side ? -c : c) : 0.0;
        }
    }
    return area;
}

PP_B3_FN double rrect_inter_area(const RRect &a, const RRect &b)
{
    const double hxa = a.hx, hya = a.hy, hxb = b.hx, hyb = b.hy;
    if (!(hxa > 0.0) || !(hya > 0.0) || !(hxb > 0.0) || !(hyb > 0.0)) return 0.0;
    const double ac = a.c, as = a.s, bc = b.c, bs = b.s;
    const double dx = (double)a.cx - (double)b.cx, dy = (double)a.cy - (double)b.cy;
    // a's axes in b's frame: ua = (cc, cs), va = (-cs, cc)
    const double cc = ac * bc + as * bs, cs = as * bc - ac * bs;
    const double ix = 0.5 / hxb, iy = 0.5 / hyb;
    const double ox = dx * bc + dy * bs, oy = dy * bc - dx * bs;       // a's centre in b's frame
    double c0[2], cu[2], cv[2];
    cu[0] = 2.0 * hxa * cc * ix;  cv[0] = -2.0 * hya * cs * ix;  c0[0] = (ox - hxa * cc + hya * cs) * ix + 0.5;
    cu[1] = 2.0 * hxa * cs * iy;  cv[1] = 2.0 * hya * cc * iy;   c0[1] = (oy - hxa * cs - hya * cc) * iy + 0.5;
    const double lo[2] = {-B3_EPS, -B3_EPS}, hi[2] = {1.0 + B3_EPS, 1.0 + B3_EPS};
    const double area = strips2_area(c0, cu, cv, lo, hi);
    return 4.0 * hxa * hya * fmin(fmax(area, 0.0), 1.0);
}

// Symmetric by construction: the pair is put into a canonical order first, so iou(a,b) == iou(b,a) bit for bit (the
// NMS decision must not depend on which box is "selected" and which "remaining").
PP_B3_ENTRY float rrect_iou(const RRect &a, const RRect &b)
{
    const bool swap = (a.cx > b.cx) || (a.cx == b.cx && (a.cy > b.cy || (a.cy == b.cy && (a.hx > b.hx ||
                      (a.hx == b.hx && (a.hy > b.hy || (a.hy == b.hy && a.s > b.s)))))));
    const RRect &p = swap ? b : a, &q = swap ? a : b;
    const double inter = rrect_inter_area(p, q);
    const double uni = 4.0 * (double)p.hx * (double)p.hy + 4.0 * (double)q.hx * (double)q.hy - inter;
    return (float)(inter / fmax(uni, 1e-6));
}

}  // namespace pp
