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

// Area of the intersection of two rotated rectangles: Sutherland-Hodgman clipping of a's corners against the four
// half-planes of b, evaluated in b's frame (where b is axis-aligned), then the shoelace formula.  FP32 CUDA cores.
__device__ __forceinline__ float rrect_inter_area(const RRect &a, const RRect &b)
{
    float px[8], py[8], qx[8], qy[8];
    const float sx[4] = {-1.f, 1.f, 1.f, -1.f}, sy[4] = {-1.f, -1.f, 1.f, 1.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float wx = a.cx + a.c * sx[i] * a.hx - a.s * sy[i] * a.hy - b.cx;
        const float wy = a.cy + a.s * sx[i] * a.hx + a.c * sy[i] * a.hy - b.cy;
        px[i] = b.c * wx + b.s * wy;
        py[i] = -b.s * wx + b.c * wy;
    }
    int n = 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool yaxis = e >= 2;
        const float sgn = (e & 1) ? -1.f : 1.f, lim = yaxis ? b.hy : b.hx;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const int j = (i + 1 == n) ? 0 : i + 1;
            const float ci = sgn * (yaxis ? py[i] : px[i]) - lim, cj = sgn * (yaxis ? py[j] : px[j]) - lim;
            if (ci <= 0.f) { qx[m] = px[i]; qy[m] = py[i]; ++m; }
            if ((ci < 0.f && cj > 0.f) || (ci > 0.f && cj < 0.f)) {
                const float t = ci / (ci - cj);
                qx[m] = px[i] + t * (px[j] - px[i]);
                qy[m] = py[i] + t * (py[j] - py[i]);
                ++m;
            }
        }
        n = m;
        for (int i = 0; i < n; ++i) { px[i] = qx[i]; py[i] = qy[i]; }
    }
    float area = 0.f;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        area += px[i] * py[j] - px[j] * py[i];
    }
    return 0.5f * fabsf(area);
}

// Symmetric by construction: the pair is put into a canonical order before clipping, so iou(a,b) == iou(b,a)
// bit for bit (the NMS decision must not depend on which box is "selected" and which "remaining").
__device__ __forceinline__ float rrect_iou(const RRect &a, const RRect &b)
{
    const bool swap = (a.cx > b.cx) || (a.cx == b.cx && (a.cy > b.cy || (a.cy == b.cy && (a.hx > b.hx ||
                      (a.hx == b.hx && (a.hy > b.hy || (a.hy == b.hy && a.s > b.s)))))));
    const float inter = swap ? rrect_inter_area(b, a) : rrect_inter_area(a, b);
    const float aa = 4.f * a.hx * a.hy, ab = 4.f * b.hx * b.hy;
    const float uni = (swap ? ab + aa : aa + ab) - inter;
    return inter / fmaxf(uni, 1e-6f);
}

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

constexpr int B3_MAXV = 12;
// keep lo <= g.q + off <= hi (Sutherland-Hodgman, two passes)
__device__ __forceinline__ int b3_clip(float (*p)[3], int n, const float g[3], float off, float lo, float hi)
{
    float q[B3_MAXV][3], s[B3_MAXV];
#pragma unroll 1
    for (int pass = 0; pass < 2 && n > 0; ++pass) {
        const float sgn = pass ? -1.f : 1.f, lim = pass ? -hi : lo;
        for (int i = 0; i < n; ++i) s[i] = sgn * (g[0] * p[i][0] + g[1] * p[i][1] + g[2] * p[i][2] + off) - lim;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const int j = (i + 1 == n) ? 0 : i + 1;
            // (a convex polygon clipped by a slab gains at most two vertices; the bound check is for inputs whose signs
            // alternate through rounding on near-degenerate, co-planar faces)
            if (s[i] >= 0.f && m < B3_MAXV) { q[m][0] = p[i][0]; q[m][1] = p[i][1]; q[m][2] = p[i][2]; ++m; }
            if (((s[i] > 0.f && s[j] < 0.f) || (s[i] < 0.f && s[j] > 0.f)) && m < B3_MAXV) {
                const float t = s[i] / (s[i] - s[j]);
                q[m][0] = p[i][0] + t * (p[j][0] - p[i][0]);
                q[m][1] = p[i][1] + t * (p[j][1] - p[i][1]);
                q[m][2] = p[i][2] + t * (p[j][2] - p[i][2]);
                ++m;
            }
        }
        n = m;
        for (int i = 0; i < n; ++i) { p[i][0] = q[i][0]; p[i][1] = q[i][1]; p[i][2] = q[i][2]; }
    }
    return n;
}
// 6 x signed volume of the cone from the origin over the polygon
__device__ __forceinline__ float b3_cone(float (*p)[3], int n)
{
    float v = 0.f;
    for (int i = 1; i + 1 < n; ++i) v += det3(p[0], p[i], p[i + 1]);
    return v;
}
// face f = 2 * axis + side of the parallelepiped (o; e), outward orientation for a right-handed basis
__device__ __forceinline__ void b3_face(const float o[3], const float e[3][3], int f, float (*p)[3])
{
    const int ax = f >> 1, side = f & 1, a = (ax + 1) % 3, b = (ax + 2) % 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = side ? i : 3 - i;
        const float ua = (k == 1 || k == 2) ? 1.f : 0.f, ub = (k >= 2) ? 1.f : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) p[i][c] = o[c] + (side ? e[ax][c] : 0.f) + ua * e[a][c] + ub * e[b][c];
    }
}

// Volume of a ∩ b: a is mapped into b's unit-cube frame (cube centre at the origin); the volume is the sum of the
// cones over a's faces clipped to the cube (closed slabs) and the cube's faces clipped to a (open slabs).
__device__ __forceinline__ float box3_inter_volume(const Box3 &a, const Box3 &b, float &va, float &vb)
{
    const float eps = 1e-6f;
    const float d1 = det3(a.e[0], a.e[1], a.e[2]), d2 = det3(b.e[0], b.e[1], b.e[2]);
    va = fabsf(d1); vb = fabsf(d2);
    if (!(va > 0.f) || !(vb > 0.f)) return 0.f;
    float g2[3][3];
    dual3(b.e, d2, g2);
    float po[3], pe[3][3];
    const float dx = a.o[0] - b.o[0], dy = a.o[1] - b.o[1], dz = a.o[2] - b.o[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        po[r] = g2[r][0] * dx + g2[r][1] * dy + g2[r][2] * dz - 0.5f;
#pragma unroll
        for (int k = 0; k < 3; ++k) pe[k][r] = g2[r][0] * a.e[k][0] + g2[r][1] * a.e[k][1] + g2[r][2] * a.e[k][2];
    }
    // quick reject: the image of a lies entirely beyond one side of the cube
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float lo = po[r] + fminf(pe[0][r], 0.f) + fminf(pe[1][r], 0.f) + fminf(pe[2][r], 0.f);
        const float hi = po[r] + fmaxf(pe[0][r], 0.f) + fmaxf(pe[1][r], 0.f) + fmaxf(pe[2][r], 0.f);
        if (lo >= 0.5f || hi <= -0.5f) return 0.f;
    }
    const float dp = det3(pe[0], pe[1], pe[2]);
    float gp[3][3];
    dual3(pe, dp, gp);
    float sum_p = 0.f, sum_c = 0.f, poly[B3_MAXV][3];
    const float ax[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
    // Per face: the slab coordinates of its four vertices decide most cases without clipping -- all four beyond the
    // same side of a slab: the face contributes nothing; all four inside every slab: its cone as it is; only faces
    // that straddle a slab boundary go through Sutherland-Hodgman, and only against the slabs they straddle.
#pragma unroll 1
    for (int f = 0; f < 6; ++f) {
        b3_face(po, pe, f, poly);
        unsigned all_out = ~0u, any_out = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned oc = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                oc |= (poly[i][k] < -0.5f - eps) ? (1u << (2 * k)) : 0u;
                oc |= (poly[i][k] > 0.5f + eps) ? (2u << (2 * k)) : 0u;
            }
            all_out &= oc;
            any_out |= oc;
        }
        if (all_out & 0x3Fu) continue;
        int n = 4;
#pragma unroll 1
        for (int k = 0; k < 3 && n > 2; ++k)
            if ((any_out >> (2 * k)) & 3u) n = b3_clip(poly, n, ax[k], 0.f, -0.5f - eps, 0.5f + eps);
        if (n > 2) sum_p += b3_cone(poly, n);
    }
    const float co[3] = {-0.5f, -0.5f, -0.5f};
    float offk[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) offk[k] = -(gp[k][0] * po[0] + gp[k][1] * po[1] + gp[k][2] * po[2]);
#pragma unroll 1
    for (int f = 0; f < 6; ++f) {
        b3_face(co, ax, f, poly);
        unsigned all_out = ~0u, any_out = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned oc = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float sk = gp[k][0] * poly[i][0] + gp[k][1] * poly[i][1] + gp[k][2] * poly[i][2] + offk[k];
                oc |= (sk < eps) ? (1u << (2 * k)) : 0u;
                oc |= (sk > 1.f - eps) ? (2u << (2 * k)) : 0u;
            }
            all_out &= oc;
            any_out |= oc;
        }
        if (all_out & 0x3Fu) continue;
        int n = 4;
#pragma unroll 1
        for (int k = 0; k < 3 && n > 2; ++k)
            if ((any_out >> (2 * k)) & 3u) n = b3_clip(poly, n, gp[k], offk[k], eps, 1.f - eps);
        if (n > 2) sum_c += b3_cone(poly, n);
    }
    float v = ((dp < 0.f ? -sum_p : sum_p) + sum_c) * (1.f / 6.f) * vb;
    v = fmaxf(v, 0.f);
    return fminf(v, fminf(va, vb));
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
__device__ __forceinline__ float box3_iou(const Box3 &a, const Box3 &b, float *vol_out)
{
    bool swap = false;
#pragma unroll
    for (int k = 2; k >= 0; --k)
        if (a.o[k] != b.o[k]) swap = a.o[k] > b.o[k];
    float va, vb;
    const float v = swap ? box3_inter_volume(b, a, vb, va) : box3_inter_volume(a, b, va, vb);
    if (vol_out) *vol_out = v;
    const float u = (swap ? vb + va : va + vb) - v;
    return u > 0.f ? v / u : 0.f;
}

}  // namespace pp
