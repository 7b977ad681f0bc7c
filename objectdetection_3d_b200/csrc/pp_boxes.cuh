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

}  // namespace pp
