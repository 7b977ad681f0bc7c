// One pillar of the fused single-layer PillarFeatureNet (model/PointPillars.py:480-526), one warp, P <= 32, C <= 7,
// U <= 64: shared by the stand-alone kernel (pp_pillar.cu) and by the voxelizer's gather kernel (pp_voxelize.cu), which
// runs it on the pillar it has just gathered, so the two paths give bit-identical features.
//
// The nine decorated channels [p, p_xyz - mean, p_xy - centre] are an affine function of the C raw ones, so the
// Linear(9 -> U) + BatchNorm(eval) collapses to C multiply-adds per point and unit plus one bias per pillar and unit:
//     W . [x y z r | x-mx y-my z-mz | x-cx y-cy]
//       = (W0+W4+W7)(x-cx) + (W1+W5+W8)(y-cy) + (W2+W6)(z-mz) + W3 r
//         + [ W0 cx + W1 cy + W2 mz + W4 (cx-mx) + W5 (cy-my) ]
// The points are recentred on the pillar centre (xy) and on the pillar mean (z) so that every product is as small as
// the reference's own; the large terms (W0 cx ...) appear once per pillar, exactly as large as the reference's W0 x.
// The BatchNorm scale is folded into the weights, the shift into the bias.  T1 (1e-5 relative) against the reference;
// the bit-exact sequential decoration is decorate_row in pp_pillar.cu (pp_decorate and the generic layer kernel).
// lane = slot while recentring (the points never leave registers until the row is staged for the broadcast reads),
// lane = unit pair (lane, lane + 32) while multiplying: no cross-lane reduction for the max over the points.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace pp {

template <int C> struct PfnWeightsC {
    float a0[C], a1[C];      // scale * folded weight of raw channel k, units lane and lane + 32
    float x0[5], x1[5];      // scale * (W0, W1, W2, W[C+0], W[C+1]): the per-pillar bias terms
    float sh0, sh1;
};
typedef PfnWeightsC<4> PfnWeights;

// W is (U, C + 5) row-major; units >= U get zeros (their outputs are never stored)
template <int C>
__device__ __forceinline__ void pfn_load_weights_c(PfnWeightsC<C> &pw, const float *__restrict__ W,
                                                   const float *__restrict__ scale, const float *__restrict__ shift,
                                                   int U, int lane)
{
    constexpr int CIN = C + 5;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int u = lane + 32 * h;
        const bool on = u < U;
        const float *w = W + (size_t)(on ? u : 0) * CIN;
        const float sc = on ? scale[u] : 0.f, sh = on ? shift[u] : 0.f;
        float *a = h ? pw.a1 : pw.a0, *x = h ? pw.x1 : pw.x0;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            float s = w[k];
            if (k < 3) s += w[C + k];
            if (k < 2) s += w[C + 3 + k];
            a[k] = on ? s * sc : 0.f;
        }
        x[0] = on ? w[0] * sc : 0.f;
        x[1] = on ? w[1] * sc : 0.f;
        x[2] = on ? w[2] * sc : 0.f;
        x[3] = on ? w[C + 0] * sc : 0.f;
        x[4] = on ? w[C + 1] * sc : 0.f;
        if (h) pw.sh1 = sh; else pw.sh0 = sh;
    }
}

__device__ __forceinline__ void pfn_load_weights(PfnWeights &pw, const float *__restrict__ W, const float *__restrict__ scale,
                                                 const float *__restrict__ shift, int U, int /*C == 4*/, int lane)
{
    pfn_load_weights_c<4>(pw, W, scale, shift, U, lane);
}

// f[0 .. C) = this lane's point (zeros for an empty slot or lane >= P); row = this warp's 32 x LD staging area
// (LD = 4 for C <= 4, else 8); out = the pillar's (U + 1) feature row.  All 32 lanes of the warp must call it.
template <int C>
__device__ __forceinline__ void pfn_pillar_c(const PfnWeightsC<C> &pw, const float (&f)[C], int P, int n, int cx, int cy,
                                             float vx, float vy, float x_off, float y_off, float *row,
                                             float *__restrict__ out, float *__restrict__ canvas_cell, int64_t plane,
                                             int U, int lane)
{
    constexpr int LD = C <= 4 ? 4 : 8;
    float sx = f[0], sy = f[1], sz = f[2];                  // zero padded slots add nothing (:493-494)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xFFFFFFFFu, sx, o);
        sy += __shfl_xor_sync(0xFFFFFFFFu, sy, o);
        sz += __shfl_xor_sync(0xFFFFFFFFu, sz, o);
    }
    const float nf = (float)n;
    const float mx_ = __fdiv_rn(sx, nf), my_ = __fdiv_rn(sy, nf), mz_ = __fdiv_rn(sz, nf);
    const float pcx = __fadd_rn(__fmul_rn((float)cx, vx), x_off);   // :500-503
    const float pcy = __fadd_rn(__fmul_rn((float)cy, vy), y_off);   // :505-508
    float d[LD];
#pragma unroll
    for (int k = 0; k < LD; ++k) d[k] = k < C ? f[k] : 0.f;
    d[0] = __fsub_rn(f[0], pcx);
    d[1] = __fsub_rn(f[1], pcy);
    d[2] = __fsub_rn(f[2], mz_);
    // slots >= n are never read below (p_end), so the padding mask (:518-521) needs no multiply
    float4 *r4 = reinterpret_cast<float4 *>(row + lane * LD);
    r4[0] = make_float4(d[0], d[1], d[2], d[3]);
    if (LD == 8) r4[1] = make_float4(d[4 % LD], d[5 % LD], d[6 % LD], d[7 % LD]);
    __syncwarp();
    // ---- per-pillar bias of the two units this lane owns
    const float ex = __fsub_rn(pcx, mx_), ey = __fsub_rn(pcy, my_);
    float b0 = pw.sh0, b1 = pw.sh1;
    b0 = __fmaf_rn(pw.x0[0], pcx, b0); b1 = __fmaf_rn(pw.x1[0], pcx, b1);
    b0 = __fmaf_rn(pw.x0[1], pcy, b0); b1 = __fmaf_rn(pw.x1[1], pcy, b1);
    b0 = __fmaf_rn(pw.x0[2], mz_, b0); b1 = __fmaf_rn(pw.x1[2], mz_, b1);
    b0 = __fmaf_rn(pw.x0[3], ex, b0);  b1 = __fmaf_rn(pw.x1[3], ex, b1);
    b0 = __fmaf_rn(pw.x0[4], ey, b0);  b1 = __fmaf_rn(pw.x1[4], ey, b1);
    // ---- multiply: lane = units (lane, lane + 32)
    const int p_end = n < P ? n : P;
    // zero-padded slots take part in the max (:403-410): they contribute relu(shift).  relu(y) >= 0, so starting
    // the running max at 0 (or relu(shift)) makes the explicit relu redundant.
    float mx0 = (p_end < P) ? fmaxf(pw.sh0, 0.f) : 0.f;
    float mx1 = (p_end < P) ? fmaxf(pw.sh1, 0.f) : 0.f;
#pragma unroll 4
    for (int p = 0; p < p_end; ++p) {
        const float4 *f4 = reinterpret_cast<const float4 *>(row + p * LD);
        const float4 fa = f4[0];
        float g[LD];
        g[0] = fa.x; g[1] = fa.y; g[2] = fa.z; g[3] = fa.w;
        if (LD == 8) {
            const float4 fb = f4[1];
            g[4 % LD] = fb.x; g[5 % LD] = fb.y; g[6 % LD] = fb.z; g[7 % LD] = fb.w;
        }
        float a0 = b0, a1 = b1;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            a0 = __fmaf_rn(g[k], pw.a0[k], a0);
            a1 = __fmaf_rn(g[k], pw.a1[k], a1);
        }
        mx0 = fmaxf(mx0, a0);
        mx1 = fmaxf(mx1, a1);
    }
    if (out) {
        if (lane < U) out[lane] = mx0;
        if (lane + 32 < U) out[lane + 32] = mx1;
        if (lane == 0) out[U] = (float)n;                                   // :526
    }
    if (canvas_cell) {
        // dense scatter of this pillar (:565-571) straight from the registers: channel c at canvas_cell[c * plane]
        if (lane < U) canvas_cell[(int64_t)lane * plane] = mx0;
        if (lane + 32 < U) canvas_cell[(int64_t)(lane + 32) * plane] = mx1;
        if (lane == 0) canvas_cell[(int64_t)U * plane] = (float)n;
    }
    __syncwarp();
}

// C == 4: the point arrives as the float4 the gather loaded
__device__ __forceinline__ void pfn_pillar4(const PfnWeights &pw, const float4 v, int P, int n, int cx, int cy, float vx,
                                            float vy, float x_off, float y_off, float *row, float *__restrict__ out,
                                            float *__restrict__ canvas_cell, int64_t plane, int U, int lane)
{
    const float f[4] = {v.x, v.y, v.z, v.w};
    pfn_pillar_c<4>(pw, f, P, n, cx, cy, vx, vy, x_off, y_off, row, out, canvas_cell, plane, U, lane);
}

}  // namespace pp
