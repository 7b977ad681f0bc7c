// One pillar of the fused single-layer PillarFeatureNet (model/PointPillars.py:480-526), one warp, P <= 32, C + 5 <= 12,
// U <= 64: shared by the stand-alone kernel (pp_pillar.cu) and by the voxelizer's gather kernel (pp_voxelize.cu), which
// runs it on the pillar it has just gathered, so the two paths give bit-identical features.
// lane = slot while decorating (the points never leave registers until the decorated row is staged for the broadcast
// reads), lane = channel pair (lane, lane + 32) while multiplying.  The mean is a shuffle tree (T1; the bit-exact
// sequential form is decorate_row in pp_pillar.cu, used by pp_decorate and the generic layer kernel).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace pp {

constexpr int PFN_LDI = 12;      // floats per staged row (decorated features, zero padded)

template <int CIN> struct PfnWeights {
    float w0[CIN], w1[CIN], sc0, sh0, sc1, sh1;
};

template <int CIN>
__device__ __forceinline__ void pfn_load_weights(PfnWeights<CIN> &pw, const float *__restrict__ W,
                                                 const float *__restrict__ scale, const float *__restrict__ shift, int U,
                                                 int lane)
{
    const int u0 = lane, u1 = lane + 32;
#pragma unroll
    for (int k = 0; k < CIN; ++k) {
        pw.w0[k] = u0 < U ? W[u0 * CIN + k] : 0.f;
        pw.w1[k] = u1 < U ? W[u1 * CIN + k] : 0.f;
    }
    pw.sc0 = u0 < U ? scale[u0] : 0.f; pw.sh0 = u0 < U ? shift[u0] : 0.f;
    pw.sc1 = u1 < U ? scale[u1] : 0.f; pw.sh1 = u1 < U ? shift[u1] : 0.f;
}

// f[0 .. C) = this lane's point (zeros for an empty slot or lane >= P); row = this warp's 32 x PFN_LDI staging area;
// out = the pillar's (U + 1) feature row.  All 32 lanes of the warp must call it.
template <int CIN>
__device__ __forceinline__ void pfn_pillar(const PfnWeights<CIN> &pw, float (&f)[PFN_LDI], int C, int P, int n, int cx,
                                           int cy, float vx, float vy, float x_off, float y_off, float *row,
                                           float *__restrict__ out, int U, int lane)
{
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
    const float x = f[0], y = f[1], z = f[2];
#pragma unroll
    for (int k = 0; k < PFN_LDI - 5; ++k)
        if (k == C) {                                        // C is 3..7 here (C + 5 == CIN)
            f[k + 0] = __fsub_rn(x, mx_);                    // :496
            f[k + 1] = __fsub_rn(y, my_);
            f[k + 2] = __fsub_rn(z, mz_);
            f[k + 3] = __fsub_rn(x, pcx);
            f[k + 4] = __fsub_rn(y, pcy);
        }
    // slots >= n are never read below (p_end), so the padding mask (:518-521) needs no multiply
    float4 *r4 = reinterpret_cast<float4 *>(row + lane * PFN_LDI);
    r4[0] = make_float4(f[0], f[1], f[2], f[3]);
    r4[1] = make_float4(f[4], f[5], f[6], f[7]);
    r4[2] = make_float4(f[8], f[9], f[10], f[11]);
    __syncwarp();
    // ---- multiply: lane = channels (lane, lane + 32)
    const int p_end = n < P ? n : P;
    // zero-padded slots take part in the max (:403-410): they contribute relu(shift).  relu(y) >= 0, so starting
    // the running max at 0 (or relu(shift)) makes the explicit relu redundant.
    float mx0 = (p_end < P) ? fmaxf(pw.sh0, 0.f) : 0.f;
    float mx1 = (p_end < P) ? fmaxf(pw.sh1, 0.f) : 0.f;
#pragma unroll 4
    for (int p = 0; p < p_end; ++p) {
        const float4 *f4 = reinterpret_cast<const float4 *>(row + p * PFN_LDI);
        const float4 fa = f4[0], fb = f4[1], fc = f4[2];
        const float g[PFN_LDI] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w, fc.x, fc.y, fc.z, fc.w};
        // scalar FMAs: on sm_100 the packed fma.rn.f32x2 issues at a quarter of the FFMA rate (measured: the
        // packed form of this loop stalled on math_pipe_throttle at 0.5 issues per cycle)
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int k = 0; k < CIN; ++k) {
            a0 = __fmaf_rn(g[k], pw.w0[k], a0);
            a1 = __fmaf_rn(g[k], pw.w1[k], a1);
        }
        mx0 = fmaxf(mx0, __fmaf_rn(a0, pw.sc0, pw.sh0));
        mx1 = fmaxf(mx1, __fmaf_rn(a1, pw.sc1, pw.sh1));
    }
    if (lane < U) out[lane] = mx0;
    if (lane + 32 < U) out[lane + 32] = mx1;
    if (lane == 0) out[U] = (float)n;                                       // :526
    __syncwarp();
}

}  // namespace pp
