// Stage 2: pillar decoration (model/PointPillars.py:490-521), PFN layers (:388-423) and the dense
// BEV scatter (:565-571).
//
// Arithmetic follows the reference op by op (separate mul / add roundings, sequential sums) so the
// FP32 results are reproducible bit for bit against the CPU oracle; the library is compiled with
// -fmad=false.  All three kernels are HBM/L2-bandwidth work: the canvas kernel writes every output
// element exactly once with 128-bit stores (zeros included), so the canvas needs no memset.
#include <math_constants.h>

#include "pp_common.cuh"
#include "pp_pillar.cuh"

namespace pp {
namespace {

constexpr int PIL_WARPS = 4;
constexpr int PIL_THREADS = PIL_WARPS * 32;

struct PillarIn {
    const float *voxels;   // (M,P,C)      [DECORATE]
    const float *in;       // (M,P,Cin)    [!DECORATE]
    const void *num;       // (M) int32 / int64
    const void *coors;     // layout per coors_kind
    int num_kind, coors_kind;
    int64_t M;
    const int32_t *m_dev;
    int P, C, Cin;
    float vx, vy, x_off, y_off;
};

__device__ __forceinline__ int load_num(const PillarIn &a, int64_t m)
{
    return a.num_kind == PP_NUM_I64 ? (int)((const int64_t *)a.num)[m] : ((const int32_t *)a.num)[m];
}

__device__ __forceinline__ void load_xy(const PillarIn &a, int64_t m, int &cx, int &cy)
{
    if (a.coors_kind == PP_COORS_XYZ_I32) {
        const int32_t *c = (const int32_t *)a.coors + m * 3;
        cx = c[0]; cy = c[1];
    } else if (a.coors_kind == PP_COORS_BZYX_I32) {
        const int32_t *c = (const int32_t *)a.coors + m * 4;
        cx = c[3]; cy = c[2];
    } else {
        const int64_t *c = (const int64_t *)a.coors + m * 4;
        cx = (int)c[3]; cy = (int)c[2];
    }
}

// Builds the decorated (P, C+5) row of pillar m in shared memory (row stride ld), one warp.  The raw points are
// staged into the row first (one coalesced read of the pillar), then decorated in place.
__device__ void decorate_row(const PillarIn &a, int64_t m, int n, float *row, int ld, int lane)
{
    const int P = a.P, C = a.C;
    const float *v = a.voxels + m * P * C;
    if (C == 4 && ((reinterpret_cast<uintptr_t>(v) & 15) == 0)) {
        for (int p = lane; p < P; p += 32) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(v) + p);
            float *o = row + p * ld;
            o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
        }
    } else {
        for (int i = lane; i < P * C; i += 32) row[(i / C) * ld + (i % C)] = __ldg(v + i);
    }
    __syncwarp();
    // mean over ALL P slots (zero padded), summed sequentially like the oracle (:493-494)
    float s = 0.f;
    if (lane < 3)
        for (int p = 0; p < P; ++p) s = __fadd_rn(s, row[p * ld + lane]);
    const float nf = (float)n;
    const float mean = __fdiv_rn(s, nf);
    const float mx_ = __shfl_sync(0xFFFFFFFFu, mean, 0);
    const float my_ = __shfl_sync(0xFFFFFFFFu, mean, 1);
    const float mz_ = __shfl_sync(0xFFFFFFFFu, mean, 2);
    int cx, cy;
    load_xy(a, m, cx, cy);
    const float pcx = __fadd_rn(__fmul_rn((float)cx, a.vx), a.x_off);   // :500-503
    const float pcy = __fadd_rn(__fmul_rn((float)cy, a.vy), a.y_off);   // :505-508
    for (int p = lane; p < P; p += 32) {
        const float mask = (n > p) ? 1.f : 0.f;                          // utils.py:456
        float *o = row + p * ld;
        const float x = o[0], y = o[1], z = o[2];
        o[0] = __fmul_rn(x, mask);
        o[1] = __fmul_rn(y, mask);
        o[2] = __fmul_rn(z, mask);
        for (int k = 3; k < C; ++k) o[k] = __fmul_rn(o[k], mask);
        o[C + 0] = __fmul_rn(__fsub_rn(x, mx_), mask);                   // :496
        o[C + 1] = __fmul_rn(__fsub_rn(y, my_), mask);
        o[C + 2] = __fmul_rn(__fsub_rn(z, mz_), mask);
        o[C + 3] = __fmul_rn(__fsub_rn(x, pcx), mask);
        o[C + 4] = __fmul_rn(__fsub_rn(y, pcy), mask);
    }
}

__global__ void __launch_bounds__(PIL_THREADS) decorate_kernel(const PillarIn a, float *__restrict__ out)
{
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int CO = a.C + 5, ld = CO | 1;
    float *row = smem + warp * a.P * ld;
    int64_t M = a.M;
    if (a.m_dev) { int64_t md = *a.m_dev; M = md < M ? md : M; }
    for (int64_t m = (int64_t)blockIdx.x * PIL_WARPS + warp; m < M; m += (int64_t)gridDim.x * PIL_WARPS) {
        decorate_row(a, m, load_num(a, m), row, ld, lane);
        __syncwarp();
        float *o = out + m * a.P * CO;
        for (int i = lane; i < a.P * CO; i += 32) o[i] = row[(i / CO) * ld + (i % CO)];
        __syncwarp();
    }
}

// One PFN layer.  DECORATE: the input row is the decoration of `voxels` (fused first layer).
// Shared memory: W as (U, ldw) + one (P, ldi) input row per warp.
template <bool DECORATE>
__global__ void __launch_bounds__(PIL_THREADS)
pfn_kernel(const PillarIn a, const float *__restrict__ W, const float *__restrict__ scale,
           const float *__restrict__ shift, int U, int last_layer, int append_num, float *__restrict__ out)
{
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = a.P, Cin = a.Cin;
    const int ldw = Cin | 1, ldi = Cin | 1;
    float *sW = smem;
    float *row = smem + U * ldw + warp * P * ldi;
    for (int i = threadIdx.x; i < U * Cin; i += PIL_THREADS) sW[(i / Cin) * ldw + (i % Cin)] = W[i];
    __syncthreads();
    int64_t M = a.M;
    if (a.m_dev) { int64_t md = *a.m_dev; M = md < M ? md : M; }
    const int out_w = last_layer ? (U + (append_num ? 1 : 0)) : 2 * U;
    for (int64_t m = (int64_t)blockIdx.x * PIL_WARPS + warp; m < M; m += (int64_t)gridDim.x * PIL_WARPS) {
        int n = P;
        if (DECORATE) {
            n = load_num(a, m);
            decorate_row(a, m, n, row, ldi, lane);
        } else {
            const float *src = a.in + m * P * Cin;
            for (int i = lane; i < P * Cin; i += 32) row[(i / Cin) * ldi + (i % Cin)] = __ldg(src + i);
        }
        __syncwarp();
        // zero-padded slots still take part in the max (:403-410): they contribute relu(shift)
        const int p_end = (DECORATE && last_layer) ? (n < P ? n : P) : P;
        for (int u0 = 0; u0 < U; u0 += 32) {
            const int u = u0 + lane;
            if (u < U) {
                const float sc = scale[u], sh = shift[u];
                const float *w = sW + u * ldw;
                float mx = (p_end < P) ? fmaxf(sh, 0.f) : -CUDART_INF_F;
                for (int p = 0; p < p_end; ++p) {
                    const float *f = row + p * ldi;
                    float acc = 0.f;
                    for (int k = 0; k < Cin; ++k) acc = __fadd_rn(acc, __fmul_rn(f[k], w[k]));
                    float y = __fadd_rn(__fmul_rn(acc, sc), sh);
                    y = y > 0.f ? y : 0.f;
                    if (!last_layer) out[(m * P + p) * out_w + u] = y;
                    mx = fmaxf(mx, y);
                }
                if (last_layer) {
                    out[m * out_w + u] = mx;
                } else {
                    for (int p = 0; p < P; ++p) out[(m * P + p) * out_w + U + u] = mx;
                }
            }
        }
        if (last_layer && append_num && lane == 0) out[m * out_w + U] = (float)n;   // :526
        __syncwarp();
    }
}

// Fast path of the fused single-layer PillarFeatureNet (C <= 7, U <= 64, P <= 32): the folded weights of the two units a
// lane owns live in registers for the whole grid-stride loop; the per-pillar work is pfn_pillar_c (pp_pillar.cuh).
template <int C>
__global__ void __launch_bounds__(PIL_THREADS, 8)
pfn_fused_small_kernel(const PillarIn a, const float *__restrict__ W, const float *__restrict__ scale,
                       const float *__restrict__ shift, int U, float *__restrict__ out)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(16) float smem[];
    constexpr int LD = C <= 4 ? 4 : 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = a.P;
    float *row = smem + warp * 32 * LD;
    PfnWeightsC<C> pw;
    pfn_load_weights_c<C>(pw, W, scale, shift, U, lane);          // weights do not depend on the predecessor
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int64_t M = a.M;
    if (a.m_dev) { int64_t md = *a.m_dev; M = md < M ? md : M; }
    const int out_w = U + 1;
    const bool vec4 = (C == 4) && ((reinterpret_cast<uintptr_t>(a.voxels) & 15) == 0);
    for (int64_t m = (int64_t)blockIdx.x * PIL_WARPS + warp; m < M; m += (int64_t)gridDim.x * PIL_WARPS) {
        const int n = load_num(a, m);
        int cx, cy;
        load_xy(a, m, cx, cy);
        float f[C];
#pragma unroll
        for (int k = 0; k < C; ++k) f[k] = 0.f;
        if (lane < P) {
            const float *v = a.voxels + (m * P + lane) * C;
            if (C == 4 && vec4) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(v));
                f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3 % C] = t.w;
            } else {
#pragma unroll
                for (int k = 0; k < C; ++k) f[k] = __ldg(v + k);
            }
        }
        pfn_pillar_c<C>(pw, f, P, n, cx, cy, a.vx, a.vy, a.x_off, a.y_off, row, out + m * out_w, nullptr, 0, U, lane);
    }
}

// ---- dense scatter ---------------------------------------------------------------------------
struct CoorsIn {
    const void *coors;
    int kind, batch_index;
};

__device__ __forceinline__ void load_bzyx(const CoorsIn &c, int64_t i, int &b, int &z, int &y, int &x)
{
    if (c.kind == PP_COORS_XYZ_I32) {
        const int32_t *p = (const int32_t *)c.coors + i * 3;
        b = c.batch_index; x = p[0]; y = p[1]; z = p[2];
    } else if (c.kind == PP_COORS_BZYX_I32) {
        const int32_t *p = (const int32_t *)c.coors + i * 4;
        b = p[0]; z = p[1]; y = p[2]; x = p[3];
    } else {
        const int64_t *p = (const int64_t *)c.coors + i * 4;
        b = (int)p[0]; z = (int)p[1]; y = (int)p[2]; x = (int)p[3];
    }
}

__global__ void __launch_bounds__(256)
scatter_map_kernel(const CoorsIn c, int64_t M, const int32_t *__restrict__ m_dev, int B, int D, int H, int W,
                   int32_t *__restrict__ map)
{
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (m_dev) { int64_t md = *m_dev; M = md < M ? md : M; }
    if (i >= M) return;
    int b, z, y, x;
    load_bzyx(c, i, b, z, y, x);
    if (b < 0 || b >= B || z < 0 || z >= D || y < 0 || y >= H || x < 0 || x >= W) return;
    atomicMax(map + (((int64_t)b * D + z) * H + y) * W + x, (int32_t)i);
}

constexpr int CANVAS_CELLS = 128;   // cells per CTA pass: 32 lanes x float4
constexpr int CANVAS_WARPS = 8;

// Each CTA owns 128 consecutive cells of one (b, z) plane and writes all C channels of them.
template <bool VEC4>
__global__ void __launch_bounds__(CANVAS_WARPS * 32)
scatter_canvas_kernel(const float *__restrict__ feat, const int32_t *__restrict__ map, int C, int D, int64_t HW,
                      int tiles_per_plane, float *__restrict__ canvas)
{
    pdl_enter();
    __shared__ int32_t s_pid[CANVAS_CELLS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t plane = blockIdx.x / tiles_per_plane;        // b * D + z
    const int64_t cell0 = (int64_t)(blockIdx.x % tiles_per_plane) * CANVAS_CELLS;
    const int64_t b = plane / D, z = plane % D;
    if (threadIdx.x < CANVAS_CELLS) {
        int64_t cell = cell0 + threadIdx.x;
        s_pid[threadIdx.x] = cell < HW ? map[plane * HW + cell] : -1;
    }
    __syncthreads();
    const int64_t c4 = cell0 + lane * 4;
    int32_t pid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pid[k] = s_pid[lane * 4 + k];
    for (int c = warp; c < C; c += CANVAS_WARPS) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = pid[k] >= 0 ? __ldg(feat + (int64_t)pid[k] * C + c) : 0.f;
        float *dst = canvas + ((b * C + c) * D + z) * HW + c4;
        if (VEC4) {
            if (c4 < HW) *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c4 + k < HW) dst[k] = v[k];
        }
    }
}

// Canvas tile of 256 float4 columns (1024 consecutive cells) x a quarter of the channels per CTA: a thread owns one
// column, so the pillar ids stay in registers and a channel step is 4 predicated loads + one 128-bit store; a CTA
// writes 4 KB contiguous per channel row.  Columns never straddle planes (HW % 4 == 0).
constexpr int CV_THREADS = 256;
constexpr int CV_CGROUPS = 4;      // channel groups (blockIdx.y): channel c belongs to group c % CV_CGROUPS

// PDL: the zeros do not depend on anything the predecessors on the stream compute, so they are written BEFORE
// griddepcontrol.wait, i.e. while the (FP32-bound) PFN kernel is still running; after the wait only the columns that
// hold a pillar (a few percent) are overwritten with features.
__global__ void __launch_bounds__(CV_THREADS, 8)
scatter_canvas_wave_kernel(const float *__restrict__ feat, const int32_t *__restrict__ map, int C, int D, int HW,
                           float *__restrict__ canvas)
{
    // grid = (column tiles of one plane, channel groups, planes): no index divisions (64-bit divisions by run-time
    // values were half of this kernel's instructions)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int col = blockIdx.x * CV_THREADS + threadIdx.x;                // float4 column inside the plane
    const int cq = blockIdx.y;
    const int plane = blockIdx.z;                                         // b * D + z
    const int b = plane / D, z = plane - b * D;
    const int cell = col << 2;
    const size_t chan_stride = (size_t)D * HW;
    float *dst0 = canvas + ((size_t)b * C * D + z) * HW + cell + (size_t)cq * chan_stride;
    const bool in = cell < HW;
    if (in) {
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        float *dst = dst0;
#pragma unroll 4
        for (int c = cq; c < C; c += CV_CGROUPS, dst += CV_CGROUPS * chan_stride) *reinterpret_cast<float4 *>(dst) = zero;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!in) return;
    const int4 pid = *reinterpret_cast<const int4 *>(map + (size_t)plane * HW + cell);
    if ((pid.x & pid.y & pid.z & pid.w) < 0) return;      // all four cells empty (the common case)
    float *dst = dst0;
    for (int c = cq; c < C; c += CV_CGROUPS, dst += CV_CGROUPS * chan_stride) {
        float4 v;
        v.x = pid.x >= 0 ? __ldg(feat + (int64_t)pid.x * C + c) : 0.f;
        v.y = pid.y >= 0 ? __ldg(feat + (int64_t)pid.y * C + c) : 0.f;
        v.z = pid.z >= 0 ? __ldg(feat + (int64_t)pid.z * C + c) : 0.f;
        v.w = pid.w >= 0 ? __ldg(feat + (int64_t)pid.w * C + c) : 0.f;
        *reinterpret_cast<float4 *>(dst) = v;
    }
}

// The same tile with 256-bit stores (sm_100: STG.256): a thread owns 8 consecutive cells, so a warp writes 1 KB
// contiguous per channel row and the zero fill issues half as many store instructions (the 128-bit form was bound by
// the store path: lg_throttle).  Needs HW % 8 == 0 and a 32-byte aligned canvas / map.
__device__ __forceinline__ void st256(float *p, float a, float b, float c, float d, float e, float f, float g, float h)
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f),
                 "f"(g), "f"(h) : "memory");
}

__global__ void __launch_bounds__(CV_THREADS, 8)
scatter_canvas_wave8_kernel(const float *__restrict__ feat, const int32_t *__restrict__ map, int C, int D, int HW,
                            float *__restrict__ canvas)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int col = blockIdx.x * CV_THREADS + threadIdx.x;                // 8-cell column inside the plane
    const int cq = blockIdx.y;
    const int plane = blockIdx.z;                                         // b * D + z
    const int b = plane / D, z = plane - b * D;
    const int cell = col << 3;
    const size_t chan_stride = (size_t)D * HW;
    float *dst0 = canvas + ((size_t)b * C * D + z) * HW + cell + (size_t)cq * chan_stride;
    const bool in = cell < HW;
    if (in) {
        float *dst = dst0;
#pragma unroll 4
        for (int c = cq; c < C; c += CV_CGROUPS, dst += CV_CGROUPS * chan_stride) st256(dst, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!in) return;
    const int4 p0 = *reinterpret_cast<const int4 *>(map + (size_t)plane * HW + cell);
    const int4 p1 = *reinterpret_cast<const int4 *>(map + (size_t)plane * HW + cell + 4);
    if ((p0.x & p0.y & p0.z & p0.w & p1.x & p1.y & p1.z & p1.w) < 0) return;      // all eight cells empty (the common case)
    float *dst = dst0;
    auto val = [&](int pid, int c) { return pid >= 0 ? __ldg(feat + (int64_t)pid * C + c) : 0.f; };
    for (int c = cq; c < C; c += CV_CGROUPS, dst += CV_CGROUPS * chan_stride)
        st256(dst, val(p0.x, c), val(p0.y, c), val(p0.z, c), val(p0.w, c), val(p1.x, c), val(p1.y, c), val(p1.z, c), val(p1.w, c));
}

int fill_pillar_in(PillarIn &a, const float *voxels, const float *in, const void *num, int num_kind, const void *coors,
                   int coors_kind, int64_t M, const int32_t *m_dev, int P, int C, int Cin, float vx, float vy,
                   float x_off, float y_off)
{
    a.voxels = voxels; a.in = in; a.num = num; a.coors = coors;
    a.num_kind = num_kind; a.coors_kind = coors_kind;
    a.M = M; a.m_dev = m_dev; a.P = P; a.C = C; a.Cin = Cin;
    a.vx = vx; a.vy = vy; a.x_off = x_off; a.y_off = y_off;
    return 0;
}

dim3 canvas_wave_grid(int64_t HW, int64_t planes) { return dim3((unsigned)ceil_div(HW / 4, CV_THREADS), CV_CGROUPS, (unsigned)planes); }
dim3 canvas_wave8_grid(int64_t HW, int64_t planes) { return dim3((unsigned)ceil_div(HW / 8, CV_THREADS), CV_CGROUPS, (unsigned)planes); }
// (measured on B200: with its 16 strided 256-bit stores per thread this form takes ~30 us for the 54.9 MB canvas against 19 us for the
// 128-bit kernel above, so it is not dispatched; the linear STG.256 fill of pp_voxelize_scatter is the fast form)
inline bool canvas_wave8_ok(int64_t, const void *, const void *) { return false; }

unsigned pillar_grid(int64_t M)
{
    int64_t g = ceil_div(M, PIL_WARPS);
    int64_t cap = 148 * 16;
    return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace
}  // namespace pp

using namespace pp;

extern "C" int pp_decorate(const float *voxels, const void *num_points, int num_kind, const void *coors,
                           int coors_kind, int64_t M, const int32_t *m_dev, int P, int C, float vx, float vy,
                           float x_off, float y_off, float *out, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(M >= 0 && P > 0 && C >= 3, "bad shape");
    if (M == 0) return PP_OK;
    PP_REQUIRE(voxels && num_points && coors && out, "null pointer");
    PP_REQUIRE(num_kind == PP_NUM_I32 || num_kind == PP_NUM_I64, "bad num_kind");
    PP_REQUIRE(coors_kind >= 0 && coors_kind <= 2, "bad coors_kind");
    PillarIn a;
    fill_pillar_in(a, voxels, nullptr, num_points, num_kind, coors, coors_kind, M, m_dev, P, C, C + 5, vx, vy, x_off,
                   y_off);
    size_t smem = (size_t)PIL_WARPS * P * ((C + 5) | 1) * sizeof(float);
    PP_REQUIRE(smem <= 200 * 1024, "P * (C+5) too large for shared memory");
    if (smem > 48 * 1024)
        PP_CUDA_TRY(cudaFuncSetAttribute(decorate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decorate_kernel<<<pillar_grid(M), PIL_THREADS, smem, (cudaStream_t)stream>>>(a, out);
    return check_launch("decorate_kernel");
}

static int launch_pfn(bool decorate, const PillarIn &a, const float *W, const float *scale, const float *shift, int U,
                      int last_layer, int append_num, float *out, cudaStream_t st)
{
    size_t smem = ((size_t)U * (a.Cin | 1) + (size_t)PIL_WARPS * a.P * (a.Cin | 1)) * sizeof(float);
    PP_REQUIRE(smem <= 200 * 1024, "PFN layer too large for shared memory");
    if (decorate) {
        if (smem > 48 * 1024)
            PP_CUDA_TRY(cudaFuncSetAttribute(pfn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pfn_kernel<true><<<pillar_grid(a.M), PIL_THREADS, smem, st>>>(a, W, scale, shift, U, last_layer, append_num, out);
    } else {
        if (smem > 48 * 1024)
            PP_CUDA_TRY(cudaFuncSetAttribute(pfn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pfn_kernel<false><<<pillar_grid(a.M), PIL_THREADS, smem, st>>>(a, W, scale, shift, U, last_layer, append_num, out);
    }
    return check_launch("pfn_kernel");
}

extern "C" int pp_pfn_layer(const float *in, int64_t M, int P, int Cin, const float *weight, const float *scale,
                            const float *shift, int U, int last_layer, float *out, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(M >= 0 && P > 0 && Cin > 0 && U > 0, "bad shape");
    if (M == 0) return PP_OK;
    PP_REQUIRE(in && weight && scale && shift && out, "null pointer");
    PillarIn a;
    fill_pillar_in(a, nullptr, in, nullptr, 0, nullptr, 0, M, nullptr, P, 0, Cin, 0, 0, 0, 0);
    return launch_pfn(false, a, weight, scale, shift, U, last_layer, 0, out, (cudaStream_t)stream);
}

extern "C" int pp_pillar_features(const float *voxels, const void *num_points, int num_kind, const void *coors,
                                  int coors_kind, int64_t M, const int32_t *m_dev, int P, int C, float vx, float vy,
                                  float x_off, float y_off, const float *weight, const float *scale,
                                  const float *shift, int U, float *feat, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(M >= 0 && P > 0 && C >= 3 && U > 0, "bad shape");
    if (M == 0) return PP_OK;
    PP_REQUIRE(voxels && num_points && coors && weight && scale && shift && feat, "null pointer");
    PP_REQUIRE(num_kind == PP_NUM_I32 || num_kind == PP_NUM_I64, "bad num_kind");
    PP_REQUIRE(coors_kind >= 0 && coors_kind <= 2, "bad coors_kind");
    PillarIn a;
    fill_pillar_in(a, voxels, nullptr, num_points, num_kind, coors, coors_kind, M, m_dev, P, C, C + 5, vx, vy, x_off,
                   y_off);
    if (C <= 7 && U <= 64 && P <= 32) {
        size_t smem = (size_t)PIL_WARPS * 32 * (C <= 4 ? 4 : 8) * sizeof(float);
        int64_t want = ceil_div(M, PIL_WARPS);
        // one resident wave that leaves room on every SM for the canvas kernel's early (pre-wait) zero fill
        const unsigned grid = (unsigned)(want < 148 * 6 ? (want > 0 ? want : 1) : 148 * 6);
        cudaStream_t st = (cudaStream_t)stream;
        switch (C) {
        case 3: launch_pdl(pfn_fused_small_kernel<3>, dim3(grid), dim3(PIL_THREADS), smem, st, a, weight, scale, shift, U, feat); break;
        case 4: launch_pdl(pfn_fused_small_kernel<4>, dim3(grid), dim3(PIL_THREADS), smem, st, a, weight, scale, shift, U, feat); break;
        case 5: launch_pdl(pfn_fused_small_kernel<5>, dim3(grid), dim3(PIL_THREADS), smem, st, a, weight, scale, shift, U, feat); break;
        case 6: launch_pdl(pfn_fused_small_kernel<6>, dim3(grid), dim3(PIL_THREADS), smem, st, a, weight, scale, shift, U, feat); break;
        default: launch_pdl(pfn_fused_small_kernel<7>, dim3(grid), dim3(PIL_THREADS), smem, st, a, weight, scale, shift, U, feat); break;
        }
        return check_launch("pfn_fused_small_kernel");
    }
    return launch_pfn(true, a, weight, scale, shift, U, 1, 1, feat, (cudaStream_t)stream);
}

extern "C" size_t pp_scatter_workspace_bytes(int B, int D, int H, int W)
{
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    return align_up((size_t)B * D * H * W * sizeof(int32_t));
}

extern "C" int pp_scatter_dense(const float *feat, const void *coors, int coors_kind, int64_t M, const int32_t *m_dev,
                                int C, int batch_index, int B, int D, int H, int W, float *canvas, void *map_ws,
                                size_t map_ws_bytes, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    PP_REQUIRE(M >= 0 && C > 0 && B > 0 && D > 0 && H > 0 && W > 0, "bad shape");
    PP_REQUIRE(canvas && map_ws, "null pointer");
    PP_REQUIRE(coors_kind >= 0 && coors_kind <= 2, "bad coors_kind");
    const size_t need = (size_t)B * D * H * W * sizeof(int32_t);
    if (map_ws_bytes < need) {
        set_error("scatter workspace too small: %zu < %zu", map_ws_bytes, need);
        return PP_ERR_WORKSPACE;
    }
    const int64_t HW = (int64_t)H * W;
    const int64_t tiles_per_plane = ceil_div(HW, CANVAS_CELLS);
    PP_REQUIRE((int64_t)B * D * tiles_per_plane < (1ll << 31), "canvas too large");
    int32_t *map = (int32_t *)map_ws;
    PP_CUDA_TRY(cudaMemsetAsync(map, 0xFF, need, st));
    prof_mark("memset");
    if (M > 0) {
        PP_REQUIRE(feat && coors, "null pointer");
        CoorsIn c{coors, coors_kind, batch_index};
        scatter_map_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(c, M, m_dev, B, D, H, W, map);
        if (int rc = check_launch("scatter_map_kernel")) return rc;
    }
    const unsigned grid = (unsigned)((int64_t)B * D * tiles_per_plane);
    const bool vec4 = (HW % 4 == 0) && ((uintptr_t)canvas % 16 == 0);
    if (vec4 && canvas_wave8_ok(HW, canvas, map) && HW < (1ll << 30) && (int64_t)B * D < 65536)
        scatter_canvas_wave8_kernel<<<canvas_wave8_grid(HW, (int64_t)B * D), CV_THREADS, 0, st>>>(feat, map, C, D, (int)HW, canvas);
    else if (vec4 && ((uintptr_t)map % 16 == 0) && HW < (1ll << 30) && (int64_t)B * D < 65536)
        scatter_canvas_wave_kernel<<<canvas_wave_grid(HW, (int64_t)B * D), CV_THREADS, 0, st>>>(feat, map, C, D, (int)HW, canvas);
    else if (vec4)
        scatter_canvas_kernel<true><<<grid, CANVAS_WARPS * 32, 0, st>>>(feat, map, C, D, HW, (int)tiles_per_plane, canvas);
    else
        scatter_canvas_kernel<false><<<grid, CANVAS_WARPS * 32, 0, st>>>(feat, map, C, D, HW, (int)tiles_per_plane, canvas);
    return check_launch("scatter_canvas_kernel");
}

extern "C" int pp_scatter_mapped(const float *feat, const int32_t *pillar_map, int C, int B, int D, int H, int W,
                                 float *canvas, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    PP_REQUIRE(C > 0 && B > 0 && D > 0 && H > 0 && W > 0, "bad shape");
    PP_REQUIRE(feat && pillar_map && canvas, "null pointer");
    const int64_t HW = (int64_t)H * W;
    const int64_t tiles_per_plane = ceil_div(HW, CANVAS_CELLS);
    PP_REQUIRE((int64_t)B * D * tiles_per_plane < (1ll << 31), "canvas too large");
    const unsigned grid = (unsigned)((int64_t)B * D * tiles_per_plane);
    const bool vec4 = (HW % 4 == 0) && ((uintptr_t)canvas % 16 == 0);
    if (vec4 && canvas_wave8_ok(HW, canvas, pillar_map) && HW < (1ll << 30) && (int64_t)B * D < 65536)
        launch_pdl(scatter_canvas_wave8_kernel, canvas_wave8_grid(HW, (int64_t)B * D), dim3(CV_THREADS), 0, st, feat, pillar_map, C, D,
                   (int)HW, canvas);
    else if (vec4 && ((uintptr_t)pillar_map % 16 == 0) && HW < (1ll << 30) && (int64_t)B * D < 65536)
        launch_pdl(scatter_canvas_wave_kernel, canvas_wave_grid(HW, (int64_t)B * D), dim3(CV_THREADS), 0, st, feat, pillar_map, C, D,
                   (int)HW, canvas);
    else if (vec4)
        launch_pdl(scatter_canvas_kernel<true>, dim3(grid), dim3(CANVAS_WARPS * 32), 0, st, feat, pillar_map, C, D, HW,
                   (int)tiles_per_plane, canvas);
    else
        launch_pdl(scatter_canvas_kernel<false>, dim3(grid), dim3(CANVAS_WARPS * 32), 0, st, feat, pillar_map, C, D, HW,
                   (int)tiles_per_plane, canvas);
    return check_launch("scatter_canvas_kernel");
}
