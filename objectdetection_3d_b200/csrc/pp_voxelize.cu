// Hard voxelization on sm_100a, bit-exact with the reference's sequential first-come pass
// (ops/ops_numba.py:171-308), restated as order-independent parallel steps with NO global sort, no claim protocol
// and no grid-wide barrier.
//
// Every point gets a unique ordering key K; the reference processes points in ascending K:
//   given order / replayed permutation : K = position p                                  (32 bit)
//   reflectance pre-order (:262)       : K = (~ordered(reflectance) << 32) | index        (64 bit)
//                                        = descending reflectance, ties by lower index
// The reference's outputs are functions of K only:
//   pillar id   = rank of the cell's smallest key among all cells' smallest keys,
//   `break`     = the (max_voxels+1)-th smallest cell minimum is the cutoff: keys >= cutoff are dropped,
//   slot        = rank of the key inside its cell, capped at max_points.
//
// A cell keeps its max_points smallest keys.  Keys are binned into 8 chunks that are monotone in K and geometric in the
// key's quantile ([0,1/128) [1/128,1/64) ... [1/2,1]; quantiles of a small key sample, or of the position), so a cell
// with any number of points finds its max_points-th key in a chunk whose prefix is at most about 2 x max_points.
// Cells are addressed directly (slot = cell) when the grid is small, through an open-addressing hash of the cell id
// otherwise (3-D grids with millions of cells); everything after the slot lookup is the same.
//
// Kernels (all with programmatic dependent launch):
//   S  vox_init_kernel     zero the chunk counters / small counters / histogram, -1 into the hash keys and the
//                          pillar map; reflectance order: 7 chunk splitters from a 128-key sample (ranked by counting)
//   A  vox_count_kernel    per point: cell -> slot, chunk(K); ticket = cnt[slot][chunk]++ : the ONE random L2 atomic a
//                          point costs; per-point record (slot, chunk, ticket, key high word).  Extra CTAs rank a
//                          512-key sample by counting -> 511 fine splitters (ranking bins), off the critical path
//   B  vox_cells_kernel    per slot: counts -> saturation chunk (the first chunk at which the cell holds max_points
//                          points), m = points in the chunks up to it; segment of m keys and compact cell id q from
//                          CTA-wide prefix sums + one atomic per CTA; the counters become write positions
//                          (segment offset + exclusive prefix over the chunks; "dropped" after the saturation chunk)
//   C  vox_place_kernel    per point: seg[position[slot][chunk] + ticket] = K -- one cached read and one store, no
//                          atomic; most points of a dense cell are dropped after the read.  Segments come out ordered
//                          by chunk
//   H  vox_rank_kernel     per cell (8 lanes): smallest key = minimum over the segment's first chunk; ranking bin of
//                          it (fine splitters, log-spaced bins), arrival index inside the bin
//   R  vox_bucket_kernel   every CTA scans the bin histogram in shared memory and puts its cells' first keys in bucket
//                          order; the last CTA to finish settles the cutoff when more than max_voxels cells are occupied
//   D  vox_gather_kernel   per cell (warp): pillar id = bucket base + smaller keys inside the bucket; the max_points
//                          smallest keys of the segment in order (ranks by counting in shared memory; bitonic merges
//                          for long segments), keys >= cutoff dropped; voxels[pid][s] = points[key.index], coors,
//                          num_points, cell -> pillar map; with the fused PillarFeatureNet the warp runs the pillar
//                          through decorate + Linear + BN + ReLU + max and writes the BEV canvas column
// Everything but the 16 B/point read and the output write is L2-resident workspace traffic.
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include "pp_common.cuh"
#include "pp_pillar.cuh"

namespace pp {
namespace {

typedef unsigned long long u64;
constexpr int VOX_THREADS = 256;
constexpr int NCH = 8;              // key chunks per cell: one 32-byte sector of counters
constexpr int NFINE = 512;          // sample intervals used to rank the cells' first keys
constexpr int NBIN = 8192;          // ranking bins
constexpr int SAMPLE = 512;         // keys sampled for the fine splitters
constexpr int CSAMPLE = 128;        // keys sampled for the chunk splitters
constexpr uint32_t DROPPED = 0xFFFFFFFFu;
enum { CTR_NQ = 0, CTR_SEG = 1, CTR_DONE = 2 };

// Development only (-DPP_TIMING): first-start / last-end %globaltimer of every kernel of a call in counters[16 ...]
#ifdef PP_TIMING
#define PP_T0(cnt, k) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMax((unsigned long long *)((cnt) + 16) + (k), ~t_); } } while (0)
#define PP_T1(cnt, k) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMax((unsigned long long *)((cnt) + 16) + (k), t_); } } while (0)
#else
#define PP_T0(cnt, k) do { } while (0)
#define PP_T1(cnt, k) do { } while (0)
#endif

template <typename K> struct KeyInf;
template <> struct KeyInf<uint32_t> { static __device__ __host__ constexpr uint32_t value() { return 0xFFFFFFFFu; } };
template <> struct KeyInf<u64> { static __device__ __host__ constexpr u64 value() { return ~0ull; } };

struct VoxParams {
    double r[3], v[3];
    float rf[3], vf[3], inv_vf[3], rv_abs[3];   // rv_abs = |r| / v: scale of the estimate's absolute error
    int g[3];
    int regime;   // 0: all f32   1: sub f32, div f64   2: all f64   (numba promotion)
    int P, max_voxels, C;
    int vec4;     // C == 4 and 16-byte aligned rows: float4 loads
    int bits;     // 32-bit keys: positions < 2^bits; chunks / fine bins are geometric in the position
};

struct VoxBuf {
    int32_t T;             // slots: the cells (direct) or a power of two >= 2 n (hash)
    int32_t hash_bits;     // 0: slot = cell
    int32_t *slot_key;     // [T] hash mode: cell of the slot, -1 empty
    uint32_t *cnt;         // [T][NCH] chunk counts; after kernel B: write position of each chunk, DROPPED after the saturation chunk
    int32_t *counters;     // CTR_*
    uint32_t *rec_slot;    // [N] (slot << 3) | chunk, ~0 outside the grid
    uint32_t *rec_tick;    // [N] arrival index inside (cell, chunk)
    uint32_t *rec_hi;      // [N] key high word (64-bit keys)
    void *seg;             // [N] keys grouped by cell, inside a cell by chunk
    int4 *qinfo;           // [Q] cell, segment offset, m, keys in the cell's first occupied chunk
    int4 *qxyz;            // [Q] x, y, z of the cell
    void *first;           // [Q] smallest key of the cell
    int32_t *bin_of_q;     // [Q]
    int32_t *hist;         // [NBIN]
    int32_t *base;         // [NBIN + 1] exclusive scan of hist
    int32_t *arr_of_q;     // [Q] arrival index inside the bin
    void *lkey;            // [Q] first keys in bucket order (bin by bin, arrival order inside a bin)
    void *cutoff;          // key
    u64 *coarse, *fine;    // splitters (64-bit keys): [NCH - 1], [NFINE - 1]
};

// Cell index along one axis, bit-identical to the reference's floor((p - r) / v) in its promotion regime: the exact
// evaluation (fp64 / IEEE-fp32 division), used for points within a guard band of a cell boundary.
// (scalar arguments and not inlined: the rare path must not make the parameter block addressable or bloat the kernel)
__device__ __noinline__ int axis_cell_exact(int regime, double r, double v, float rf, float vf, int g, float p)
{
    double cd;
    if (regime == 2) cd = floor(((double)p - r) / v);
    else if (regime == 1) cd = floor((double)__fsub_rn(p, rf) / v);
    else cd = (double)floorf(__fdiv_rn(__fsub_rn(p, rf), vf));
    if (!(cd >= 0.0) || cd >= (double)g) return -1;           // also rejects NaN
    return (int)cd;
}

// Linear cell of a point, or -1 outside the grid.  An fp32 reciprocal-multiply estimate decides every point that is
// not within a guard band of a cell boundary on any axis (the band covers the rounding of r, 1/v and the two fp32
// operations; outside it the estimate's floor and the reference's agree); only those points take the exact path.
// Six instructions per axis on the common path: subtract, multiply, round, distance to the nearest integer, band, compare.
__device__ __forceinline__ int32_t point_cell(const VoxParams &q, float x, float y, float z)
{
    const float p[3] = {x, y, z};
    int c[3];
    bool near = false, inside = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float est = (p[j] - q.rf[j]) * q.inv_vf[j];
        const float dist = fabsf(est - rintf(est));                // distance to the nearest cell boundary
        const float band = 1e-6f * (fabsf(est) + q.rv_abs[j]) + 1e-6f;
        near = near || !(dist > band);                             // also true for NaN / inf
        c[j] = __float2int_rd(est);
        inside = inside && ((unsigned)c[j] < (unsigned)q.g[j]);
    }
    if (near) {
#pragma unroll
        for (int j = 0; j < 3; ++j) c[j] = axis_cell_exact(q.regime, q.r[j], q.v[j], q.rf[j], q.vf[j], q.g[j], p[j]);
        if ((c[0] | c[1] | c[2]) < 0) return -1;
    } else if (!inside) {
        return -1;
    }
    // cell linearisation (z*gy + y)*gx + x = the (D,H,W) order of the BEV canvas
    return (c[2] * q.g[1] + c[1]) * q.g[0] + c[0];
}

// High word of a reflectance key: descending reflectance in ascending unsigned order; -0.0 orders like +0.0 (the
// reference's argsort compares them equal, ops/ops_numba.py:262)
__device__ __forceinline__ uint32_t refl_key_hi(float refl)
{
    uint32_t u = __float_as_uint(refl);
    if (u == 0x80000000u) u = 0u;
    return ~(u ^ ((u & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u));
}

// Geometric binning of a position u < 2^bits into NOCT octaves x 2^SUB sub-bins (monotone in u).  Octave 0 is
// [0, 2^(bits-NOCT+1)), octave k >= 1 is [2^(bits-NOCT+k), 2^(bits-NOCT+k+1)); each octave is split evenly.
template <int NOCT, int SUB>
__device__ __forceinline__ int geo_bin(uint32_t u, int bits)
{
    const int l0 = bits - (NOCT - 1);                     // log2 of octave 0's width
    if (l0 < SUB) return (int)min(u >> max(bits - (31 - __clz(NOCT << SUB)), 0), (uint32_t)((NOCT << SUB) - 1));   // tiny inputs: uniform
    const uint32_t top = u >> l0;
    if (top == 0) return (int)(u >> (l0 - SUB));
    const int k = 31 - __clz(top);                        // octave k + 1
    return ((k + 1) << SUB) | (int)((u - (1u << (l0 + k))) >> (l0 + k - SUB));
}

constexpr int C_OCT = 8, C_SUB = 0;      // 8 chunks: binary-geometric in the position

// rank += (a < b) as compare + predicated increment (three instructions for a 64-bit key)
__device__ __forceinline__ void count_if_less(int &rank, u64 a, u64 b)
{
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u64 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(rank) : "l"(a), "l"(b));
}
__device__ __forceinline__ void count_if_less(int &rank, uint32_t a, uint32_t b)
{
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(rank) : "r"(a), "r"(b));
}

__device__ __forceinline__ u64 sample_key(const float *__restrict__ points, int64_t n, int C, int i, int nsample)
{
    // evenly spaced sample; short inputs are padded with +inf keys
    const int64_t idx = (n >= nsample) ? (int64_t)i * (n / nsample) : i;
    return idx < n ? (((u64)refl_key_hi(points[idx * C + 3]) << 32) | (uint32_t)idx) : ~0ull - (u64)(nsample - i);   // (distinct pads)
}

// ---- canvas zero fill, spread over the per-point kernels ---------------------------------------------------------------
// The fused frame call (pp_voxelize_scatter) writes the BEV canvas from the gather kernel; the zeros of the other ~95 % of
// the canvas do not depend on anything, so halves of them are written by kernels A and C BEFORE their dependency wait,
// i.e. while the predecessor is still running: linear 256-bit stores (STG.256), fire and forget.  The init kernel
// orders wait -> trigger, so none of this starts before the caller's earlier work on the stream is complete.
struct FillArgs {
    float *base;          // nullptr: nothing to fill
    int64_t units;        // 32-byte units in the canvas
};
constexpr int FILL_SLICES = 2;

// (CTA `block` of `nblocks` taking part)
__device__ __forceinline__ void fill_slice(const FillArgs &fa, int slice, int block, int nblocks)
{
    if (!fa.base) return;
    const int64_t per = (fa.units + FILL_SLICES - 1) / FILL_SLICES;
    const int64_t lo = per * slice, hi = lo + per < fa.units ? lo + per : fa.units;
    const int64_t stride = (int64_t)nblocks * blockDim.x;
    for (int64_t i = lo + (int64_t)block * blockDim.x + threadIdx.x; i < hi; i += stride)
        asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(fa.base + i * 8), "f"(0.f) : "memory");
}

// ---- S: workspace initialisation, and (reflectance order) chunk splitters from a key sample ---------------------------
// One launch replaces the memsets.  CTA 0 also ranks a 128-key sample by counting (no sort, no barrier chain) and
// publishes the keys of rank 1, 2, 4 ... 64 as the starts of chunks 1 ... 7.
struct InitArgs {
    int4 *ff_ptr[3];  int64_t ff_n[3];     // regions filled with 0xFF (16-byte units)
    int4 *z_ptr[2];   int64_t z_n[2];      // regions filled with 0
};

__global__ void __launch_bounds__(1024)
vox_init_kernel(const float *__restrict__ points, int64_t n, int C, int wide, u64 *__restrict__ coarse, const InitArgs ia)
{
    // First kernel of the call: wait for everything earlier on the stream, THEN let the per-point kernel start -- its
    // CTAs read `points` before their own dependency wait, which is only safe once the producer of the points is done.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x;
    __shared__ u64 s[CSAMPLE];
    if (wide && blockIdx.x == 0) {
        if (tid < CSAMPLE) s[tid] = sample_key(points, n, C, tid, CSAMPLE);
        __syncthreads();
        if (tid < CSAMPLE) {
            const u64 k = s[tid];
            int rank = 0;
#pragma unroll 8
            for (int j = 0; j < CSAMPLE; ++j) rank += (s[j] < k) ? 1 : 0;       // keys are unique (index in the low word)
            if (rank >= 1 && (rank & (rank - 1)) == 0 && rank < CSAMPLE)        // 1 2 4 ... 64 -> splitter 0 ... 6
                coarse[31 - __clz(rank)] = k;
        }
    }
    const int4 ff = make_int4(-1, -1, -1, -1), zz = make_int4(0, 0, 0, 0);
#pragma unroll
    for (int r = 0; r < 3; ++r)
        for (int64_t i = (int64_t)blockIdx.x * 1024 + tid; i < ia.ff_n[r]; i += (int64_t)gridDim.x * 1024) ia.ff_ptr[r][i] = ff;
#pragma unroll
    for (int r = 0; r < 2; ++r)
        for (int64_t i = (int64_t)blockIdx.x * 1024 + tid; i < ia.z_n[r]; i += (int64_t)gridDim.x * 1024) ia.z_ptr[r][i] = zz;
}

// ---- slot of a cell ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_cell(uint32_t cell, int bits) { return (cell * 0x9E3779B1u) >> (32 - bits); }

// Open addressing, linear probing.  A slot's key goes from -1 to a cell once and never changes, so the CAS that
// inserts a key is also its publication: nobody waits for anybody.
__device__ __forceinline__ int hash_insert(const VoxBuf &w, int32_t cell)
{
    const uint32_t mask = (1u << w.hash_bits) - 1u;
    uint32_t h = hash_cell((uint32_t)cell, w.hash_bits);
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        int32_t k = __ldcg(w.slot_key + h);
        if (k == -1) k = atomicCAS(w.slot_key + h, -1, cell);
        if (k == -1 || k == cell) return (int)h;
        h = (h + 1) & mask;
    }
    return -1;                                            // (a table of 2 n slots cannot fill up)
}

// The fine splitters of the ranking bins, needed three kernels later, come from SPL_CTAS extra CTAs of the per-point count
// kernel, off the critical path: each stages the key sample in shared memory and ranks its share of it by counting
// (no sorting network, no barrier chain); the key of rank r is splitter r - 1.
constexpr int SPL_CTAS = SAMPLE / VOX_THREADS;

__device__ void fine_splitters(const float *__restrict__ points, int64_t n, int C, u64 *__restrict__ fine, int part)
{
    __shared__ u64 s[SAMPLE];
    const int tid = threadIdx.x;
    for (int i = tid; i < SAMPLE; i += VOX_THREADS) s[i] = sample_key(points, n, C, i, SAMPLE);
    __syncthreads();
    const int me = part * VOX_THREADS + tid;
    const u64 k = s[me];
    int rank = 0;
#pragma unroll 8
    for (int j = 0; j < SAMPLE; ++j) {
        count_if_less(rank, s[j], k);                         // (the sample keys are distinct, pads included)
    }
    if (rank >= 1) fine[rank - 1] = k;
}

// ---- A: per point, count ---------------------------------------------------------------------------------------
constexpr int CNT_IT = 4;      // points per thread: the four point loads, then the four atomics, are in flight together

template <typename K, bool HASH, bool VEC4, bool PERM>
__global__ void __launch_bounds__(VOX_THREADS, 7)      // 7 x 148 = 1036 CTAs resident: 1e6 points (977 CTAs) are one wave
vox_count_kernel(const float *__restrict__ points, int n, const VoxParams prm, const int32_t *__restrict__ perm,
                 int pt_blocks, const VoxBuf w, const FillArgs fa)
{
    // PDL: the points are read and binned into cells BEFORE the dependency wait, i.e. while the init kernel (workspace
    // fill, sample ranking) is still running; nothing the init kernel writes is touched before it.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr bool WIDE = sizeof(K) == 8;
    // the extra CTAs (64-bit keys) come first in the grid so that they start with the first wave; they read only caller memory
    const int nspl = (int)gridDim.x - pt_blocks, blk = (int)blockIdx.x - nspl;
    if (blk < 0) {
        fine_splitters(points, n, prm.C, w.fine, (int)blockIdx.x);
        return;
    }
    __shared__ u64 s_split[NCH];
    const int p0 = blk * (VOX_THREADS * CNT_IT) + threadIdx.x;
    int32_t cell[CNT_IT];
    float refl[CNT_IT];
#pragma unroll
    for (int k = 0; k < CNT_IT; ++k) {
        const int p = p0 + k * VOX_THREADS;
        cell[k] = -1;
        refl[k] = 0.f;
        if (p >= n) continue;
        const uint32_t idx = PERM ? (uint32_t)perm[p] : (uint32_t)p;
        float x, y, z;
        if (VEC4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(points) + idx);
            x = v.x; y = v.y; z = v.z; refl[k] = v.w;
        } else {
            const float *pt = points + (size_t)idx * prm.C;
            x = __ldg(pt); y = __ldg(pt + 1); z = __ldg(pt + 2);
            if (WIDE) refl[k] = __ldg(pt + 3);
        }
        cell[k] = point_cell(prm, x, y, z);
    }
    fill_slice(fa, 0, blk, pt_blocks);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    PP_T0(w.counters, 1);
    if (WIDE) {
        if (threadIdx.x < NCH) s_split[threadIdx.x] = threadIdx.x < NCH - 1 ? w.coarse[threadIdx.x] : ~0ull;
        __syncthreads();
    }
    uint32_t r[CNT_IT], hi[CNT_IT], tick[CNT_IT];
#pragma unroll
    for (int k = 0; k < CNT_IT; ++k) {
        const int p = p0 + k * VOX_THREADS;
        r[k] = 0xFFFFFFFFu; hi[k] = 0u; tick[k] = 0u;
        if (cell[k] < 0) continue;
        int chunk;
        if (WIDE) {
            hi[k] = refl_key_hi(refl[k]);
            const u64 key = ((u64)hi[k] << 32) | (uint32_t)p;
            chunk = 0;                                    // number of splitters <= key (entry 7 is +inf): three steps
#pragma unroll
            for (int step = NCH / 2; step > 0; step >>= 1) chunk += (s_split[chunk + step - 1] <= key) ? step : 0;
        } else {
            chunk = geo_bin<C_OCT, C_SUB>((uint32_t)p, prm.bits);
        }
        const int slot = HASH ? hash_insert(w, cell[k]) : cell[k];
        if (slot < 0) continue;
        tick[k] = atomicAdd(w.cnt + (size_t)slot * NCH + chunk, 1u);
        r[k] = ((uint32_t)slot << 3) | (uint32_t)chunk;
    }
#pragma unroll
    for (int k = 0; k < CNT_IT; ++k) {
        const int p = p0 + k * VOX_THREADS;
        if (p >= n) continue;
        w.rec_slot[p] = r[k];
        if (r[k] != 0xFFFFFFFFu) {
            w.rec_tick[p] = tick[k];
            if (WIDE) w.rec_hi[p] = hi[k];
        }
    }
    PP_T1(w.counters, 2);
}

// ---- B: per slot -------------------------------------------------------------------------------------------------
// counts -> saturation chunk (the first chunk at which the cell holds max_points points; later chunks are dropped
// unseen) and m = the points of the chunks up to it.  Segment offsets and the compact ids q come from CTA-wide prefix
// sums and one atomic per CTA and counter; the eight counters of the slot become the write positions of its chunks.
constexpr int CELLS_THREADS = 1024;

__global__ void __launch_bounds__(CELLS_THREADS) vox_cells_kernel(const VoxParams prm, const VoxBuf w)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    PP_T0(w.counters, 4);
    __shared__ u64 s_warp[CELLS_THREADS / 32];
    __shared__ u64 s_cta;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t = (int64_t)blockIdx.x * CELLS_THREADS + tid;
    uint32_t c[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) c[k] = 0u;
    if (t < w.T) {
        const uint4 a = *reinterpret_cast<const uint4 *>(w.cnt + (size_t)t * NCH);
        const uint4 b = *reinterpret_cast<const uint4 *>(w.cnt + (size_t)t * NCH + 4);
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    }
    uint32_t total = 0, m = 0, n0 = 0;
    int sat = NCH - 1;
    u64 mine = 0, incl = 0;
    // (nine warps in ten hold only empty slots on the benchmark tile: they skip the counting and the warp scan)
    if (__any_sync(0xFFFFFFFFu, (c[0] | c[1] | c[2] | c[3] | c[4] | c[5] | c[6] | c[7]) != 0u)) {
        bool found = false;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            if (n0 == 0) n0 = c[k];                       // keys in the first occupied chunk: the smallest key is one of them
            total += c[k];
            if (!found && total >= (uint32_t)prm.P) { found = true; sat = k; m = total; }
        }
        if (!found) m = total;
        // exclusive prefix over the CTA of (m, occupied) packed in one 64-bit word (m < 2^31, at most 1024 cells per CTA)
        mine = ((u64)m << 11) | (total > 0 ? 1ull : 0ull);
        incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned lo = __shfl_up_sync(0xFFFFFFFFu, (unsigned)incl, o), hi = __shfl_up_sync(0xFFFFFFFFu, (unsigned)(incl >> 32), o);
            if (lane >= o) incl += ((u64)hi << 32) | lo;
        }
    }
    const bool occ = total > 0;
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u64 v = s_warp[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned lo = __shfl_up_sync(0xFFFFFFFFu, (unsigned)iv, o), hi = __shfl_up_sync(0xFFFFFFFFu, (unsigned)(iv >> 32), o);
            if (lane >= o) iv += ((u64)hi << 32) | lo;
        }
        s_warp[lane] = iv - v;
        if (lane == 31) {
            const uint32_t cells = (uint32_t)(iv & 0x7FFull), keys = (uint32_t)(iv >> 11);
            uint32_t qb = 0, sb = 0;
            if (cells) {
                qb = (uint32_t)atomicAdd(w.counters + CTR_NQ, (int)cells);
                sb = (uint32_t)atomicAdd(w.counters + CTR_SEG, (int)keys);
            }
            s_cta = ((u64)sb << 32) | qb;
        }
    }
    __syncthreads();
    PP_T1(w.counters, 5);
    if (!occ) return;
    const u64 excl = s_warp[warp] + incl - mine;
    const uint32_t q = (uint32_t)s_cta + (uint32_t)(excl & 0x7FFull);
    const uint32_t off = (uint32_t)(s_cta >> 32) + (uint32_t)(excl >> 11);
    uint32_t pos[NCH], run = off;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        pos[k] = k <= sat ? run : DROPPED;
        run += c[k];
    }
    *reinterpret_cast<uint4 *>(w.cnt + (size_t)t * NCH) = make_uint4(pos[0], pos[1], pos[2], pos[3]);
    *reinterpret_cast<uint4 *>(w.cnt + (size_t)t * NCH + 4) = make_uint4(pos[4], pos[5], pos[6], pos[7]);
    const int cell = w.hash_bits ? w.slot_key[t] : (int)t;
    w.qinfo[q] = make_int4(cell, (int)off, (int)m, (int)n0);
    const int cx = cell % prm.g[0], tt = cell / prm.g[0];
    w.qxyz[q] = make_int4(cx, tt % prm.g[1], tt / prm.g[1], 0);
}

// ---- C: per point, place -------------------------------------------------------------------------------------------
constexpr int PLACE_IT = 4;

template <typename K>
__global__ void __launch_bounds__(VOX_THREADS)
vox_place_kernel(int n, const VoxBuf w, const FillArgs fa)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr bool WIDE = sizeof(K) == 8;
    fill_slice(fa, 1, blockIdx.x, gridDim.x);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    PP_T0(w.counters, 7);
    const int p0 = blockIdx.x * (VOX_THREADS * PLACE_IT) + threadIdx.x;
    // the three record streams are read up front (coalesced, independent); the only dependent access of a point is
    // the write position of its (cell, chunk)
    uint32_t r[PLACE_IT], tick[PLACE_IT], hi[PLACE_IT];
#pragma unroll
    for (int k = 0; k < PLACE_IT; ++k) {
        const int p = p0 + k * VOX_THREADS;
        r[k] = 0xFFFFFFFFu; tick[k] = 0u; hi[k] = 0u;
        if (p < n) {
            r[k] = __ldcs(w.rec_slot + p);
            tick[k] = __ldcs(w.rec_tick + p);
            if (WIDE) hi[k] = __ldcs(w.rec_hi + p);
        }
    }
    uint32_t pos[PLACE_IT];
#pragma unroll
    for (int k = 0; k < PLACE_IT; ++k)          // (slot << 3 | chunk) is the index of the chunk's write position
        pos[k] = r[k] != 0xFFFFFFFFu ? __ldg(w.cnt + r[k]) : DROPPED;
#pragma unroll
    for (int k = 0; k < PLACE_IT; ++k) {
        if (pos[k] == DROPPED) continue;
        const int p = p0 + k * VOX_THREADS;
        ((K *)w.seg)[pos[k] + tick[k]] = WIDE ? (K)(((u64)hi[k] << 32) | (uint32_t)p) : (K)(uint32_t)p;
    }
    PP_T1(w.counters, 8);
}

// ---- H: per cell, smallest key, ranking bin; the last CTA: bucket order and cutoff ---------------------------------
// Ranking bins, monotone in the key.  A key is first located on a scale of NFINE quantile intervals: interval i and a
// linear position frac in [0, 1) inside it (64-bit keys: the 1023 sample splitters + interpolation on the whole key, so
// keys that tie on the reflectance spread by their index; 32-bit keys: the position itself).  A cell with m points has
// its smallest key near the 1/m quantile and cell sizes span decades, so the bins are laid out logarithmically in
// x = i + frac + x0 (x0 = the mean 1/m in intervals): interval i still owns at least one bin.  The logarithm is the
// float's own bit pattern (monotone and exact in integer arithmetic: pillar ids must not depend on rounding).
struct BinPlan {
    float x0;
    int l0;
    unsigned long long mul;
};
__device__ __forceinline__ BinPlan bin_plan(int nq, int64_t n_points)
{
    BinPlan b;
    float x0 = (float)NFINE * (float)nq / (float)(n_points > 0 ? n_points : 1);
    b.x0 = x0 < 0.5f ? 0.5f : x0;
    b.l0 = __float_as_int(b.x0);
    const unsigned long long span = (unsigned long long)(__float_as_int((float)NFINE + b.x0) - b.l0);
    b.mul = ((unsigned long long)(NBIN - NFINE - 1) << 32) / span;
    return b;
}
__device__ __forceinline__ int log_bin(const BinPlan bp, int i, float frac)
{
    frac = frac < 0.f ? 0.f : (frac > 0.99999994f ? 0.99999994f : frac);
    const float x = ((float)i + bp.x0) + frac;                          // monotone in (i, frac)
    const unsigned long long l = (unsigned long long)(__float_as_int(x) - bp.l0);
    return i + (int)((l * bp.mul) >> 32);
}

__device__ __forceinline__ int wide_bin(const u64 *s_fine, const BinPlan bp, u64 f)
{
    int lo = 0, hi = NFINE - 1;                       // number of splitters <= f
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_fine[mid] <= f) lo = mid + 1; else hi = mid;
    }
    const int i = lo;
    // interval i = [bot, top): bot = splitter i - 1 (or an extrapolation below the first), top = splitter i (or +inf)
    u64 bot, top;
    // (the outermost intervals are open: their far edge is extrapolated by twice the mean width of the 16 nearest
    // intervals -- single intervals of a sample are far too irregular for that)
    if (i > 0) bot = s_fine[i - 1];
    else { const u64 wd = (s_fine[16] - s_fine[0]) / 8; bot = s_fine[0] > wd ? s_fine[0] - wd : 0ull; }
    if (i < NFINE - 1) top = s_fine[i];
    else { const u64 wd = (s_fine[NFINE - 2] - s_fine[NFINE - 18]) / 8; top = bot + wd > bot ? bot + wd : ~0ull; }
    float frac = 0.f;
    if (f > bot) frac = top > bot ? (float)(f - bot) / ((float)(top - bot) + 1.0f) : 0.f;      // monotone in f
    return log_bin(bp, i, frac);
}

__device__ __forceinline__ int narrow_bin(const BinPlan bp, uint32_t p, int bits)
{
    // positions < 2^bits: the top 9 bits are the interval (NFINE = 2^9), the rest the fraction
    if (bits <= 9) return log_bin(bp, (int)(p << (9 - bits)), 0.f);
    const int sh = bits - 9;
    return log_bin(bp, (int)(p >> sh), (float)(p & ((1u << sh) - 1u)) / (float)(1u << sh));
}

constexpr int RANK_THREADS = 1024;

template <typename K>
__global__ void __launch_bounds__(RANK_THREADS)
vox_rank_kernel(const VoxParams prm, const VoxBuf w, int64_t n_points)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    PP_T0(w.counters, 10);
    constexpr bool WIDE = sizeof(K) == 8;
    __shared__ u64 s_fine[NFINE];
    const int tid = threadIdx.x, lane = tid & 31;
    const int nq = w.counters[CTR_NQ];
    const K *seg = (const K *)w.seg;
    constexpr int GL = 8, CPB = RANK_THREADS / GL;    // 8 lanes per cell: the keys of the first chunk are read side by side
    if ((int64_t)blockIdx.x * CPB < nq) {
        if (WIDE) {
            for (int i = tid; i < NFINE - 1; i += RANK_THREADS) s_fine[i] = w.fine[i];
            if (tid == 0) s_fine[NFINE - 1] = ~0ull;
            __syncthreads();
        }
        const BinPlan bp = bin_plan(nq, n_points);
        const int sub = lane & (GL - 1);
        for (int q0 = blockIdx.x * CPB; q0 < nq; q0 += gridDim.x * CPB) {
            const int q = q0 + tid / GL;
            K mn = KeyInf<K>::value();
            if (q < nq) {
                const int4 info = w.qinfo[q];
                for (int j = sub; j < info.w; j += GL) {      // the segment is ordered by chunk: the first chunk holds the minimum
                    const K k = seg[info.y + j];
                    mn = k < mn ? k : mn;
                }
            }
#pragma unroll
            for (int o = GL / 2; o > 0; o >>= 1) {
                K x;
                if (WIDE) {
                    const unsigned lo = __shfl_xor_sync(0xFFFFFFFFu, (unsigned)mn, o), hi = __shfl_xor_sync(0xFFFFFFFFu, (unsigned)((u64)mn >> 32), o);
                    x = (K)(((u64)hi << 32) | lo);
                } else {
                    x = (K)__shfl_xor_sync(0xFFFFFFFFu, (unsigned)mn, o);
                }
                mn = x < mn ? x : mn;
            }
            if (q < nq && sub == 0) {
                const int bin = WIDE ? wide_bin(s_fine, bp, (u64)mn) : narrow_bin(bp, (uint32_t)mn, prm.bits);
                ((K *)w.first)[q] = mn;
                w.bin_of_q[q] = bin;
                w.arr_of_q[q] = atomicAdd(w.hist + bin, 1);
            }
        }
    }
    PP_T1(w.counters, 11);
}

// ---- R: cells in bucket order; cutoff -------------------------------------------------------------------------------
// Every cell's first key goes to lkey[base[bin] + arrival index]: the cells of a bin are contiguous, so the gather
// kernel ranks a cell against its bin with coalesced reads whatever the distribution of the keys.  The last CTA to
// finish settles the cutoff.
constexpr int BUCKET_THREADS = 1024;

template <typename K>
__global__ void __launch_bounds__(BUCKET_THREADS)
vox_bucket_kernel(const VoxParams prm, const VoxBuf w, int32_t *__restrict__ voxel_num)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    PP_T0(w.counters, 13);
    __shared__ int s_base[NBIN + 1];
    __shared__ int s_warp[BUCKET_THREADS / 32];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nq = w.counters[CTR_NQ];
    const bool work = (int64_t)blockIdx.x * BUCKET_THREADS < nq;
    if (work || blockIdx.x == 0) {
        // exclusive scan of the bin histogram, every CTA for itself (warp w scans bins [w, w + 1) * PERW, coalesced)
        constexpr int PERW = NBIN / (BUCKET_THREADS / 32), ITS = PERW / 32;
        int ex[ITS], carry = 0;
#pragma unroll
        for (int it = 0; it < ITS; ++it) {
            const int v = w.hist[warp * PERW + it * 32 + lane];
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            ex[it] = carry + incl - v;
            carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (lane == 0) s_warp[warp] = carry;
        __syncthreads();
        int wbase = 0;
        for (int k = 0; k < warp; ++k) wbase += s_warp[k];
#pragma unroll
        for (int it = 0; it < ITS; ++it) s_base[warp * PERW + it * 32 + lane] = wbase + ex[it];
        if (tid == BUCKET_THREADS - 1) s_base[NBIN] = wbase + carry;
        __syncthreads();
        if (blockIdx.x == 0) {
            for (int i = tid; i <= NBIN; i += BUCKET_THREADS) w.base[i] = s_base[i];
            if (tid == 0) *voxel_num = nq < prm.max_voxels ? nq : prm.max_voxels;
        }
    }
    K *lkey = (K *)w.lkey;
    for (int q = blockIdx.x * BUCKET_THREADS + tid; q < nq; q += gridDim.x * BUCKET_THREADS)
        lkey[s_base[w.bin_of_q[q]] + w.arr_of_q[q]] = ((const K *)w.first)[q];
    PP_T1(w.counters, 14);
    if (nq <= prm.max_voxels) {
        if (blockIdx.x == 0 && tid == 0) *(K *)w.cutoff = KeyInf<K>::value();
        return;
    }
    // The reference breaks at the first point that would open pillar max_voxels + 1 (:223, :291): its key -- the cell
    // minimum of rank max_voxels -- is the cutoff; it sits in the bin that holds that rank.  The last CTA to finish
    // (one that scanned the histogram: CTA 0 always does) looks it up.
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(w.counters + CTR_DONE, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (!(work || blockIdx.x == 0)) {                 // (a CTA without cells did not scan: read CTA 0's copy)
        for (int i = tid; i <= NBIN; i += BUCKET_THREADS) s_base[i] = __ldcg(w.base + i);
        __syncthreads();
    }
    int lo = 0, hi = NBIN;                               // largest bin b with base[b] <= max_voxels
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_base[mid] <= prm.max_voxels) lo = mid; else hi = mid - 1;
    }
    // base[NBIN] = nq > max_voxels, so lo < NBIN and base[lo + 1] > max_voxels: bin lo holds the rank max_voxels
    const int b0 = s_base[lo], b1 = s_base[lo + 1], target = prm.max_voxels - b0;
    for (int i = b0 + tid; i < b1; i += BUCKET_THREADS) {
        const K f = __ldcg(lkey + i);
        int c = 0;
        for (int j = b0; j < b1; ++j) c += (__ldcg(lkey + j) < f) ? 1 : 0;
        if (c == target) *(K *)w.cutoff = f;
    }
}

// ---- D: gather -----------------------------------------------------------------------------------------------------
template <typename K>
__device__ __forceinline__ K shfl_xor_key(K v, int m)
{
    if (sizeof(K) == 8) {
        unsigned lo = __shfl_xor_sync(0xFFFFFFFFu, (unsigned)v, m), hi = __shfl_xor_sync(0xFFFFFFFFu, (unsigned)((u64)v >> 32), m);
        return (K)(((u64)hi << 32) | lo);
    }
    return (K)__shfl_xor_sync(0xFFFFFFFFu, (unsigned)v, m);
}

template <typename K>
__device__ __forceinline__ K shfl_idx_key(K v, int src)
{
    if (sizeof(K) == 8) {
        unsigned lo = __shfl_sync(0xFFFFFFFFu, (unsigned)v, src), hi = __shfl_sync(0xFFFFFFFFu, (unsigned)((u64)v >> 32), src);
        return (K)(((u64)hi << 32) | lo);
    }
    return (K)__shfl_sync(0xFFFFFFFFu, (unsigned)v, src);
}

// ascending bitonic sort of one key per lane
template <typename K>
__device__ __forceinline__ K warp_sort32(K k, int lane)
{
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int j = size >> 1; j > 0; j >>= 1) {
            const K o = shfl_xor_key<K>(k, j);
            const bool take_min = ((lane & size) == 0) != ((lane & j) != 0);
            k = take_min ? (k < o ? k : o) : (k < o ? o : k);
        }
    }
    return k;
}

// ascending merge of a bitonic sequence held one key per lane
template <typename K>
__device__ __forceinline__ K warp_bitonic_merge(K m, int lane)
{
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const K x = shfl_xor_key<K>(m, j);
        m = (lane & j) ? (m < x ? x : m) : (m < x ? m : x);
    }
    return m;
}

// a, c: two ascending 32-key sequences -> a = the 32 smallest, c = the 32 largest, both ascending
template <typename K>
__device__ __forceinline__ void warp_merge_split(K &a, K &c, int lane)
{
    const K o = shfl_idx_key<K>(c, 31 - lane);
    const K lo = a < o ? a : o, hi = a < o ? o : a;       // both bitonic
    a = warp_bitonic_merge<K>(lo, lane);
    c = warp_bitonic_merge<K>(hi, lane);
}

constexpr int SEL_SMEM = 128;        // segments up to this many keys are ranked by counting in shared memory

// Segment of at most NK * 32 keys: every key's rank = number of smaller keys of the segment (broadcast reads of the
// staged segment: no shuffles, no sorting network), then the keys are put in rank order through shared memory.
template <typename K, int R, int NK>
__device__ __forceinline__ void rank_select(const K *__restrict__ seg, int m, int lane, K *s_keys, const K (&pre)[2], K (&b)[R])
{
    K k[NK];
#pragma unroll
    for (int t = 0; t < NK; ++t) {
        k[t] = t < 2 ? pre[t] : ((t * 32 + lane < m) ? seg[t * 32 + lane] : KeyInf<K>::value());    // keys 0 .. 63 were prefetched
        s_keys[t * 32 + lane] = k[t];
    }
    __syncwarp();
    int rank[NK];
#pragma unroll
    for (int t = 0; t < NK; ++t) rank[t] = 0;
#pragma unroll 4
    for (int j = 0; j < m; ++j) {
        const K o = s_keys[j];
#pragma unroll
        for (int t = 0; t < NK; ++t) count_if_less(rank[t], o, k[t]);
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NK; ++t)
        if (t * 32 + lane < m && rank[t] < R * 32) s_keys[rank[t]] = k[t];      // ranks are a permutation: no collisions
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (r * 32 + lane < m) b[r] = s_keys[r * 32 + lane];
    __syncwarp();
}


// The R * 32 smallest keys of seg[0, m), ascending: element e = r * 32 + lane is b[r].
//   m <= SEL_SMEM  every key's rank = number of smaller keys of the segment (broadcast reads of the staged segment; no
//                  shuffles, no sorting network); the keys are then put in rank order through shared memory
//   longer         batches of 32 keys are sorted and merged down the registers; a batch with no key below the current
//                  last element is skipped after one ballot
template <typename K, int R>
__device__ __forceinline__ void select_smallest(const K *__restrict__ seg, int m, int lane, K *s_keys, const K (&pre)[2], K (&b)[R])
{
#pragma unroll
    for (int r = 0; r < R; ++r) b[r] = KeyInf<K>::value();
    if (m <= 32) { rank_select<K, R, 1>(seg, m, lane, s_keys, pre, b); return; }
    if (m <= 64) { rank_select<K, R, 2>(seg, m, lane, s_keys, pre, b); return; }
    if (m <= SEL_SMEM) { rank_select<K, R, SEL_SMEM / 32>(seg, m, lane, s_keys, pre, b); return; }
    for (int i0 = 0; i0 < m; i0 += 32) {
        K c = (i0 + lane < m) ? seg[i0 + lane] : KeyInf<K>::value();
        if (i0 == 0) {
            b[0] = warp_sort32<K>(c, lane);
            continue;
        }
        const K thr = shfl_idx_key<K>(b[R - 1], 31);
        if (__ballot_sync(0xFFFFFFFFu, c < thr) == 0u) continue;
        c = warp_sort32<K>(c, lane);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const K cmin = shfl_idx_key<K>(c, 0), bmax = shfl_idx_key<K>(b[r], 31);
            if (cmin == KeyInf<K>::value()) break;
            if (!(cmin < bmax)) continue;                  // the whole batch lies after this register
            warp_merge_split<K>(b[r], c, lane);
        }
    }
}

struct PfnArgs {
    const float *W, *scale, *shift;
    float *feat;          // (rows, U + 1) or nullptr
    float *canvas;        // (1, (U + 1) * D, H, W) of this frame or nullptr
    int64_t plane;        // D * H * W: elements between two PFN channels of the canvas
    int U;
    float vx, vy, x_off, y_off;
};

struct GatherOut {
    float *voxels;
    int32_t *coors, *num_points, *pillar_map;
};

constexpr int GP_THREADS = 128;      // 4 pillars per CTA

// pillar id of cell q = base of its bin + number of smaller first keys in the bin (the whole warp counts)
template <typename K>
__device__ __forceinline__ int pillar_rank(const VoxBuf &w, int q, int lane)
{
    const K f = ((const K *)w.first)[q];
    const int bin = w.bin_of_q[q];
    const int b0 = w.base[bin], b1 = w.base[bin + 1];
    if (b1 - b0 == 1) return b0;                          // alone in its bin (the common case)
    const K *lkey = (const K *)w.lkey;
    int cnt = 0;
    for (int j = b0 + lane; j < b1; j += 32) cnt += (lkey[j] < f) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    return b0 + cnt;
}

__device__ __forceinline__ void write_coors(const GatherOut &out, int pid, int cell, const int4 xyz)
{
    out.coors[pid * 3 + 0] = xyz.x;
    out.coors[pid * 3 + 1] = xyz.y;
    out.coors[pid * 3 + 2] = xyz.z;
    if (out.pillar_map) out.pillar_map[cell] = pid;
}

// One warp per occupied cell, max_points <= 32 * R.  PFN (R == 1, C == 4): the warp that has just gathered a pillar
// holds its points in registers in exactly the layout the PillarFeatureNet starts from (lane = slot), so it also runs
// the pillar through pfn_pillar: no PFN launch, no re-read of the voxels.
template <typename K, int R, bool PFN>
__global__ void __launch_bounds__(GP_THREADS, PFN ? 8 : 4)
vox_gather_kernel(const float *__restrict__ points, const int32_t *__restrict__ perm, const VoxParams prm, const VoxBuf w,
                  const GatherOut out, const PfnArgs pa)
{
    // the weights do not depend on the predecessor: load them before the dependency wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ __align__(16) float s_row[PFN ? (GP_THREADS / 32) * 32 * 4 : 4];
    __shared__ __align__(16) K s_sel[(GP_THREADS / 32) * SEL_SMEM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PfnWeights pw;
    if (PFN) pfn_load_weights(pw, pa.W, pa.scale, pa.shift, pa.U, 4, lane);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    PP_T0(w.counters, 16);
    const int nq = w.counters[CTR_NQ];
    const K cutoff = *(const K *)w.cutoff;
    const int P = prm.P, C = prm.C;
    for (int q = blockIdx.x * (GP_THREADS / 32) + warp; q < nq; q += gridDim.x * (GP_THREADS / 32)) {
        const int4 info = w.qinfo[q], xyz = w.qxyz[q];
        // the first 64 keys of the segment are requested before the ranking loads: the two chains of dependent reads
        // (cell -> keys, cell -> bin -> bucket) overlap
        const K *seg = (const K *)w.seg + info.y;
        K pre[2];
        pre[0] = lane < info.z ? seg[lane] : KeyInf<K>::value();
        pre[1] = lane + 32 < info.z ? seg[lane + 32] : KeyInf<K>::value();
        const int pid = pillar_rank<K>(w, q, lane);
        if (pid >= prm.max_voxels) continue;
        K b[R];
        select_smallest<K, R>(seg, info.z, lane, s_sel + warp * SEL_SMEM, pre, b);
        int n = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            // keys at or after the cutoff are dropped: they are the largest keys of their rows
            const bool valid = r * 32 + lane < P && b[r] < cutoff;
            n += __popc(__ballot_sync(0xFFFFFFFFu, valid));
            if (!valid) b[r] = KeyInf<K>::value();
        }
        if (lane == 0) out.num_points[pid] = n;
        if (lane == 1) write_coors(out, pid, info.x, xyz);
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int sl = r * 32 + lane;
            if (sl >= P) continue;
            const bool valid = b[r] != KeyInf<K>::value();
            int64_t idx = 0;
            if (valid) {
                const uint32_t pos = (uint32_t)b[r];                 // low word = position / original index
                idx = perm ? (int64_t)(uint32_t)perm[pos] : (int64_t)pos;
            }
            const int64_t t = (int64_t)pid * P + sl;
            if (prm.vec4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) v = __ldg(reinterpret_cast<const float4 *>(points) + idx);
                reinterpret_cast<float4 *>(out.voxels)[t] = v;
                if (r == 0) v0 = v;
            } else {
                for (int c = 0; c < C; ++c) out.voxels[t * C + c] = valid ? __ldg(points + idx * C + c) : 0.f;
            }
        }
        if (PFN) {
            const int cell = info.x;
            // channel c of cell (z, y, x) sits at canvas[(c * D + z) * H * W + y * W + x] = canvas[c * plane + cell]
            pfn_pillar4(pw, v0, P, n, xyz.x, xyz.y, pa.vx, pa.vy, pa.x_off, pa.y_off,
                        s_row + warp * 32 * 4, pa.feat ? pa.feat + (int64_t)pid * (pa.U + 1) : nullptr,
                        pa.canvas ? pa.canvas + cell : nullptr, pa.plane, pa.U, lane);
        }
    }
    PP_T1(w.counters, 17);
}

// Any max_points: the rank of every key of the segment by counting (quadratic in the segment, which the chunk filter
// keeps near 2 x max_points).
template <typename K>
__global__ void __launch_bounds__(GP_THREADS)
vox_gather_any_kernel(const float *__restrict__ points, const int32_t *__restrict__ perm, const VoxParams prm,
                      const VoxBuf w, const GatherOut out)
{
    pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nq = w.counters[CTR_NQ];
    const K cutoff = *(const K *)w.cutoff;
    const int P = prm.P, C = prm.C;
    for (int q = blockIdx.x * (GP_THREADS / 32) + warp; q < nq; q += gridDim.x * (GP_THREADS / 32)) {
        const int4 info = w.qinfo[q], xyz = w.qxyz[q];
        const int pid = pillar_rank<K>(w, q, lane);
        if (pid >= prm.max_voxels) continue;
        const K *seg = (const K *)w.seg + info.y;
        const int m = info.z;
        int kept = 0;
        for (int i0 = 0; i0 < m; i0 += 32) {
            const int i = i0 + lane;
            const K k = i < m ? seg[i] : KeyInf<K>::value();
            int rank = 0;
            for (int j = 0; j < m; ++j) rank += (seg[j] < k) ? 1 : 0;
            const bool valid = i < m && rank < P && k < cutoff;
            kept += __popc(__ballot_sync(0xFFFFFFFFu, valid));
            if (!valid) continue;
            const uint32_t pos = (uint32_t)k;
            const int64_t idx = perm ? (int64_t)(uint32_t)perm[pos] : (int64_t)pos;
            float *dst = out.voxels + ((int64_t)pid * P + rank) * C;
            for (int c = 0; c < C; ++c) dst[c] = __ldg(points + idx * C + c);
        }
        // ranks [0, kept) were written (the kept keys are the smallest); the padding is zero
        for (int i = kept * C + lane; i < P * C; i += 32) out.voxels[(int64_t)pid * P * C + i] = 0.f;
        if (lane == 0) out.num_points[pid] = kept;
        if (lane == 1) write_coors(out, pid, info.x, xyz);
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
int64_t max_rows_of(int64_t n, const pp_voxel_cfg *c)
{
    int64_t cells = (int64_t)c->grid[0] * c->grid[1] * c->grid[2];
    int64_t r = c->max_voxels;
    if (n < r) r = n;
    if (cells < r) r = cells;
    return r > 0 ? r : 1;
}

struct Carve {
    VoxBuf b;
    InitArgs ia;
    int64_t Q;
};

inline int64_t units16(size_t bytes) { return (int64_t)(align_up(bytes, 16) / 16); }

// Everything here depends on (n, cfg, key width) only, so the workspace size and the launch sequence agree.
Carve carve(void *ws, int64_t n, const pp_voxel_cfg *c, bool wide, size_t *total)
{
    Carve r;
    const int64_t n1 = n > 0 ? n : 1;
    const int64_t cells = (int64_t)c->grid[0] * c->grid[1] * c->grid[2];
    r.Q = n1 < cells ? n1 : cells;
    const size_t ksz = wide ? 8 : 4;
    // direct addressing when the grid is not larger than the hash table would be (2 n slots, rounded up to a power of two)
    int hb = 10;
    while (((int64_t)1 << hb) < 2 * n1) ++hb;
    const bool direct = cells <= ((int64_t)1 << hb);
    r.b.hash_bits = direct ? 0 : hb;
    r.b.T = (int32_t)(direct ? cells : ((int64_t)1 << hb));
    Arena a(ws, (size_t)-1);
    r.b.slot_key = a.take<int32_t>(direct ? 4 : (size_t)r.b.T);
    r.b.cutoff = a.take<u64>(2);
    const size_t ff0_bytes = a.off;                          // hash keys + cutoff, contiguous, 0xFF
    r.b.counters = a.take<int32_t>(64);
    r.b.hist = a.take<int32_t>(NBIN);
    const size_t z0_off = (size_t)((char *)r.b.counters - (char *)ws), z0_bytes = a.off - z0_off;
    r.b.cnt = a.take<uint32_t>((size_t)r.b.T * NCH);
    const size_t z1_bytes = (size_t)r.b.T * NCH * sizeof(uint32_t);
    r.b.base = a.take<int32_t>(NBIN + 1);
    r.b.rec_slot = a.take<uint32_t>((size_t)n1);
    r.b.rec_tick = a.take<uint32_t>((size_t)n1);
    r.b.rec_hi = wide ? a.take<uint32_t>((size_t)n1) : nullptr;
    r.b.seg = a.take<char>((size_t)n1 * ksz);
    r.b.qinfo = a.take<int4>((size_t)r.Q);
    r.b.qxyz = a.take<int4>((size_t)r.Q);
    r.b.first = a.take<char>((size_t)r.Q * ksz);
    r.b.bin_of_q = a.take<int32_t>((size_t)r.Q);
    r.b.arr_of_q = a.take<int32_t>((size_t)r.Q);
    r.b.lkey = a.take<char>((size_t)r.Q * ksz);
    r.b.coarse = a.take<u64>(NCH);
    r.b.fine = a.take<u64>(NFINE);
    *total = align_up(a.off);
    // every array starts 256-byte aligned, so rounding the fills up to 16 bytes stays inside the padding
    r.ia.ff_ptr[0] = (int4 *)r.b.slot_key;  r.ia.ff_n[0] = units16(ff0_bytes);
    r.ia.ff_ptr[1] = nullptr;               r.ia.ff_n[1] = 0;      // optional pillar_map, set by the caller
    r.ia.ff_ptr[2] = nullptr;               r.ia.ff_n[2] = 0;
    r.ia.z_ptr[0] = (int4 *)((char *)ws + z0_off);  r.ia.z_n[0] = units16(z0_bytes);
    r.ia.z_ptr[1] = (int4 *)r.b.cnt;        r.ia.z_n[1] = units16(z1_bytes);
    return r;
}

template <typename K, bool HASH>
void launch_count(int pt_blocks, cudaStream_t st, const float *points, int n, const VoxParams &prm, const int32_t *perm,
                  const VoxBuf &w, const FillArgs &fa)
{
    const dim3 blk(VOX_THREADS), grid(pt_blocks + (sizeof(K) == 8 ? SPL_CTAS : 0));
    if (prm.vec4) {
        if (perm) launch_pdl(vox_count_kernel<K, HASH, true, true>, grid, blk, 0, st, points, n, prm, perm, pt_blocks, w, fa);
        else launch_pdl(vox_count_kernel<K, HASH, true, false>, grid, blk, 0, st, points, n, prm, perm, pt_blocks, w, fa);
    } else {
        if (perm) launch_pdl(vox_count_kernel<K, HASH, false, true>, grid, blk, 0, st, points, n, prm, perm, pt_blocks, w, fa);
        else launch_pdl(vox_count_kernel<K, HASH, false, false>, grid, blk, 0, st, points, n, prm, perm, pt_blocks, w, fa);
    }
}

template <typename K>
int run(const float *points, int64_t n, const VoxParams &prm, const int32_t *perm, const Carve &cv, float *voxels,
        int32_t *coors, int32_t *num_points, int32_t *voxel_num, int32_t *pillar_map, const PfnArgs *pfn, cudaStream_t st)
{
    const VoxBuf &w = cv.b;
    constexpr bool WIDE = sizeof(K) == 8;
    FillArgs fa = {nullptr, 0};
    if (pfn && pfn->canvas) {
        fa.base = pfn->canvas;
        fa.units = (int64_t)(pfn->U + 1) * pfn->plane / 8;
    }
    int64_t fill_units = 0;
    for (int r = 0; r < 3; ++r) fill_units += cv.ia.ff_n[r];
    for (int r = 0; r < 2; ++r) fill_units += cv.ia.z_n[r];
    int init_blocks = (int)ceil_div(fill_units, 1024 * 4);
    init_blocks = init_blocks < 1 ? 1 : (init_blocks > 148 * 2 ? 148 * 2 : init_blocks);
    launch_pdl(vox_init_kernel, dim3(init_blocks), dim3(1024), 0, st, points, n, prm.C, WIDE ? 1 : 0, w.coarse, cv.ia);
    if (int rc = check_launch("vox_init_kernel")) return rc;
    const int pt_blocks = (int)ceil_div(n, VOX_THREADS * CNT_IT);
    if (w.hash_bits) launch_count<K, true>(pt_blocks, st, points, (int)n, prm, perm, w, fa);
    else launch_count<K, false>(pt_blocks, st, points, (int)n, prm, perm, w, fa);
    if (int rc = check_launch("vox_count_kernel")) return rc;
    launch_pdl(vox_cells_kernel, dim3((unsigned)ceil_div(w.T, CELLS_THREADS)), dim3(CELLS_THREADS), 0, st, prm, w);
    if (int rc = check_launch("vox_cells_kernel")) return rc;
    const int pl_blocks = (int)ceil_div(n, VOX_THREADS * PLACE_IT);
    launch_pdl(vox_place_kernel<K>, dim3(pl_blocks), dim3(VOX_THREADS), 0, st, (int)n, w, fa);
    if (int rc = check_launch("vox_place_kernel")) return rc;
    // the cell count is only known on the device: grid-stride grids sized for the worst case, capped at a few waves
    auto capped = [](int64_t blocks, int64_t cap) { return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap); };
    launch_pdl(vox_rank_kernel<K>, dim3(capped(ceil_div(cv.Q, RANK_THREADS / 8), 148 * 2)), dim3(RANK_THREADS), 0, st, prm, w, n);
    if (int rc = check_launch("vox_rank_kernel")) return rc;
    launch_pdl(vox_bucket_kernel<K>, dim3(capped(ceil_div(cv.Q, BUCKET_THREADS), 148)), dim3(BUCKET_THREADS), 0, st, prm, w, voxel_num);
    if (int rc = check_launch("vox_bucket_kernel")) return rc;
    GatherOut out = {voxels, coors, num_points, pillar_map};
    PfnArgs pa = {};
    if (pfn) pa = *pfn;
    const dim3 gg(capped(ceil_div(cv.Q, GP_THREADS / 32), 148 * 16));
    if (pfn) {
        launch_pdl(vox_gather_kernel<K, 1, true>, gg, dim3(GP_THREADS), 0, st, points, perm, prm, w, out, pa);
        return check_launch("vox_gather_pfn_kernel");
    }
    if (prm.P <= 32)
        launch_pdl(vox_gather_kernel<K, 1, false>, gg, dim3(GP_THREADS), 0, st, points, perm, prm, w, out, pa);
    else if (prm.P <= 64)
        launch_pdl(vox_gather_kernel<K, 2, false>, gg, dim3(GP_THREADS), 0, st, points, perm, prm, w, out, pa);
    else
        launch_pdl(vox_gather_any_kernel<K>, gg, dim3(GP_THREADS), 0, st, points, perm, prm, w, out);
    return check_launch("vox_gather_kernel");
}

}  // namespace
}  // namespace pp

using namespace pp;

extern "C" int64_t pp_voxelize_max_rows(int64_t n_points, const pp_voxel_cfg *cfg)
{
    if (!cfg) return 0;
    return max_rows_of(n_points, cfg);
}

extern "C" size_t pp_voxelize_workspace_bytes(int64_t n_points, const pp_voxel_cfg *cfg, int order)
{
    if (!cfg) return 0;
    size_t total;
    carve(nullptr, n_points, cfg, order == PP_ORDER_REFLECTANCE_DESC, &total);
    return total;
}

static int voxelize_impl(const float *points, int64_t n, const pp_voxel_cfg *cfg, int order, const int32_t *perm,
                         float *voxels, int32_t *coors, int32_t *num_points, int32_t *voxel_num, int32_t *pillar_map,
                         const pp_pfn_fused *pfn, float *canvas, void *workspace, size_t workspace_bytes,
                         pp_stream_t stream);

extern "C" int pp_voxelize(const float *points, int64_t n, const pp_voxel_cfg *cfg, int order, const int32_t *perm,
                           float *voxels, int32_t *coors, int32_t *num_points, int32_t *voxel_num,
                           int32_t *pillar_map, void *workspace, size_t workspace_bytes, pp_stream_t stream)
{
    return voxelize_impl(points, n, cfg, order, perm, voxels, coors, num_points, voxel_num, pillar_map, nullptr, nullptr,
                         workspace, workspace_bytes, stream);
}

extern "C" int pp_voxelize_features(const float *points, int64_t n, const pp_voxel_cfg *cfg, int order,
                                    const int32_t *perm, float *voxels, int32_t *coors, int32_t *num_points,
                                    int32_t *voxel_num, int32_t *pillar_map, const pp_pfn_fused *pfn, void *workspace,
                                    size_t workspace_bytes, pp_stream_t stream)
{
    PP_REQUIRE(pfn && pfn->weight && pfn->scale && pfn->shift && pfn->feat, "null PFN arguments");
    PP_REQUIRE(cfg && cfg->num_feats == 4 && cfg->max_points <= 32 && pfn->units >= 1 && pfn->units <= 64,
               "the fused form needs C == 4, max_points <= 32, units <= 64 (else pp_voxelize + pp_pillar_features)");
    PP_REQUIRE(((uintptr_t)points % 16 == 0) && ((uintptr_t)voxels % 16 == 0), "points / voxels must be 16-byte aligned");
    return voxelize_impl(points, n, cfg, order, perm, voxels, coors, num_points, voxel_num, pillar_map, pfn, nullptr,
                         workspace, workspace_bytes, stream);
}

extern "C" int pp_voxelize_scatter(const float *points, int64_t n, const pp_voxel_cfg *cfg, int order,
                                   const int32_t *perm, float *voxels, int32_t *coors, int32_t *num_points,
                                   int32_t *voxel_num, int32_t *pillar_map, const pp_pfn_fused *pfn, float *canvas,
                                   void *workspace, size_t workspace_bytes, pp_stream_t stream)
{
    PP_REQUIRE(pfn && pfn->weight && pfn->scale && pfn->shift && canvas, "null PFN arguments / canvas");
    PP_REQUIRE(cfg && cfg->num_feats == 4 && cfg->max_points <= 32 && pfn->units >= 1 && pfn->units <= 64,
               "the fused form needs C == 4, max_points <= 32, units <= 64 (else pp_voxelize + pp_pillar_features + pp_scatter_dense)");
    PP_REQUIRE(((uintptr_t)points % 16 == 0) && ((uintptr_t)voxels % 16 == 0), "points / voxels must be 16-byte aligned");
    const int64_t plane = (int64_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
    PP_REQUIRE(((uintptr_t)canvas % 32 == 0) && (((int64_t)(pfn->units + 1) * plane) % 8 == 0),
               "canvas must be 32-byte aligned and a multiple of 32 bytes (else the unfused calls)");
    return voxelize_impl(points, n, cfg, order, perm, voxels, coors, num_points, voxel_num, pillar_map, pfn, canvas,
                         workspace, workspace_bytes, stream);
}

static int voxelize_impl(const float *points, int64_t n, const pp_voxel_cfg *cfg, int order, const int32_t *perm,
                         float *voxels, int32_t *coors, int32_t *num_points, int32_t *voxel_num, int32_t *pillar_map,
                         const pp_pfn_fused *pfn, float *canvas, void *workspace, size_t workspace_bytes,
                         pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    PP_REQUIRE(cfg && voxel_num, "null cfg / voxel_num");
    PP_REQUIRE(n >= 0 && n < (1ll << 27), "n_points out of range (< 2^27)");
    PP_REQUIRE(cfg->num_feats >= 3, "points need at least x, y, z");
    PP_REQUIRE(order == PP_ORDER_GIVEN || order == PP_ORDER_REFLECTANCE_DESC || order == PP_ORDER_PERM, "bad order");
    PP_REQUIRE(order != PP_ORDER_REFLECTANCE_DESC || cfg->num_feats >= 4, "reflectance order needs >= 4 features");
    PP_REQUIRE(order != PP_ORDER_PERM || perm, "PP_ORDER_PERM needs perm");
    PP_REQUIRE(cfg->max_points > 0 && cfg->max_voxels >= 0, "bad caps");
    PP_REQUIRE(cfg->grid[0] > 0 && cfg->grid[1] > 0 && cfg->grid[2] > 0, "empty grid");
    const int64_t cells = (int64_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
    PP_REQUIRE(cells < (1ll << 31), "grid too large (>= 2^31 cells)");
    if (n == 0 || cfg->max_voxels == 0) {
        PP_CUDA_TRY(cudaMemsetAsync(voxel_num, 0, sizeof(int32_t), st));
        if (pillar_map) PP_CUDA_TRY(cudaMemsetAsync(pillar_map, 0xFF, (size_t)cells * 4, st));
        if (canvas) PP_CUDA_TRY(cudaMemsetAsync(canvas, 0, (size_t)(pfn->units + 1) * cells * sizeof(float), st));
        return PP_OK;
    }
    PP_REQUIRE(points && voxels && coors && num_points && workspace, "null pointer");

    const bool wide = order == PP_ORDER_REFLECTANCE_DESC;
    size_t total;
    Carve cv = carve(workspace, n, cfg, wide, &total);
    if (workspace_bytes < total) {
        set_error("voxelize workspace too small: %zu < %zu", workspace_bytes, total);
        return PP_ERR_WORKSPACE;
    }

    VoxParams q;
    for (int j = 0; j < 3; ++j) {
        q.r[j] = cfg->range[j];
        q.v[j] = cfg->vsize[j];
        q.rf[j] = (float)cfg->range[j];
        q.vf[j] = (float)cfg->vsize[j];
        q.inv_vf[j] = 1.0f / q.vf[j];
        q.rv_abs[j] = (float)(fabs(q.r[j]) / q.v[j]);
        q.g[j] = cfg->grid[j];
    }
    q.regime = cfg->range_is_f64 ? 2 : (cfg->vsize_is_f64 ? 1 : 0);
    q.P = cfg->max_points;
    q.max_voxels = cfg->max_voxels;
    q.C = cfg->num_feats;
    q.vec4 = (q.C == 4 && ((uintptr_t)points % 16 == 0) && ((uintptr_t)voxels % 16 == 0)) ? 1 : 0;
    int bits = 0;
    while (((int64_t)1 << bits) < n) ++bits;                 // positions < 2^bits
    q.bits = bits;

    if (pillar_map) {          // filled with -1 by the init kernel (needs 16-byte alignment; else a memset)
        if (((uintptr_t)pillar_map % 16 == 0) && (cells % 4 == 0)) {
            cv.ia.ff_ptr[1] = (int4 *)pillar_map;
            cv.ia.ff_n[1] = cells / 4;
        } else {
            PP_CUDA_TRY(cudaMemsetAsync(pillar_map, 0xFF, (size_t)cells * 4, st));
            prof_mark("memset");
        }
    }
    const int32_t *order_perm = order == PP_ORDER_PERM ? perm : nullptr;
    PfnArgs pa, *pap = nullptr;
    if (pfn) {
        pa.W = pfn->weight; pa.scale = pfn->scale; pa.shift = pfn->shift; pa.feat = pfn->feat; pa.U = pfn->units;
        pa.canvas = canvas; pa.plane = cells;
        pa.vx = pfn->vx; pa.vy = pfn->vy; pa.x_off = pfn->x_off; pa.y_off = pfn->y_off;
        pap = &pa;
    }
    if (wide)
        return run<u64>(points, n, q, order_perm, cv, voxels, coors, num_points, voxel_num, pillar_map, pap, st);
    return run<uint32_t>(points, n, q, order_perm, cv, voxels, coors, num_points, voxel_num, pillar_map, pap, st);
}
