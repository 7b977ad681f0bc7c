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
// key's quantile ([0,1/128) [1/128,1/64) ... [1/2,1]; quantiles of a sorted key sample, or of the position), so a cell
// with any number of points finds its max_points-th key in a chunk whose prefix is at most about 2 x max_points.
// Cells are addressed directly (slot = cell) when the grid is small, through an open-addressing hash of the cell id
// otherwise (3-D grids with millions of cells); everything after the slot lookup is the same.
//
// Kernels (all with programmatic dependent launch; each point costs ONE random L2 atomic per per-point pass):
//   S  vox_init_kernel     zero the chunk counters / small counters / histogram, -1 into the hash keys and the
//                          pillar map; reflectance order: one CTA sorts a 1024-key sample -> 7 chunk splitters + 1023
//                          fine splitters (ranking bins)
//   A  vox_count_kernel    per point: cell -> slot, chunk(K); cnt[slot][chunk]++ (no return value: a reduction);
//                          per-point record (slot, chunk, key high word)
//   B  vox_cells_kernel    per slot: counts -> saturation chunk, m = points in the chunks up to it; segment of m keys
//                          and compact cell id q allocated with one atomic per CTA
//   C  vox_place_kernel    per point: dropped when its chunk lies after the cell's saturation chunk (most points of a
//                          dense cell: one cached read); else seg[cursor[slot]++] = K
//   F  vox_first_kernel    per cell (8 lanes): first[q] = smallest key of the segment; ranking bin of it (fine
//                          splitters + adaptive linear sub-bins) and arrival index inside the bin
//   R  vox_bucket_kernel   every CTA scans the bin histogram in shared memory; cells in bucket order; the last CTA to
//                          finish settles the cutoff when more than max_voxels cells are occupied
//   D  vox_gather_kernel   per cell (warp): pillar id = bucket base + smaller keys inside the bucket; the max_points
//                          smallest keys of the segment, sorted (bitonic merges in registers), keys >= cutoff dropped;
//                          voxels[pid][s] = points[key.index], coors, num_points, cell -> pillar map; with the fused
//                          PillarFeatureNet the warp runs the pillar through decorate + Linear + BN + ReLU + max
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
constexpr int NFINE = 1024;         // sample intervals used to rank the cells' first keys
constexpr int NSUB = 8;             // bins per sample interval on average
constexpr int NBIN = NFINE * NSUB;  // ranking bins
constexpr int SAMPLE = 1024;        // keys sampled for the quantile splitters (one per sorting thread)
enum { CTR_NQ = 0, CTR_SEG = 1, CTR_DONE = 2 };

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
    uint32_t *cnt;         // [T][NCH] chunk counts; after kernel B: [0] segment cursor, [1] saturation chunk
    int32_t *counters;     // CTR_*
    uint32_t *rec;         // [N] (slot << 3) | chunk, ~0 outside the grid      (32-bit keys)
    uint2 *rec2;           // [N] the same + the key's high word                (64-bit keys)
    void *seg;             // [N] keys grouped by cell
    int4 *qinfo;           // [Q] cell, segment offset, m, points in the cell
    void *first;           // [Q] smallest key of the cell
    int32_t *bin_of_q;     // [Q]
    int32_t *arr_of_q;     // [Q] arrival index inside the bin
    int32_t *hist;         // [NBIN]
    int32_t *base;         // [NBIN + 1] exclusive scan of hist
    void *lkey;            // [Q] first keys in bucket order
    int32_t *lq;           // [Q] their cells q
    void *cutoff;          // key
    u64 *coarse, *fine;    // splitters (64-bit keys): [NCH - 1], [NFINE - 1]
};

// Cell index along one axis, bit-identical to the reference's floor((p - r) / v) in its promotion regime: the exact
// evaluation (fp64 / IEEE-fp32 division), used for points within a guard band of a cell boundary.
__device__ __forceinline__ bool axis_cell_exact(const VoxParams &q, int j, float p, int &c)
{
    double cd;
    if (q.regime == 2) cd = floor(((double)p - q.r[j]) / q.v[j]);
    else if (q.regime == 1) cd = floor((double)__fsub_rn(p, q.rf[j]) / q.v[j]);
    else cd = (double)floorf(__fdiv_rn(__fsub_rn(p, q.rf[j]), q.vf[j]));
    if (!(cd >= 0.0) || cd >= (double)q.g[j]) return false;   // also rejects NaN
    c = (int)cd;
    return true;
}

// Linear cell of a point, or -1 outside the grid.  An fp32 reciprocal-multiply estimate decides every point that is
// not within a guard band of a cell boundary on any axis (the band covers the rounding of r, 1/v and the two fp32
// operations; outside it both floors agree); only those points take the exact path.  One branch per point.
__device__ __forceinline__ int32_t point_cell(const VoxParams &q, float x, float y, float z)
{
    const float p[3] = {x, y, z};
    float fl[3];
    bool near = false, inside = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float est = (p[j] - q.rf[j]) * q.inv_vf[j];
        fl[j] = floorf(est);
        const float fr = est - fl[j];
        const float band = 1e-6f * (fabsf(est) + q.rv_abs[j]) + 1e-6f;
        near = near || !(fr > band && fr < 1.0f - band);          // also true for NaN / huge values
        inside = inside && (fl[j] >= 0.f) && (fl[j] < (float)q.g[j]);
    }
    int cx, cy, cz;
    if (near) {
        if (!(axis_cell_exact(q, 0, x, cx) && axis_cell_exact(q, 1, y, cy) && axis_cell_exact(q, 2, z, cz))) return -1;
    } else {
        if (!inside) return -1;
        cx = (int)fl[0]; cy = (int)fl[1]; cz = (int)fl[2];
    }
    // cell linearisation (z*gy + y)*gx + x = the (D,H,W) order of the BEV canvas
    return (cz * q.g[1] + cy) * q.g[0] + cx;
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
constexpr int F_OCT = 8, F_SUB = 10;     // 8192 ranking bins (32-bit keys): 8 octaves x 1024 (cell minima crowd at low positions)

// The same geometric chunk layout expressed as quantiles of the sorted sample: index of the lower edge of chunk b >= 1
__device__ __forceinline__ int chunk_sample_index(int b) { return SAMPLE >> (NCH - b); }      // 8 16 32 ... 512

// ---- canvas zero fill, spread over the per-point kernels ---------------------------------------------------------------
// The fused frame call (pp_voxelize_scatter) writes the BEV canvas from the gather kernel; the zeros of the other ~95 % of
// the canvas do not depend on anything, so slices of them are written by kernels A, C and F BEFORE their dependency wait,
// i.e. while the predecessor is still running: linear 256-bit stores (STG.256), fire and forget.  The init kernel
// orders wait -> trigger, so none of this starts before the caller's earlier work on the stream is complete.
struct FillArgs {
    float *base;          // nullptr: nothing to fill
    int64_t units;        // 32-byte units in the canvas
};
constexpr int FILL_SLICES = 3;

__device__ __forceinline__ void fill_slice(const FillArgs &fa, int slice)
{
    if (!fa.base) return;
    const int64_t per = (fa.units + FILL_SLICES - 1) / FILL_SLICES;
    const int64_t lo = per * slice, hi = lo + per < fa.units ? lo + per : fa.units;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride)
        asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(fa.base + i * 8), "f"(0.f) : "memory");
}

// ---- S: workspace initialisation, and (reflectance order) quantile splitters from a key sample ----------------
// One launch replaces the memsets: CTA 0 sorts the key sample (bitonic, shared memory) while the other CTAs fill.
struct InitArgs {
    int4 *ff_ptr[3];  int64_t ff_n[3];     // regions filled with 0xFF (16-byte units)
    int4 *z_ptr[2];   int64_t z_n[2];      // regions filled with 0
};

__global__ void __launch_bounds__(1024)
vox_init_kernel(const float *__restrict__ points, int64_t n, int C, int wide, u64 *__restrict__ coarse,
                u64 *__restrict__ fine, const InitArgs ia)
{
    // First kernel of the call: wait for everything earlier on the stream, THEN let the per-point kernel start -- its
    // CTAs read `points` before their own dependency wait, which is only safe once the producer of the points is done.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x;
    if (blockIdx.x > 0 || !wide) {
        const int64_t nb = gridDim.x - (wide ? 1 : 0), b = blockIdx.x - (wide ? 1 : 0);
        const int4 ff = make_int4(-1, -1, -1, -1), zz = make_int4(0, 0, 0, 0);
#pragma unroll
        for (int r = 0; r < 3; ++r)
            for (int64_t i = b * 1024 + tid; i < ia.ff_n[r]; i += nb * 1024) ia.ff_ptr[r][i] = ff;
#pragma unroll
        for (int r = 0; r < 2; ++r)
            for (int64_t i = b * 1024 + tid; i < ia.z_n[r]; i += nb * 1024) ia.z_ptr[r][i] = zz;
        return;
    }
    __shared__ u64 s[SAMPLE];
    for (int i = tid; i < SAMPLE; i += 1024) {
        // evenly spaced sample; short inputs are padded with +inf keys
        int64_t idx = (n >= SAMPLE) ? (int64_t)i * (n / SAMPLE) : i;
        u64 k = ~0ull;
        if (idx < n) k = ((u64)refl_key_hi(points[idx * C + 3]) << 32) | (uint32_t)idx;
        s[i] = k;
    }
    __syncthreads();
    {
        // bitonic sort, one key per thread (SAMPLE == blockDim): partners less than a warp apart exchange with
        // shuffles (40 of the 55 steps), the others through shared memory
        u64 k = s[tid];
        for (int size = 2; size <= SAMPLE; size <<= 1) {
            const bool up = (tid & size) == 0;
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                u64 o;
                if (stride >= 32) {
                    __syncthreads();
                    s[tid] = k;
                    __syncthreads();
                    o = s[tid ^ stride];
                } else {
                    const unsigned lo = __shfl_xor_sync(0xFFFFFFFFu, (unsigned)k, stride);
                    const unsigned hi = __shfl_xor_sync(0xFFFFFFFFu, (unsigned)(k >> 32), stride);
                    o = ((u64)hi << 32) | lo;
                }
                const bool lower = (tid & stride) == 0;
                const bool take_min = lower == up;
                k = take_min ? (k < o ? k : o) : (k < o ? o : k);
            }
        }
        __syncthreads();
        s[tid] = k;
        __syncthreads();
    }
    // chunk b >= 1 starts at the geometric quantile chunk_sample_index(b); splitter i is the start of chunk i + 1
    if (tid < NCH - 1) coarse[tid] = s[chunk_sample_index(tid + 1)];
    for (int i = tid; i < NFINE - 1; i += 1024) fine[i] = s[(i + 1) * (SAMPLE / NFINE)];
}

// ---- slot of a cell ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_cell(uint32_t cell, int bits) { return (cell * 0x9E3779B1u) >> (32 - bits); }

// Open addressing, linear probing.  A slot's key goes from -1 to a cell once and never changes, so the CAS that
// inserts a key is also its publication: nobody waits for anybody.
template <bool INSERT>
__device__ __forceinline__ int hash_slot(const VoxBuf &w, int32_t cell)
{
    const uint32_t mask = (1u << w.hash_bits) - 1u;
    uint32_t h = hash_cell((uint32_t)cell, w.hash_bits);
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        int32_t k = __ldcg(w.slot_key + h);
        if (k == cell) return (int)h;
        if (k == -1) {
            if (!INSERT) return -1;                       // cannot happen: every cell was inserted by kernel A
            k = atomicCAS(w.slot_key + h, -1, cell);
            if (k == -1 || k == cell) return (int)h;
        }
        h = (h + 1) & mask;
    }
    return -1;
}

// ---- A: per point, count ---------------------------------------------------------------------------------------
constexpr int CNT_IT = 4;      // points per thread: the four point loads are in flight together; the counter updates
                               // return nothing, so nothing else in this kernel waits on memory

template <typename K, bool HASH>
__global__ void __launch_bounds__(VOX_THREADS)
vox_count_kernel(const float *__restrict__ points, int64_t n, const VoxParams prm, const int32_t *__restrict__ perm,
                 const VoxBuf w, const FillArgs fa)
{
    // PDL: the points are read and binned into cells BEFORE the dependency wait, i.e. while the init kernel (workspace
    // fill, sample sort) is still running; nothing the init kernel writes is touched before it.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr bool WIDE = sizeof(K) == 8;
    __shared__ u64 s_split[NCH];
    const int64_t p0 = (int64_t)blockIdx.x * (VOX_THREADS * CNT_IT) + threadIdx.x;
    int32_t cell[CNT_IT];
    float refl[CNT_IT];
#pragma unroll
    for (int k = 0; k < CNT_IT; ++k) {
        const int64_t p = p0 + k * VOX_THREADS;
        cell[k] = -1;
        refl[k] = 0.f;
        if (p >= n) continue;
        const int64_t idx = perm ? (int64_t)(uint32_t)perm[p] : p;
        float x, y, z;
        if (prm.vec4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(points) + idx);
            x = v.x; y = v.y; z = v.z; refl[k] = v.w;
        } else {
            const float *pt = points + idx * prm.C;
            x = __ldg(pt); y = __ldg(pt + 1); z = __ldg(pt + 2);
            if (WIDE) refl[k] = __ldg(pt + 3);
        }
        cell[k] = point_cell(prm, x, y, z);
    }
    fill_slice(fa, 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (WIDE) {
        if (threadIdx.x < NCH - 1) s_split[threadIdx.x] = w.coarse[threadIdx.x];
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < CNT_IT; ++k) {
        const int64_t p = p0 + k * VOX_THREADS;
        if (p >= n) continue;
        uint32_t r = 0xFFFFFFFFu, hi = 0u;
        if (cell[k] >= 0) {
            int chunk = 0;
            if (WIDE) {
                hi = refl_key_hi(refl[k]);
                const u64 key = ((u64)hi << 32) | (uint32_t)p;
#pragma unroll
                for (int c = 0; c < NCH - 1; ++c) chunk += (s_split[c] <= key) ? 1 : 0;
            } else {
                chunk = geo_bin<C_OCT, C_SUB>((uint32_t)p, prm.bits);
            }
            const int slot = HASH ? hash_slot<true>(w, cell[k]) : cell[k];
            if (slot >= 0) {                                  // (a hash table of 2 n slots cannot fill up)
                atomicAdd(w.cnt + (size_t)slot * NCH + chunk, 1u);
                r = ((uint32_t)slot << 3) | (uint32_t)chunk;
            }
        }
        if (WIDE) w.rec2[p] = make_uint2(r, hi);
        else w.rec[p] = r;
    }
}

// ---- B: per slot -------------------------------------------------------------------------------------------------
// counts -> saturation chunk (the first chunk at which the cell holds max_points points; later chunks are dropped
// unseen) and m = the points of the chunks up to it.  Segment offsets and the compact ids q come from CTA-wide prefix
// sums and one atomic per CTA and counter.
constexpr int CELLS_THREADS = 1024;

__global__ void __launch_bounds__(CELLS_THREADS) vox_cells_kernel(const VoxParams prm, const VoxBuf w)
{
    pdl_enter();
    __shared__ u64 s_warp[CELLS_THREADS / 32];
    __shared__ u64 s_cta;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t = (int64_t)blockIdx.x * CELLS_THREADS + tid;
    uint32_t total = 0, m = 0;
    int sat = NCH - 1;
    if (t < w.T) {
        const uint4 a = *reinterpret_cast<const uint4 *>(w.cnt + (size_t)t * NCH);
        const uint4 b = *reinterpret_cast<const uint4 *>(w.cnt + (size_t)t * NCH + 4);
        const uint32_t c[NCH] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        bool found = false;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            total += c[k];
            if (!found && total >= (uint32_t)prm.P) { found = true; sat = k; m = total; }
        }
        if (!found) m = total;
    }
    const bool occ = total > 0;
    // exclusive prefix over the CTA of (m, occupied) packed in one 64-bit word (m < 2^31, at most 1024 cells per CTA)
    const u64 mine = ((u64)m << 11) | (occ ? 1ull : 0ull);
    u64 incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned lo = __shfl_up_sync(0xFFFFFFFFu, (unsigned)incl, o), hi = __shfl_up_sync(0xFFFFFFFFu, (unsigned)(incl >> 32), o);
        if (lane >= o) incl += ((u64)hi << 32) | lo;
    }
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
    if (!occ) return;
    const u64 excl = s_warp[warp] + incl - mine;
    const uint32_t q = (uint32_t)s_cta + (uint32_t)(excl & 0x7FFull);
    const uint32_t off = (uint32_t)(s_cta >> 32) + (uint32_t)(excl >> 11);
    *reinterpret_cast<uint2 *>(w.cnt + (size_t)t * NCH) = make_uint2(off, (uint32_t)sat);
    const int cell = w.hash_bits ? w.slot_key[t] : (int)t;
    w.qinfo[q] = make_int4(cell, (int)off, (int)m, (int)total);
}

// ---- C: per point, place -------------------------------------------------------------------------------------------
constexpr int PLACE_IT = 4;

template <typename K>
__global__ void __launch_bounds__(VOX_THREADS) vox_place_kernel(int64_t n, const VoxBuf w, const FillArgs fa)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    fill_slice(fa, 1);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    constexpr bool WIDE = sizeof(K) == 8;
    const int64_t p0 = (int64_t)blockIdx.x * (VOX_THREADS * PLACE_IT) + threadIdx.x;
    uint32_t r[PLACE_IT], hi[PLACE_IT];
#pragma unroll
    for (int k = 0; k < PLACE_IT; ++k) {
        const int64_t p = p0 + k * VOX_THREADS;
        r[k] = 0xFFFFFFFFu;
        hi[k] = 0u;
        if (p < n) {
            if (WIDE) { const uint2 v = w.rec2[p]; r[k] = v.x; hi[k] = v.y; }
            else r[k] = w.rec[p];
        }
    }
    uint32_t sat[PLACE_IT];
#pragma unroll
    for (int k = 0; k < PLACE_IT; ++k)
        sat[k] = r[k] != 0xFFFFFFFFu ? __ldcg(w.cnt + (size_t)(r[k] >> 3) * NCH + 1) : 0u;
    uint32_t pos[PLACE_IT];
#pragma unroll
    for (int k = 0; k < PLACE_IT; ++k) {
        pos[k] = 0xFFFFFFFFu;
        if (r[k] != 0xFFFFFFFFu && (r[k] & 7u) <= sat[k]) pos[k] = atomicAdd(w.cnt + (size_t)(r[k] >> 3) * NCH, 1u);
    }
#pragma unroll
    for (int k = 0; k < PLACE_IT; ++k) {
        if (pos[k] == 0xFFFFFFFFu) continue;
        const int64_t p = p0 + k * VOX_THREADS;
        ((K *)w.seg)[pos[k]] = WIDE ? (K)(((u64)hi[k] << 32) | (uint32_t)p) : (K)(uint32_t)p;
    }
}

// ---- F: per cell, smallest key and its ranking bin ---------------------------------------------------------------
// Bins (monotone in the key).  64-bit keys: sample interval of the key, then a linear position inside the interval on
// the reflectance bits.  A cell with m points has its smallest key near the 1/m quantile, so the smallest keys crowd
// into the first ~ nq / n of the key space: the intervals below `istar` share most of the bins.
struct BinPlan {
    int istar, nsub;
};
__device__ __forceinline__ BinPlan bin_plan(int nq, int64_t n_points)
{
    BinPlan b;
    int istar = (int)(((int64_t)NFINE * 3 * nq) / (n_points > 0 ? n_points : 1)) + 8;
    b.istar = istar > NFINE - 1 ? NFINE - 1 : istar;
    b.nsub = (NBIN - NFINE) / b.istar;
    return b;
}

__device__ __forceinline__ int wide_bin(const u64 *s_fine, const BinPlan bp, u64 f)
{
    int lo = 0, hi = NFINE - 1;                       // number of splitters <= f
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_fine[mid] <= f) lo = mid + 1; else hi = mid;
    }
    const int i = lo;
    if (i >= bp.istar) return bp.istar * bp.nsub + (i - bp.istar);
    const uint32_t hi_w = (uint32_t)(s_fine[i] >> 32);
    uint32_t lo_w;
    if (i > 0) lo_w = (uint32_t)(s_fine[i - 1] >> 32);
    else { const uint32_t a = (uint32_t)(s_fine[0] >> 32), b = (uint32_t)(s_fine[1] >> 32), wd = b - a; lo_w = a > wd ? a - wd : 0u; }
    const uint32_t fw = (uint32_t)(f >> 32);
    int sub = 0;
    if (fw > lo_w) {
        // monotone in f: conversion, multiplication by a positive constant and truncation are all monotone
        const float inv = (float)bp.nsub / ((float)(hi_w - lo_w) + 1.0f);
        sub = (int)((float)(fw - lo_w) * inv);
    }
    return i * bp.nsub + (sub < bp.nsub - 1 ? sub : bp.nsub - 1);
}

constexpr int FIRST_THREADS = 256, FIRST_GL = 8;      // 8 lanes per cell

template <typename K>
__global__ void __launch_bounds__(FIRST_THREADS)
vox_first_kernel(const VoxParams prm, const VoxBuf w, int64_t n_points, const FillArgs fa)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    fill_slice(fa, 2);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    constexpr bool WIDE = sizeof(K) == 8;
    __shared__ u64 s_fine[NFINE];
    if (WIDE) {
        for (int i = threadIdx.x; i < NFINE - 1; i += FIRST_THREADS) s_fine[i] = w.fine[i];
        if (threadIdx.x == 0) s_fine[NFINE - 1] = ~0ull;
        __syncthreads();
    }
    const int nq = w.counters[CTR_NQ];
    const BinPlan bp = bin_plan(nq, n_points);
    const int lane = threadIdx.x & 31, sub = lane & (FIRST_GL - 1);
    const K *seg = (const K *)w.seg;
    const int groups = gridDim.x * (FIRST_THREADS / FIRST_GL);
    for (int q0 = blockIdx.x * (FIRST_THREADS / FIRST_GL); q0 < nq; q0 += groups) {
        const int q = q0 + threadIdx.x / FIRST_GL;
        K mn = KeyInf<K>::value();
        if (q < nq) {
            const int4 info = w.qinfo[q];
            for (int j = sub; j < info.z; j += FIRST_GL) {
                const K k = seg[info.y + j];
                mn = k < mn ? k : mn;
            }
        }
#pragma unroll
        for (int o = FIRST_GL / 2; o > 0; o >>= 1) {
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
            const int bin = WIDE ? wide_bin(s_fine, bp, (u64)mn) : geo_bin<F_OCT, F_SUB>((uint32_t)mn, prm.bits);
            ((K *)w.first)[q] = mn;
            w.bin_of_q[q] = bin;
            w.arr_of_q[q] = atomicAdd(w.hist + bin, 1);
        }
    }
}

// ---- R: cells in bucket order; cutoff ------------------------------------------------------------------------------
constexpr int BUCKET_THREADS = 1024;

template <typename K>
__global__ void __launch_bounds__(BUCKET_THREADS)
vox_bucket_kernel(const VoxParams prm, const VoxBuf w, int32_t *__restrict__ voxel_num)
{
    pdl_enter();
    __shared__ int s_base[NBIN + 1];
    __shared__ int s_warp[BUCKET_THREADS / 32];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nq = w.counters[CTR_NQ];
    {
        constexpr int PER = NBIN / BUCKET_THREADS;
        int v[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { v[k] = w.hist[tid * PER + k]; sum += v[k]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int base = incl - sum;
        for (int k = 0; k < warp; ++k) base += s_warp[k];
#pragma unroll
        for (int k = 0; k < PER; ++k) { s_base[tid * PER + k] = base; base += v[k]; }
        if (tid == BUCKET_THREADS - 1) s_base[NBIN] = base;
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        for (int i = tid; i <= NBIN; i += BUCKET_THREADS) w.base[i] = s_base[i];
        if (tid == 0) *voxel_num = nq < prm.max_voxels ? nq : prm.max_voxels;
    }
    for (int q = blockIdx.x * BUCKET_THREADS + tid; q < nq; q += gridDim.x * BUCKET_THREADS) {
        const int slot = s_base[w.bin_of_q[q]] + w.arr_of_q[q];
        w.lq[slot] = q;
        ((K *)w.lkey)[slot] = ((const K *)w.first)[q];
    }
    // The reference breaks at the first point that would open pillar max_voxels + 1 (:223, :291): its key -- the cell
    // minimum of rank max_voxels -- is the cutoff.  The last CTA to finish finds it in the bucket that holds that rank.
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(w.counters + CTR_DONE, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (nq <= prm.max_voxels) {
        if (tid == 0) *(K *)w.cutoff = KeyInf<K>::value();
        return;
    }
    int lo = 0, hi = NBIN;                               // largest bin b with base[b] <= max_voxels
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_base[mid] <= prm.max_voxels) lo = mid; else hi = mid - 1;
    }
    // base[NBIN] = nq > max_voxels, so lo < NBIN and base[lo + 1] > max_voxels: bin lo holds the rank max_voxels
    const int b0 = s_base[lo], b1 = s_base[lo + 1];
    const int target = prm.max_voxels - b0;
    const K *lkey = (const K *)w.lkey;
    for (int i = b0 + tid; i < b1; i += BUCKET_THREADS) {
        const K f = __ldcg(lkey + i);
        int cnt = 0;
        for (int j = b0; j < b1; ++j) cnt += (__ldcg(lkey + j) < f) ? 1 : 0;
        if (cnt == target) *(K *)w.cutoff = f;
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

// The R * 32 smallest keys of seg[0, m), ascending: element e = r * 32 + lane is b[r].  Batches of 32 keys are sorted
// and merged down the registers; a batch with no key below the current last element is skipped after one ballot.
template <typename K, int R>
__device__ __forceinline__ void select_smallest(const K *__restrict__ seg, int m, int lane, K (&b)[R])
{
#pragma unroll
    for (int r = 0; r < R; ++r) b[r] = KeyInf<K>::value();
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

// pillar id of bucket slot s = bucket base + number of smaller keys inside the bucket (the whole warp counts)
template <typename K>
__device__ __forceinline__ int pillar_rank(const VoxBuf &w, int s, int lane, int &q)
{
    const K *lkey = (const K *)w.lkey;
    q = w.lq[s];
    const K f = lkey[s];
    const int bin = w.bin_of_q[q];
    const int b0 = w.base[bin], b1 = w.base[bin + 1];
    int cnt = 0;
    for (int j = b0 + lane; j < b1; j += 32) cnt += (lkey[j] < f) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    return b0 + cnt;
}

__device__ __forceinline__ void write_coors(const VoxParams &prm, const GatherOut &out, int pid, int cell)
{
    const int cx = cell % prm.g[0], tt = cell / prm.g[0];
    out.coors[pid * 3 + 0] = cx;
    out.coors[pid * 3 + 1] = tt % prm.g[1];
    out.coors[pid * 3 + 2] = tt / prm.g[1];
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
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PfnWeights pw;
    if (PFN) pfn_load_weights(pw, pa.W, pa.scale, pa.shift, pa.U, 4, lane);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int nq = w.counters[CTR_NQ];
    const K cutoff = *(const K *)w.cutoff;
    const int P = prm.P, C = prm.C;
    for (int s = blockIdx.x * (GP_THREADS / 32) + warp; s < nq; s += gridDim.x * (GP_THREADS / 32)) {
        int q;
        const int pid = pillar_rank<K>(w, s, lane, q);
        if (pid >= prm.max_voxels) continue;
        const int4 info = w.qinfo[q];
        K b[R];
        select_smallest<K, R>((const K *)w.seg + info.y, info.z, lane, b);
        int n = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            // keys at or after the cutoff are dropped: they are the largest keys of their rows
            const bool valid = r * 32 + lane < P && b[r] < cutoff;
            n += __popc(__ballot_sync(0xFFFFFFFFu, valid));
            if (!valid) b[r] = KeyInf<K>::value();
        }
        if (lane == 0) out.num_points[pid] = n;
        if (lane == 1) write_coors(prm, out, pid, info.x);
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
            pfn_pillar4(pw, v0, P, n, cell % prm.g[0], (cell / prm.g[0]) % prm.g[1], pa.vx, pa.vy, pa.x_off, pa.y_off,
                        s_row + warp * 32 * 4, pa.feat ? pa.feat + (int64_t)pid * (pa.U + 1) : nullptr,
                        pa.canvas ? pa.canvas + cell : nullptr, pa.plane, pa.U, lane);
        }
    }
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
    for (int s = blockIdx.x * (GP_THREADS / 32) + warp; s < nq; s += gridDim.x * (GP_THREADS / 32)) {
        int q;
        const int pid = pillar_rank<K>(w, s, lane, q);
        if (pid >= prm.max_voxels) continue;
        const int4 info = w.qinfo[q];
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
        if (lane == 1) write_coors(prm, out, pid, info.x);
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
    r.b.rec = wide ? nullptr : a.take<uint32_t>((size_t)n1);
    r.b.rec2 = wide ? a.take<uint2>((size_t)n1) : nullptr;
    r.b.seg = a.take<char>((size_t)n1 * ksz);
    r.b.qinfo = a.take<int4>((size_t)r.Q);
    r.b.first = a.take<char>((size_t)r.Q * ksz);
    r.b.bin_of_q = a.take<int32_t>((size_t)r.Q);
    r.b.arr_of_q = a.take<int32_t>((size_t)r.Q);
    r.b.lkey = a.take<char>((size_t)r.Q * ksz);
    r.b.lq = a.take<int32_t>((size_t)r.Q);
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

template <typename K>
int run(const float *points, int64_t n, const VoxParams &prm, const int32_t *perm, const Carve &cv, float *voxels,
        int32_t *coors, int32_t *num_points, int32_t *voxel_num, int32_t *pillar_map, int64_t max_rows, const PfnArgs *pfn,
        cudaStream_t st)
{
    const VoxBuf &w = cv.b;
    FillArgs fa = {nullptr, 0};
    if (pfn && pfn->canvas) {
        fa.base = pfn->canvas;
        fa.units = (int64_t)(pfn->U + 1) * pfn->plane / 8;
    }
    constexpr bool WIDE = sizeof(K) == 8;
    int64_t fill_units = 0;
    for (int r = 0; r < 3; ++r) fill_units += cv.ia.ff_n[r];
    for (int r = 0; r < 2; ++r) fill_units += cv.ia.z_n[r];
    int init_blocks = (int)ceil_div(fill_units, 1024 * 4);
    init_blocks = init_blocks < 1 ? 1 : (init_blocks > 148 * 2 ? 148 * 2 : init_blocks);
    launch_pdl(vox_init_kernel, dim3(init_blocks + (WIDE ? 1 : 0)), dim3(1024), 0, st, points, n, prm.C, WIDE ? 1 : 0, w.coarse, w.fine, cv.ia);
    if (int rc = check_launch("vox_init_kernel")) return rc;
    const unsigned pt_blocks = (unsigned)ceil_div(n, VOX_THREADS * CNT_IT);
    if (w.hash_bits)
        launch_pdl(vox_count_kernel<K, true>, dim3(pt_blocks), dim3(VOX_THREADS), 0, st, points, n, prm, perm, w, fa);
    else
        launch_pdl(vox_count_kernel<K, false>, dim3(pt_blocks), dim3(VOX_THREADS), 0, st, points, n, prm, perm, w, fa);
    if (int rc = check_launch("vox_count_kernel")) return rc;
    launch_pdl(vox_cells_kernel, dim3((unsigned)ceil_div(w.T, CELLS_THREADS)), dim3(CELLS_THREADS), 0, st, prm, w);
    if (int rc = check_launch("vox_cells_kernel")) return rc;
    launch_pdl(vox_place_kernel<K>, dim3((unsigned)ceil_div(n, VOX_THREADS * PLACE_IT)), dim3(VOX_THREADS), 0, st, n, w, fa);
    if (int rc = check_launch("vox_place_kernel")) return rc;
    // the cell count is only known on the device: grid-stride grids sized for the worst case, capped at a few waves
    auto capped = [](int64_t blocks, int64_t cap) { return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap); };
    launch_pdl(vox_first_kernel<K>, dim3(capped(ceil_div(cv.Q, FIRST_THREADS / FIRST_GL), 148 * 8)), dim3(FIRST_THREADS), 0, st, prm, w, n, fa);
    if (int rc = check_launch("vox_first_kernel")) return rc;
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
    const int64_t max_rows = max_rows_of(n, cfg);
    PfnArgs pa, *pap = nullptr;
    if (pfn) {
        pa.W = pfn->weight; pa.scale = pfn->scale; pa.shift = pfn->shift; pa.feat = pfn->feat; pa.U = pfn->units;
        pa.canvas = canvas; pa.plane = cells;
        pa.vx = pfn->vx; pa.vy = pfn->vy; pa.x_off = pfn->x_off; pa.y_off = pfn->y_off;
        pap = &pa;
    }
    if (wide)
        return run<u64>(points, n, q, order_perm, cv, voxels, coors, num_points, voxel_num, pillar_map, max_rows, pap, st);
    return run<uint32_t>(points, n, q, order_perm, cv, voxels, coors, num_points, voxel_num, pillar_map, max_rows, pap, st);
}
