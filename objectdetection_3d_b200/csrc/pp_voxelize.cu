// Hard voxelization on sm_100a, bit-exact with the reference's sequential first-come pass
// (ops/ops_numba.py:171-308), restated as order-independent parallel steps with NO global sort.
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
// Kernels (64 key chunks per cell, monotone in K: geometric in the position, or geometric quantiles of a key sample):
//   S  vox_init_kernel        workspace fill; reflectance order: one CTA sorts a 1024-key sample -> 63 coarse chunk
//                             splitters + 1023 fine splitters
//   A0 vox_preclaim_kernel    every 16th point: claim the row of its cell (nobody waits), so that A finds the rows published
//   A  vox_scatter_kernel     per point: cell -> compact cell row q (claimed on first touch, atomicCAS on a dense map,
//                             looked up through L1 afterwards), ticket = cnt[q][chunk(K)]++
//   Q  vox_cell_prefix_kernel per cell (warp): counts -> inclusive prefix over the chunks, saturation chunk
//   C  vox_place_kernel       per point: window [prefix(chunk-1), prefix(chunk)) of its cell's row.  Untruncated
//                             windows are filled in arrival (ticket) order, the truncated one by a lock-free
//                             atomicMin insertion chain; the points of a cell's lowest chunk also set first[q] = min K
//   R  vox_rank_kernel        cooperative launch, two grid barriers: bin the cells by first[q] | bucket | pillar id =
//                             bucket base + rank inside the bucket; coors, pillar_map, cutoff, voxel_num
//   D  vox_gather*_kernel     per pillar: sort the row, drop keys >= cutoff, voxels[m][s] = points[row[s]]
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
constexpr int NCHUNK = 64;          // coarse key chunks per cell
constexpr int NFINE = 1024;         // sample intervals used to rank the cells' first keys
constexpr int NSUB = 8;             // linear sub-bins per sample interval (64-bit keys)
constexpr int NBIN = NFINE * NSUB;  // ranking bins
constexpr int SAMPLE = 1024;        // keys sampled for the quantile splitters (one per sorting thread)

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
    int ticket;   // max_points <= 64: untruncated windows are filled in arrival order and sorted by the gather kernel
    int bits;     // 32-bit keys: positions < 2^bits; chunks / fine bins are geometric in the position (geo_bin)
};

struct VoxBuf {
    int32_t *map;          // [cells]  cell -> q, -1 empty, -2 being claimed
    int32_t *counters;     // [0] nq
    int32_t r_rows;        // cnt rows [0, r_rows) are zeroed up front, later rows by the claiming thread
    int32_t *base;         // [NFINE + 1] exclusive scan of hist (written by the bucket kernel)
    void *cutoff;          // key
    int32_t *cell_of_q;    // [Q]
    void *first;           // [Q] key
    int32_t *cnt;          // [Q][NCHUNK] counts, then inclusive prefix
    void *rows;            // [Q][P] keys, sorted ascending, INF padded
    int32_t *pid_of_q;     // [Q]
    int32_t *bin_of_q;     // [Q]
    uint8_t *sat_of_q;     // [Q] first chunk whose inclusive prefix reaches max_points (NCHUNK if none)
    uint16_t *tick;        // [N] (chunk << 8) | arrival index of the point inside its (cell, chunk), saturated at 255
    int32_t *q_of_point;   // [N] row of the point's cell, -1 outside the grid (32-bit keys)
    int2 *qk_of_point;     // [N] (row, primary key) in one 8-byte record (64-bit keys)
    int32_t *hist, *fill;  // [NFINE] each
    int32_t *list;         // [Q] cells in bucket order
    void *lkey;            // [Q] their first keys, same order
    int32_t *q_of_pid;     // [rows]
    u64 *coarse, *fine;    // splitters (64-bit mode): [NCHUNK-1], [NFINE-1]
    // partitioned front end (vox_part_kernel / vox_select_kernel): cell = (local id << pt_lg) | group
    int32_t *pt_cursor;    // [G] records appended to each group's bin
    int32_t *pt_cell;      // [G][pt_cap] cell of the record
    void *pt_key;          // [G][pt_cap] key of the record
    int32_t *pt_flag;      // != 0: a bin overflowed -> the per-point kernels A / Q / C run instead
    int32_t pt_on, pt_lg, pt_cap, pt_D;   // enabled; log2 G; bin capacity; local ids per group = ceil(cells / G)
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

// number of splitters <= k (upper bound): a monotone map key -> [0, n]
__device__ __forceinline__ int upper_bound_u64(const u64 *s, int n, u64 k)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (s[mid] <= k) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Geometric binning of a position u < 2^bits into NOCT octaves x 2^SUB sub-bins (monotone in u).  Octave 0 is
// [0, 2^(bits-NOCT+1)), octave k >= 1 is [2^(bits-NOCT+k), 2^(bits-NOCT+k+1)); each octave is split evenly.
// A cell with n points keeps its max_points smallest keys, i.e. the key quantiles below max_points / n: the
// resolution is spent where the dense cells need it, so windows stay a few points wide whatever n is.
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

// The same geometric layout expressed as quantiles of a sorted sample: index of the lower edge of bin b
template <int NOCT, int SUB>
__device__ __forceinline__ int geo_sample_index(int b, int sample)
{
    const int oct = b >> SUB, sub = b & ((1 << SUB) - 1);
    // octave 0 covers sample / 2^(NOCT-1) elements, octave k >= 1 covers sample / 2^(NOCT-k)
    const int w = oct == 0 ? sample >> (NOCT - 1) : sample >> (NOCT - oct);
    const int lo = oct == 0 ? 0 : sample >> (NOCT - oct);
    return lo + ((w * sub) >> SUB);
}

constexpr int C_OCT = 8, C_SUB = 3;      // 64 coarse chunks
constexpr int F_OCT = 16, F_SUB = 6;     // 1024 fine bins (32-bit keys); 64-bit keys use uniform sample quantiles

__device__ __forceinline__ uint32_t key_min(uint32_t *p, uint32_t v) { return atomicMin(p, v); }
__device__ __forceinline__ u64 key_min(u64 *p, u64 v) { return atomicMin(p, v); }

// ---- S: workspace initialisation, and (reflectance order) quantile splitters from a key sample ----------------
// One launch replaces the memsets: CTA 0 sorts the key sample (bitonic, shared memory) while the other CTAs
// fill map = -1, the first `r_init` cell rows (cnt = 0, first = rows = +inf) and the small counters.
struct InitArgs {
    int4 *ff_ptr[4];  int64_t ff_n[4];     // regions filled with 0xFF (16-byte units)
    int4 *z_ptr[2];   int64_t z_n[2];      // regions filled with 0
};

__global__ void __launch_bounds__(1024)
vox_init_kernel(const float *__restrict__ points, int64_t n, int C, int wide, u64 *__restrict__ coarse,
                u64 *__restrict__ fine, const InitArgs ia)
{
    // First kernel of the call: wait for everything earlier on the stream, THEN let the scatter kernel start -- its
    // CTAs read `points` before their own dependency wait, which is only safe once the producer of the points is done.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x;
    if (blockIdx.x > 0 || !wide) {
        const int64_t nb = gridDim.x - (wide ? 1 : 0), b = blockIdx.x - (wide ? 1 : 0);
        const int4 ff = make_int4(-1, -1, -1, -1), zz = make_int4(0, 0, 0, 0);
#pragma unroll
        for (int r = 0; r < 4; ++r)
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
        if (idx < n) k = ((u64)(~ordered_bits(points[idx * C + 3])) << 32) | (uint32_t)idx;
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
    // coarse chunk b starts at the geometric quantile geo_sample_index(b); splitter i is the start of chunk i + 1
    for (int i = tid; i < NCHUNK - 1; i += 1024) coarse[i] = s[geo_sample_index<C_OCT, C_SUB>(i + 1, SAMPLE)];
    for (int i = tid; i < NFINE - 1; i += 1024) fine[i] = s[(i + 1) * (SAMPLE / NFINE)];
}

// ---- A: per point ------------------------------------------------------------------------------------------
// Row of a cell.  A published row id never changes, so the first look goes through L1 (plain load: a stale line can
// only still say "empty", which falls through to the coherent path); only the first touch of a cell claims a row.
__device__ __forceinline__ int claim_row(const VoxBuf &w, int32_t cell)
{
    int q = __ldcg(w.map + cell);
    while (q < 0) {
        if (q == -1) {
            int old = atomicCAS(w.map + cell, -1, -2);
            if (old == -1) {
                // first touch: allocate a compact row; counter rows >= r_rows were not zeroed by vox_init_kernel
                q = atomicAdd(w.counters, 1);
                w.cell_of_q[q] = cell;
                if (q >= w.r_rows) {
                    int4 *c4 = reinterpret_cast<int4 *>(w.cnt + (size_t)q * NCHUNK);
#pragma unroll
                    for (int k = 0; k < NCHUNK / 4; ++k) c4[k] = make_int4(0, 0, 0, 0);
                    __threadfence();
                }
                atomicExch(w.map + cell, q);
                return q;
            }
            q = old;
        } else {
            q = *((volatile int32_t *)(w.map + cell));     // another thread is publishing the row
        }
    }
    return q;
}

// ---- A0: rows for most cells before the per-point pass ------------------------------------------------------------------
// At the start of kernel A nearly every cell is unclaimed: thousands of threads lose the CAS on the same map entries and
// wait for the winners' two further round trips (a third of A's stall samples).  This small kernel claims the rows of
// the cells of every PRECLAIM_STRIDE-th point first; nobody waits here (a thread that loses the CAS is done), so when A
// runs, the rows of all but the sparsest cells are already published.
constexpr int PRECLAIM_STRIDE = 16;      // measured at 1e6 points: 4 / 8 / 32 / 64 are all slower in flight (54.4 - 56.3 against 52.6 us per frame)

__global__ void __launch_bounds__(VOX_THREADS)
vox_preclaim_kernel(const float *__restrict__ points, int64_t n, const VoxParams prm, const VoxBuf w)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int64_t p = ((int64_t)blockIdx.x * VOX_THREADS + threadIdx.x) * PRECLAIM_STRIDE;
    int32_t cell = -1;
    if (p < n) {          // the points may be read before the wait (the init kernel orders wait -> trigger)
        const float *pt = points + p * prm.C;
        cell = point_cell(prm, __ldg(pt), __ldg(pt + 1), __ldg(pt + 2));
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (cell < 0 || __ldcg(w.map + cell) != -1) return;
    if (atomicCAS(w.map + cell, -1, -2) != -1) return;
    const int q = atomicAdd(w.counters, 1);
    w.cell_of_q[q] = cell;
    if (q >= w.r_rows) {          // counter rows beyond r_rows were not zeroed by vox_init_kernel
        int4 *c4 = reinterpret_cast<int4 *>(w.cnt + (size_t)q * NCHUNK);
#pragma unroll
        for (int k = 0; k < NCHUNK / 4; ++k) c4[k] = make_int4(0, 0, 0, 0);
        __threadfence();
    }
    atomicExch(w.map + cell, q);
}

// Chunk of a 64-bit key = number of coarse splitters <= key.  The search runs on the splitters' high words (the
// reflectance part); the low words only matter when a splitter shares the key's reflectance bits.
struct CoarseTable {
    uint32_t hi[NCHUNK], lo[NCHUNK];
};
__device__ __forceinline__ void load_coarse(CoarseTable &t, const u64 *coarse)
{
    if (threadIdx.x < NCHUNK) {
        const u64 v = threadIdx.x < NCHUNK - 1 ? coarse[threadIdx.x] : ~0ull;
        t.hi[threadIdx.x] = (uint32_t)(v >> 32);
        t.lo[threadIdx.x] = (uint32_t)v;
    }
}
__device__ __forceinline__ int coarse_chunk(const CoarseTable &t, uint32_t prim, uint32_t idx)
{
    // upper bound over the 63 real splitters (entry 63 is +inf), six fixed steps, no branches
    int lo = 0;
#pragma unroll
    for (int step = NCHUNK / 2; step > 0; step >>= 1) lo += (t.hi[lo + step - 1] <= prim) ? step : 0;
    while (lo > 0 && t.hi[lo - 1] == prim && t.lo[lo - 1] > idx) --lo;     // equal reflectance bits: order by index
    return lo;
}

constexpr int SC_IT = 1;      // points per thread and iteration.  Measured with the rows pre-claimed (kernel A0), 24 frames in
                              // flight: 8 / 4 / 2 / 1 points give 57.5 / 53.0 / 51.9 / 51.5 us per frame and 31 / 25 / 22 / 20 us alone;
                              // more warps hide the map -> counter round trips better than interleaved chains of one thread

template <typename K>
__global__ void __launch_bounds__(VOX_THREADS, 8)
vox_scatter_kernel(const float *__restrict__ points, int64_t n, const VoxParams prm, const int32_t *__restrict__ perm,
                   const VoxBuf w)
{
    // PDL: the first batch of points is read and binned into cells BEFORE the dependency wait, i.e. while the init
    // kernel (workspace fill, sample sort) is still running; nothing the init kernel writes is touched before it.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ CoarseTable s_ct;
    constexpr bool WIDE = sizeof(K) == 8;
    bool waited = false;
    if (w.pt_on) {
        // behind the partitioned front end this kernel only runs when a bin overflowed
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (__ldcg(w.pt_flag) == 0) return;
        if (WIDE) {
            load_coarse(s_ct, w.coarse);
            __syncthreads();
        }
        waited = true;
    }
    for (int64_t b0 = (int64_t)blockIdx.x * (VOX_THREADS * SC_IT); b0 < n; b0 += (int64_t)gridDim.x * (VOX_THREADS * SC_IT)) {
        const int64_t p0 = b0 + threadIdx.x;
        int32_t cell[SC_IT];
        float refl[SC_IT];
#pragma unroll
        for (int k = 0; k < SC_IT; ++k) {
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
        if (!waited) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            if (WIDE) {
                load_coarse(s_ct, w.coarse);
                __syncthreads();
            }
            waited = true;
        }
        int q[SC_IT];
#pragma unroll
        for (int k = 0; k < SC_IT; ++k) q[k] = cell[k] >= 0 ? w.map[cell[k]] : -1;      // L1 look (see claim_row)
        int t[SC_IT], chk[SC_IT];
        uint32_t prim[SC_IT];
#pragma unroll
        for (int k = 0; k < SC_IT; ++k) {
            const int64_t p = p0 + k * VOX_THREADS;
            t[k] = 0;
            chk[k] = 0;
            prim[k] = 0;
            if (cell[k] < 0) continue;
            if (q[k] < 0) q[k] = claim_row(w, cell[k]);
            if (WIDE) {
                prim[k] = ~ordered_bits(refl[k]);
                chk[k] = coarse_chunk(s_ct, prim[k], (uint32_t)p);
            } else {
                chk[k] = geo_bin<C_OCT, C_SUB>((uint32_t)p, prm.bits);
            }
            t[k] = atomicAdd(w.cnt + (size_t)q[k] * NCHUNK + chk[k], 1);
        }
        // per point: (row, primary key) as one 8-byte record, and (chunk, ticket) as one 16-bit word
#pragma unroll
        for (int k = 0; k < SC_IT; ++k) {
            const int64_t p = p0 + k * VOX_THREADS;
            if (p >= n) continue;
            const int qq = cell[k] >= 0 ? q[k] : -1;
            if (WIDE) w.qk_of_point[p] = make_int2(qq, (int)prim[k]);
            else w.q_of_point[p] = qq;
            if (cell[k] >= 0) w.tick[p] = (uint16_t)((chk[k] << 8) | (t[k] < 255 ? t[k] : 255));
        }
    }
}

// ---- Q1: per occupied cell, before the placement -----------------------------------------------------------------
constexpr int Q1_THREADS = 1024;   // 32 cells per CTA, one warp each

// One warp per cell: counts -> inclusive prefix over the 64 chunks (two chunks per lane, coalesced), the saturation
// chunk, first[q] = +inf, and +inf in the slots the placement fills by sorted insertion.
template <typename K>
__global__ void __launch_bounds__(Q1_THREADS) vox_cell_prefix_kernel(const VoxParams prm, const VoxBuf w)
{
    pdl_enter();
    if (w.pt_on && __ldcg(w.pt_flag) == 0) return;
    const int nq = w.counters[0];
    const int lane = threadIdx.x & 31;
    const int P = prm.P;
    for (int q = blockIdx.x * (Q1_THREADS / 32) + (threadIdx.x >> 5); q < nq; q += gridDim.x * (Q1_THREADS / 32)) {
        int2 *c2 = reinterpret_cast<int2 *>(w.cnt + (size_t)q * NCHUNK) + lane;
        int2 v = *c2;
        v.y += v.x;
        int incl = v.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        const int excl = incl - v.y;
        v.x += excl;
        v.y += excl;
        *c2 = v;
        // first chunk at which the cell already holds max_points points: later chunks are dropped unseen
        int sat = v.x >= P ? 2 * lane : (v.y >= P ? 2 * lane + 1 : NCHUNK);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sat = min(sat, __shfl_xor_sync(0xFFFFFFFFu, sat, o));
        K *row = (K *)w.rows + (size_t)q * P;
        if (prm.ticket) {
            // only a truncated window (more points than free slots) is filled by sorted insertion
            if (sat < NCHUNK) {
                const int src = sat > 0 ? (sat - 1) >> 1 : 0;
                const int bx = __shfl_sync(0xFFFFFFFFu, v.x, src), by = __shfl_sync(0xFFFFFFFFu, v.y, src);
                const int wbase = sat == 0 ? 0 : (((sat - 1) & 1) ? by : bx);
                for (int s = wbase + lane; s < P; s += 32) row[s] = KeyInf<K>::value();
            }
        } else {
            for (int s = lane; s < P; s += 32) row[s] = KeyInf<K>::value();
        }
        if (lane == 0) {
            w.sat_of_q[q] = (uint8_t)sat;
            ((K *)w.first)[q] = KeyInf<K>::value();
        }
    }
}

template <int N> __device__ __forceinline__ bool any_active(const bool (&a)[N])
{
    bool r = false;
#pragma unroll
    for (int k = 0; k < N; ++k) r = r || a[k];
    return r;
}

// ---- C: per point, slot inside the pillar ------------------------------------------------------------------------
constexpr int PLACE_THREADS = 512;
constexpr int PLACE_IT = 2;                 // points per thread and iteration
constexpr int PLACE_TABLE = 40 * 1024;      // cells whose saturation chunk is cached in shared memory

// Persistent CTAs: the per-cell saturation chunk (1 byte per cell) is staged in shared memory once per CTA, so the
// majority of the points -- those of already full pillars -- are rejected without any random global access.
// The points of a cell's lowest occupied chunk (window base 0) also settle the cell's smallest key, first[q].
template <typename K>
__global__ void __launch_bounds__(PLACE_THREADS) vox_place_kernel(int64_t n, const VoxParams prm, const VoxBuf w)
{
    pdl_enter();
    if (w.pt_on && __ldcg(w.pt_flag) == 0) return;
    __shared__ __align__(16) uint8_t s_sat[PLACE_TABLE];
    constexpr bool WIDE = sizeof(K) == 8;
    const int nq = w.counters[0];
    const int ntab = nq < PLACE_TABLE ? nq : PLACE_TABLE;
    for (int i = threadIdx.x; i * 16 < ntab; i += PLACE_THREADS)
        reinterpret_cast<uint4 *>(s_sat)[i] = reinterpret_cast<const uint4 *>(w.sat_of_q)[i];
    __syncthreads();
    const int P = prm.P;
    // PLACE_IT points per thread and iteration (2: 20.2 us against 24.5 with 4 and 20.9 with 1): their loads, and later
    // their insertion chains, are issued back to back
    // (an in-order warp stalls at the first use of a result, so one point at a time would serialise every L2 round
    // trip of the chain).  The scatter kernel left (row, primary key) and (chunk, ticket) per point.
    constexpr int IT = PLACE_IT;
    const int64_t stride = (int64_t)gridDim.x * PLACE_THREADS;
    for (int64_t p0 = (int64_t)blockIdx.x * PLACE_THREADS + threadIdx.x; p0 < n; p0 += stride * IT) {
        int q[IT], tk[IT];
        uint32_t prim[IT];
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            const int64_t p = p0 + k * stride;
            q[k] = -1;
            tk[k] = 0;
            prim[k] = 0u;
            if (p < n) {
                if (WIDE) { const int2 r = w.qk_of_point[p]; q[k] = r.x; prim[k] = (uint32_t)r.y; }
                else q[k] = w.q_of_point[p];
                tk[k] = (int)w.tick[p];
            }
        }
        K key[IT];
        int ch[IT], base[IT], end[IT];
        bool act[IT];
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            const int64_t p = p0 + k * stride;
            act[k] = q[k] >= 0;
            base[k] = end[k] = 0;
            key[k] = 0;
            ch[k] = tk[k] >> 8;
            tk[k] &= 0xFF;
            if (!act[k]) continue;
            key[k] = WIDE ? (K)(((u64)prim[k] << 32) | (uint32_t)p) : (K)(uint32_t)p;
            const int sat = q[k] < PLACE_TABLE ? (int)s_sat[q[k]] : (int)w.sat_of_q[q[k]];
            if (ch[k] > sat) { act[k] = false; continue; }      // the pillar is full before this chunk (:303)
            const int32_t *incl = w.cnt + (size_t)q[k] * NCHUNK;
            base[k] = ch[k] ? incl[ch[k] - 1] : 0;
            end[k] = incl[ch[k]];
        }
        K *slot[IT];
        K carry[IT];
        int left[IT];
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            slot[k] = nullptr;
            carry[k] = key[k];
            left[k] = 0;
            if (!act[k]) continue;
            const int wend = end[k] < P ? end[k] : P;
            K *row = (K *)w.rows + (size_t)q[k] * P;
            if (end[k] - base[k] == 1) {                     // alone in its window: the slot is known
                row[base[k]] = key[k];
                if (base[k] == 0) ((K *)w.first)[q[k]] = key[k];      // ... and so is the cell's smallest key
                act[k] = false;
                continue;
            }
            if (base[k] == 0) key_min((K *)w.first + q[k], key[k]);  // lowest occupied chunk of the cell
            if (prm.ticket && end[k] <= P) {                 // the whole window is kept: arrival order now, the
                row[base[k] + tk[k]] = key[k];               // gather kernel sorts the row
                act[k] = false;
                continue;
            }
            // truncated window (more points than free slots): skip keys that can no longer enter it
            if (wend < end[k] && __ldcg(row + wend - 1) < key[k]) { act[k] = false; continue; }
            slot[k] = row + base[k];
            left[k] = wend - base[k];
        }
        // A few points share a window: sorted insertion with a lock-free atomicMin chain.  Slots only decrease and
        // every displaced key is pushed one slot down, so the final window is sorted for any interleaving.
        while (any_active(act)) {
            K old[IT];
#pragma unroll
            for (int k = 0; k < IT; ++k)
                if (act[k]) old[k] = key_min(slot[k], carry[k]);
#pragma unroll
            for (int k = 0; k < IT; ++k) {
                if (!act[k]) continue;
                if (old[k] == KeyInf<K>::value() || --left[k] == 0) { act[k] = false; continue; }
                carry[k] = old[k] > carry[k] ? old[k] : carry[k];
                ++slot[k];
            }
        }
    }
}

// ---- R: rank the cells by their smallest key -----------------------------------------------------------------------
// One cooperative launch, two grid barriers:
//   1  fine bin of first[q]; arrival index inside the bin (global histogram)
//   2  exclusive scan of the histogram (every CTA, shared memory); cells bucketed by bin, keys in bucket order
//   3  pillar id = bucket base + number of smaller keys in the bucket; coors, pillar_map, cutoff, voxel_num
constexpr int RANK_THREADS = 1024;

__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (*((volatile unsigned *)bar) < target) { }
        __threadfence();
    }
    __syncthreads();
}

template <typename K>
__global__ void __launch_bounds__(RANK_THREADS)
vox_rank_kernel(const VoxParams prm, const VoxBuf w, int32_t *__restrict__ coors, int32_t *__restrict__ voxel_num,
                int32_t *__restrict__ pillar_map, int64_t n_points)
{
    constexpr bool WIDE = sizeof(K) == 8;
    extern __shared__ __align__(16) unsigned char rk_smem[];
    int *s_base = reinterpret_cast<int *>(rk_smem);                          // [NBIN + 1]
    u64 *s_fine = reinterpret_cast<u64 *>(rk_smem + (NBIN + 4) * sizeof(int));   // [NFINE], later the staged keys
    __shared__ int s_warp[RANK_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_enter();
    const int nq = w.counters[0];
    unsigned *bar = (unsigned *)(w.counters + 8);
    if (WIDE) {
        for (int i = tid; i < NFINE - 1; i += RANK_THREADS) s_fine[i] = w.fine[i];
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid == 0) {
        *voxel_num = nq < prm.max_voxels ? nq : prm.max_voxels;
        *(K *)w.cutoff = KeyInf<K>::value();
    }
    // ---- 1
    // intervals [0, istar) get nsub sub-bins each, the others one bin: istar * nsub + (NFINE - istar) <= NBIN
    int istar = (int)(((int64_t)NFINE * 3 * nq) / (n_points > 0 ? n_points : 1)) + 8;
    istar = istar > NFINE - 1 ? NFINE - 1 : istar;
    const int nsub = (NBIN - NFINE) / istar;
    for (int q = blockIdx.x * RANK_THREADS + tid; q < nq; q += gridDim.x * RANK_THREADS) {
        const K f = ((const K *)w.first)[q];
        int bin;
        if (WIDE) {
            // Sample interval of the key, then a linear position inside the interval (on the reflectance bits).  A
            // cell with m points has its smallest key near the 1/m quantile, so the smallest keys crowd into the
            // first ~ nq / n of the key space: the intervals below `istar` share most of the bins.
            const int i = upper_bound_u64(s_fine, NFINE - 1, (u64)f);
            if (i >= istar) {
                bin = istar * nsub + (i - istar);
            } else {
                u64 lo, hi = s_fine[i];
                if (i > 0) lo = s_fine[i - 1];
                else { const u64 wd = s_fine[1] - s_fine[0]; lo = s_fine[0] > wd ? s_fine[0] - wd : 0ull; }
                int sub = 0;
                if ((u64)f > lo) {
                    const u64 num = (((u64)f - lo) >> 32) * (u64)nsub, den = ((hi - lo) >> 32) + 1ull;
                    sub = (int)(num / den);
                }
                bin = i * nsub + (sub < nsub - 1 ? sub : nsub - 1);
            }
        } else {
            bin = geo_bin<F_OCT, F_SUB>((uint32_t)f, prm.bits);
        }
        w.bin_of_q[q] = bin;
        w.pid_of_q[q] = atomicAdd(w.hist + bin, 1);            // arrival index inside the bin, until step 3
    }
    grid_barrier(bar, gridDim.x);
    // ---- 2
    {
        constexpr int PER = NBIN / RANK_THREADS;
        int v[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { v[k] = __ldcg(w.hist + tid * PER + k); sum += v[k]; }
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
        if (tid == RANK_THREADS - 1) s_base[NBIN] = base;
        __syncthreads();
    }
    for (int q = blockIdx.x * RANK_THREADS + tid; q < nq; q += gridDim.x * RANK_THREADS) {
        const int slot = s_base[w.bin_of_q[q]] + w.pid_of_q[q];
        w.list[slot] = q;
        ((K *)w.lkey)[slot] = ((const K *)w.first)[q];         // keys in bucket order: contiguous reads when ranking
    }
    grid_barrier(bar, 2 * gridDim.x);
    // ---- 3: eight lanes per bucket slot.  The cells' smallest keys crowd into the lowest bins (a cell with n points
    // has its minimum near the 1/n quantile), so buckets are long and shared by many slots: a CTA takes 128
    // consecutive slots per step and stages the key range of their buckets in shared memory.
    constexpr int GL = 8, SL = RANK_THREADS / GL, CH = 1024;
    K *s_keys = reinterpret_cast<K *>(s_fine);        // the splitters are no longer needed
    __shared__ int s_rng[2];
    const K *lkey = (const K *)w.lkey;
    const int sub = lane & (GL - 1);
    const unsigned gmask = 0xFFu << (lane & ~(GL - 1));
    for (int slot0 = blockIdx.x * SL; slot0 < nq; slot0 += gridDim.x * SL) {
        const int t = slot0 + tid / GL;
        const bool valid = t < nq;
        int q = 0, b0 = 0, b1 = 0;
        K f = 0;
        if (valid) {
            q = __ldcg(w.list + t);
            f = __ldcg(lkey + t);
            const int bin = __ldcg(w.bin_of_q + q);
            b0 = s_base[bin];
            b1 = s_base[bin + 1];
        }
        const int t_last = min(slot0 + SL, nq) - 1;
        if (tid == 0) s_rng[0] = b0;                                  // slots are in bucket order
        if (t == t_last && sub == 0) s_rng[1] = b1;
        __syncthreads();
        const int r0 = s_rng[0], r1 = s_rng[1];
        int cnt = 0;
        for (int c0 = r0; c0 < r1; c0 += CH) {
            const int cn = min(CH, r1 - c0);
            for (int i = tid; i < cn; i += RANK_THREADS) s_keys[i] = __ldcg(lkey + c0 + i);
            __syncthreads();
            const int j0 = max(b0, c0), j1 = min(b1, c0 + cn);
            for (int j = j0 + sub; j < j1; j += GL) cnt += (s_keys[j - c0] < f) ? 1 : 0;
            __syncthreads();
        }
#pragma unroll
        for (int o = GL / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(gmask, cnt, o);
        if (!valid || sub != 0) continue;
        const int rank = b0 + cnt;
        if (rank < prm.max_voxels) {
            const int c = w.cell_of_q[q];
            w.pid_of_q[q] = rank;
            w.q_of_pid[rank] = q;
            const int cx = c % prm.g[0], tt = c / prm.g[0];
            coors[rank * 3 + 0] = cx;
            coors[rank * 3 + 1] = tt % prm.g[1];
            coors[rank * 3 + 2] = tt / prm.g[1];
            if (pillar_map) pillar_map[c] = rank;
        } else {
            w.pid_of_q[q] = -1;
            if (rank == prm.max_voxels) *(K *)w.cutoff = f;    // the reference breaks here (:223, :291)
        }
    }
}

// ---- D: gather ---------------------------------------------------------------------------------------------------
// Keys at or after the cutoff (the first key of the pillar the reference breaks on) are dropped here: they are the
// largest keys of their rows, so the slots before them are unaffected.
template <typename K, bool VEC4>
__global__ void __launch_bounds__(VOX_THREADS)
vox_gather_kernel(const float *__restrict__ points, const int32_t *__restrict__ perm, const VoxBuf w,
                  const int32_t *__restrict__ voxel_num, int64_t max_rows, int P, int C, float *__restrict__ voxels,
                  int32_t *__restrict__ num_points)
{
    pdl_enter();
    int64_t t = (int64_t)blockIdx.x * VOX_THREADS + threadIdx.x;   // one thread per (pillar, slot)
    if (t >= max_rows * P) return;
    const int64_t m = t / P;
    const int s = (int)(t - m * P);
    if (m >= *voxel_num) return;
    const K cutoff = *(const K *)w.cutoff;
    const K *row = (const K *)w.rows + (size_t)w.q_of_pid[m] * P;
    const K key = row[s];
    const bool valid = key < cutoff;                         // +inf (empty slot) is never below the cutoff
    if (valid && (s == P - 1 || !(row[s + 1] < cutoff))) num_points[m] = s + 1;
    int64_t idx = 0;
    if (valid) {
        const uint32_t pos = (uint32_t)key;                  // low word = position / original index
        idx = perm ? (int64_t)(uint32_t)perm[pos] : (int64_t)pos;
    }
    if (VEC4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) v = __ldg(reinterpret_cast<const float4 *>(points) + idx);
        reinterpret_cast<float4 *>(voxels)[t] = v;
    } else {
        for (int c = 0; c < C; ++c) voxels[t * C + c] = valid ? __ldg(points + idx * C + c) : 0.f;
    }
}

// P <= 64: one warp per pillar.  The row's keys (windows are in order, the keys inside a window are not) are sorted
// in registers with a bitonic network, then lane s writes slot s.
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
__device__ __forceinline__ K shfl_down_key(K v, int d)
{
    if (sizeof(K) == 8) {
        unsigned lo = __shfl_down_sync(0xFFFFFFFFu, (unsigned)v, d), hi = __shfl_down_sync(0xFFFFFFFFu, (unsigned)((u64)v >> 32), d);
        return (K)(((u64)hi << 32) | lo);
    }
    return (K)__shfl_down_sync(0xFFFFFFFFu, (unsigned)v, d);
}
template <typename K>
__device__ __forceinline__ K shfl_up_key(K v, int d)
{
    if (sizeof(K) == 8) {
        unsigned lo = __shfl_up_sync(0xFFFFFFFFu, (unsigned)v, d), hi = __shfl_up_sync(0xFFFFFFFFu, (unsigned)((u64)v >> 32), d);
        return (K)(((u64)hi << 32) | lo);
    }
    return (K)__shfl_up_sync(0xFFFFFFFFu, (unsigned)v, d);
}

template <typename K, bool TWO>
__global__ void __launch_bounds__(VOX_THREADS)
vox_gather_sorted_kernel(const float *__restrict__ points, const int32_t *__restrict__ perm, const VoxBuf w,
                         const int32_t *__restrict__ voxel_num, int P, int C, int vec4, float *__restrict__ voxels,
                         int32_t *__restrict__ num_points)
{
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int nvox = *voxel_num;
    const K cutoff = *(const K *)w.cutoff;
    for (int m = blockIdx.x * (VOX_THREADS / 32) + (threadIdx.x >> 5); m < nvox; m += gridDim.x * (VOX_THREADS / 32)) {
        const int q = w.q_of_pid[m];
        const K *row = (const K *)w.rows + (size_t)q * P;
        const int total = w.cnt[(size_t)q * NCHUNK + NCHUNK - 1];      // inclusive prefix of the last chunk
        const int nk = total < P ? total : P;                          // slots [0, nk) were written by the placement
        K k0 = lane < nk ? row[lane] : KeyInf<K>::value();
        K k1 = (TWO && lane + 32 < nk) ? row[lane + 32] : KeyInf<K>::value();
        if (!(k0 < cutoff)) k0 = KeyInf<K>::value();
        if (!(k1 < cutoff)) k1 = KeyInf<K>::value();
        // bitonic sort of 32 (or 64) keys, element index e = r * 32 + lane
#pragma unroll
        for (int size = 2; size <= (TWO ? 64 : 32); size <<= 1) {
#pragma unroll
            for (int j = size >> 1; j > 0; j >>= 1) {
                if (j == 32) {
                    // partner is the other register of the same lane; e & size == 0 for both (size == 64): ascending
                    const K lo = k0 < k1 ? k0 : k1, hi = k0 < k1 ? k1 : k0;
                    k0 = lo; k1 = hi;
                } else {
                    const bool upper = (lane & j) != 0;
                    {
                        const K o = shfl_xor_key<K>(k0, j);
                        const bool asc = (lane & size) == 0;             // e = lane for register 0
                        const bool take_min = asc != upper;
                        k0 = take_min ? (k0 < o ? k0 : o) : (k0 < o ? o : k0);
                    }
                    if (TWO) {
                        const K o = shfl_xor_key<K>(k1, j);
                        const bool asc = ((lane + 32) & size) == 0;      // e = lane + 32 for register 1
                        const bool take_min = asc != upper;
                        k1 = take_min ? (k1 < o ? k1 : o) : (k1 < o ? o : k1);
                    }
                }
            }
        }
        const unsigned v0 = __ballot_sync(0xFFFFFFFFu, k0 != KeyInf<K>::value());
        const unsigned v1 = TWO ? __ballot_sync(0xFFFFFFFFu, k1 != KeyInf<K>::value()) : 0u;
        if (lane == 0) num_points[m] = __popc(v0) + __popc(v1);
#pragma unroll
        for (int r = 0; r < (TWO ? 2 : 1); ++r) {
            const int s = lane + 32 * r;
            if (s >= P) continue;
            const K key = r ? k1 : k0;
            const bool valid = key != KeyInf<K>::value();
            int64_t idx = 0;
            if (valid) {
                const uint32_t pos = (uint32_t)key;                  // low word = position / original index
                idx = perm ? (int64_t)(uint32_t)perm[pos] : (int64_t)pos;
            }
            const int64_t t = (int64_t)m * P + s;
            if (vec4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) v = __ldg(reinterpret_cast<const float4 *>(points) + idx);
                reinterpret_cast<float4 *>(voxels)[t] = v;
            } else {
                for (int c = 0; c < C; ++c) voxels[t * C + c] = valid ? __ldg(points + idx * C + c) : 0.f;
            }
        }
    }
}

// Gather fused with the single-layer PillarFeatureNet (C == 4, P <= 32, float4 points): the warp that has just sorted
// and gathered a pillar holds its points in registers in exactly the layout the PFN's decoration starts from (lane =
// slot), so it writes `voxels` / `num_points` AND runs the pillar through pfn_pillar: the PFN kernel, its launch and its
// re-read of the voxels disappear from the frame.
struct PfnArgs {
    const float *W, *scale, *shift;
    float *feat;
    int U;
    float vx, vy, x_off, y_off;
    int gx, gy;
};

constexpr int GP_THREADS = 128;      // 4 pillars per CTA

template <typename K>
__global__ void __launch_bounds__(GP_THREADS, 8)
vox_gather_pfn_kernel(const float *__restrict__ points, const int32_t *__restrict__ perm, const VoxBuf w,
                      const int32_t *__restrict__ voxel_num, int P, float *__restrict__ voxels,
                      int32_t *__restrict__ num_points, const PfnArgs pa)
{
    // the weights do not depend on the predecessor: load them before the dependency wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ __align__(16) float s_row[(GP_THREADS / 32) * 32 * PFN_LDI];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PfnWeights<9> pw;
    pfn_load_weights<9>(pw, pa.W, pa.scale, pa.shift, pa.U, lane);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int nvox = *voxel_num;
    const K cutoff = *(const K *)w.cutoff;
    float *row = s_row + warp * 32 * PFN_LDI;
    for (int m = blockIdx.x * (GP_THREADS / 32) + warp; m < nvox; m += gridDim.x * (GP_THREADS / 32)) {
        const int q = w.q_of_pid[m];
        const K *krow = (const K *)w.rows + (size_t)q * P;
        const int total = w.cnt[(size_t)q * NCHUNK + NCHUNK - 1];
        const int cell = w.cell_of_q[q];
        const int nk = total < P ? total : P;
        K k0 = lane < nk ? krow[lane] : KeyInf<K>::value();
        if (!(k0 < cutoff)) k0 = KeyInf<K>::value();
        // The placement leaves a row sorted up to the order inside a key-chunk window (a few keys in arrival order), so
        // a few odd-even transposition rounds finish it; the bitonic network (15 steps) is the fallback for rows that
        // are not nearly sorted.
        bool sorted = false;
#pragma unroll 1
        for (int round = 0; round < 5; ++round) {
            const K nx = shfl_down_key<K>(k0, 1);
            sorted = __ballot_sync(0xFFFFFFFFu, lane < 31 && nx < k0) == 0u;
            if (sorted) break;
            const K o = shfl_xor_key<K>(k0, 1);                       // pairs (0,1) (2,3) ...
            k0 = (lane & 1) ? (k0 < o ? o : k0) : (k0 < o ? k0 : o);
            const K up = shfl_down_key<K>(k0, 1), dn = shfl_up_key<K>(k0, 1);      // pairs (1,2) (3,4) ...
            if (lane & 1) { if (lane < 31) k0 = k0 < up ? k0 : up; }
            else if (lane > 0) k0 = k0 < dn ? dn : k0;
        }
        if (!sorted) {
#pragma unroll
            for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
                for (int j = size >> 1; j > 0; j >>= 1) {
                    const K o = shfl_xor_key<K>(k0, j);
                    const bool take_min = ((lane & size) == 0) != ((lane & j) != 0);
                    k0 = take_min ? (k0 < o ? k0 : o) : (k0 < o ? o : k0);
                }
            }
        }
        const bool valid = k0 != KeyInf<K>::value();
        const int n = __popc(__ballot_sync(0xFFFFFFFFu, valid));
        if (lane == 0) num_points[m] = n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            const uint32_t pos = (uint32_t)k0;                   // low word = position / original index
            const int64_t idx = perm ? (int64_t)(uint32_t)perm[pos] : (int64_t)pos;
            v = __ldg(reinterpret_cast<const float4 *>(points) + idx);
        }
        if (lane < P) reinterpret_cast<float4 *>(voxels)[(int64_t)m * P + lane] = v;
        float f[PFN_LDI];
#pragma unroll
        for (int k = 0; k < PFN_LDI; ++k) f[k] = 0.f;
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
        pfn_pillar<9>(pw, f, 4, P, n, cell % pa.gx, (cell / pa.gx) % pa.gy, pa.vx, pa.vy, pa.x_off, pa.y_off, row,
                      pa.feat + (int64_t)m * (pa.U + 1), pa.U, lane);
    }
}

// ---- partitioned front end, OPT-IN (PP_VOX_PATH=partition; replaces A0 / A / Q / C when every bin fits) ----------------
// Measured slower than the per-point kernels on B200 and therefore not the product path; it is
// bit-exact and covered by the front-end tests.
// The per-point kernels A and C pay two dependent random L2 round trips per point.  Here the points are first
// partitioned by cell group (group = low bits of the cell, so the cells of a dense cluster spread over the groups):
//   P1 vox_part_kernel    per tile of 4096 points: cell, key; rank inside the tile's share of each group with shared-
//                         memory atomics, ONE global atomic per (tile, group) to reserve bin space, records (cell, key)
//                         written in runs
//   P2 vox_select_kernel  one CTA per group, everything in shared memory: counting sort of the group's keys by local
//                         cell id (direct table, ceil(cells / G) entries; the keys of cells with more than 32 points are
//                         also counted per key chunk), then one warp per cell keeps the max_points smallest keys
//                         (unsorted: the gather kernels sort a row): rows[q], first[q], cell_of_q[q] -- what the
//                         ranking and gather kernels read
// A bin that overflows (a group with more than pt_cap points) raises pt_flag: P2 returns and the kernels A / Q / C,
// which otherwise exit at once, do the frame.
constexpr int PT_THREADS = 512, PT_IT = 8, PT_TILE = PT_THREADS * PT_IT;

template <typename K>
__global__ void __launch_bounds__(PT_THREADS)
vox_part_kernel(const float *__restrict__ points, int64_t n, const VoxParams prm, const int32_t *__restrict__ perm,
                const VoxBuf w)
{
    // PDL as in kernel A: the first tile is read and binned before the dependency wait (the init kernel orders
    // wait -> trigger, so the producer of the points is complete)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ int pt_smem[];
    constexpr bool WIDE = sizeof(K) == 8;
    const int G = 1 << w.pt_lg, gmask = G - 1, cap = w.pt_cap;
    int *s_hist = pt_smem, *s_base = pt_smem + G;
    const int tid = threadIdx.x;
    bool waited = false;
    for (int64_t t0 = (int64_t)blockIdx.x * PT_TILE; t0 < n; t0 += (int64_t)gridDim.x * PT_TILE) {
        for (int g = tid; g < G; g += PT_THREADS) s_hist[g] = 0;
        __syncthreads();
        int32_t cell[PT_IT];
        int r[PT_IT];
        uint32_t prim[PT_IT];
#pragma unroll
        for (int k = 0; k < PT_IT; ++k) {
            const int64_t p = t0 + k * PT_THREADS + tid;
            cell[k] = -1;
            prim[k] = 0u;
            r[k] = 0;
            if (p >= n) continue;
            const int64_t idx = perm ? (int64_t)(uint32_t)perm[p] : p;
            float x, y, z, refl = 0.f;
            if (prm.vec4) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(points) + idx);
                x = v.x; y = v.y; z = v.z; refl = v.w;
            } else {
                const float *pt = points + idx * prm.C;
                x = __ldg(pt); y = __ldg(pt + 1); z = __ldg(pt + 2);
                if (WIDE) refl = __ldg(pt + 3);
            }
            cell[k] = point_cell(prm, x, y, z);
            if (WIDE) prim[k] = ~ordered_bits(refl);
        }
#pragma unroll
        for (int k = 0; k < PT_IT; ++k)
            if (cell[k] >= 0) r[k] = atomicAdd(s_hist + (cell[k] & gmask), 1);
        if (!waited) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            waited = true;
        }
        __syncthreads();
        for (int g = tid; g < G; g += PT_THREADS) {
            const int c = s_hist[g];
            if (c) s_base[g] = atomicAdd(w.pt_cursor + g, c);
        }
        __syncthreads();
        bool over = false;
#pragma unroll
        for (int k = 0; k < PT_IT; ++k) {
            if (cell[k] < 0) continue;
            const int64_t p = t0 + k * PT_THREADS + tid;
            const int g = cell[k] & gmask;
            const int pos = s_base[g] + r[k];
            if (pos >= cap) { over = true; continue; }
            const size_t at = (size_t)g * cap + pos;
            w.pt_cell[at] = cell[k];
            ((K *)w.pt_key)[at] = WIDE ? (K)(((u64)prim[k] << 32) | (uint32_t)p) : (K)(uint32_t)p;
        }
        if (over) *w.pt_flag = 1;
        __syncthreads();
    }
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

// the 32 smallest of two ascending 32-key sequences, ascending
template <typename K>
__device__ __forceinline__ K warp_merge_low(K a, K b, int lane)
{
    const K o = shfl_idx_key<K>(b, 31 - lane);
    K m = a < o ? a : o;                                  // bitonic
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const K x = shfl_xor_key<K>(m, j);
        m = (lane & j) ? (m < x ? x : m) : (m < x ? m : x);
    }
    return m;
}

// key chunk (0 .. NCHUNK-1, monotone in the key): sample quantiles (64-bit keys) or geometric in the position
template <typename K>
__device__ __forceinline__ int key_chunk(const CoarseTable &t, K key, int bits)
{
    if (sizeof(K) == 8) return coarse_chunk(t, (uint32_t)((u64)key >> 32), (uint32_t)key);
    return geo_bin<C_OCT, C_SUB>((uint32_t)key, bits);
}

constexpr int SEL_THREADS = 512, SEL_RPT = 16, SEL_WARPS = SEL_THREADS / 32;   // pt_cap <= SEL_THREADS * SEL_RPT
constexpr int SEL_PEND = 64;

template <typename K>
__global__ void __launch_bounds__(SEL_THREADS, 2)
vox_select_kernel(const VoxParams prm, const VoxBuf w)
{
    pdl_enter();
    if (__ldcg(w.pt_flag) != 0) return;
    extern __shared__ __align__(16) unsigned char sel_smem[];
    const int cap = w.pt_cap, D = w.pt_D, lg = w.pt_lg, G = 1 << lg, P = prm.P;
    K *s_seg = reinterpret_cast<K *>(sel_smem);                               // [cap] keys, grouped by cell
    K *s_pend = s_seg + cap;                                                  // [SEL_WARPS][SEL_PEND]
    int *s_off = reinterpret_cast<int *>(s_pend + SEL_WARPS * SEL_PEND);      // [D + 1] counts, then exclusive offsets
    const int nbig_max = cap / 33 + 1;                                        // cells with more than 32 points
    uint32_t *s_cc = reinterpret_cast<uint32_t *>(s_off + D + 1);             // [nbig_max][NCHUNK / 2] chunk counts, 16 bit
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_cc + nbig_max * (NCHUNK / 2));   // [D] occupied local ids
    uint8_t *s_aux = reinterpret_cast<uint8_t *>(s_list + D);                 // [cap] chunk of the key at that position
    uint8_t *s_big = s_aux + cap;                                             // [D] index of a big cell's counters
    __shared__ int s_warp[SEL_WARPS], s_qbase, s_next, s_nbig;
    __shared__ CoarseTable s_ct;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (D + SEL_THREADS - 1) / SEL_THREADS;                      // table entries per thread (<= 16)
    if (sizeof(K) == 8) load_coarse(s_ct, w.coarse);                          // visible after the first barrier below
    for (int g = blockIdx.x; g < G; g += gridDim.x) {
        const int R = __ldcg(w.pt_cursor + g);
        if (R == 0) continue;
        for (int i = tid; i <= D; i += SEL_THREADS) s_off[i] = 0;
        for (int i = tid; i < nbig_max * (NCHUNK / 2); i += SEL_THREADS) s_cc[i] = 0u;
        if (tid == 0) s_nbig = 0;
        __syncthreads();
        const int32_t *gcell = w.pt_cell + (size_t)g * cap;
        const K *gkey = (const K *)w.pt_key + (size_t)g * cap;
        // A: count per local cell id; a record remembers (local id, arrival index)
        uint32_t rec[SEL_RPT];
#pragma unroll
        for (int j = 0; j < SEL_RPT; ++j) {
            if (j * SEL_THREADS >= R) break;              // uniform: a group rarely needs all SEL_RPT rounds
            const int i = tid + j * SEL_THREADS;
            rec[j] = 0;
            if (i < R) {
                const int l = __ldcg(gcell + i) >> lg;
                rec[j] = ((uint32_t)l << 16) | (uint32_t)atomicAdd(s_off + l, 1);
            }
        }
        __syncthreads();
        // exclusive scan of the counts (high half) and of the occupied flags (low half), `per` entries per thread
        int cnt[16];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) cnt[k] = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k >= per) break;
            const int l = tid * per + k;
            cnt[k] = l < D ? s_off[l] : 0;
            sum += (cnt[k] << 16) + (cnt[k] ? 1 : 0);
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int base = incl - sum, total = 0;
#pragma unroll
        for (int k = 0; k < SEL_WARPS; ++k) {
            const int v = s_warp[k];
            if (k < warp) base += v;
            total += v;
        }
        const int ncell = total & 0xFFFF;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k >= per) break;
            const int l = tid * per + k;
            if (l < D) {
                s_off[l] = base >> 16;
                if (cnt[k]) s_list[base & 0xFFFF] = (uint16_t)l;
                s_big[l] = cnt[k] > 32 ? (uint8_t)atomicAdd(&s_nbig, 1) : (uint8_t)0xFF;
                base += (cnt[k] << 16) + (cnt[k] ? 1 : 0);
            }
        }
        if (tid == 0) {
            s_off[D] = R;
            s_qbase = atomicAdd(w.counters, ncell);
            s_next = 0;
        }
        __syncthreads();
        // B: keys into their cell's segment; the keys of big cells (more than 32 points) are also counted per key chunk
#pragma unroll
        for (int j = 0; j < SEL_RPT; ++j) {
            if (j * SEL_THREADS >= R) break;
            const int i = tid + j * SEL_THREADS;
            if (i < R) {
                const int l = (int)(rec[j] >> 16);
                const int pos = s_off[l] + (int)(rec[j] & 0xFFFFu);
                const K key = __ldcg(gkey + i);
                s_seg[pos] = key;
                const int bg = s_big[l];
                if (bg != 0xFF) {
                    const int ch = key_chunk<K>(s_ct, key, prm.bits);
                    s_aux[pos] = (uint8_t)ch;
                    atomicAdd(s_cc + bg * (NCHUNK / 2) + (ch >> 1), 1u << ((ch & 1) * 16));
                }
            }
        }
        __syncthreads();
        // C: one warp per cell (dynamic assignment).  Rows are written UNSORTED (the gather kernels sort a row in
        // registers); what matters here is which keys are kept, and the cell's smallest key.
        //   n <= P      every key is kept
        //   n <= 32     one bitonic sort, the first P
        //   else        the old algorithm, in shared memory: histogram of the cell's keys over the 64 key chunks ->
        //               saturation chunk; keys of lower chunks are kept as they come, the r free slots go to the r
        //               smallest keys of the saturation chunk's window (a handful of keys).  A window wider than a warp
        //               (heavily tied or concentrated keys) falls back to a streaming top-P selection over the cell.
        const int qbase = s_qbase;
        K *pend = s_pend + warp * SEL_PEND;
        for (;;) {
            int c = 0;
            if (lane == 0) c = atomicAdd(&s_next, 1);
            c = __shfl_sync(0xFFFFFFFFu, c, 0);
            if (c >= ncell) break;
            const int l = s_list[c];
            const int off = s_off[l], nn = s_off[l + 1] - off;
            const K *seg = s_seg + off;
            const int q = qbase + c;
            K *row = (K *)w.rows + (size_t)q * P;
            K kmin = KeyInf<K>::value();                  // lane-local minimum of the kept keys
            if (nn <= P) {
                if (lane < nn) kmin = seg[lane];
                if (lane < P) row[lane] = kmin;
            } else if (nn <= 32) {
                K best = lane < nn ? seg[lane] : KeyInf<K>::value();
                best = warp_sort32<K>(best, lane);
                if (lane < P) row[lane] = best;
                kmin = best;
            } else {
                const uint32_t cw = s_cc[(int)s_big[l] * (NCHUNK / 2) + lane];
                const int c0 = (int)(cw & 0xFFFFu), c1 = (int)(cw >> 16);
                const uint8_t *aux = s_aux + off;
                int incl = c0 + c1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += t;
                }
                const int e0 = incl - c0 - c1, e1 = e0 + c0;          // exclusive prefixes of chunks 2*lane, 2*lane + 1
                int sat = e1 >= P ? 2 * lane : (incl >= P ? 2 * lane + 1 : NCHUNK);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sat = min(sat, __shfl_xor_sync(0xFFFFFFFFu, sat, o));
                const int base = __shfl_sync(0xFFFFFFFFu, (sat & 1) ? e1 : e0, sat >> 1);      // keys below the window
                const int wn = __shfl_sync(0xFFFFFFFFu, (sat & 1) ? c1 : c0, sat >> 1);        // keys in the window
                const int r = P - base;                                                       // 1 <= r <= wn
                __syncwarp();
                if (wn > 32) {
                    K best = KeyInf<K>::value(), thr = KeyInf<K>::value();
                    int np = 0;
                    for (int b = 0; b < nn; b += 32) {
                        const K k = b + lane < nn ? seg[b + lane] : KeyInf<K>::value();
                        const bool pass = k < thr;
                        const unsigned mask = __ballot_sync(0xFFFFFFFFu, pass);
                        if (!mask) continue;
                        if (pass) pend[np + __popc(mask & lanemask_lt())] = k;
                        np += __popc(mask);
                        if (np >= 32) {
                            __syncwarp();
                            K cnd = pend[lane];
                            const K rest = pend[32 + lane];
                            __syncwarp();
                            np -= 32;
                            if (lane < np) pend[lane] = rest;
                            __syncwarp();
                            cnd = warp_sort32<K>(cnd, lane);
                            best = warp_merge_low<K>(best, cnd, lane);
                            thr = shfl_idx_key<K>(best, P - 1);
                        }
                    }
                    if (np > 0) {
                        __syncwarp();
                        K cnd = lane < np ? pend[lane] : KeyInf<K>::value();
                        cnd = warp_sort32<K>(cnd, lane);
                        best = warp_merge_low<K>(best, cnd, lane);
                    }
                    __syncwarp();
                    if (lane < P) row[lane] = best;
                    kmin = best;
                } else {
                    int nk = 0, nw = 0;
                    for (int b = 0; b < nn; b += 32) {
                        const bool in = b + lane < nn;
                        const K k = in ? seg[b + lane] : KeyInf<K>::value();
                        const int ch = in ? (int)aux[b + lane] : NCHUNK;
                        const unsigned mk = __ballot_sync(0xFFFFFFFFu, ch < sat), mw = __ballot_sync(0xFFFFFFFFu, ch == sat);
                        if (ch < sat) {
                            row[nk + __popc(mk & lanemask_lt())] = k;
                            kmin = k < kmin ? k : kmin;
                        } else if (ch == sat) {
                            pend[nw + __popc(mw & lanemask_lt())] = k;
                        }
                        nk += __popc(mk);
                        nw += __popc(mw);
                    }
                    __syncwarp();
                    // the r smallest of the window's wn keys: rank by counting
                    const K kw = lane < wn ? pend[lane] : KeyInf<K>::value();
                    int rank = 0;
                    for (int j = 0; j < wn; ++j) rank += (pend[j] < kw) ? 1 : 0;
                    if (lane < wn && rank < r) {
                        row[base + rank] = kw;
                        kmin = kw < kmin ? kw : kmin;
                    }
                    __syncwarp();
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const K x = shfl_xor_key<K>(kmin, o);
                kmin = x < kmin ? x : kmin;
            }
            if (lane == 0) {
                ((K *)w.first)[q] = kmin;
                w.cell_of_q[q] = (l << lg) | g;
                w.cnt[(size_t)q * NCHUNK + NCHUNK - 1] = nn;      // the gather kernels read the cell's point count here
            }
        }
        __syncthreads();
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

// Partitioned front end: G = 2^lg groups of about <= 1024 points, ceil(cells / G) <= 8192 local cell ids per group (the
// direct table of vox_select_kernel), bins of `cap` records (4 x the mean, at most what one CTA sorts in shared memory).
// Everything here depends on (n, cfg) only, so the workspace size and the launch sequence agree.
constexpr int PT_MAX_LG = 13, PT_MAX_D = 8192;
struct PartPlan {
    int on, lg, cap, D;
};
PartPlan plan_part(int64_t n, const pp_voxel_cfg *c)
{
    PartPlan p = {0, 0, 0, 0};
    // Opt-in (PP_VOX_PATH=partition): measured on B200 at 1e6 points this front end is slower than the per-point
    // kernels (P1 15 us + P2 36 us against 22 + 17 us, and 74 against 57 us per frame with 24 frames in flight), see
    // DESIGN.md section 7; it stays as a tested alternative, not as the product path.
    const char *path = getenv("PP_VOX_PATH");
    if (!path || strcmp(path, "partition") != 0) return p;
    if (c->max_points > 32 || n < 1) return p;
    const int64_t cells = (int64_t)c->grid[0] * c->grid[1] * c->grid[2];
    int lg = 3;
    while (lg < PT_MAX_LG && (n >> lg) > 1024) ++lg;
    while (lg < PT_MAX_LG && ((cells + ((int64_t)1 << lg) - 1) >> lg) > PT_MAX_D) ++lg;
    const int64_t D = (cells + ((int64_t)1 << lg) - 1) >> lg;
    if (D > PT_MAX_D) return p;
    int64_t cap = 4 * ((n + ((int64_t)1 << lg) - 1) >> lg);
    cap = (cap + 511) / 512 * 512;
    cap = cap < 1024 ? 1024 : (cap > SEL_THREADS * SEL_RPT ? SEL_THREADS * SEL_RPT : cap);
    const char *ecap = getenv("PP_VOX_CAP");               // tests: a small capacity forces the overflow path
    if (ecap && atoi(ecap) > 0 && atoi(ecap) < cap) cap = atoi(ecap);
    p.on = 1; p.lg = lg; p.cap = (int)cap; p.D = (int)D;
    return p;
}

inline int64_t units16(size_t bytes) { return (int64_t)(align_up(bytes, 16) / 16); }

Carve carve(void *ws, int64_t n, const pp_voxel_cfg *c, bool wide, size_t *total)
{
    Carve r;
    const int64_t n1 = n > 0 ? n : 1;
    const int64_t cells = (int64_t)c->grid[0] * c->grid[1] * c->grid[2];
    r.Q = n1 < cells ? n1 : cells;
    const size_t ksz = wide ? 8 : 4;
    // counter rows zeroed up front: a bit more than the pillar cap; beyond that the claiming thread zeroes its row
    int64_t r_rows = (int64_t)c->max_voxels + c->max_voxels / 4 + 1024;
    if (r_rows > r.Q) r_rows = r.Q;
    r.b.r_rows = (int32_t)r_rows;
    Arena a(ws, (size_t)-1);
    r.b.map = a.take<int32_t>((size_t)cells);
    r.b.cutoff = a.take<u64>(8);
    const size_t ff0_bytes = a.off;                          // map + cutoff, contiguous, 0xFF
    r.b.counters = a.take<int32_t>(64);
    r.b.hist = a.take<int32_t>(NBIN);
    r.b.fill = a.take<int32_t>(16);
    const PartPlan pl = plan_part(n, c);
    r.b.pt_on = pl.on; r.b.pt_lg = pl.lg; r.b.pt_cap = pl.cap; r.b.pt_D = pl.D;
    r.b.pt_flag = r.b.counters + 16;
    r.b.pt_cursor = a.take<int32_t>(pl.on ? ((size_t)1 << pl.lg) : 4);      // zeroed with the counters
    const size_t z0_off = (size_t)((char *)r.b.counters - (char *)ws), z0_bytes = a.off - z0_off;
    r.b.base = a.take<int32_t>(NFINE + 1);
    r.b.cell_of_q = a.take<int32_t>((size_t)r.Q);
    r.b.first = a.take<char>((size_t)r.Q * ksz);
    r.b.cnt = a.take<int32_t>((size_t)r.Q * NCHUNK);
    r.b.rows = a.take<char>((size_t)r.Q * c->max_points * ksz);
    r.b.pid_of_q = a.take<int32_t>((size_t)r.Q);
    r.b.bin_of_q = a.take<int32_t>((size_t)r.Q);
    r.b.sat_of_q = a.take<uint8_t>((size_t)r.Q);
    r.b.q_of_point = wide ? nullptr : a.take<int32_t>((size_t)n1);
    r.b.qk_of_point = wide ? a.take<int2>((size_t)n1) : nullptr;
    r.b.tick = a.take<uint16_t>((size_t)n1);
    r.b.list = a.take<int32_t>((size_t)r.Q);
    r.b.lkey = a.take<char>((size_t)r.Q * ksz);
    r.b.q_of_pid = a.take<int32_t>((size_t)max_rows_of(n, c));
    r.b.coarse = a.take<u64>(NCHUNK);
    r.b.fine = a.take<u64>(NFINE);
    const size_t recs = pl.on ? ((size_t)pl.cap << pl.lg) : 4;
    r.b.pt_cell = a.take<int32_t>(recs);
    r.b.pt_key = a.take<char>(recs * ksz);
    *total = align_up(a.off);
    // every array starts 256-byte aligned, so rounding the fills up to 16 bytes stays inside the padding
    r.ia.ff_ptr[0] = (int4 *)r.b.map;    r.ia.ff_n[0] = units16(ff0_bytes);
    r.ia.ff_ptr[1] = nullptr;            r.ia.ff_n[1] = 0;
    r.ia.ff_ptr[2] = nullptr;            r.ia.ff_n[2] = 0;
    r.ia.ff_ptr[3] = nullptr;            r.ia.ff_n[3] = 0;      // optional pillar_map, set by the caller
    r.ia.z_ptr[0] = (int4 *)((char *)ws + z0_off);  r.ia.z_n[0] = units16(z0_bytes);
    r.ia.z_ptr[1] = (int4 *)r.b.cnt;     r.ia.z_n[1] = units16((size_t)r_rows * NCHUNK * 4);
    return r;
}

template <typename K>
int run(const float *points, int64_t n, const VoxParams &prm, const int32_t *perm, const Carve &cv, float *voxels,
        int32_t *coors, int32_t *num_points, int32_t *voxel_num, int32_t *pillar_map, int64_t max_rows, const PfnArgs *pfn,
        cudaStream_t st)
{
    const VoxBuf &w = cv.b;
    constexpr bool WIDE = sizeof(K) == 8;
    int64_t fill_units = 0;
    for (int r = 0; r < 4; ++r) fill_units += cv.ia.ff_n[r];
    for (int r = 0; r < 2; ++r) fill_units += cv.ia.z_n[r];
    int init_blocks = (int)ceil_div(fill_units, 1024 * 4);
    init_blocks = init_blocks < 1 ? 1 : (init_blocks > 148 * 2 ? 148 * 2 : init_blocks);
    launch_pdl(vox_init_kernel, dim3(init_blocks + (WIDE ? 1 : 0)), dim3(1024), 0, st, points, n, prm.C, WIDE ? 1 : 0, w.coarse, w.fine, cv.ia);
    if (int rc = check_launch("vox_init_kernel")) return rc;
    // persistent grid: exactly the CTAs that are resident at once (no partial last wave)
    static int sc_resident = 0;
    if (!sc_resident) {
        int per_sm = 0, dev = 0, sms = 0;
        PP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vox_scatter_kernel<K>, VOX_THREADS, 0));
        PP_CUDA_TRY(cudaGetDevice(&dev));
        PP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        sc_resident = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
    }
    if (w.pt_on) {
        const int G = 1 << w.pt_lg;
        const size_t p1_smem = (size_t)2 * G * sizeof(int);
        const size_t p2_smem = ((size_t)w.pt_cap + SEL_WARPS * SEL_PEND) * sizeof(K) + ((size_t)w.pt_D + 1) * sizeof(int) +
                               (size_t)(w.pt_cap / 33 + 1) * (NCHUNK / 2) * sizeof(uint32_t) + (size_t)w.pt_D * sizeof(uint16_t) +
                               (size_t)w.pt_cap + (size_t)w.pt_D + 16;
        static size_t p1_set = 48 * 1024, p2_set = 48 * 1024;
        if (p1_smem > p1_set) {
            PP_CUDA_TRY(cudaFuncSetAttribute(vox_part_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p1_smem));
            p1_set = p1_smem;
        }
        if (p2_smem > p2_set) {
            PP_CUDA_TRY(cudaFuncSetAttribute(vox_select_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2_smem));
            p2_set = p2_smem;
        }
        const int64_t tiles = ceil_div(n, PT_TILE);
        launch_pdl(vox_part_kernel<K>, dim3((unsigned)(tiles < 148 * 2 ? tiles : 148 * 2)), dim3(PT_THREADS), p1_smem, st, points,
                   n, prm, perm, w);
        if (int rc = check_launch("vox_part_kernel")) return rc;
        launch_pdl(vox_select_kernel<K>, dim3((unsigned)(G < 148 * 4 ? G : 148 * 4)), dim3(SEL_THREADS), p2_smem, st, prm, w);
        if (int rc = check_launch("vox_select_kernel")) return rc;
    }
    if (!w.pt_on && n >= 65536) {          // small inputs: the extra launch costs more than the claims
        launch_pdl(vox_preclaim_kernel, dim3((unsigned)ceil_div(ceil_div(n, PRECLAIM_STRIDE), VOX_THREADS)), dim3(VOX_THREADS), 0, st,
                   points, n, prm, w);
        if (int rc = check_launch("vox_preclaim_kernel")) return rc;
    }
    const int64_t sc_blocks = ceil_div(n, VOX_THREADS * SC_IT);
    launch_pdl(vox_scatter_kernel<K>, dim3((unsigned)(sc_blocks < sc_resident ? sc_blocks : sc_resident)), dim3(VOX_THREADS), 0, st,
               points, n, prm, perm, w);
    if (int rc = check_launch("vox_scatter_kernel")) return rc;
    const unsigned cap = 148 * 8;    // persistent-style grids: the cell count is only known on the device
    auto capped = [&](int64_t blocks) { return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap); };
    launch_pdl(vox_cell_prefix_kernel<K>, dim3(capped(ceil_div(cv.Q, Q1_THREADS / 32)) / 4 + 1), dim3(Q1_THREADS), 0, st, prm, w);
    if (int rc = check_launch("vox_cell_prefix_kernel")) return rc;
    const int place_per_sm = 4;
    {
        const int64_t want = ceil_div(n, PLACE_THREADS * PLACE_IT), cap_p = 148 * place_per_sm;
        launch_pdl(vox_place_kernel<K>, dim3((unsigned)(want < cap_p ? want : cap_p)), dim3(PLACE_THREADS), 0, st, n, prm, w);
    }
    if (int rc = check_launch("vox_place_kernel")) return rc;
    {
        // cooperative: the CTAs meet at two grid barriers, so they must all be resident
        int64_t rb = ceil_div(cv.Q, RANK_THREADS);
        rb = rb < 1 ? 1 : (rb > 148 ? 148 : rb);
        const size_t rk_smem = (NBIN + 4) * sizeof(int) + NFINE * sizeof(u64);
        static bool rk_attr = false;
        if (!rk_attr) {
            PP_CUDA_TRY(cudaFuncSetAttribute(vox_rank_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rk_smem));
            rk_attr = true;
        }
        void *args[] = {(void *)&prm, (void *)&w, (void *)&coors, (void *)&voxel_num, (void *)&pillar_map, (void *)&n};
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)rb);
        cfg.blockDim = dim3(RANK_THREADS);
        cfg.dynamicSmemBytes = rk_smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 2;
        PP_CUDA_TRY(cudaLaunchKernelExC(&cfg, (const void *)vox_rank_kernel<K>, args));
        if (int rc = check_launch("vox_rank_kernel")) return rc;
    }
    const bool vec4 = prm.vec4 && ((uintptr_t)voxels % 16 == 0);
    if (pfn) {
        const int64_t want = ceil_div(max_rows, GP_THREADS / 32);
        launch_pdl(vox_gather_pfn_kernel<K>, dim3((unsigned)(want < 148 * 8 ? want : 148 * 8)), dim3(GP_THREADS), 0, st, points,
                   perm, w, (const int32_t *)voxel_num, prm.P, voxels, num_points, *pfn);
        return check_launch("vox_gather_pfn_kernel");
    }
    if (prm.ticket) {
        const unsigned gs = (unsigned)(ceil_div(max_rows, VOX_THREADS / 32) < 148 * 8 ? ceil_div(max_rows, VOX_THREADS / 32) : 148 * 8);
        if (prm.P <= 32)
            launch_pdl(vox_gather_sorted_kernel<K, false>, dim3(gs), dim3(VOX_THREADS), 0, st, points, perm, w,
                       (const int32_t *)voxel_num, prm.P, prm.C, vec4 ? 1 : 0, voxels, num_points);
        else
            launch_pdl(vox_gather_sorted_kernel<K, true>, dim3(gs), dim3(VOX_THREADS), 0, st, points, perm, w,
                       (const int32_t *)voxel_num, prm.P, prm.C, vec4 ? 1 : 0, voxels, num_points);
        return check_launch("vox_gather_sorted_kernel");
    }
    const int64_t slots = max_rows * prm.P;
    const unsigned gb = (unsigned)ceil_div(slots, VOX_THREADS);
    if (vec4)
        launch_pdl(vox_gather_kernel<K, true>, dim3(gb), dim3(VOX_THREADS), 0, st, points, perm, w, (const int32_t *)voxel_num, max_rows, prm.P, prm.C, voxels, num_points);
    else
        launch_pdl(vox_gather_kernel<K, false>, dim3(gb), dim3(VOX_THREADS), 0, st, points, perm, w, (const int32_t *)voxel_num, max_rows, prm.P, prm.C, voxels, num_points);
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
                         const pp_pfn_fused *pfn, void *workspace, size_t workspace_bytes, pp_stream_t stream);

extern "C" int pp_voxelize(const float *points, int64_t n, const pp_voxel_cfg *cfg, int order, const int32_t *perm,
                           float *voxels, int32_t *coors, int32_t *num_points, int32_t *voxel_num,
                           int32_t *pillar_map, void *workspace, size_t workspace_bytes, pp_stream_t stream)
{
    return voxelize_impl(points, n, cfg, order, perm, voxels, coors, num_points, voxel_num, pillar_map, nullptr, workspace,
                         workspace_bytes, stream);
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
    return voxelize_impl(points, n, cfg, order, perm, voxels, coors, num_points, voxel_num, pillar_map, pfn, workspace,
                         workspace_bytes, stream);
}

static int voxelize_impl(const float *points, int64_t n, const pp_voxel_cfg *cfg, int order, const int32_t *perm,
                         float *voxels, int32_t *coors, int32_t *num_points, int32_t *voxel_num, int32_t *pillar_map,
                         const pp_pfn_fused *pfn, void *workspace, size_t workspace_bytes, pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    PP_REQUIRE(cfg && voxel_num, "null cfg / voxel_num");
    PP_REQUIRE(n >= 0 && n < (1ll << 30), "n_points out of range");
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
    q.vec4 = (q.C == 4 && ((uintptr_t)points % 16 == 0)) ? 1 : 0;
    int bits = 0;
    while (((int64_t)1 << bits) < n) ++bits;                 // positions < 2^bits
    q.bits = bits;
    q.ticket = q.P <= 64 ? 1 : 0;

    if (pillar_map) {          // filled with -1 by the init kernel (needs 16-byte alignment; else a memset)
        if (((uintptr_t)pillar_map % 16 == 0) && (cells % 4 == 0)) {
            cv.ia.ff_ptr[3] = (int4 *)pillar_map;
            cv.ia.ff_n[3] = cells / 4;
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
        pa.vx = pfn->vx; pa.vy = pfn->vy; pa.x_off = pfn->x_off; pa.y_off = pfn->y_off;
        pa.gx = cfg->grid[0]; pa.gy = cfg->grid[1];
        pap = &pa;
    }
    if (wide)
        return run<u64>(points, n, q, order_perm, cv, voxels, coors, num_points, voxel_num, pillar_map, max_rows, pap, st);
    return run<uint32_t>(points, n, q, order_perm, cv, voxels, coors, num_points, voxel_num, pillar_map, max_rows, pap, st);
}
