// Hard voxelization on sm_100a, bit-exact with the reference's sequential first-come pass
// (ops/ops_numba.py:171-308), restated as order-independent parallel steps:
//
//   position p        = index of a point in processing order (given / reflectance-desc / perm)
//   K1 cell           : cell id of every position; first[cell] = atomicMin(position)
//   K2 assign         : a position is a "first arrival" iff first[cell] == p.  The pillar id is the
//                       number of first arrivals before p (single-pass decoupled look-back scan);
//                       the first arrival with id == max_voxels is the reference's `break`
//                       (:223, :291): its position is the cutoff, everything at or after it is dropped.
//   K3 rank           : every surviving position inserts itself into its pillar's sorted row of the
//                       P smallest positions (lock-free atomicMin insertion chain; the final row is
//                       independent of thread scheduling) -> slot = arrival rank of the reference.
//   K4 gather         : voxels[m][s] = points[row[m][s]], zero padded; num_points[m] = filled slots.
//
// Everything but the 16 B/point read and the output write is L2-resident workspace traffic.
#include "pp_common.cuh"
#include "pp_sort.cuh"

namespace pp {
namespace {

constexpr int VOX_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = VOX_THREADS * SCAN_ITEMS;

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VAL_MASK = ~FLAG_MASK;

struct VoxParams {
    double r[3], v[3];
    float rf[3], vf[3];
    int g[3];
    int regime;   // 0: all f32   1: sub f32, div f64   2: all f64   (numba promotion)
    int P, max_voxels, C;
    int vec4;     // C == 4 and 16-byte aligned rows: float4 loads
};

__device__ __forceinline__ bool axis_cell(const VoxParams &q, int j, float p, int &c)
{
    double cd;
    if (q.regime == 2) {
        cd = floor(((double)p - q.r[j]) / q.v[j]);
    } else if (q.regime == 1) {
        float d = __fsub_rn(p, q.rf[j]);
        cd = floor((double)d / q.v[j]);
    } else {
        float d = __fsub_rn(p, q.rf[j]);
        cd = (double)floorf(__fdiv_rn(d, q.vf[j]));
    }
    if (!(cd >= 0.0) || cd >= (double)q.g[j]) return false;   // also rejects NaN
    c = (int)cd;
    return true;
}

// key for the reflectance pre-order: ascending key == descending reflectance
__global__ void __launch_bounds__(VOX_THREADS) vox_refl_key_kernel(const float *__restrict__ points, int64_t n, int C,
                                                                    uint32_t *__restrict__ keys)
{
    int64_t i = (int64_t)blockIdx.x * VOX_THREADS + threadIdx.x;
    if (i < n) keys[i] = ~ordered_bits(points[i * C + 3]);
}

// K1: cell of every position + first arrival per cell.  Cell linearisation is (z*gy + y)*gx + x so
// that the pillar map can be consumed directly by the (D,H,W) canvas scatter.
__global__ void __launch_bounds__(VOX_THREADS)
vox_cell_kernel(const float *__restrict__ points, int64_t n, const VoxParams q, const int32_t *__restrict__ perm,
                int32_t *__restrict__ cell_of_pos, int32_t *__restrict__ first)
{
    int64_t p = (int64_t)blockIdx.x * VOX_THREADS + threadIdx.x;
    if (p >= n) return;
    int64_t idx = perm ? (int64_t)(uint32_t)perm[p] : p;
    float x, y, z;
    if (q.vec4) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(points) + idx);
        x = v.x; y = v.y; z = v.z;
    } else {
        const float *pt = points + idx * q.C;
        x = __ldg(pt); y = __ldg(pt + 1); z = __ldg(pt + 2);
    }
    int cx, cy, cz;
    int32_t cell = -1;
    if (axis_cell(q, 0, x, cx) && axis_cell(q, 1, y, cy) && axis_cell(q, 2, z, cz)) {
        cell = (cz * q.g[1] + cy) * q.g[0] + cx;
        atomicMin(first + cell, (int32_t)p);
    }
    cell_of_pos[p] = cell;
}

// K2: pillar ids by an ordered scan of the first-arrival flags (decoupled look-back, one launch).
__global__ void __launch_bounds__(VOX_THREADS)
vox_assign_kernel(const int32_t *__restrict__ cell_of_pos, int64_t n, const int32_t *__restrict__ first,
                  const VoxParams q, int32_t *__restrict__ pid_of_cell, int32_t *__restrict__ coors,
                  int32_t *__restrict__ cutoff, int32_t *__restrict__ voxel_num, uint32_t *status, uint32_t *ticket,
                  int num_tiles)
{
    __shared__ uint32_t s_tile, s_excl;
    __shared__ uint32_t warp_sum[VOX_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;

    int32_t cell[SCAN_ITEMS];
    uint32_t flags = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t p = base + k;
        cell[k] = (p < n) ? cell_of_pos[p] : -1;
        if (cell[k] >= 0 && first[cell[k]] == (int32_t)p) flags |= 1u << k;
    }
    uint32_t cnt = __popc(flags);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < VOX_THREADS / 32; ++w) {
        uint32_t s = warp_sum[w];
        if (w < warp) wbase += s;
        block_total += s;
    }
    if (tid == 0) {
        uint32_t excl = 0;
        if (tile == 0) {
            atomicExch(status, FLAG_PREFIX | block_total);
        } else {
            atomicExch(status + tile, FLAG_AGG | block_total);
            int64_t t = (int64_t)tile - 1;
            while (true) {
                uint32_t s = *((volatile uint32_t *)(status + t));
                if ((s & FLAG_MASK) == 0) continue;
                excl += s & VAL_MASK;
                if ((s & FLAG_MASK) == FLAG_PREFIX) break;
                --t;
            }
            atomicExch(status + tile, FLAG_PREFIX | (excl + block_total));
        }
        s_excl = excl;
        if ((int)tile == num_tiles - 1) {
            uint32_t total = excl + block_total;
            *voxel_num = (int32_t)(total < (uint32_t)q.max_voxels ? total : (uint32_t)q.max_voxels);
        }
    }
    __syncthreads();
    uint32_t vid = s_excl + wbase + incl - cnt;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (flags & (1u << k)) {
            if (vid < (uint32_t)q.max_voxels) {
                int c = cell[k];
                pid_of_cell[c] = (int32_t)vid;
                int cx = c % q.g[0];
                int t = c / q.g[0];
                coors[vid * 3 + 0] = cx;
                coors[vid * 3 + 1] = t % q.g[1];
                coors[vid * 3 + 2] = t / q.g[1];
            } else if (vid == (uint32_t)q.max_voxels) {
                *cutoff = (int32_t)(base + k);   // the reference breaks here
            }
            ++vid;
        }
    }
}

// K3: insert position p into the sorted row of the P smallest positions of its pillar.
// Slots only ever decrease, so "row[j-1] < p was observed" stays true forever and the insertion
// chain may start at j; every displaced value is pushed one slot down with atomicMin.  The final
// row is the sorted set of the P smallest positions whatever the interleaving.
__global__ void __launch_bounds__(VOX_THREADS)
vox_rank_kernel(const int32_t *__restrict__ cell_of_pos, int64_t n, const int32_t *__restrict__ pid_of_cell,
                const int32_t *__restrict__ cutoff, int P, int32_t *rows)
{
    int64_t p64 = (int64_t)blockIdx.x * VOX_THREADS + threadIdx.x;
    if (p64 >= n) return;
    const int32_t p = (int32_t)p64;
    if (p >= *cutoff) return;
    int32_t cell = cell_of_pos[p];
    if (cell < 0) return;
    int32_t *row = rows + (int64_t)pid_of_cell[cell] * P;
    if (ld_cg(row + P - 1) < p) return;           // already P smaller positions: dropped (:303)
    int lo = 0, hi = P - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ld_cg(row + mid) < p) lo = mid + 1; else hi = mid;
    }
    int32_t carry = p;
    for (int k = lo; k < P; ++k) {
        int32_t old = atomicMin(row + k, carry);
        if (old == PP_INF_POS) break;
        carry = old > carry ? old : carry;
    }
}

// K4: gather the kept points into (M, P, C), zero padded, and count them.
template <bool VEC4>
__global__ void __launch_bounds__(VOX_THREADS)
vox_gather_kernel(const float *__restrict__ points, const int32_t *__restrict__ perm, const int32_t *__restrict__ rows,
                  const int32_t *__restrict__ voxel_num, int64_t max_rows, int P, int C, float *__restrict__ voxels,
                  int32_t *__restrict__ num_points)
{
    int64_t t = (int64_t)blockIdx.x * VOX_THREADS + threadIdx.x;   // one thread per (pillar, slot)
    if (t >= max_rows * P) return;
    int64_t m = t / P;
    int s = (int)(t - m * P);
    if (m >= *voxel_num) return;
    int32_t pos = rows[t];
    bool valid = pos != PP_INF_POS;
    if (valid && (s == P - 1 || rows[t + 1] == PP_INF_POS)) num_points[m] = s + 1;
    int64_t idx = 0;
    if (valid) idx = perm ? (int64_t)(uint32_t)perm[pos] : (int64_t)pos;
    if (VEC4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) v = __ldg(reinterpret_cast<const float4 *>(points) + idx);
        reinterpret_cast<float4 *>(voxels)[t] = v;
    } else {
        for (int c = 0; c < C; ++c) voxels[t * C + c] = valid ? __ldg(points + idx * C + c) : 0.f;
    }
}

struct VoxWs {
    // 0x7F-filled
    int32_t *first, *rows, *cutoff;
    size_t fill7f_bytes;
    // zero-filled
    uint32_t *status, *ticket;
    size_t zero_off, zero_bytes;
    // uninitialised
    int32_t *cell_of_pos, *pid_of_cell;
    uint32_t *keys, *keys_sorted, *perm;
    void *sort_ws;
    size_t sort_ws_bytes;
    int num_tiles;
    int64_t max_rows, cells;
};

int64_t max_rows_of(int64_t n, const pp_voxel_cfg *c)
{
    int64_t cells = (int64_t)c->grid[0] * c->grid[1] * c->grid[2];
    int64_t r = c->max_voxels;
    if (n < r) r = n;
    if (cells < r) r = cells;
    return r > 0 ? r : 1;
}

VoxWs carve(void *ws, int64_t n, const pp_voxel_cfg *c, int order, bool need_pid, size_t *total)
{
    VoxWs w;
    int64_t n1 = n > 0 ? n : 1;
    w.cells = (int64_t)c->grid[0] * c->grid[1] * c->grid[2];
    w.max_rows = max_rows_of(n, c);
    w.num_tiles = (int)ceil_div(n1, SCAN_TILE);
    Arena a(ws, (size_t)-1);
    w.first = a.take<int32_t>((size_t)w.cells);
    w.rows = a.take<int32_t>((size_t)w.max_rows * c->max_points);
    w.cutoff = a.take<int32_t>(64);
    w.fill7f_bytes = a.off;
    w.zero_off = align_up(a.off);
    w.status = a.take<uint32_t>((size_t)w.num_tiles);
    w.ticket = a.take<uint32_t>(64);
    w.zero_bytes = a.off - w.zero_off;
    w.cell_of_pos = a.take<int32_t>((size_t)n1);
    w.pid_of_cell = need_pid ? a.take<int32_t>((size_t)w.cells) : nullptr;
    w.keys = w.keys_sorted = w.perm = nullptr;
    w.sort_ws = nullptr;
    w.sort_ws_bytes = 0;
    if (order == PP_ORDER_REFLECTANCE_DESC) {
        w.keys = a.take<uint32_t>((size_t)n1);
        w.keys_sorted = a.take<uint32_t>((size_t)n1);
        w.perm = a.take<uint32_t>((size_t)n1);
        w.sort_ws_bytes = sort_workspace_bytes(n1);
        w.sort_ws = a.take<char>(w.sort_ws_bytes);
    }
    *total = align_up(a.off);
    return w;
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
    carve(nullptr, n_points, cfg, order, true, &total);
    return total;
}

extern "C" int pp_voxelize(const float *points, int64_t n, const pp_voxel_cfg *cfg, int order, const int32_t *perm,
                           float *voxels, int32_t *coors, int32_t *num_points, int32_t *voxel_num,
                           int32_t *pillar_map, void *workspace, size_t workspace_bytes, pp_stream_t stream)
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
    int64_t cells = (int64_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
    PP_REQUIRE(cells < (1ll << 31), "grid too large (>= 2^31 cells)");
    if (n == 0 || cfg->max_voxels == 0) {
        PP_CUDA_TRY(cudaMemsetAsync(voxel_num, 0, sizeof(int32_t), st));
        if (pillar_map) PP_CUDA_TRY(cudaMemsetAsync(pillar_map, 0xFF, (size_t)cells * 4, st));
        return PP_OK;
    }
    PP_REQUIRE(points && voxels && coors && num_points && workspace, "null pointer");

    size_t total;
    VoxWs w = carve(workspace, n, cfg, order, pillar_map == nullptr, &total);
    if (workspace_bytes < total) {
        set_error("voxelize workspace too small: %zu < %zu", workspace_bytes, total);
        return PP_ERR_WORKSPACE;
    }
    int32_t *pid_of_cell = pillar_map ? pillar_map : w.pid_of_cell;

    VoxParams q;
    for (int j = 0; j < 3; ++j) {
        q.r[j] = cfg->range[j];
        q.v[j] = cfg->vsize[j];
        q.rf[j] = (float)cfg->range[j];
        q.vf[j] = (float)cfg->vsize[j];
        q.g[j] = cfg->grid[j];
    }
    q.regime = cfg->range_is_f64 ? 2 : (cfg->vsize_is_f64 ? 1 : 0);
    q.P = cfg->max_points;
    q.max_voxels = cfg->max_voxels;
    q.C = cfg->num_feats;
    q.vec4 = (q.C == 4 && ((uintptr_t)points % 16 == 0)) ? 1 : 0;

    PP_CUDA_TRY(cudaMemsetAsync(workspace, 0x7F, w.fill7f_bytes, st));
    PP_CUDA_TRY(cudaMemsetAsync((char *)workspace + w.zero_off, 0, w.zero_bytes, st));
    if (pillar_map) PP_CUDA_TRY(cudaMemsetAsync(pillar_map, 0xFF, (size_t)cells * 4, st));
    prof_mark("memset");

    const int32_t *order_perm = nullptr;
    if (order == PP_ORDER_PERM) {
        order_perm = perm;
    } else if (order == PP_ORDER_REFLECTANCE_DESC) {
        vox_refl_key_kernel<<<(unsigned)ceil_div(n, VOX_THREADS), VOX_THREADS, 0, st>>>(points, n, q.C, w.keys);
        if (int rc = check_launch("vox_refl_key_kernel")) return rc;
        if (int rc = sort_pairs_u32(w.keys, nullptr, w.keys_sorted, w.perm, n, w.sort_ws, w.sort_ws_bytes, st)) return rc;
        order_perm = (const int32_t *)w.perm;
    }

    const unsigned nb = (unsigned)ceil_div(n, VOX_THREADS);
    vox_cell_kernel<<<nb, VOX_THREADS, 0, st>>>(points, n, q, order_perm, w.cell_of_pos, w.first);
    if (int rc = check_launch("vox_cell_kernel")) return rc;
    vox_assign_kernel<<<w.num_tiles, VOX_THREADS, 0, st>>>(w.cell_of_pos, n, w.first, q, pid_of_cell, coors, w.cutoff,
                                                          voxel_num, w.status, w.ticket, w.num_tiles);
    if (int rc = check_launch("vox_assign_kernel")) return rc;
    vox_rank_kernel<<<nb, VOX_THREADS, 0, st>>>(w.cell_of_pos, n, pid_of_cell, w.cutoff, q.P, w.rows);
    if (int rc = check_launch("vox_rank_kernel")) return rc;
    const int64_t slots = w.max_rows * q.P;
    const bool vec4 = q.vec4 && ((uintptr_t)voxels % 16 == 0);
    if (vec4)
        vox_gather_kernel<true><<<(unsigned)ceil_div(slots, VOX_THREADS), VOX_THREADS, 0, st>>>(
            points, order_perm, w.rows, voxel_num, w.max_rows, q.P, q.C, voxels, num_points);
    else
        vox_gather_kernel<false><<<(unsigned)ceil_div(slots, VOX_THREADS), VOX_THREADS, 0, st>>>(
            points, order_perm, w.rows, voxel_num, w.max_rows, q.P, q.C, voxels, num_points);
    return check_launch("vox_gather_kernel");
}
