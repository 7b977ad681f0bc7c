// Stable LSD radix sort of (u32 key, u32 value) pairs: one histogram pass + 4 single-sweep digit
// passes with decoupled look-back ("onesweep" structure, hand-written; no CUB).
//
// Used by the head's top-k (model/PointPillars.py:1056-1065) and by NMS score ordering (model/utils.py:398).
// Stability is what defines the documented tie rule (lower original index first).  HBM-bound at large n:
// 4 x (read keys+vals, write keys+vals); latency-bound at the NMS sizes, where the caller can zero the counters with
// a memset of its own, supply the digit histograms from the kernel that produced the keys, and have the last pass
// move up to four float4 payload arrays with the keys (pp_sort.cuh): three launches and a memset less per NMS call.
// <= 4096 keys: one CTA, one launch, everything in shared memory.
#include "pp_common.cuh"
#include "pp_sort.cuh"

namespace pp {

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int NPASS = 4;

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VAL_MASK = ~FLAG_MASK;

// All four digit histograms in one pass over the keys.
__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(const uint32_t *__restrict__ keys, int64_t n,
                                                                  uint32_t *__restrict__ hist /* [4][256] */)
{
    __shared__ uint32_t sh[NPASS * RADIX];
    for (int i = threadIdx.x; i < NPASS * RADIX; i += SORT_THREADS) sh[i] = 0;
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * SORT_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * SORT_THREADS + threadIdx.x; i < n; i += stride) {
        uint32_t k = keys[i];
#pragma unroll
        for (int p = 0; p < NPASS; ++p) atomicAdd(&sh[p * RADIX + ((k >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NPASS * RADIX; i += SORT_THREADS)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

template <int ITEMS, bool TAIL>
__global__ void __launch_bounds__(SORT_THREADS)
sort_pass_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                 uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, int pass,
                 const uint32_t *__restrict__ hist /* this pass: [256] */, uint32_t *status /* [tiles][256] */,
                 uint32_t *ticket, bool direct, const SortTail tail)
{
    constexpr int TILE = SORT_THREADS * ITEMS;
    __shared__ uint32_t warp_hist[SORT_WARPS][RADIX + 1];
    __shared__ uint32_t digit_off[RADIX];   // global offset of this tile's first key of each digit
    __shared__ uint32_t scan_tmp[SORT_WARPS];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < SORT_WARPS * (RADIX + 1); i += SORT_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t tile_base = (int64_t)tile * TILE;
    const int64_t warp_base = tile_base + (int64_t)warp * (32 * ITEMS);
    const int shift = pass * RADIX_BITS;

    uint32_t key[ITEMS];
    uint32_t rank[ITEMS];   // rank of the key among equal digits inside its warp (stable)
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        int64_t idx = warp_base + r * 32 + lane;
        bool valid = idx < n;
        key[r] = valid ? keys_in[idx] : 0xFFFFFFFFu;
        uint32_t d = valid ? ((key[r] >> shift) & (RADIX - 1)) : (uint32_t)RADIX;
        unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
        uint32_t pre = warp_hist[warp][d];
        __syncwarp();
        if ((peers & lanemask_lt()) == 0) warp_hist[warp][d] = pre + __popc(peers);
        __syncwarp();
        rank[r] = pre + __popc(peers & lanemask_lt());
    }
    __syncthreads();

    // thread d owns digit d: exclusive scan over the warps, then decoupled look-back over tiles
    {
        const int d = tid;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            uint32_t c = warp_hist[w][d];
            warp_hist[w][d] = run;
            run += c;
        }
        const uint32_t tile_count = run;
        uint32_t *my = status + (size_t)tile * RADIX + d;
        uint32_t excl = 0;
        if (tile == 0) {
            atomicExch(my, FLAG_PREFIX | tile_count);
        } else if (direct) {
            // few tiles (the NMS sizes): every tile publishes its own count at once and sums the counts of ALL its
            // predecessors with independent loads -- one L2 round trip instead of a chain of up to `tile` dependent ones
            atomicExch(my, FLAG_AGG | tile_count);
            for (int t0 = 0; t0 < (int)tile; t0 += 8) {
                uint32_t sv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    sv[j] = 1u << 30;        // "published, count 0" for slots past the last predecessor
                    if (t0 + j < (int)tile) sv[j] = *((volatile uint32_t *)(status + (size_t)(t0 + j) * RADIX + d));
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    while ((sv[j] & FLAG_MASK) == 0) sv[j] = *((volatile uint32_t *)(status + (size_t)(t0 + j) * RADIX + d));
                    excl += sv[j] & VAL_MASK;
                }
            }
        } else {
            atomicExch(my, FLAG_AGG | tile_count);
            int64_t t = (int64_t)tile - 1;
            while (true) {
                uint32_t s = *((volatile uint32_t *)(status + (size_t)t * RADIX + d));
                if ((s & FLAG_MASK) == 0) continue;   // predecessor not published yet
                excl += s & VAL_MASK;
                if ((s & FLAG_MASK) == FLAG_PREFIX) break;
                --t;
            }
            atomicExch(my, FLAG_PREFIX | (excl + tile_count));
        }
        // exclusive scan of the global digit histogram (256 values, one per thread)
        uint32_t h = hist[d];
        uint32_t incl = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) scan_tmp[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += scan_tmp[w];
        digit_off[d] = wbase + incl - h + excl;
    }
    __syncthreads();

#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        int64_t idx = warp_base + r * 32 + lane;
        if (idx < n) {
            uint32_t d = (key[r] >> shift) & (RADIX - 1);
            uint32_t dst = digit_off[d] + warp_hist[warp][d] + rank[r];
            keys_out[dst] = key[r];
            const uint32_t val = vals_in ? vals_in[idx] : (uint32_t)idx;
            vals_out[dst] = val;
            if (TAIL) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if (a < tail.arrays) tail.out[a][dst] = tail.in[a][val];
            }
        }
    }
    if (TAIL && tile == 0 && tid == 0 && tail.count_out) {
        const int32_t c = *tail.count_in;
        *tail.count_out = c < tail.count_cap ? c : tail.count_cap;
    }
}

// n <= SMALL_TILE: the whole array is one tile of one CTA, so there is no look-back chain and no separate histogram
// pass; the four digit passes run inside ONE launch and the data never leaves shared memory between them (keys as
// u32, values as u16 positions): global memory is read once and written once, both coalesced.  A single SM cannot
// scatter 4-byte stores to L2 fast enough (one transaction per clock), shared memory can.
// The reference's nms_pre = 500 candidates (config.yaml:61): one launch instead of a memset + 5 launches.
constexpr int SMALL_THREADS = 512;
constexpr int SMALL_WARPS = SMALL_THREADS / 32;
constexpr int SMALL_ITEMS = 8;
constexpr int SMALL_TILE = SMALL_THREADS * SMALL_ITEMS;      // 4096; beyond, one SM ranks more slowly than the multi-CTA passes
constexpr int SMALL_MAX = SMALL_TILE;                        // (measured: 20k keys take 89 us in one CTA, 60 us in the passes)
constexpr size_t SMALL_SMEM = (size_t)SMALL_TILE * 4 + (size_t)SMALL_TILE * 2 + (size_t)SMALL_WARPS * (RADIX + 1) * 4;

__global__ void __launch_bounds__(SMALL_THREADS)
sort_small_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                  uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int n)
{
    extern __shared__ __align__(16) unsigned char ss_smem[];
    uint32_t *s_key = reinterpret_cast<uint32_t *>(ss_smem);                                   // [SMALL_TILE]
    uint16_t *s_pos = reinterpret_cast<uint16_t *>(ss_smem + (size_t)SMALL_TILE * 4);           // [SMALL_TILE]
    uint32_t(*warp_hist)[RADIX + 1] = reinterpret_cast<uint32_t(*)[RADIX + 1]>(ss_smem + (size_t)SMALL_TILE * 6);
    __shared__ uint32_t digit_off[RADIX];
    __shared__ uint32_t scan_tmp[RADIX / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rounds = (n + SMALL_THREADS - 1) / SMALL_THREADS;      // keys per thread; a warp owns 32 * rounds consecutive keys
    const int warp_base = warp * (32 * rounds);
    for (int pass = 0; pass < NPASS; ++pass) {
        const int shift = pass * RADIX_BITS;
        for (int i = tid; i < SMALL_WARPS * (RADIX + 1); i += SMALL_THREADS) (&warp_hist[0][0])[i] = 0;
        uint32_t key[SMALL_ITEMS], pos2[SMALL_ITEMS / 2], rank2[SMALL_ITEMS / 2];
#pragma unroll
        for (int r = 0; r < SMALL_ITEMS; ++r) {
            if (r >= rounds) break;
            const int idx = warp_base + r * 32 + lane;
            uint32_t k = 0xFFFFFFFFu, p = (uint32_t)idx;
            if (idx < n) {
                if (pass == 0) k = keys_in[idx];
                else { k = s_key[idx]; p = s_pos[idx]; }
            }
            key[r] = k;
            if (r & 1) pos2[r >> 1] |= p << 16; else pos2[r >> 1] = p;
        }
        __syncthreads();      // every thread holds its slice: the shared arrays may be overwritten below
#pragma unroll
        for (int r = 0; r < SMALL_ITEMS; ++r) {
            if (r >= rounds) break;
            const int idx = warp_base + r * 32 + lane;
            const uint32_t d = idx < n ? ((key[r] >> shift) & (RADIX - 1)) : (uint32_t)RADIX;
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
            const uint32_t pre = warp_hist[warp][d];
            __syncwarp();
            if ((peers & lanemask_lt()) == 0) warp_hist[warp][d] = pre + __popc(peers);
            __syncwarp();
            const uint32_t rk = pre + __popc(peers & lanemask_lt());
            if (r & 1) rank2[r >> 1] |= rk << 16; else rank2[r >> 1] = rk;
        }
        __syncthreads();
        if (tid < RADIX) {
            // thread d owns digit d: exclusive scan over the warps, then over the digits
            uint32_t run = 0;
#pragma unroll 8
            for (int w = 0; w < SMALL_WARPS; ++w) {
                const uint32_t c = warp_hist[w][tid];
                warp_hist[w][tid] = run;
                run += c;
            }
            uint32_t incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) scan_tmp[warp] = incl;
            digit_off[tid] = incl - run;
        }
        __syncthreads();
        if (tid < RADIX) {
            uint32_t wbase = 0;
            for (int w = 0; w < warp; ++w) wbase += scan_tmp[w];
            digit_off[tid] += wbase;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SMALL_ITEMS; ++r) {
            if (r >= rounds) break;
            const int idx = warp_base + r * 32 + lane;
            if (idx < n) {
                const uint32_t d = (key[r] >> shift) & (RADIX - 1);
                const uint32_t rk = (r & 1) ? (rank2[r >> 1] >> 16) : (rank2[r >> 1] & 0xFFFFu);
                const uint32_t dst = digit_off[d] + warp_hist[warp][d] + rk;
                s_key[dst] = key[r];
                s_pos[dst] = (uint16_t)((r & 1) ? (pos2[r >> 1] >> 16) : (pos2[r >> 1] & 0xFFFFu));
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += SMALL_THREADS) {
        const uint32_t p = s_pos[i];
        keys_out[i] = s_key[i];
        vals_out[i] = vals_in ? vals_in[p] : p;
    }
}

struct SortWs {
    uint32_t *hist, *tickets, *status;
    uint32_t *keys_tmp, *vals_tmp;
    size_t zero_bytes;
    int tiles;
};

// Keys per thread of a digit pass.  Small inputs (the NMS and top-k sizes) are latency-bound -- 20 k keys are 10 tiles
// of 2048, each thread ranking 8 keys one after the other -- so they get 4x as many tiles of a quarter the size; the
// look-back of <= SORT_DIRECT_TILES tiles is one round of independent loads either way.
constexpr int SORT_ITEMS = 8, SORT_ITEMS_SMALL = 2;
constexpr int64_t SORT_SMALL_N = 32768;
constexpr int SORT_DIRECT_TILES = 64;
inline int sort_items(int64_t n) { return n <= SORT_SMALL_N ? SORT_ITEMS_SMALL : SORT_ITEMS; }

SortWs carve(void *ws, int64_t n, size_t *total)
{
    SortWs s;
    s.tiles = (int)ceil_div(n > 0 ? n : 1, (int64_t)SORT_THREADS * sort_items(n));
    Arena a(ws, (size_t)-1);
    s.hist = a.take<uint32_t>(NPASS * RADIX);
    s.tickets = a.take<uint32_t>(64);
    s.status = a.take<uint32_t>((size_t)NPASS * s.tiles * RADIX);
    s.zero_bytes = a.off;
    s.keys_tmp = a.take<uint32_t>((size_t)(n > 0 ? n : 1));
    s.vals_tmp = a.take<uint32_t>((size_t)(n > 0 ? n : 1));
    *total = align_up(a.off);
    return s;
}

}  // namespace

size_t sort_workspace_bytes(int64_t n)
{
    size_t total;
    carve(nullptr, n, &total);
    return total;
}

bool sort_runs_tail(int64_t n) { return n > SMALL_MAX; }

uint32_t *sort_hist(void *ws, int64_t n)
{
    if (n <= SMALL_MAX) return nullptr;
    size_t total;
    return carve(ws, n, &total).hist;
}

size_t sort_zero_bytes(int64_t n)
{
    size_t total;
    return carve(nullptr, n, &total).zero_bytes;
}

int sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                   int64_t n, void *ws, size_t ws_bytes, cudaStream_t st, bool ws_zeroed, bool have_hist,
                   const SortTail *tail)
{
    if (n <= 0) return PP_OK;
    PP_REQUIRE(n < (1ll << 30), "n must be < 2^30");
    size_t total;
    SortWs s = carve(ws, n, &total);
    if (ws_bytes < total) {
        set_error("sort workspace too small: %zu < %zu", ws_bytes, total);
        return PP_ERR_WORKSPACE;
    }
    if (n <= SMALL_MAX) {
        // per device and context, cheap: set on every call (a process may drive several GPUs, from several threads)
        PP_CUDA_TRY(cudaFuncSetAttribute(sort_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM));
        sort_small_kernel<<<1, SMALL_THREADS, SMALL_SMEM, st>>>(keys_in, vals_in, keys_out, vals_out, (int)n);
        return check_launch("sort_small_kernel");
    }
    if (!ws_zeroed) {
        PP_CUDA_TRY(cudaMemsetAsync(ws, 0, s.zero_bytes, st));
        prof_mark("memset");
    }
    if (!(have_hist && ws_zeroed)) {
        const int64_t per_block = n <= SORT_SMALL_N ? SORT_THREADS * 4 : SORT_THREADS * 16;
        int hist_blocks = (int)(ceil_div(n, per_block) < 148 * 4 ? ceil_div(n, per_block) : 148 * 4);
        sort_hist_kernel<<<hist_blocks, SORT_THREADS, 0, st>>>(keys_in, n, s.hist);
        if (int rc = check_launch("sort_hist_kernel")) return rc;
    }
    const uint32_t *kin = keys_in, *vin = vals_in;
    for (int p = 0; p < NPASS; ++p) {
        uint32_t *kout = (p & 1) ? keys_out : s.keys_tmp;
        uint32_t *vout = (p & 1) ? vals_out : s.vals_tmp;
        const bool with_tail = tail && p == NPASS - 1;
        const SortTail t = with_tail ? *tail : SortTail();
#define PP_PASS(IT, TL)                                                                                               \
    sort_pass_kernel<IT, TL><<<s.tiles, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, p, s.hist + p * RADIX,        \
                                                               s.status + (size_t)p * s.tiles * RADIX, s.tickets + p, \
                                                               s.tiles <= SORT_DIRECT_TILES, t)
        if (sort_items(n) == SORT_ITEMS_SMALL) { if (with_tail) PP_PASS(SORT_ITEMS_SMALL, true); else PP_PASS(SORT_ITEMS_SMALL, false); }
        else { if (with_tail) PP_PASS(SORT_ITEMS, true); else PP_PASS(SORT_ITEMS, false); }
#undef PP_PASS
        if (int rc = check_launch("sort_pass_kernel")) return rc;
        kin = kout;
        vin = vout;
    }
    return PP_OK;
}

}  // namespace pp

extern "C" size_t pp_sort_workspace_bytes(int64_t n) { return pp::sort_workspace_bytes(n); }

extern "C" int pp_sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out,
                                 uint32_t *vals_out, int64_t n, void *workspace, size_t workspace_bytes,
                                 pp_stream_t stream)
{
    pp::enter((cudaStream_t)stream);
    PP_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return PP_OK;
    PP_REQUIRE(keys_in && keys_out && vals_out && workspace, "null pointer");
    return pp::sort_pairs_u32(keys_in, vals_in, keys_out, vals_out, n, workspace, workspace_bytes,
                              (cudaStream_t)stream);
}
