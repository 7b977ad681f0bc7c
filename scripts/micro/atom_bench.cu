// Micro-benchmark (measurement only): what 1e6 random L2 atomics / reductions / loads cost on B200, for the access
// pattern of the voxelizer's per-point passes (11.5k hot cells x 8 counters inside a 214k x 8 table, geometric chunk skew).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atom_bench atom_bench.cu ; run: ./atom_bench
#include <cstdio>
#include <cstdint>
#include <vector>
#include <random>
#include <cuda_runtime.h>

template <int MODE, int IT>   // 0 RED, 1 ATOM + store ticket, 2 load + store, 3 ATOM no store dependency (result summed)
__global__ void k(const uint32_t *__restrict__ idx, int n, uint32_t *tab, uint32_t *out)
{
    const int p0 = blockIdx.x * (blockDim.x * IT) + threadIdx.x;
    uint32_t a[IT], r[IT];
#pragma unroll
    for (int i = 0; i < IT; ++i) { const int p = p0 + i * blockDim.x; a[i] = p < n ? idx[p] : 0xFFFFFFFFu; }
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        r[i] = 0;
        if (a[i] == 0xFFFFFFFFu) continue;
        if (MODE == 0) atomicAdd(tab + a[i], 1u);
        else if (MODE == 1 || MODE == 3) r[i] = atomicAdd(tab + a[i], 1u);
        else r[i] = __ldcg(tab + a[i]);
    }
    if (MODE == 0) return;
#pragma unroll
    for (int i = 0; i < IT; ++i) { const int p = p0 + i * blockDim.x; if (p < n) out[p] = r[i]; }
}

template <int MODE, int IT> float run(const uint32_t *idx, int n, uint32_t *tab, size_t tab_bytes, uint32_t *out, int threads)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = (n + threads * IT - 1) / (threads * IT);
    float best = 1e9f;
    for (int rep = 0; rep < 12; ++rep) {
        cudaMemsetAsync(tab, 0, tab_bytes);
        cudaEventRecord(e0);
        k<MODE, IT><<<blocks, threads>>>(idx, n, tab, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    return best * 1e3f;
}

int main()
{
    const int n = 1000000, T = 214272, NC = 11500;
    std::mt19937 rng(1);
    std::vector<uint32_t> cells(NC);
    for (auto &c : cells) c = rng() % T;
    std::lognormal_distribution<double> ln(0.0, 1.0);
    std::vector<double> w(NC);
    for (auto &x : w) x = ln(rng);
    std::discrete_distribution<int> pick(w.begin(), w.end());
    std::uniform_real_distribution<double> u(0, 1);
    std::vector<uint32_t> h(n), hu(n);
    for (int i = 0; i < n; ++i) {
        const double q = u(rng);
        int chunk = 0;
        for (int c = 1; c < 8; ++c) if (q >= 1.0 / (1 << (8 - c))) chunk = c;
        h[i] = cells[pick(rng)] * 8 + chunk;
        hu[i] = (rng() % T) * 8 + (rng() % 8);
    }
    uint32_t *idx, *tab, *out;
    const size_t tb = (size_t)T * 8 * 4;
    cudaMalloc(&idx, n * 4); cudaMalloc(&tab, tb); cudaMalloc(&out, n * 4);
    for (int pat = 0; pat < 2; ++pat) {
        cudaMemcpy(idx, pat ? hu.data() : h.data(), n * 4, cudaMemcpyHostToDevice);
        printf("pattern %s\n", pat ? "uniform over the table" : "D1M-like (11.5k hot cells, geometric chunks)");
        printf("  RED   it1 %6.2f us  it4 %6.2f us  it4/128thr %6.2f us\n", run<0, 1>(idx, n, tab, tb, out, 256), run<0, 4>(idx, n, tab, tb, out, 256), run<0, 4>(idx, n, tab, tb, out, 128));
        printf("  ATOM  it1 %6.2f us  it4 %6.2f us  it4/128thr %6.2f us\n", run<1, 1>(idx, n, tab, tb, out, 256), run<1, 4>(idx, n, tab, tb, out, 256), run<1, 4>(idx, n, tab, tb, out, 128));
        printf("  LOAD  it1 %6.2f us  it4 %6.2f us  it8 %6.2f us\n", run<2, 1>(idx, n, tab, tb, out, 256), run<2, 4>(idx, n, tab, tb, out, 256), run<2, 8>(idx, n, tab, tb, out, 256));
    }
    // launch overhead reference: an empty kernel between two events
    printf("  (event-to-event time of an empty 1-CTA kernel: ");
    { cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e9f;
      for (int r = 0; r < 10; ++r) { cudaEventRecord(e0); k<2, 1><<<1, 32>>>(idx, 0, tab, out); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
      printf("%.2f us)\n", best * 1e3f); }
    return 0;
}
