// Micro-benchmark (measurement only): issue rate of FP64 arithmetic against FP32 on this GPU -- DFMA, the
// compare+select pair an fp64 min/max compiles to, and FFMA -- with 8 independent chains per thread.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_bench fp64_bench.cu ; run: ./fp64_bench
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int MODE>   // 0 fma, 1 fma + min/max
__global__ void k(T *out, T a, T b, int iters)
{
    T x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = x[i] * a + b;
            if (MODE == 1) x[i] = x[i] < (T)1e6 ? x[i] : (T)3;
        }
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T, int MODE> void run(const char *name)
{
    T *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(T));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<T, MODE><<<148 * 8, 256>>>(out, (T)1.0000001, (T)0.5, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double ops = 148.0 * 8 * 256 * 8 * iters;
    printf("%-28s %8.3f ms  %8.2f T inner-steps/s\n", name, best, ops / (best * 1e-3) * 1e-12);
    cudaFree(out);
}

int main()
{
    run<float, 0>("f32 fma");
    run<double, 0>("f64 fma");
    run<float, 1>("f32 fma + compare/select");
    run<double, 1>("f64 fma + compare/select");
    return 0;
}
