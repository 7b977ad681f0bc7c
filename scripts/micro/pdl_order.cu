// Micro-test (measurement only): what waits for what around a kernel that triggers its dependents early
// (griddepcontrol.launch_dependents at its top).  A spins ~100 us after the trigger, then writes 1.
//   case 1: A, then cudaMemsetAsync(buf, 0): buf must end 0 if the memset waits for A's completion
//   case 2: A, then a plain kernel that writes 2: must end 2
//   case 3: A, then a kernel launched programmatically WITHOUT griddepcontrol.wait (expected to race: ends 1)
//   case 4: memset(0) of a counter, then a programmatic kernel that adds 1 after griddepcontrol.wait: must end 1
//   case 5: event timing around A alone: must be >= the spin
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pdl_order pdl_order.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void spin_then_write(int *buf, int value, long long spin_ns, bool trigger, bool wait)
{
    if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < spin_ns);
    buf[0] = value;
}
__global__ void add_one(int *buf, bool wait)
{
    if (wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    atomicAdd(buf, 1);
}
template <typename... A> void launch(bool pdl, void (*k)(A...), cudaStream_t st, A... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(1); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k, args...);
}
int main()
{
    int *buf; cudaMalloc(&buf, 256);
    cudaStream_t st; cudaStreamCreate(&st);
    int h;
    const long long spin = 100000;
    for (int pdl_a = 0; pdl_a < 2; ++pdl_a) {
        printf("--- A launched %s\n", pdl_a ? "programmatically" : "plainly");
        cudaMemsetAsync(buf, 0, 4, st); cudaStreamSynchronize(st);
        launch(pdl_a, spin_then_write, st, buf, 1, spin, true, true);
        cudaMemsetAsync(buf, 0, 4, st);
        cudaMemcpyAsync(&h, buf, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        printf("case 1 (memset after A): %d (0 = waited)\n", h);
        launch(pdl_a, spin_then_write, st, buf, 1, spin, true, true);
        launch(false, spin_then_write, st, buf, 2, 0ll, false, false);
        cudaMemcpyAsync(&h, buf, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        printf("case 2 (plain kernel after A): %d (2 = waited)\n", h);
        launch(pdl_a, spin_then_write, st, buf, 1, spin, true, true);
        launch(true, spin_then_write, st, buf, 2, 0ll, false, false);
        cudaMemcpyAsync(&h, buf, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        printf("case 3 (programmatic kernel, no wait, after A): %d (1 = ran ahead, expected)\n", h);
        launch(pdl_a, spin_then_write, st, buf, 1, spin, true, true);
        launch(true, spin_then_write, st, buf, 2, 0ll, false, true);
        cudaMemcpyAsync(&h, buf, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        printf("case 3b (programmatic kernel with wait, after A): %d (2 = waited)\n", h);
    }
    cudaMemsetAsync(buf, 0xFF, 4, st); cudaStreamSynchronize(st);
    launch(false, spin_then_write, st, buf + 8, 1, spin, true, false);     // keeps the stream busy
    cudaMemsetAsync(buf, 0, 4, st);
    launch(true, add_one, st, buf, true);
    cudaMemcpyAsync(&h, buf, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
    printf("case 4 (memset then programmatic add): %d (1 = the add waited for the memset)\n", h);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    launch(true, spin_then_write, st, buf, 1, spin, true, true);
    cudaEventRecord(e1, st); cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("case 5 (events around A): %.1f us (>= %.0f)\n", ms * 1e3, spin * 1e-3);
    // A -> B -> memset, B programmatic and short: does the memset wait for A through B?
    cudaMemsetAsync(buf, 0, 8, st); cudaStreamSynchronize(st);
    launch(false, spin_then_write, st, buf, 1, spin, true, false);
    launch(true, spin_then_write, st, buf + 1, 1, 0ll, true, true);
    cudaMemsetAsync(buf, 0, 8, st);
    int h2[2];
    cudaMemcpyAsync(h2, buf, 8, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
    printf("case 6 (A, programmatic B, memset): %d %d (0 0 = waited)\n", h2[0], h2[1]);
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
