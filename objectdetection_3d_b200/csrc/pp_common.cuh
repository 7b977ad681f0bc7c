// Shared helpers for libpp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pp_b200.h"

#define PP_INF_POS 0x7F7F7F7F  // "empty" marker that cudaMemsetAsync(0x7F) produces for int32

namespace pp {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
void prof_mark(const char *name);   // no-op unless pp_profile_enable(1)
void enter(cudaStream_t st);        // call first in every launching entry point

inline int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return PP_ERR_CUDA;
    }
    count_launch();
    prof_mark(what);
    return PP_OK;
}

#define PP_CUDA_TRY(expr)                                                        \
    do {                                                                         \
        cudaError_t e__ = (expr);                                                \
        if (e__ != cudaSuccess) {                                                \
            pp::set_error("%s: %s", #expr, cudaGetErrorString(e__));             \
            return PP_ERR_CUDA;                                                  \
        }                                                                        \
    } while (0)

#define PP_REQUIRE(cond, msg)                                                    \
    do {                                                                         \
        if (!(cond)) {                                                           \
            pp::set_error("%s: %s", __func__, msg);                              \
            return PP_ERR_INVALID;                                               \
        }                                                                        \
    } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct Arena {
    char *base;
    size_t off, cap;
    Arena(void *p, size_t bytes) : base((char *)p), off(0), cap(bytes) {}
    template <typename T> T *take(size_t n)
    {
        size_t o = align_up(off);
        off = o + n * sizeof(T);
        return (T *)(base + o);
    }
    bool ok() const { return off <= cap; }
};

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// float -> uint32 whose unsigned order equals the float order (finite values; -0 < +0)
__device__ __forceinline__ uint32_t ordered_bits(float f)
{
    uint32_t u = __float_as_uint(f);
    return u ^ ((u & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u);
}

// L2-coherent loads (bypass the non-coherent L1) for data other CTAs are updating
__device__ __forceinline__ int ld_cg(const int *p) { return __ldcg(p); }

// Programmatic dependent launch: a kernel launched with launch_pdl may be scheduled while its predecessor on
// the stream is still draining; it must execute pdl_enter() before it touches anything the predecessor wrote.
// pdl_enter() also lets the NEXT kernel start being scheduled (safe: a dependent grid is only launched once
// every CTA of this grid has started).  This hides most of the ~3 us launch gap between the small kernels.
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// The same, for kernels that read what the predecessor wrote through `const T *__restrict__` parameters: such loads are
// ld.global.nc, which the compiler may move ABOVE the wait (observed on the NMS kernels, whose first load -- the
// candidate count -- was hoisted over both instructions: the "memory" clobber does not order loads of data the
// compiler is told nobody writes; scripts/pdl_audit.py lists what precedes the wait in the built objects).  Passing the pointers
// through an empty asm after the wait makes every load through them depend on it.
template <typename P> __device__ __forceinline__ void pdl_launder(P &p) { asm volatile("" : "+l"(p)); }
template <typename... P> __device__ __forceinline__ void pdl_enter(P &...ptrs)
{
    pdl_enter();
    (pdl_launder(ptrs), ...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace pp
