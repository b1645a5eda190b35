// pdl.h — programmatic dependent launch for chains of short kernels on one stream (the PPO update: ~140 launches of 5-30 us).
// A kernel launched with launch_pdl() may be scheduled while the kernel before it on the stream is still running: its blocks take the
// SMs that kernel's blocks leave, run their prologue (barrier init, TMEM allocation, index arithmetic) and stop at pdl_wait() until the
// earlier kernel has completed and its memory is visible — stream order is kept, only launch latency and prologues move under the tail
// of the previous kernel.  Every kernel launched this way calls pdl_enter() before it touches global memory.
// RLG_PDL=0 launches the same kernels the ordinary way (A/B).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the kernel after this one may be scheduled from now on
    asm volatile("griddepcontrol.wait;" ::: "memory");               // everything before this kernel on the stream is complete and visible
}

inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("RLG_PDL"); return !(e && atoi(e) == 0); }();
    return on;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
