// Programmatic dependent launch (sm_90+) for the kernels a rollout step chains on one stream (actor-critic ->
// env step -> VecNormalize x 3 -> episode flags): a kernel launched through rl_launch() may be set up while
// its predecessor in the stream drains, and rl_grid_dependency_wait() -- the FIRST statement of every such
// kernel, before it touches global memory -- holds it until the predecessor has completed and its writes are
// visible.  Nothing is read or written early, so the results cannot differ; what is saved is the launch latency
// between back-to-back kernels (2-3 us each).  A kernel launched the ordinary way passes the wait at once.
// RL_B200_PDL=0 in the environment launches everything the ordinary way (for A/B measurements).
#pragma once
#include <cstdlib>
#include <utility>

#include <cuda_runtime.h>

namespace rl {

// (An explicit griddepcontrol.launch_dependents right after the wait -- the next kernel's CTAs resident while this
// kernel's last wave still runs -- measured slightly slower than the implicit trigger at CTA exit: 2.77e9 against
// 2.80e9 env-steps/s of rollout collection; not kept.)
__device__ __forceinline__ void rl_grid_dependency_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline bool pdl_enabled() {
    static const bool on = [] {
        const char* t = getenv("RL_B200_PDL");
        return !(t && t[0] == '0');
    }();
    return on;
}

template <class... Params, class... Args>
inline cudaError_t rl_launch(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace rl
