// rt_kernels_exact.cu — the bit-exact render kernel.  MUST be compiled with --fmad=false
// (and the nvcc defaults -prec-div=true -prec-sqrt=true -ftz=false): every multiply, add,
// divide and square root is then a correctly rounded IEEE binary32 operation in the
// reference's association order, so the frame equals the CPU arithmetic bit for bit.
#define RT_TU_EXACT 1
#include "rt_kernels.cuh"

namespace rt {

cudaError_t launch_render_exact(const RtFrameParams& P, const RtSceneView& G, int grid, size_t smem_limit,
                                cudaStream_t stream)
{
    return launch_render<false>(P, G, grid, smem_limit, stream);
}

cudaError_t occupancy_exact(size_t hot_bytes, size_t smem_limit, int* blocks_per_sm)
{
    return render_occupancy<false>(hot_bytes, smem_limit, blocks_per_sm);
}

}   // namespace rt
