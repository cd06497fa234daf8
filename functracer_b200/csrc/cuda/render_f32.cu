// FP32 product kernels (FMA contraction on).
#include "render.cuh"

namespace ftb {
template <>
cudaError_t launch_render<float>(const DevScene<float>& s, const DevFrame<float>& f, bool stats, int sm_count, cudaStream_t stream, int* launches)
{
    return launch_render_impl<float>(s, f, stats, sm_count, stream, launches);
}
}  // namespace ftb
