// FP64 verification kernels: same source as the product kernels, instantiated for double and
// compiled with --fmad=false so that a*b+c rounds twice like the reference's JIT and the oracle.
#include "render.cuh"

namespace ftb {
template <>
cudaError_t launch_render<double>(const DevScene<double>& s, const DevFrame<double>& f, bool stats, int sm_count, cudaStream_t stream, int* launches)
{
    return launch_render_impl<double>(s, f, stats, sm_count, stream, launches);
}
}  // namespace ftb
