// placeholder replaced by the marching-cubes implementation
#include "common.cuh"
namespace p2s {
void marching_cubes(const float*, int, float, float*, int64_t, int32_t*, int64_t, int64_t*, int64_t*, cudaStream_t) {
    throw Error("marching cubes not built yet");
}
}  // namespace p2s
