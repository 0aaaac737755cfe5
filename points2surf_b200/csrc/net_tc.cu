// placeholder replaced by the tcgen05 implementation
#include "model.cuh"
namespace p2s {
struct TcWeights {};
void tc_build(Model&) {}
void tc_destroy(Model&) {}
void forward_tc(Model&, const float*, const float*, const float*, int64_t, float*, cudaStream_t) {
    throw Error("tensor-core path not built yet");
}
}  // namespace p2s
