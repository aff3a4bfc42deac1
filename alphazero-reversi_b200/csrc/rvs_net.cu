// rvs_net.cu -- K4: AlphaZeroNetwork inference (placeholder until the tcgen05 tower lands).
#include "rvs_engine.cuh"

using namespace rvs;

int rvs_net_search(rvs_engine*, int32_t, int32_t, cudaStream_t) {
    return fail(-5, "RVS_EVAL_NN: network kernels are not built into this library yet");
}
void rvs_net_destroy(rvs::NetState*) {}

extern "C" {
int rvs_engine_load_weights(rvs_engine*, const float*, int64_t, int, void*) {
    return fail(-5, "rvs_engine_load_weights: network kernels are not built into this library yet");
}
int rvs_engine_predict(rvs_engine*, const uint64_t*, const uint64_t*, const uint8_t*, int64_t, float*, float*, int, void*) {
    return fail(-5, "rvs_engine_predict: network kernels are not built into this library yet");
}
}
