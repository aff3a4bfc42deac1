// rvs_conv_tc.cuh -- 3x3 convolution C->C as an implicit GEMM on tcgen05 (declarations).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvs {

struct ConvTcPlan {
    bool valid = false;
    int C = 0;
    int64_t max_batch = 0;
    void* impl = nullptr;
};

// builds the TMA descriptors for one layer's folded weights [9][C][C] bf16
int conv_tc_plan(ConvTcPlan& plan, const __nv_bfloat16* w, int C, int64_t max_batch);
// out = relu(conv3x3(in) + bias [+ residual]) on B boards, NHWC bf16
int conv_tc_launch(const ConvTcPlan& plan, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out,
                   const float* bias, int64_t B, cudaStream_t s);
void conv_tc_destroy(ConvTcPlan& plan);

}  // namespace rvs
