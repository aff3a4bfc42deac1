// rvs_conv_tc.cuh -- 3x3 convolution C->C as an implicit GEMM on tcgen05 (declarations).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvs {

// Activation layout in HBM, shared by every K4 kernel: [tile][y][board-in-tile][x][c] bf16 with
// tile = board / 2.  Row index (in units of C channels) of pixel px = y*8+x of `board`:
__host__ __device__ inline size_t act_row(int64_t board, int px) {
    return (size_t)(board >> 1) * 128 + (size_t)(px >> 3) * 16 + (size_t)(board & 1) * 8 + (size_t)(px & 7);
}

struct ConvTcPlan {
    bool valid = false;
    int C = 0;
    int64_t max_batch = 0;
    void* impl = nullptr;
};

// builds the TMA descriptors for one layer's folded weights [9][C][cin] bf16 (cin = C by default)
int conv_tc_plan(ConvTcPlan& plan, const __nv_bfloat16* w, int C, int64_t max_batch, int cin = 0);
// out = relu(conv3x3(in) + bias [+ residual]) on B boards, NHWC bf16
// head_host / feat (optional, 2-CTA kernel only): fuse the policy/value 1x1 convolutions into the epilogue:
// head_host = host copy of [3][C] folded weights + 3 biases, feat = device [B][192] f32 output; `out` is then unused
// n_dev (optional, 2-CTA kernels only): device int with the actual number of boards (<= B, which then only sizes
// the grid): compacted leaf batches whose size never visits the host
int conv_tc_launch(const ConvTcPlan& plan, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out,
                   const float* bias, int64_t B, cudaStream_t s, const float* head_host = nullptr, float* feat = nullptr,
                   const int* n_dev = nullptr);
bool conv_tc_can_fuse_head(const ConvTcPlan& plan);
void conv_tc_destroy(ConvTcPlan& plan);

}  // namespace rvs
