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

// folded policy / value 1x1 convolutions (network.py:104-105, 111-112) for the fused last-layer epilogue:
// rows 0,1 = policy planes, row 2 = value plane, padded to 128 channels; b = the three folded BN biases.
// Passed to the kernel BY VALUE (kernel parameter space), owned per network (never shared between handles).
struct ConvHeadW {
    float w[3][128];
    float b[4];
};

// builds the TMA descriptors for one layer's folded weights [9][C][cin] bf16 (cin = C by default)
int conv_tc_plan(ConvTcPlan& plan, const __nv_bfloat16* w, int C, int64_t max_batch, int cin = 0);
// out = relu(conv3x3(in) + bias [+ residual]) on B boards, NHWC bf16
// head / feat (optional): fuse the policy/value 1x1 convolutions into the epilogue: feat = device [B][192] f32
// output; `out` is then unused
// n_dev (optional): device int with the actual number of boards (<= B, which then only sizes the grid):
// compacted leaf batches whose size never visits the host
// count_is_fresh: *n_dev is written by the kernel launched just before this one (re-read after griddepcontrol.wait)
// max_ctas (optional): cap of the persistent grid (default: all 148 SMs)
// cap_boards (optional): boards addressable from `in` (default plan.max_batch): sizes the TMA descriptor when `in`
// points into the middle of a buffer (second half-batch of a pipelined search)
// rev: walk the tiles in DESCENDING order.  Consecutive layers alternate the direction, so a layer starts on the
// tiles its predecessor wrote last -- the part of the previous output (and of the residual) that is still in L2
// (a 4096-board activation tensor is 62 MB, two of them fill the L2: in one direction every read missed)
int conv_tc_launch(const ConvTcPlan& plan, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out,
                   const float* bias, int64_t B, cudaStream_t s, const ConvHeadW* head = nullptr, float* feat = nullptr,
                   const int* n_dev = nullptr, int max_ctas = 0, int64_t cap_boards = 0, int rev = 0, int count_is_fresh = 0);
bool conv_tc_can_fuse_head(const ConvTcPlan& plan);

// The whole residual tower (2 x blocks layers, C = 128) as ONE persistent launch: a CTA pair walks its own tiles
// through every layer (tiles are whole boards, so layers never exchange data between CTAs), the next layer's weights
// replace the current ones in shared memory half by half under the last tile's MMAs.  Bit-identical to the chain of
// conv_tc_launch calls it replaces.  plan.valid stays false for shapes it does not cover (callers keep the
// per-layer path).  w_slab: the layers' folded weights back to back [n_layers][9][C][C]; bias_slab [n_layers][C].
// x: tower input (also the first block's residual), t / y: scratch of the same size; the last layer writes the
// fused heads' planes to feat (required) instead of an activation tensor.
struct ConvTowerPlan {
    bool valid = false;
    int C = 0, n_layers = 0;
    int64_t max_batch = 0;
    const float* bias = nullptr;
    void* impl = nullptr;
};
// w0 (optional): the network's first layer, folded [9][C][64] with its bias in the K dimension (conv0_bias_in_k_kernel);
// a launch with x0 (its input tiles [tile][y][board][x][64]) then runs it as layer 0 and writes its output to x.
int conv_tower_plan(ConvTowerPlan& plan, const __nv_bfloat16* w_slab, const float* bias_slab, int C, int n_layers, int64_t max_batch,
                    const __nv_bfloat16* w0 = nullptr);
int conv_tower_launch(const ConvTowerPlan& plan, __nv_bfloat16* x, __nv_bfloat16* t, __nv_bfloat16* y, int64_t B, cudaStream_t s,
                      const ConvHeadW& head, float* feat, const int* n_dev = nullptr, int max_ctas = 0, int64_t cap_boards = 0,
                      const __nv_bfloat16* x0 = nullptr);
void conv_tower_destroy(ConvTowerPlan& plan);
void conv_tc_destroy(ConvTcPlan& plan);

}  // namespace rvs
