// rvs_net.cu -- K4: AlphaZeroNetwork inference (reference: src/model/network.py:33-117) and the
// NN-evaluated lockstep search (MCTS._process_batch with model.predict, src/mcts/mcts.py:544-623).
//
// Data layout: activations NHWC bf16 [B][8][8][C]; conv weights BN-folded bf16 [tap][Cout][Cin]
// (tap = ky*3+kx, K-major for both GEMM operands); biases and the two small heads in f32.
//   conv3x3_tc2*     3x3 convolutions C->C: implicit GEMM on tcgen05 (rvs_conv_tc.cu)
//   conv0_bits       first layer (3->C, K = 27) for 64 / 256 filters: CUDA cores, fused with the leaf encoding
//   heads_kernel     policy 1x1 conv + FC + softmax, value 1x1 conv + FC + FC + tanh (fp32)
#include "rvs_engine.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "rvs_conv_tc.cuh"

namespace rvs {

// -DRVS_TIMELINE (debug builds only, never in the shipped library): CUDA events at the phase boundaries of a
// few waves of rvs_net_search_w1 / net_forward, printed as offsets from the first one (tools/probe_nn_search.py)
#ifdef RVS_TIMELINE
struct TimelineRec { const char* what; int wave, half; };
static std::vector<TimelineRec> g_tl;
static unsigned long long* g_tl_buf = nullptr;  // device: %globaltimer stamps
static bool g_tl_on = false;
static int g_tl_wave = 0, g_tl_half = 0;
__global__ void tl_stamp_kernel(unsigned long long* out) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *out = t;
}
static void tl_mark(const char* what, cudaStream_t s) {
    if (!g_tl_on || g_tl.size() >= 4096) return;
    if (!g_tl_buf) cudaMalloc(&g_tl_buf, 4096 * 8);
    tl_stamp_kernel<<<1, 1, 0, s>>>(g_tl_buf + g_tl.size());
    g_tl.push_back({what, g_tl_wave, g_tl_half});
}
static void tl_dump() {
    cudaDeviceSynchronize();
    std::vector<unsigned long long> t(g_tl.size());
    cudaMemcpy(t.data(), g_tl_buf, t.size() * 8, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < g_tl.size(); ++i)
        fprintf(stderr, "TL wave %d half %d %-12s %9.1f us\n", g_tl[i].wave, g_tl[i].half, g_tl[i].what, (double)(t[i] - t[0]) * 1e-3);
    g_tl.clear();
}
#define TL_MARK(what, s) tl_mark(what, s)
#else
#define TL_MARK(what, s) do { } while (0)
#endif

struct ConvLayer {
    __nv_bfloat16* w = nullptr;  // [9][Cout][Cin]
    float* bias = nullptr;       // [Cout]
    int cin = 0, cout = 0;
    ConvTcPlan tc;               // tensor-core plan (TMA descriptors), valid when cin == cout
};

struct NetState {
    int blocks = 0, C = 0;
    int64_t max_batch = 0;
    ConvLayer conv0;
    ConvLayer* tower = nullptr;  // 2*blocks layers (weights / biases are slices of the two slabs below)
    __nv_bfloat16* tower_w = nullptr;  // [2*blocks][9][C][C]: one slab, so that ONE TMA descriptor covers every layer
    float* tower_bias = nullptr;       // [2*blocks][C]
    ConvTowerPlan tower_tc;            // persistent whole-network kernel (C = 128), rvs_conv_tc.cu
    // heads (BN folded), f32
    float *pw = nullptr, *pb = nullptr;      // policy conv [2][C], bias [2]
    float *pfw = nullptr, *pfb = nullptr;    // policy fc TRANSPOSED [128][65], [65]
    float *vw = nullptr, *vb = nullptr;      // value conv [C], bias [1]
    float *v1w = nullptr, *v1b = nullptr;    // fc1 TRANSPOSED [64][256], [256]
    float *w0f = nullptr, *b0f = nullptr;    // first conv for the bit-plane kernel: [27][C] f32, [C]
    uint64_t* bits = nullptr;                // [B][3] own / opponent / legal bit planes (K3 output)
    int* n_valid = nullptr;                  // [8] boards in the compacted leaf batch of the current wave ([0]: lockstep path;
                                             // [2 + 2 * half + parity]: the pipelined wave-1 path, see rvs_net_search_w1)
    int* rows = nullptr;                     // [G] pipelined path: row of each game's pending leaf in its half's batch (-1: none)
    cudaStream_t side = nullptr;             // second stream of the pipelined search (the caller's stream is the first)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int* inv = nullptr;                      // [B] wave slot -> row of the compacted batch (-1: none, <= -2: same as slot -2-inv of its game)
    __nv_bfloat16* x0 = nullptr;             // [tile][y][board][x][64] bf16 input planes (3 used) for the tcgen05 first layer
    float *v2w = nullptr, *v2b = nullptr;    // fc2 [256], [1]
    // activations
    __nv_bfloat16 *a = nullptr, *b = nullptr, *c = nullptr;  // [tile][y][board][x][C] (act_row)
    float *probs = nullptr, *logits = nullptr, *values = nullptr;          // [B][65], [B][65], [B]
    float* feat = nullptr;                   // [B][192] head planes written by the fused last-layer epilogue
    ConvHeadW head;                          // folded 1x1 head weights of THIS network for the fused last-layer epilogue
    float* flat = nullptr;  // staging of the raw state_dict
    bool loaded = false;
    std::vector<void*> allocs;  // 2 per tower layer: 80 for the 20-block network
    // one wave of the NN search (select -> encode -> tower -> heads -> expand/backup, ~16 launches) captured
    // as a CUDA graph and replayed for the remaining waves of a search
    cudaGraphExec_t wave_exec = nullptr;
    EngineView wave_view;        // kernel arguments baked into the graph: re-capture when they change
    int wave_k = 0;
    int wave_kernels = 0;        // kernels per replay (launch accounting)
    bool graph_ok = true;        // cleared when capture is unavailable: plain launches from then on
};

namespace {

template <typename T>
int nalloc(NetState* n, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
    if (e != cudaSuccess) return fail(-100 - (int)e, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    cudaMemset(q, 0, count * sizeof(T));
    n->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

// ---- weight folding (eval-mode BatchNorm, eps 1e-5, network.py:19-21,48) ---------------------
// w [Cout][Cin][3][3] f32 + bn{gamma,beta,mean,var}[Cout] -> wf [9][Cout][CinPad] bf16, bias[Cout]
__global__ void fold_conv3x3_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, int cout, int cin, int cin_pad,
                                    __nv_bfloat16* __restrict__ wf, float* __restrict__ bias) {
    const int total = 9 * cout * cin_pad;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int ci = t % cin_pad, co = (t / cin_pad) % cout, tap = t / (cin_pad * cout);
        const float scale = gamma[co] / sqrtf(var[co] + 1e-5f);
        float v = 0.f;
        if (ci < cin) v = w[((size_t)co * cin + ci) * 9 + tap] * scale;
        wf[t] = __float2bfloat16(v);
        if (tap == 0 && ci == 0) bias[co] = beta[co] - mean[co] * scale;
    }
}
// 1x1 conv [Cout][Cin] + bn -> f32 folded
__global__ void fold_conv1x1_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, int cout, int cin, float* __restrict__ wf,
                                    float* __restrict__ bias) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cout * cin; t += gridDim.x * blockDim.x) {
        const int co = t / cin;
        const float scale = gamma[co] / sqrtf(var[co] + 1e-5f);
        wf[t] = w[t] * scale;
        if (t % cin == 0) bias[co] = beta[co] - mean[co] * scale;
    }
}

// ---- heads (network.py:103-117) + softmax (mcts.py:596) ---------------------------------------
// One CTA (256 threads) evaluates kHB = 8 boards: (1) the three 1x1 convolutions + BN + ReLU, one
// thread per pixel streaming its C channels with 16-byte loads (skipped when the last tower layer's
// epilogue already produced the planes); (2) policy_fc (65 rows) and value_fc1 (256 rows) with
// TRANSPOSED weights so that consecutive threads read consecutive addresses, each weight reused for
// the boards of the thread; (3) one warp per board: softmax over the 65 logits (no legal-move
// masking: mcts.py:596 applies the plain softmax), value_fc2 + tanh.
// The FC loops fetch their weights 16 rows at a time into registers before the multiply-adds: the
// first version loaded one weight per iteration and was a serial chain of 128 L2 latencies (ncu:
// 42 us for 4096 boards, issue-active 16 %, long-scoreboard stalls); features sit in shared memory
// board-minor ([feature][board]) so that one 16-byte load serves four boards.  The summation order
// over the features is unchanged (ascending), so outputs are bit-identical to the first version.
constexpr int kHB = 8;
constexpr int kFcBatch = 16;
__global__ void __launch_bounds__(256) heads_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ feat_in, int C, int64_t B,
                                                     const float* __restrict__ pw, const float* __restrict__ pb,
                                                     const float* __restrict__ pfwT, const float* __restrict__ pfb,
                                                     const float* __restrict__ vw, const float* __restrict__ vb,
                                                     const float* __restrict__ v1wT, const float* __restrict__ v1b,
                                                     const float* __restrict__ v2w, const float* __restrict__ v2b,
                                                     float* __restrict__ logits, float* __restrict__ probs,
                                                     float* __restrict__ values, const int* __restrict__ n_dev,
                                                     unsigned long long* __restrict__ nn_evals) {
    pdl_trigger();    // the tree step that follows may become resident while the heads drain
    pdl_grid_wait();  // activations / head planes come from the tower kernel launched just before
    if (n_dev) {
        B = *n_dev;
        if (nn_evals && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(nn_evals, (unsigned long long)B);
        if ((int64_t)blockIdx.x * kHB >= B) return;
    }
    __shared__ float w1x1[3 * 256];                    // [3][C] (C <= 256)
    __shared__ __align__(16) float featT[192][kHB];    // [0,128): policy planes channel-major (ch*64+px); [128,192): value plane
    __shared__ float vpart[8][kHB];                    // value_fc2 partial sums per warp
    __shared__ float lg[kHB][68];
    // (11.5 KB of shared memory and <= 48 registers: a CTA of this kernel fits beside a resident whole-network CTA, so the
    // heads of one half-batch can run under the tower of the other, rvs_net_search_w1)
    const int t = threadIdx.x;
    const int64_t board0 = (int64_t)blockIdx.x * kHB;
    if (feat_in) {  // (1') the last tower layer already produced the three head planes (fused epilogue)
#pragma unroll
        for (int r = 0; r < kHB * 192 / 256; ++r) {
            const int i = t + 256 * r;
            const int bi = i / 192, j = i - bi * 192;
            featT[j][bi] = board0 + bi < B ? feat_in[(size_t)(board0 + bi) * 192 + j] : 0.f;
        }
    } else {
        for (int i = t; i < 3 * C; i += 256) w1x1[i] = i < 2 * C ? pw[i] : vw[i - 2 * C];
    }
    __syncthreads();
    for (int p = t; p < kHB * 64 && !feat_in; p += 256) {  // (1) 1x1 convs
        const int bi = p >> 6, px = p & 63;
        const int64_t board = board0 + bi;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        if (board < B) {
            const uint4* xr = reinterpret_cast<const uint4*>(x + act_row(board, px) * C);
            for (int v = 0; v < C / 8; ++v) {
                const uint4 q = xr[v];
                const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __bfloat1622float2(q2[i]);
                    const int c = v * 8 + 2 * i;
                    a0 = fmaf(f.x, w1x1[c], a0); a0 = fmaf(f.y, w1x1[c + 1], a0);
                    a1 = fmaf(f.x, w1x1[C + c], a1); a1 = fmaf(f.y, w1x1[C + c + 1], a1);
                    a2 = fmaf(f.x, w1x1[2 * C + c], a2); a2 = fmaf(f.y, w1x1[2 * C + c + 1], a2);
                }
            }
        }
        featT[px][bi] = fmaxf(a0 + pb[0], 0.f);
        featT[64 + px][bi] = fmaxf(a1 + pb[1], 0.f);
        featT[128 + px][bi] = fmaxf(a2 + vb[0], 0.f);
    }
    __syncthreads();
    {  // (2a) value_fc1 + ReLU: row t of [256][64], transposed weights [64][256]
        float acc[kHB];
#pragma unroll
        for (int b = 0; b < kHB; ++b) acc[b] = v1b[t];
#pragma unroll
        for (int i0 = 0; i0 < 64; i0 += kFcBatch) {
            float w[kFcBatch];
#pragma unroll
            for (int u = 0; u < kFcBatch; ++u) w[u] = v1wT[(i0 + u) * 256 + t];
#pragma unroll
            for (int u = 0; u < kFcBatch; ++u) {
                const float4 f0 = *reinterpret_cast<const float4*>(&featT[128 + i0 + u][0]);
                const float4 f1 = *reinterpret_cast<const float4*>(&featT[128 + i0 + u][4]);
                acc[0] = fmaf(w[u], f0.x, acc[0]); acc[1] = fmaf(w[u], f0.y, acc[1]);
                acc[2] = fmaf(w[u], f0.z, acc[2]); acc[3] = fmaf(w[u], f0.w, acc[3]);
                acc[4] = fmaf(w[u], f1.x, acc[4]); acc[5] = fmaf(w[u], f1.y, acc[5]);
                acc[6] = fmaf(w[u], f1.z, acc[6]); acc[7] = fmaf(w[u], f1.w, acc[7]);
            }
        }
        // value_fc2 (network.py:114) folded in: this thread's hidden unit times its fc2 weight, summed over the warp here
        // and over the eight warps in (3) -- a fixed order, and no [boards][256] hidden buffer in shared memory
        const float w2 = v2w[t];
#pragma unroll
        for (int b = 0; b < kHB; ++b) {
            float pv = w2 * fmaxf(acc[b], 0.f);
            for (int o = 16; o; o >>= 1) pv += __shfl_xor_sync(0xffffffffu, pv, o);
            if ((t & 31) == 0) vpart[t >> 5][b] = pv;
        }
    }
    if (t < 130) {  // (2b) policy_fc: view(batch,-1) is channel-major (network.py:107); weights [128][65];
                    // thread = (logit o, half of the boards)
        const int o = t < 65 ? t : t - 65, bh = t < 65 ? 0 : 4;
        float acc[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[b] = pfb[o];
#pragma unroll 2
        for (int i0 = 0; i0 < 128; i0 += kFcBatch) {
            float w[kFcBatch];
#pragma unroll
            for (int u = 0; u < kFcBatch; ++u) w[u] = pfwT[(i0 + u) * 65 + o];
#pragma unroll
            for (int u = 0; u < kFcBatch; ++u) {
                const float4 f = *reinterpret_cast<const float4*>(&featT[i0 + u][bh]);
                acc[0] = fmaf(w[u], f.x, acc[0]); acc[1] = fmaf(w[u], f.y, acc[1]);
                acc[2] = fmaf(w[u], f.z, acc[2]); acc[3] = fmaf(w[u], f.w, acc[3]);
            }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) lg[bh + b][o] = acc[b];
    }
    __syncthreads();
    {  // (3) warp w <-> board w
        const int w = t >> 5, l = t & 31;
        const int64_t board = board0 + w;
        if (board < B) {
            float m = fmaxf(lg[w][l], lg[w][l + 32]);
            if (l == 0) m = fmaxf(m, lg[w][64]);
            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            const float e0 = expf(lg[w][l] - m), e1 = expf(lg[w][l + 32] - m), e2 = l == 0 ? expf(lg[w][64] - m) : 0.f;
            float s = e0 + e1 + e2;
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (probs) {
                float* pr = probs + (size_t)board * 65;
                pr[l] = e0 / s;
                pr[l + 32] = e1 / s;
                if (l == 0) pr[64] = e2 / s;
            }
            if (logits) {
                float* lo = logits + (size_t)board * 65;
                lo[l] = lg[w][l];
                lo[l + 32] = lg[w][l + 32];
                if (l == 0) lo[64] = lg[w][64];
            }
            if (l == 0) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc += vpart[i][w];
                values[board] = tanhf(acc + v2b[0]);
            }
        }
    }
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < rows * cols; t += gridDim.x * blockDim.x)
        out[(t % cols) * rows + t / cols] = in[t];
}

// first layer folded for the bit-plane kernel: w [C][3][3][3] -> wf [tap*3+plane][C] f32
__global__ void fold_conv0_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                                  const float* __restrict__ mean, const float* __restrict__ var, int C,
                                  float* __restrict__ wf, float* __restrict__ bias) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 27 * C; t += gridDim.x * blockDim.x) {
        const int co = t % C, tp = t / C, tap = tp / 3, plane = tp % 3;
        const float scale = gamma[co] / sqrtf(var[co] + 1e-5f);
        wf[t] = w[((size_t)co * 3 + plane) * 9 + tap] * scale;
        if (tp == 0) bias[co] = beta[co] - mean[co] * scale;
    }
}

// tensor-core first layer: the folded bias becomes weights of the two constant-one input channels (3: the bf16
// head of the bias, 4: the bf16 remainder; together 16 significant bits) at the centre tap, which never reads padding
__global__ void conv0_bias_in_k_kernel(__nv_bfloat16* __restrict__ w /*[9][C][64]*/, const float* __restrict__ bias, int C) {
    for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < C; co += gridDim.x * blockDim.x) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(bias[co]);
        const __nv_bfloat16 lo = __float2bfloat16_rn(bias[co] - __bfloat162float(hi));
        w[((size_t)4 * C + co) * 64 + 3] = hi;
        w[((size_t)4 * C + co) * 64 + 4] = lo;
    }
}

// K3 + first convolution fused (network.py:97): the network input is never materialised.  The
// three input planes are bit masks (own discs, opponent discs, legal squares), so the 3->C 3x3
// convolution is, per output pixel, a sum of at most 27 weight rows selected by neighbour bits.
// One warp per output pixel, lanes across the output channels (C/32 each): the bit tests are
// warp-uniform (no divergence), only set bits cost work, and a warp writes one contiguous row.
template <int C>
__global__ void __launch_bounds__(256) conv0_bits_kernel(const uint64_t* __restrict__ bits /*[B][3]*/, int64_t B,
                                                          const float* __restrict__ wf, const float* __restrict__ bias,
                                                          __nv_bfloat16* __restrict__ out, const int* __restrict__ n_dev) {
    constexpr int PER = C / 32;  // couts per lane: 2, 4 or 8
    if (n_dev) {
        B = *n_dev;
        if ((int64_t)blockIdx.x * 2 >= B) return;
    }
    __shared__ __align__(16) float sw[27 * C];
    for (int i = threadIdx.x; i < 27 * C; i += 256) sw[i] = wf[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float bv[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) bv[i] = bias[lane * PER + i];
    // one CTA per tile of two boards = 128 rows; warp w handles rows w, w+8, ...
    const int64_t tile = blockIdx.x;
    const uint64_t* b0 = bits + tile * 6;
    const bool h0 = tile * 2 < B, h1 = tile * 2 + 1 < B;
    const uint64_t a0 = h0 ? b0[0] : 0ULL, a1 = h0 ? b0[1] : 0ULL, a2 = h0 ? b0[2] : 0ULL;
    const uint64_t c0 = h1 ? b0[3] : 0ULL, c1 = h1 ? b0[4] : 0ULL, c2 = h1 ? b0[5] : 0ULL;
    for (int row = warp; row < 128; row += 8) {
        const int y = row >> 4, b = (row >> 3) & 1, x = row & 7;
        const uint64_t pl[3] = {b ? c0 : a0, b ? c1 : a1, b ? c2 : a2};
        float acc[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) acc[i] = bv[i];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
            if (yy < 0 || yy > 7 || xx < 0 || xx > 7) continue;  // warp-uniform
            const int sq = yy * 8 + xx;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                if ((pl[p] >> sq) & 1) {                         // warp-uniform
                    const float* wr = sw + (tap * 3 + p) * C + lane * PER;
#pragma unroll
                    for (int i = 0; i < PER; ++i) acc[i] += wr[i];
                }
            }
        }
        __nv_bfloat16* o = out + ((size_t)tile * 128 + row) * C + lane * PER;
#pragma unroll
        for (int i = 0; i < PER; i += 2)
            *reinterpret_cast<__nv_bfloat162*>(o + i) = __floats2bfloat162_rn(fmaxf(acc[i], 0.f), fmaxf(acc[i + 1], 0.f));
    }
}

// bit planes -> bf16 input tiles [tile][y][board][x][64] (channels 0..2 used) for the tensor-core first
// layer: one thread per pixel row, 128 B = 4 x 256-bit... written as 8 x uint4
__global__ void __launch_bounds__(256) planes_tiles_kernel(const uint64_t* __restrict__ bits, int64_t B, int64_t n_tiles,
                                                            uint4* __restrict__ out, const int* __restrict__ n_dev) {
    if (n_dev) {
        B = *n_dev;
        n_tiles = (B + 1) >> 1;
    }
    const int64_t total = n_tiles * 128;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tile = r >> 7;
        const int row = (int)(r & 127), y = row >> 4, b = (row >> 3) & 1, x = row & 7;
        const int64_t board = tile * 2 + b;
        uint32_t c01 = 0, c23 = 0, c45 = 0;
        if (board < B) {
            const int sq = y * 8 + x;
            const uint32_t one = 0x3F80u;
            c01 = (((bits[board * 3] >> sq) & 1) ? one : 0u) | ((((bits[board * 3 + 1] >> sq) & 1) ? one : 0u) << 16);
            // channels 3 and 4 are constant 1 on every real pixel: the first layer's folded bias rides on them as two
            // bf16 weights (hi + lo) of the centre tap (conv0_bias_in_k_kernel), so its epilogue adds nothing
            c23 = (((bits[board * 3 + 2] >> sq) & 1) ? one : 0u) | (one << 16);
            c45 = one;
        }
        // only channels 0..15 (the first 32 bytes of the 128-byte row) are ever read by the first layer's single
        // K = 16 step (Cfg2::KSTEPS); the rest of the buffer stays at its initial zero
        uint4* o = out + r * 8;
        o[0] = make_uint4(c01, c23, c45, 0u);
        o[1] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// K3 with compaction: only leaves that need a network evaluation enter the batch, and the leaves of
// one game's wave that sit on the SAME node are evaluated once (the reference's wave sends most of its
// simulations down one path, SURVEY.md 0.3, and evaluates every copy: mcts.py:586-597).  Rows are
// appended with one atomic per unique leaf; their order is arbitrary, which cannot change any result
// because the network treats boards independently.
__global__ void __launch_bounds__(256) encode_compact_kernel(EngineView ev, int k, uint64_t* __restrict__ out,
                                                              int* __restrict__ inv, int* __restrict__ n_valid) {
    const int64_t total = (int64_t)ev.G * k;
    for (int64_t slot = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; slot < total; slot += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(slot / k), j = (int)(slot - (int64_t)g * k);
        const size_t o = (size_t)g * ev.kmax + j;
        const int node = ev.w_node[o];
        const uint64_t lm = node < 0 ? 0ULL : ev.w_lm[o];
        int code = -1;
        if (lm) {
            int rep = -1;
            for (int q = 0; q < j; ++q)
                if (ev.w_node[(size_t)g * ev.kmax + q] == node) { rep = q; break; }
            if (rep >= 0) {
                code = -2 - rep;
            } else {
                code = atomicAdd(n_valid, 1);
                const bool blk = (ev.w_sf[o] & 0xFF) == 1;
                out[(size_t)code * 3] = blk ? ev.w_black[o] : ev.w_white[o];
                out[(size_t)code * 3 + 1] = blk ? ev.w_white[o] : ev.w_black[o];
                out[(size_t)code * 3 + 2] = lm;
            }
        }
        inv[slot] = code;
    }
}

template <int RULES>
__global__ void __launch_bounds__(256) encode_positions_kernel(const uint64_t* __restrict__ black,
                                                                const uint64_t* __restrict__ white,
                                                                const uint8_t* __restrict__ side, int64_t n,
                                                                uint64_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool blk = side[i] == 1;
        const uint64_t P = blk ? black[i] : white[i], O = blk ? white[i] : black[i];
        out[i * 3] = P; out[i * 3 + 1] = O; out[i * 3 + 2] = legal_moves<RULES>(P, O);
    }
}

}  // namespace

// forward pass on B boards whose bit planes are already in n->bits (+ off boards); results in n->probs / logits /
// values (+ off).  `off` (even: whole tiles) and `cap` select a sub-range of every buffer: the half-batches of a
// pipelined search run their own forward passes on their own streams.
int net_forward(rvs_engine* h, int64_t B, bool want_logits, cudaStream_t s, const int* n_dev = nullptr, int64_t off = 0,
                int64_t cap = 0, bool tiles_written = false, bool pdl_heads = true, int max_ctas = -1) {
    NetState* n = h->net;
    if (!n || !n->loaded) return fail(-7, "network weights not loaded: call rvs_engine_load_weights first");
    if (cap <= 0) cap = n->max_batch - off;
    if (B > cap || off + cap > n->max_batch + 1 || (off & 1)) return fail(-7, "network batch %lld (+%lld) exceeds capacity %lld", (long long)B, (long long)off, (long long)n->max_batch);
    if (B == 0) return 0;
    int rc;
    const int C = n->C;
    const int mc = max_ctas >= 0 ? max_ctas : h->net_max_ctas;
    const uint64_t* bits = n->bits + off * 3;
    __nv_bfloat16* x0 = n->x0 + off * 64 * 64;
    __nv_bfloat16 *x = n->a + off * 64 * C, *t = n->b + off * 64 * C, *y = n->c + off * 64 * C;
    float* feat = n->feat + off * 192;
    const bool whole_net = n->tower_tc.valid && h->net_tower && n->conv0.tc.valid;  // first layer + tower + head planes in ONE launch
    if (n->conv0.tc.valid) {  // first layer on the tensor cores: planes -> bf16 tiles -> tcgen05
        const int64_t tiles = (B + 1) / 2;
        // (the wave-1 search's tree step kernel writes the tiles itself: tiles_written)
        if (!tiles_written) RVS_LAUNCH(planes_tiles_kernel, grid_for(tiles * 128, 256), 256, 0, s, bits, B, tiles, (uint4*)x0, n_dev);
        TL_MARK("planes", s);
        if (!whole_net && (rc = conv_tc_launch(n->conv0.tc, x0, nullptr, x, n->conv0.bias, B, s, nullptr, nullptr, n_dev, mc, cap, 0, tiles_written ? 1 : 0))) return rc;
        TL_MARK("conv0", s);
    } else {  // network.py:97, fused with the leaf encoding (CUDA cores)
        const int tiles = (int)((B + 1) / 2);
        if (C == 64) RVS_LAUNCH(conv0_bits_kernel<64>, tiles, 256, 0, s, bits, B, n->w0f, n->b0f, x, n_dev);
        else if (C == 128) RVS_LAUNCH(conv0_bits_kernel<128>, tiles, 256, 0, s, bits, B, n->w0f, n->b0f, x, n_dev);
        else RVS_LAUNCH(conv0_bits_kernel<256>, tiles, 256, 0, s, bits, B, n->w0f, n->b0f, x, n_dev);
    }
    bool fused_head = false;
    if (whole_net) {  // first layer + all ResBlocks (network.py:97-101) + the heads' 1x1 convs in one persistent launch
        if ((rc = conv_tower_launch(n->tower_tc, x, t, y, B, s, n->head, feat, n_dev, mc, cap, x0))) return rc;
        fused_head = true;
    } else
    for (int i = 0; i < n->blocks; ++i) {  // ResBlock (network.py:23-28)
        const ConvLayer& c1 = n->tower[2 * i];
        const ConvLayer& c2 = n->tower[2 * i + 1];
        if ((rc = conv_tc_launch(c1.tc, x, nullptr, t, c1.bias, B, s, nullptr, nullptr, n_dev, mc, cap, 1))) return rc;  // layers alternate the tile direction (L2 reuse)
        if (i == n->blocks - 1 && conv_tc_can_fuse_head(c2.tc)) {  // last layer: heads' 1x1 convs in the epilogue
            if ((rc = conv_tc_launch(c2.tc, t, x, y, c2.bias, B, s, &n->head, feat, n_dev, mc, cap, 0))) return rc;
            fused_head = true;
        } else if ((rc = conv_tc_launch(c2.tc, t, x, y, c2.bias, B, s, nullptr, nullptr, n_dev, mc, cap, 0))) return rc;
        __nv_bfloat16* tmp = x; x = y; y = tmp;
    }
    TL_MARK("tower", s);
    if (pdl_heads) {
        RVS_LAUNCH_PDL(heads_kernel, (int)((B + kHB - 1) / kHB), 256, 0, s, (const __nv_bfloat16*)x, fused_head ? (const float*)feat : (const float*)nullptr, C, B,
                       (const float*)n->pw, (const float*)n->pb, (const float*)n->pfw, (const float*)n->pfb, (const float*)n->vw, (const float*)n->vb,
                       (const float*)n->v1w, (const float*)n->v1b, (const float*)n->v2w, (const float*)n->v2b,
                       want_logits ? n->logits + off * 65 : (float*)nullptr, n->probs + off * 65, n->values + off, n_dev,
                       n_dev ? (unsigned long long*)(h->v.stats + ST_NNEVALS) : (unsigned long long*)nullptr);
    } else {  // pipelined half-batches: see rvs_engine_nn_step
        RVS_LAUNCH(heads_kernel, (int)((B + kHB - 1) / kHB), 256, 0, s, (const __nv_bfloat16*)x, fused_head ? (const float*)feat : (const float*)nullptr, C, B,
                       (const float*)n->pw, (const float*)n->pb, (const float*)n->pfw, (const float*)n->pfb, (const float*)n->vw, (const float*)n->vb,
                       (const float*)n->v1w, (const float*)n->v1b, (const float*)n->v2w, (const float*)n->v2b,
                       want_logits ? n->logits + off * 65 : (float*)nullptr, n->probs + off * 65, n->values + off, n_dev,
                       n_dev ? (unsigned long long*)(h->v.stats + ST_NNEVALS) : (unsigned long long*)nullptr);
    }
    TL_MARK("heads", s);
    return 0;
}

int net_create(rvs_engine* h) {
    if (h->net) return 0;
    const int blocks = h->cfg.net_blocks, C = h->cfg.net_filters;
    if (blocks < 1 || blocks > 40 || (C != 64 && C != 128 && C != 256))
        return fail(-1, "network: net_blocks in [1,40] and net_filters in {64,128,256} required (got %d, %d)", blocks, C);
    NetState* n = new NetState();
    h->net = n;
    n->blocks = blocks;
    n->C = C;
    n->max_batch = (int64_t)h->v.G * h->cfg.max_wave;
    const size_t B = (size_t)((n->max_batch + 1) / 2) * 2 + 2;  // whole tiles of two boards (+ one tile: the two half-batches of a pipelined search start on tile boundaries)
    int rc = 0;
    n->tower = new ConvLayer[2 * blocks];
    if ((rc = nalloc(n, &n->conv0.w, (size_t)9 * C * 64)) || (rc = nalloc(n, &n->conv0.bias, (size_t)C)) ||
        (rc = nalloc(n, &n->x0, B * 64 * 64)))
        return rc;
    n->conv0.cin = 64; n->conv0.cout = C;
    // one contiguous slab for the tower weights so that a single TMA descriptor covers every layer
    if ((rc = nalloc(n, &n->tower_w, (size_t)2 * blocks * 9 * C * C)) || (rc = nalloc(n, &n->tower_bias, (size_t)2 * blocks * C))) return rc;
    for (int i = 0; i < 2 * blocks; ++i) {
        ConvLayer& L = n->tower[i];
        L.cin = C; L.cout = C;
        L.w = n->tower_w + (size_t)i * 9 * C * C;
        L.bias = n->tower_bias + (size_t)i * C;
    }
    if ((rc = nalloc(n, &n->pw, (size_t)2 * C)) || (rc = nalloc(n, &n->pb, 2)) || (rc = nalloc(n, &n->pfw, 65 * 128)) ||
        (rc = nalloc(n, &n->pfb, 65)) || (rc = nalloc(n, &n->vw, (size_t)C)) || (rc = nalloc(n, &n->vb, 1)) ||
        (rc = nalloc(n, &n->v1w, 256 * 64)) || (rc = nalloc(n, &n->v1b, 256)) || (rc = nalloc(n, &n->v2w, 256)) ||
        (rc = nalloc(n, &n->v2b, 1)) || (rc = nalloc(n, &n->bits, B * 3)) || (rc = nalloc(n, &n->n_valid, 8)) || (rc = nalloc(n, &n->rows, (size_t)h->v.G)) || (rc = nalloc(n, &n->inv, B)) || (rc = nalloc(n, &n->w0f, (size_t)27 * C)) || (rc = nalloc(n, &n->b0f, (size_t)C)) || (rc = nalloc(n, &n->a, B * 64 * C)) ||
        (rc = nalloc(n, &n->b, B * 64 * C)) || (rc = nalloc(n, &n->c, B * 64 * C)) || (rc = nalloc(n, &n->probs, B * 65)) ||
        (rc = nalloc(n, &n->logits, B * 65)) || (rc = nalloc(n, &n->values, B)) || (rc = nalloc(n, &n->feat, B * 192)))
        return rc;
    return 0;
}

int64_t net_param_floats(int blocks, int C) {
    int64_t n = (int64_t)C * 27 + 4 * C;
    n += (int64_t)blocks * 2 * ((int64_t)C * C * 9 + 4 * C);
    n += 2 * C + 8 + 65 * 128 + 65;
    n += C + 4 + 256 * 64 + 256 + 256 + 1;
    return n;
}

}  // namespace rvs

using namespace rvs;

void rvs_net_destroy(rvs::NetState* n) {
    if (!n) return;
    if (n->wave_exec) cudaGraphExecDestroy(n->wave_exec);
    if (n->side) cudaStreamDestroy(n->side);
    if (n->ev_fork) cudaEventDestroy(n->ev_fork);
    if (n->ev_join) cudaEventDestroy(n->ev_join);
    for (void* q : n->allocs) cudaFree(q);
    if (n->flat) cudaFree(n->flat);
    for (int i = 0; i < 2 * n->blocks; ++i) conv_tc_destroy(n->tower[i].tc);
    conv_tower_destroy(n->tower_tc);
    conv_tc_destroy(n->conv0.tc);
    delete[] n->tower;
    delete n;
}

// MCTS.search with the built-in network: per wave  select -> encode (K3) -> tower + heads (K4)
// -> expand/backup with the softmax priors (K2).  No host round trip inside the loop.
static int net_wave(rvs_engine* h, int k, cudaStream_t s) {
    int rc;
    if ((rc = rvs_engine_select(h, k, s))) return rc;
    const int64_t B = (int64_t)h->v.G * k;  // capacity; the batch itself is compacted on the device
    RVS_CUDA(cudaMemsetAsync(h->net->n_valid, 0, sizeof(int), s));
    RVS_LAUNCH(encode_compact_kernel, grid_for(B, 256), 256, 0, s, h->v, k, h->net->bits, h->net->inv, h->net->n_valid);
    h->launches++;
    if ((rc = net_forward(h, B, false, s, h->net->n_valid))) return rc;
    return rvs_engine_process_mapped(h, h->net->probs, h->net->values, h->net->inv, s);
}

// MCTS.search with the built-in network and batch_size 1 (the configs[2] / configs[3] path): per wave and per
// half-batch  [process pending leaf | select | encode] (ONE tree kernel, rvs_engine_nn_step) -> tower -> heads.
// With RVS_OPT_NET_PIPELINE (default) the games are split into two half-batches that ping-pong on two streams:
// while the tcgen05 tower of one half owns the SMs' shared memory, the tree kernel of the other half (no shared
// memory, few registers) runs beside it, so selection / expansion / backup leave the critical path.  Per-game
// results do not depend on the split: every game's tree is touched by its own warp only, and the network
// treats boards independently.
static int rvs_net_search_w1(rvs_engine* h, int32_t num_sims, cudaStream_t s) {
    NetState* n = h->net;
    int rc;
    const int G = h->v.G;
    // Two half-batches pay when the network of a half is ONE launch (the whole-network kernel) and the halves are big enough
    // to fill the tensor-core grid; that grid then leaves 12 SMs to the other half's tree step / heads (measured on B200,
    // 5x128, 4096 games: lockstep 6.85, halves on 148 / 140 / 136 / 132 / 124 CTAs 6.74 / 7.03 / 7.24 / 7.16 / 7.03 M sims/s --
    // beside a resident whole-network CTA the small kernels make little progress, on SMs of their own they do).
    const bool two = h->net_pipeline && G >= 2048 && n->tower_tc.valid && h->net_tower && n->conv0.tc.valid;
    const int mc = h->net_max_ctas > 0 ? h->net_max_ctas : (two ? kNumSMs - 12 : 0);
    if (two && !n->side) {
        RVS_CUDA(cudaStreamCreateWithFlags(&n->side, cudaStreamNonBlocking));
        RVS_CUDA(cudaEventCreateWithFlags(&n->ev_fork, cudaEventDisableTiming));
        RVS_CUDA(cudaEventCreateWithFlags(&n->ev_join, cudaEventDisableTiming));
    }
    const int nh = two ? 2 : 1;
    const int split = two ? ((G / 2) & ~1) : G;  // even: the second half starts on a tile boundary
    const int g0[2] = {0, split}, g1[2] = {split, G};
    cudaStream_t st[2] = {s, two ? n->side : s};
    RVS_CUDA(cudaMemsetAsync(n->n_valid, 0, 8 * sizeof(int), s));
    if (two) {
        RVS_CUDA(cudaEventRecord(n->ev_fork, s));
        RVS_CUDA(cudaStreamWaitEvent(n->side, n->ev_fork, 0));
    }
    const bool fast = h->v.mode == RVS_MODE_FAST;
    (void)fast;  // wave 1: the FAST schedule (first wave = one simulation) is the ordinary one
    for (int w = 0; w <= num_sims; ++w) {
        for (int hf = 0; hf < nh; ++hf) {
#ifdef RVS_TIMELINE
            g_tl_on = (w >= 40 && w < 43);
            g_tl_wave = w; g_tl_half = hf;
            TL_MARK("begin", st[hf]);
#endif
            const int64_t off = g0[hf], cap = g1[hf] - g0[hf];
            int* cur = n->n_valid + 2 + 2 * hf + (w & 1);
            int* nxt = n->n_valid + 2 + 2 * hf + ((w + 1) & 1);
            const int flags = (w > 0 ? 1 : 0) | (w < num_sims ? 2 : 0) | (w == 1 ? 4 : 0);
            void* tiles = n->conv0.tc.valid ? (void*)(n->x0 + off * 64 * 64) : nullptr;  // first layer on the tensor cores
            if ((rc = rvs_engine_nn_step(h, g0[hf], g1[hf], flags, n->probs + off * 65, n->values + off, n->rows, n->bits + off * 3,
                                         cur, nxt, tiles, st[hf], !two)))
                return rc;
            TL_MARK("tree", st[hf]);
            if (w < num_sims && (rc = net_forward(h, cap, false, st[hf], cur, off, cap, tiles != nullptr, !two, mc))) return rc;
        }
    }
#ifdef RVS_TIMELINE
    g_tl_on = false;
    if (!g_tl.empty()) tl_dump();
#endif
    if (two) {
        RVS_CUDA(cudaEventRecord(n->ev_join, n->side));
        RVS_CUDA(cudaStreamWaitEvent(s, n->ev_join, 0));
    }
    h->cur_k = 0;
    h->searching = false;
    return 0;
}

// MCTS.search with the built-in network: per wave  select -> encode (K3) -> tower + heads (K4)
// -> expand/backup with the softmax priors (K2).  No host round trip inside the loop.  Optionally
// the first wave (root expansion, root noise, one-time kernel setup) is launched
// directly and the following full waves replay one captured CUDA graph (RVS_OPT_NET_GRAPH).
int rvs_net_search(rvs_engine* h, int32_t num_sims, int32_t wave, cudaStream_t s) {
    int rc;
    if ((rc = net_create(h))) return rc;
    NetState* n = h->net;
    if (!n->loaded) return fail(-7, "RVS_EVAL_NN: call rvs_engine_load_weights before rvs_engine_search");
    if ((rc = rvs_engine_begin_search(h, s))) return rc;
    if (wave == 1 && !h->net_graph) return rvs_net_search_w1(h, num_sims, s);
    // Measured on B200 (5x128, 4096 games, 100 waves): 81.5 ms with the graph, 80.1 ms with plain launches --
    // the wave loop is not launch bound (PDL already chains the tower), so replay is opt-in
    const bool use_graph = h->net_graph != 0;  // RVS_OPT_NET_GRAPH
    const bool fast = h->v.mode == RVS_MODE_FAST;
    for (int start = 0, k = 0; start < num_sims; start += k) {
        k = (fast && start == 0) ? 1 : (num_sims - start < wave ? num_sims - start : wave);  // FAST: wave 0 expands the root alone
        const bool replayable = use_graph && n->graph_ok && start > 0 && k == wave && num_sims / wave >= 4;
        if (!replayable) {
            if ((rc = net_wave(h, k, s))) return rc;
            continue;
        }
        if (n->wave_exec && (n->wave_k != k || memcmp(&n->wave_view, &h->v, sizeof(EngineView)) != 0)) {
            cudaGraphExecDestroy(n->wave_exec);
            n->wave_exec = nullptr;
        }
        if (!n->wave_exec) {  // capture this wave (nothing runs while capturing), then fall through to the replay
            const int64_t l0 = g_launches.load();
            const int64_t hl0 = h->launches;
            const int wd0 = h->waves_done;
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError();
                n->graph_ok = false;
                if ((rc = net_wave(h, k, s))) return rc;
                continue;
            }
            rc = net_wave(h, k, s);
            const cudaError_t ce = cudaStreamEndCapture(s, &graph);
            n->wave_kernels = (int)(g_launches.load() - l0);
            g_launches.store(l0);  // captured, not run
            h->launches = hl0;
            h->waves_done = wd0;
            if (rc || ce != cudaSuccess || !graph || cudaGraphInstantiate(&n->wave_exec, graph, 0) != cudaSuccess) {
                cudaGetLastError();
                if (graph) cudaGraphDestroy(graph);
                n->wave_exec = nullptr;
                n->graph_ok = false;
                if ((rc = net_wave(h, k, s))) return rc;
                continue;
            }
            cudaGraphDestroy(graph);
            n->wave_view = h->v;
            n->wave_k = k;
        }
        RVS_CUDA(cudaGraphLaunch(n->wave_exec, s));
        g_launches.fetch_add(n->wave_kernels, std::memory_order_relaxed);
        h->launches += n->wave_kernels;
        h->waves_done++;
    }
    h->cur_k = 0;
    h->searching = false;
    return 0;
}

extern "C" {

int rvs_engine_load_weights(rvs_engine* h, const float* flat, int64_t n_floats, int mem, void* stream) {
    RVS_ENTER(h);
    int rc;
    if ((rc = net_create(h))) return rc;
    NetState* n = h->net;
    const int C = n->C, blocks = n->blocks;
    const int64_t need = net_param_floats(blocks, C);
    if (!flat || n_floats != need)
        return fail(-1, "rvs_engine_load_weights: expected %lld floats for %dx%d, got %lld", (long long)need, blocks, C, (long long)n_floats);
    cudaStream_t s = (cudaStream_t)stream;
    if (n->wave_exec) {  // the fused-head weights travel as kernel parameters baked into the captured wave
        cudaGraphExecDestroy(n->wave_exec);
        n->wave_exec = nullptr;
    }
    if (!n->flat) RVS_CUDA(cudaMalloc(&n->flat, need * sizeof(float)));
    RVS_CUDA(cudaMemcpyAsync(n->flat, flat, need * sizeof(float), mem != RVS_MEM_DEVICE ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    const float* p = n->flat;
    auto take = [&](int64_t k) { const float* q = p; p += k; return q; };
    {
        const float* w = take((int64_t)C * 27);
        const float *g = take(C), *b = take(C), *m = take(C), *v = take(C);
        RVS_LAUNCH(fold_conv3x3_kernel, 64, 256, 0, s, w, g, b, m, v, C, 3, 64, n->conv0.w, n->conv0.bias);
        RVS_LAUNCH(fold_conv0_kernel, 16, 256, 0, s, w, g, b, m, v, C, n->w0f, n->b0f);
        if (C == 128) RVS_LAUNCH(conv0_bias_in_k_kernel, 1, 128, 0, s, n->conv0.w, n->conv0.bias, C);  // tensor-core first layer
    }
    for (int i = 0; i < 2 * blocks; ++i) {
        const float* w = take((int64_t)C * C * 9);
        const float *g = take(C), *b = take(C), *m = take(C), *v = take(C);
        RVS_LAUNCH(fold_conv3x3_kernel, 592, 256, 0, s, w, g, b, m, v, C, C, C, n->tower[i].w, n->tower[i].bias);
    }
    {
        const float* w = take(2 * C);
        const float *g = take(2), *b = take(2), *m = take(2), *v = take(2);
        RVS_LAUNCH(fold_conv1x1_kernel, 4, 128, 0, s, w, g, b, m, v, 2, C, n->pw, n->pb);
        RVS_LAUNCH(transpose_kernel, 32, 256, 0, s, take(65 * 128), n->pfw, 65, 128);
        RVS_CUDA(cudaMemcpyAsync(n->pfb, take(65), 65 * 4, cudaMemcpyDeviceToDevice, s));
    }
    {
        const float* w = take(C);
        const float *g = take(1), *b = take(1), *m = take(1), *v = take(1);
        RVS_LAUNCH(fold_conv1x1_kernel, 2, 128, 0, s, w, g, b, m, v, 1, C, n->vw, n->vb);
        RVS_LAUNCH(transpose_kernel, 64, 256, 0, s, take(256 * 64), n->v1w, 256, 64);
        RVS_CUDA(cudaMemcpyAsync(n->v1b, take(256), 256 * 4, cudaMemcpyDeviceToDevice, s));
        RVS_CUDA(cudaMemcpyAsync(n->v2w, take(256), 256 * 4, cudaMemcpyDeviceToDevice, s));
        RVS_CUDA(cudaMemcpyAsync(n->v2b, take(1), 4, cudaMemcpyDeviceToDevice, s));
    }
    for (int i = 0; i < 2 * blocks; ++i) {
        if ((rc = conv_tc_plan(n->tower[i].tc, n->tower[i].w, C, n->max_batch))) return rc;
    }
    if (C == 128) {  // 128 filters: the first layer runs on the tensor cores too (64 -> 128 variant, one K = 16 step per tap)
        if ((rc = conv_tc_plan(n->conv0.tc, n->conv0.w, C, n->max_batch, 64))) return rc;
    }
    if ((rc = conv_tower_plan(n->tower_tc, n->tower_w, n->tower_bias, C, 2 * blocks, n->max_batch, C == 128 ? n->conv0.w : nullptr))) return rc;  // valid for C = 128
    RVS_CUDA(cudaStreamSynchronize(s));
    if (C <= 128) {  // host copy of the folded 1x1 head weights for the fused last-layer epilogue (kernel parameter)
        float hw[3 * 128 + 4];
        RVS_CUDA(cudaMemcpy(hw, n->pw, (size_t)2 * C * 4, cudaMemcpyDeviceToHost));
        RVS_CUDA(cudaMemcpy(hw + 2 * C, n->vw, (size_t)C * 4, cudaMemcpyDeviceToHost));
        RVS_CUDA(cudaMemcpy(hw + 3 * C, n->pb, 8, cudaMemcpyDeviceToHost));
        RVS_CUDA(cudaMemcpy(hw + 3 * C + 2, n->vb, 4, cudaMemcpyDeviceToHost));
        for (int j = 0; j < 3; ++j)
            for (int c = 0; c < 128; ++c) n->head.w[j][c] = c < C ? hw[j * C + c] : 0.f;
        for (int j = 0; j < 3; ++j) n->head.b[j] = hw[3 * C + j];
        n->head.b[3] = 0.f;
    }
    n->loaded = true;
    return 0;
}

// AlphaZeroNetwork.predict on packed positions; any of out_logits / out_probs may be null
static int predict_impl(rvs_engine* h, const uint64_t* black, const uint64_t* white, const uint8_t* side, int64_t n_pos,
                        float* out_logits, float* out_probs, float* out_value, int mem, void* stream, const char* who) {
    RVS_ENTER(h);
    if (!h->net || !h->net->loaded) return fail(-7, "%s: weights not loaded", who);
    if (n_pos < 0 || (n_pos > 0 && (!black || !white || !side || !(out_logits || out_probs) || !out_value))) return fail(-1, "%s: bad arguments", who);
    if (mem < RVS_MEM_DEVICE || mem > RVS_MEM_HOST_ASYNC) return fail(-1, "%s: bad mem %d", who, mem);
    NetState* n = h->net;
    cudaStream_t s = (cudaStream_t)stream;
    const bool host = mem != RVS_MEM_DEVICE;  // RVS_MEM_HOST_ASYNC is treated like RVS_MEM_HOST here
    const int hmem = host ? RVS_MEM_HOST : RVS_MEM_DEVICE;
    std::unique_lock<std::mutex> lk(g_stage_mu, std::defer_lock);
    if (host) lk.lock();
    for (int64_t off = 0; off < n_pos; off += n->max_batch) {
        const int64_t B = n_pos - off < n->max_batch ? n_pos - off : n->max_batch;
        Arg ab, aw, as;
        int rc;
        if ((rc = arg_in(ab, black + off, B * 8, hmem, 0, s)) || (rc = arg_in(aw, white + off, B * 8, hmem, 1, s)) ||
            (rc = arg_in(as, side + off, B, hmem, 2, s)))
            return rc;
        if (h->cfg.rules == RVS_RULES_STRICT)
            RVS_LAUNCH(encode_positions_kernel<RULES_STRICT>, grid_for(B, 256), 256, 0, s, (const uint64_t*)ab.dev,
                       (const uint64_t*)aw.dev, (const uint8_t*)as.dev, B, n->bits);
        else
            RVS_LAUNCH(encode_positions_kernel<RULES_REF>, grid_for(B, 256), 256, 0, s, (const uint64_t*)ab.dev,
                       (const uint64_t*)aw.dev, (const uint8_t*)as.dev, B, n->bits);
        if ((rc = net_forward(h, B, out_logits != nullptr, s))) return rc;
        const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        if (out_logits) RVS_CUDA(cudaMemcpyAsync(out_logits + off * 65, n->logits, B * 65 * 4, kind, s));
        if (out_probs) RVS_CUDA(cudaMemcpyAsync(out_probs + off * 65, n->probs, B * 65 * 4, kind, s));
        RVS_CUDA(cudaMemcpyAsync(out_value + off, n->values, B * 4, kind, s));
        if (host) RVS_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int rvs_engine_predict(rvs_engine* h, const uint64_t* black, const uint64_t* white, const uint8_t* side, int64_t n_pos,
                       float* out_logits, float* out_value, int mem, void* stream) {
    if (!out_logits) return fail(-1, "rvs_engine_predict: bad arguments");
    return predict_impl(h, black, white, side, n_pos, out_logits, nullptr, out_value, mem, stream, "rvs_engine_predict");
}

int rvs_engine_predict_probs(rvs_engine* h, const uint64_t* black, const uint64_t* white, const uint8_t* side, int64_t n_pos,
                             float* out_probs, float* out_value, int mem, void* stream) {
    if (!out_probs) return fail(-1, "rvs_engine_predict_probs: bad arguments");
    return predict_impl(h, black, white, side, n_pos, nullptr, out_probs, out_value, mem, stream, "rvs_engine_predict_probs");
}

}  // extern "C"
