// rvs_engine.cuh -- host-side engine object shared by rvs_engine.cu (tree kernels + C ABI) and
// rvs_net.cu (K4 network).  All pointers are device memory owned by the engine.
#pragma once
#include "rvs_common.cuh"
#include "rvs_tree.cuh"
#include "rvs_treeg.cuh"
#include "rvs_noise.cuh"

namespace rvs {

enum : int { ST_SIMS = 0, ST_EVALS, ST_STEPS, ST_FINISHED, ST_SAMPLES, ST_OVERFLOW, ST_DROPPED, ST_STALLED, ST_BYTES, ST_NODES, ST_NNEVALS, ST_BADPOS, ST_COUNT };

struct NetState;  // rvs_net.cu

// POD view passed to kernels by value
struct EngineView {
    int G, cap, kmax;
    float c_puct;
    float noise_eps;     // Dirichlet root noise (rvs_noise.cuh); 0 = off
    double noise_alpha;
    uint64_t seed;
    int mode;             // RVS_MODE_REF (reference wave semantics) / RVS_MODE_FAST (virtual-loss PUCT, rvs_tree.cuh)
    uint64_t game_limit;  // self-play: a slot restarts only while its next game id < game_limit (0 = unlimited)
    // game state per slot
    uint64_t* black; uint64_t* white; uint8_t* side; uint8_t* flags;
    uint64_t* game_id; int* ply; uint8_t* live; uint8_t* finished;
    // trees
    int4* hot; int4* cold; int* n_nodes;
    int4* brd;   // wave-1 group kernels with a built-in evaluator: stored positions of the visited nodes (rvs_treeg.cuh); else 1 row
    int* order;  // slots sorted by game phase (disc count): the four games of a warp have similar rollout lengths
    // wave scratch [G*kmax]
    int* w_node; int* w_plen; int* w_path; uint64_t* w_black; uint64_t* w_white; uint16_t* w_sf;
    uint64_t* w_lm; float* w_val; uint64_t* w_sides;
    // per-slot sample staging [G*64]
    uint64_t* s_black; uint64_t* s_white; uint8_t* s_side; float* s_pi;  // s_pi [G*64*65]
    // completed-sample ring [ring_cap]
    int64_t ring_cap;
    uint64_t* r_black; uint64_t* r_white; uint8_t* r_side; int8_t* r_z; float* r_pi;
    unsigned long long* ring_count;
    unsigned long long* ply_counter;  // game-plies claimed in the current persistent self-play launch
    unsigned long long* stats;  // [kStatStripes][kStatStride] counters, striped (stat_at): summed by rvs_engine_stats_get
};

// Counters are striped over kStatStripes 128-byte lines chosen by the CTA index: thousands of warps add to them at
// the end of every tree kernel, and same-address atomics serialise in one L2 slice (ncu on the NN step kernel: ~15 %
// of its stall samples sat on these six atomics when all of them hit one 32-byte sector).
constexpr int kStatStripes = 64;
constexpr int kStatStride = 16;  // >= ST_COUNT, 128 bytes
static_assert(ST_COUNT <= kStatStride, "stat stripe too small");
__device__ __forceinline__ unsigned long long* stat_at(const EngineView& ev, int k) {
    return ev.stats + (size_t)(blockIdx.x & (kStatStripes - 1)) * kStatStride + k;
}

// Dirichlet noise into the priors of the root's children (rvs_noise.cuh), by ONE thread, right
// after the expansion of the root; children still have N == 0, so no cached score is stale
static __device__ __noinline__ void root_noise_apply(const EngineView& ev, int g, uint64_t game_id, uint64_t search_id) {
    int4* cold = ev.cold + (size_t)g * ev.cap;
    const int4 c = cold[0];
    const int nc = c.z & 0xFF, fc = c.y;
    if (nc == 0 || nc > 64) return;
    float eta[64];
    noise_dirichlet(ev.noise_alpha, nc, stream_seed(ev.seed, game_id, 0xD1000000ULL + search_id), eta);
    for (int i = 0; i < nc; ++i) {
        float* P = reinterpret_cast<float*>(&cold[fc + i]);
        *P = noise_mix(*P, eta[i], ev.noise_eps);
    }
}

__device__ __forceinline__ WaveScratch scratch_of(const EngineView& ev, int g) {
    const size_t o = (size_t)g * ev.kmax;
    return WaveScratch{ev.w_node + o, ev.w_plen + o, ev.w_path + o * kMaxPath, ev.w_black + o, ev.w_white + o,
                       ev.w_sf + o,   ev.w_lm + o,   ev.w_val + o,  ev.w_sides + o};
}

}  // namespace rvs

struct rvs_engine {
    rvs_engine_config cfg;
    rvs::EngineView v;
    int cur_k = 0;          // wave size of the last select (external path)
    int lanes_per_game = 0; // wave-1 kernels: 0 = choose by the number of games (rvs_engine_set_lanes_per_game)
    int net_graph = 0;      // RVS_OPT_NET_GRAPH
    int net_max_ctas = 0;   // RVS_OPT_NET_MAX_CTAS
    int net_tower = 1;      // RVS_OPT_NET_TOWER (persistent whole-network kernel where the shape allows it)
    int net_pipeline = 1;   // RVS_OPT_NET_PIPELINE (takes effect where the whole-network kernel runs: rvs_net_search_w1)
    uint64_t epoch = 0;     // set_positions calls since create / reset: game id of slot g = g + epoch * G
    unsigned long long* pinned_count = nullptr;  // pinned host word for the sample count of the synchronous drains
    int waves_done = 0;     // waves processed since begin_search (root noise goes in after the first)
    bool searching = false;
    float* ext_probs = nullptr;   // staging for host-side probs/values/planes of the external path
    float* ext_values = nullptr;
    float* ext_planes = nullptr;
    uint8_t* ext_valid = nullptr;
    int32_t* visits = nullptr;    // [G*65] scratch for root_visits with host output
    uint8_t* moves = nullptr;     // [G]
    void* io_stage = nullptr;     // staging for set/get positions, drain (grow-only)
    size_t io_stage_cap = 0;
    int64_t launches = 0;
    rvs::NetState* net = nullptr;
    void* allocs[64];
    int n_allocs = 0;
};

namespace rvs {
// restores the caller's current CUDA device when an entry point returns (a torchrun rank that called
// torch.cuda.set_device(r) must not find its device switched by a handle that lives elsewhere)
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    int enter(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev) {
            cudaError_t e = cudaSetDevice(dev);
            if (e != cudaSuccess) return fail(-100 - (int)e, "cudaSetDevice(%d) failed: %s", dev, cudaGetErrorString(e));
            switched = true;
        }
        return 0;
    }
    ~DeviceGuard() {
        if (switched && prev >= 0) cudaSetDevice(prev);
    }
};
}  // namespace rvs

// every engine entry point starts with this: validates the handle and makes its device current for the call
#define RVS_ENTER(h)                                             \
    if (!(h)) return ::rvs::fail(-1, "null engine handle");      \
    ::rvs::DeviceGuard _dg;                                      \
    if (int _rc = _dg.enter((h)->cfg.device)) return _rc

// K4 hooks implemented in rvs_net.cu
int rvs_net_search(rvs_engine* h, int32_t num_sims, int32_t wave, cudaStream_t s);
// rvs_engine_process with device probs / values addressed through the compaction map of rvs_net.cu
int rvs_engine_process_mapped(rvs_engine* h, const float* probs, const float* values, const int* inv, cudaStream_t s);
void rvs_net_destroy(rvs::NetState* n);
// one fused tree step (process pending leaf | select next | encode) of the wave-1 NN search for games [g0, g1);
// tiles_out (optional): bf16 input tiles of the tensor-core first layer, written by the same kernel
int rvs_engine_nn_step(rvs_engine* h, int g0, int g1, int flags, const float* probs, const float* values, int* rows,
                       uint64_t* bits_out, int* n_cur, int* n_next, void* tiles_out, cudaStream_t s, bool pdl = true);
