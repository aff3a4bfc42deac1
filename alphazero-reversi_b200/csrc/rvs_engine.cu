// rvs_engine.cu -- K2 kernels (lockstep batched MCTS), self-play ply / sample kernels, and
// the engine C ABI.  One warp per game everywhere; grids are G/4 CTAs of 128 threads.
#include "rvs_engine.cuh"

#include <stdlib.h>

#include <new>

namespace rvs {

constexpr int kWarpsPerBlock = 4;
constexpr int kBlock = 32 * kWarpsPerBlock;

__device__ __forceinline__ void flush_stats(const EngineView& ev, const TreeCtx& cx, unsigned long long lane_steps) {
    for (int o = 16; o; o >>= 1) lane_steps += __shfl_down_sync(kFull, lane_steps, o);
    if (cx.lane == 0) {
        atomicAdd(stat_at(ev, ST_SIMS), (unsigned long long)cx.sims);
        atomicAdd(stat_at(ev, ST_EVALS), (unsigned long long)cx.evals);
        atomicAdd(stat_at(ev, ST_STEPS), (unsigned long long)cx.steps + lane_steps);
        atomicAdd(stat_at(ev, ST_BYTES), (unsigned long long)cx.bytes);
        atomicAdd(stat_at(ev, ST_NODES), (unsigned long long)cx.created);
        if (cx.overflow) atomicAdd(stat_at(ev, ST_OVERFLOW), 1ULL);
    }
}

__device__ __forceinline__ void init_root(TreeCtx& cx, int side) {
    // MCTSNode(1.0, game.current_player, ...) (mcts.py:334-341)
    if (cx.lane == 0) {
        cx.hot[0] = make_int4(0, 0, 0, 0);
        cx.cold[0] = make_int4(__float_as_int(1.0f), -1, (255 << 8) | (side << 16), 0);
    }
    cx.n_nodes = 1;
    __syncwarp();
}

// evaluator for the fused kernels: fills ws.lm / ws.val.
// Small waves (k <= kCoopWave): the warp evaluates the leaves one after another with the
// direction-sliced board ops (every lane busy on one rollout).  Large waves: one leaf per lane.
constexpr int kCoopWave = 4;

template <int RULES, int EVAL>
__device__ __forceinline__ void eval_wave(const EngineView& ev, TreeCtx& cx, const WaveScratch& ws, int k, uint64_t game_id,
                                          uint64_t search_id, int sim_base, unsigned long long& lane_steps) {
    if (EVAL == RVS_EVAL_ROLLOUT && k <= kCoopWave) {
        for (int j = 0; j < k; ++j) {
            if (ws.node[j] < 0) continue;  // warp-uniform
            const uint16_t sf = ws.sf[j];
            CoopBoard b = coop_load(cx.dir, Board{ws.black[j], ws.white[j], (uint8_t)(sf & 0xFF), (uint8_t)(sf >> 8)});
            const uint64_t lm = coop_legal(cx.dir, b);
            float v = 0.0f;
            if (lm) {
                const int leaf_side = b.side;
                const uint64_t st = stream_seed(ev.seed, game_id, (search_id << 16) | (uint64_t)(sim_base + j));
                cx.steps += (unsigned)coop_random_playout(cx.dir, b, lm, st, cx.lane);
                const int w = (b.flags & F_WIN_MASK) >> F_WIN_SHIFT;
                v = (!(b.flags & F_OVER) || w == 0) ? 0.0f : (w == leaf_side ? 1.0f : -1.0f);
            }
            if (cx.lane == 0) { ws.lm[j] = lm; ws.val[j] = v; }
        }
        __syncwarp();
        return;
    }
    for (int j = cx.lane; j < k; j += 32) {
        if (ws.node[j] < 0) continue;
        const uint16_t sf = ws.sf[j];
        Board b{ws.black[j], ws.white[j], (uint8_t)(sf & 0xFF), (uint8_t)(sf >> 8)};
        const uint64_t lm = board_legal<RULES>(b);
        float v = 0.0f;
        if (lm) {
            if (EVAL == RVS_EVAL_E0) {
                const int nb = popc64(b.black), nw = popc64(b.white);
                const int d = b.side == 1 ? nb - nw : nw - nb;
                v = __fdiv_rn((float)d, 64.0f);
            } else {
                const int leaf_side = b.side;
                const uint64_t st = stream_seed(ev.seed, game_id, (search_id << 16) | (uint64_t)(sim_base + j));
                lane_steps += (unsigned long long)random_playout<RULES>(b, st);
                const int w = winner_of(b);
                v = (!is_over(b) || w == 0) ? 0.0f : (w == leaf_side ? 1.0f : -1.0f);
            }
        }
        ws.lm[j] = lm;
        ws.val[j] = v;
    }
    __syncwarp();
}

// MCTS.search (mcts.py:322-407) with a built-in evaluator, whole search in one launch.
template <int RULES, int EVAL>
__global__ void __launch_bounds__(kBlock) search_fused_kernel(EngineView ev, int S, int K) {
    const int g = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (g >= ev.G) return;
    TreeCtx cx{ev.hot + (size_t)g * ev.cap, ev.cold + (size_t)g * ev.cap, ev.cap, 1, ev.c_puct, (int)(threadIdx.x & 31), 0, 0, 0, 0, 0, 0, make_dir<RULES>(threadIdx.x & 7)};
    const Board root{ev.black[g], ev.white[g], ev.side[g], ev.flags[g]};
    const uint64_t game_id = ev.game_id[g];
    const uint64_t search_id = (uint64_t)ev.ply[g];
    const WaveScratch ws = scratch_of(ev, g);
    unsigned long long lane_steps = 0;
    init_root(cx, root.side);
    const bool fast = ev.mode == RVS_MODE_FAST;
    for (int start = 0; start < S;) {
        // FAST: the first wave is a single simulation (it expands the root)
        const int k = (fast && start == 0) ? 1 : ((S - start) < K ? (S - start) : K);
        if (fast) select_wave_fast(cx, root, ws, k);
        else select_wave(cx, root, ws, k);
        eval_wave<RULES, EVAL>(ev, cx, ws, k, game_id, search_id, start, lane_steps);
        if (fast) process_wave_fast(cx, ws, k, [](int, int) { return 1.0f / 65.0f; });
        else process_wave(cx, ws, k, [](int, int) { return 1.0f / 65.0f; });
        if (start == 0 && ev.noise_eps > 0.0f) {
            __syncwarp();
            if (cx.lane == 0) root_noise_apply(ev, g, game_id, search_id);
            __syncwarp();
        }
        start += k;
    }
    if (cx.lane == 0) ev.n_nodes[g] = cx.n_nodes;
    flush_stats(ev, cx, lane_steps);
}

__global__ void __launch_bounds__(kBlock) begin_search_kernel(EngineView ev) {
    const int g = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (g >= ev.G) return;
    TreeCtx cx{ev.hot + (size_t)g * ev.cap, ev.cold + (size_t)g * ev.cap, ev.cap, 1, ev.c_puct, (int)(threadIdx.x & 31), 0, 0, 0, 0, 0, 0, DirLane{}};
    init_root(cx, ev.side[g]);
    if (cx.lane == 0) ev.n_nodes[g] = 1;
}

// MCTS._traverse for k simulations per game (external / NN evaluator path)
template <int RULES>
__global__ void __launch_bounds__(kBlock) select_kernel(EngineView ev, int k) {
    const int g = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (g >= ev.G) return;
    TreeCtx cx{ev.hot + (size_t)g * ev.cap, ev.cold + (size_t)g * ev.cap, ev.cap, ev.n_nodes[g], ev.c_puct, (int)(threadIdx.x & 31), 0, 0, 0, 0, 0, 0, make_dir<RULES>(threadIdx.x & 7)};
    const Board root{ev.black[g], ev.white[g], ev.side[g], ev.flags[g]};
    const WaveScratch ws = scratch_of(ev, g);
    if (ev.mode == RVS_MODE_FAST) select_wave_fast(cx, root, ws, k);
    else select_wave(cx, root, ws, k);
    // legal masks of the leaves (reused by leaf_planes and process)
    for (int j = cx.lane; j < k; j += 32) {
        if (ws.node[j] < 0) { ws.lm[j] = 0; continue; }
        const uint16_t sf = ws.sf[j];
        const Board b{ws.black[j], ws.white[j], (uint8_t)(sf & 0xFF), (uint8_t)(sf >> 8)};
        ws.lm[j] = board_legal<RULES>(b);
    }
    flush_stats(ev, cx, 0);
}

// MCTS._process_batch with caller-supplied softmax outputs (mcts.py:596-623).
// probs [G*k,65], values [G*k], slot = g*k + j.
// `inv` (optional): slot -> row of probs / values (compacted batch of rvs_net.cu); -1 = no evaluation,
// <= -2 = same row as slot -2-inv of this game's wave
__global__ void __launch_bounds__(kBlock) process_probs_kernel(EngineView ev, int k, const float* __restrict__ probs,
                                                                const float* __restrict__ values, int first_wave,
                                                                const int* __restrict__ inv) {
    const int g = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (g >= ev.G) return;
    TreeCtx cx{ev.hot + (size_t)g * ev.cap, ev.cold + (size_t)g * ev.cap, ev.cap, ev.n_nodes[g], ev.c_puct, (int)(threadIdx.x & 31), 0, 0, 0, 0, 0, 0, DirLane{}};
    const WaveScratch ws = scratch_of(ev, g);
    const size_t slot0 = (size_t)g * k;
    auto row_of = [&](int j) -> long long {
        if (!inv) return (long long)(slot0 + j);
        int c = inv[slot0 + j];
        if (c <= -2) c = inv[slot0 + (-2 - c)];
        return (long long)c;
    };
    for (int j = cx.lane; j < k; j += 32) {
        const long long r = row_of(j);
        ws.val[j] = r >= 0 ? values[r] : 0.0f;
    }
    __syncwarp();
    if (ev.mode == RVS_MODE_FAST) process_wave_fast(cx, ws, k, [&](int j, int sq) { return probs[row_of(j) * 65 + sq]; });
    else process_wave(cx, ws, k, [&](int j, int sq) { return probs[row_of(j) * 65 + sq]; });
    if (first_wave && ev.noise_eps > 0.0f) {
        __syncwarp();
        if (cx.lane == 0) root_noise_apply(ev, g, ev.game_id[g], (uint64_t)ev.ply[g]);
        __syncwarp();
    }
    if (cx.lane == 0) ev.n_nodes[g] = cx.n_nodes;
    flush_stats(ev, cx, 0);
}

// One wave-1 step of the NN-evaluated search for the games [g0, g1), everything the tree does between two network
// passes in ONE kernel: (flags & 1) MCTS._process_batch of the pending leaf with the network's outputs (row `rows[g]` of
// probs / values; mcts.py:596-623), (flags & 4) root noise after the root expansion, then (flags & 2) the next
// MCTS._traverse (mcts.py:409-444), the leaf's legal mask and K3 with compaction: a leaf that needs an evaluation
// appends its three bit planes to the half-batch (one atomic) and remembers its row.  n_next is the counter of the
// following wave, zeroed here because nothing else uses it while this kernel runs.
template <int RULES>
__global__ void __launch_bounds__(kBlock, 7) nn_step_kernel(EngineView ev, int g0, int g1, int flags, const float* __restrict__ probs,
                                                          const float* __restrict__ values, int* __restrict__ rows,
                                                          uint64_t* __restrict__ bits_out, int* __restrict__ n_cur,
                                                          int* __restrict__ n_next, uint4* __restrict__ tiles_out) {
    pdl_trigger();    // the first layer's CTAs may load their weights while this kernel runs
    pdl_grid_wait();  // probs / values come from the heads kernel launched just before
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_next) *n_next = 0;
    const int g = g0 + blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (g >= g1) return;
    TreeCtx cx{ev.hot + (size_t)g * ev.cap, ev.cold + (size_t)g * ev.cap, ev.cap, ev.n_nodes[g], ev.c_puct, (int)(threadIdx.x & 31), 0, 0, 0, 0, 0, 0, make_dir<RULES>(threadIdx.x & 7)};
    const WaveScratch ws = scratch_of(ev, g);
    const bool fast = ev.mode == RVS_MODE_FAST;
    if (flags & 1) {
        const int row = rows[g];
        if (cx.lane == 0) ws.val[0] = row >= 0 ? values[row] : 0.0f;
        __syncwarp();
        auto prior = [&](int, int sq) { return probs[(size_t)row * 65 + sq]; };  // only called for evaluated leaves (row >= 0)
        if (fast) process_wave_fast(cx, ws, 1, prior);
        else process_wave(cx, ws, 1, prior);
        if ((flags & 4) && ev.noise_eps > 0.0f) {
            __syncwarp();
            if (cx.lane == 0) root_noise_apply(ev, g, ev.game_id[g], (uint64_t)ev.ply[g]);
            __syncwarp();
        }
    }
    if (flags & 2) {
        const Board root{ev.black[g], ev.white[g], ev.side[g], ev.flags[g]};
        // lane 0 ends up with: the leaf's legal mask, its row in the batch (-1: nothing to evaluate), its position
        uint64_t lm = 0;
        int row = -1;
        Board lb{0, 0, 1, 0};
        if (fast) {
            select_wave_fast(cx, root, ws, 1);
            if (cx.lane == 0 && ws.node[0] >= 0) {
                const uint16_t sf = ws.sf[0];
                lb = Board{ws.black[0], ws.white[0], (uint8_t)(sf & 0xFF), (uint8_t)(sf >> 8)};
                lm = board_legal<RULES>(lb);
            }
        } else {
            // select_wave(k = 1) with the leaf kept in registers: its legal mask comes from the direction-sliced
            // warp-cooperative scan instead of one lane re-reading the wave scratch and scanning 8 directions alone
            CoopBoard b = coop_load(cx.dir, root);
            int p0, p1, plen, vlf;
            const int node = select_one(cx, b, p0, p1, plen, vlf);
            ++cx.sims;
            if (vlf & kTerminal) {  // warp-uniform (mcts.py:364-366)
                backup_path(cx, p0, p1, plen, term_value_of(vlf));
                if (cx.lane == 0) ws.node[0] = -1;
            } else {
                lb = coop_store(cx.dir, b);
                store_leaf(ws, 0, cx.lane, node, p0, p1, plen, lb);
                lm = coop_legal(cx.dir, b);
            }
        }
        const bool blk = lb.side == 1;
        const uint64_t own = blk ? lb.black : lb.white, opp = blk ? lb.white : lb.black;
        if (cx.lane == 0) {
            if (lm) {
                row = atomicAdd(n_cur, 1);
                bits_out[(size_t)row * 3] = own;
                bits_out[(size_t)row * 3 + 1] = opp;
                bits_out[(size_t)row * 3 + 2] = lm;
            }
            ws.lm[0] = lm;
            rows[g] = row;
        }
        if (tiles_out) {
            // K3 fused: the warp writes the leaf's three input planes (+ the two constant-one bias channels) straight
            // into the bf16 input tile of the tensor-core first layer -- the same 32 bytes per pixel that
            // planes_tiles_kernel (rvs_net.cu) derives from the bit planes, without the extra kernel in every wave
            row = __shfl_sync(kFull, row, 0);
            if (row >= 0) {
                const uint64_t P = __shfl_sync(kFull, own, 0), O = __shfl_sync(kFull, opp, 0), M = __shfl_sync(kFull, lm, 0);
                const uint32_t one = 0x3F80u;
#pragma unroll
                for (int hp = 0; hp < 2; ++hp) {
                    const int sq = cx.lane + 32 * hp, y = sq >> 3, x = sq & 7;
                    uint4* o = tiles_out + ((size_t)(row >> 1) * 128 + y * 16 + (row & 1) * 8 + x) * 8;
                    o[0] = make_uint4((((P >> sq) & 1) ? one : 0u) | ((((O >> sq) & 1) ? one : 0u) << 16),
                                      (((M >> sq) & 1) ? one : 0u) | (one << 16), one, 0u);
                    o[1] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
    }
    if (cx.lane == 0) ev.n_nodes[g] = cx.n_nodes;
    flush_stats(ev, cx, 0);
}

// canonical planes of the selected leaves (game.py:131-162), 48 threads x float4 per slot
__global__ void __launch_bounds__(256) leaf_planes_kernel(EngineView ev, int k, float4* __restrict__ out,
                                                           uint8_t* __restrict__ valid) {
    const int64_t total = (int64_t)ev.G * k * 48;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t slot = t / 48;
        const int r = (int)(t - slot * 48);
        const int g = (int)(slot / k), j = (int)(slot - (int64_t)g * k);
        const size_t o = (size_t)g * ev.kmax + j;
        const uint64_t lm = ev.w_node[o] < 0 ? 0ULL : ev.w_lm[o];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lm) {
            const bool blk = (ev.w_sf[o] & 0xFF) == 1;
            const uint64_t P = blk ? ev.w_black[o] : ev.w_white[o], O = blk ? ev.w_white[o] : ev.w_black[o];
            const int plane = r >> 4, q = r & 15;
            const uint64_t src = plane == 0 ? P : (plane == 1 ? O : lm);
            const uint32_t nib = (uint32_t)(src >> (4 * q)) & 15u;
            v = make_float4((float)(nib & 1), (float)((nib >> 1) & 1), (float)((nib >> 2) & 1), (float)((nib >> 3) & 1));
        }
        out[t] = v;
        if (valid && r == 0) valid[slot] = lm ? 1 : 0;
    }
}

// dict MCTS.search returns (mcts.py:406-407) as [G,65] int32
__global__ void __launch_bounds__(kBlock) root_visits_kernel(EngineView ev, int32_t* __restrict__ out, int n) {
    const int g = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (g >= n) return;
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < 65; i += 32) out[(size_t)g * 65 + i] = 0;
    __syncwarp();
    const int4* hot = ev.hot + (size_t)g * ev.cap;
    const int4* cold = ev.cold + (size_t)g * ev.cap;
    const int4 c = cold[0];
    const int nc = c.z & 0xFF;
    for (int i = lane; i < nc; i += 32) {
        const int mv = (cold[c.y + i].z >> 8) & 0xFF;
        out[(size_t)g * 65 + mv] = hot[c.y + i].x;
    }
}

// numpy pairwise sum of 65 doubles (see oracle/rvs_oracle.c np_sum65)
__device__ inline double np_sum65(const double* a) {
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    for (int i = 8; i < 64; i += 8)
        for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], a[i + j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    return __dadd_rn(res, a[64]);
}

// One self-play ply of slot g by ONE thread (self_play.py:80-101, mcts.py:642-694): pi from the
// root visit counts, move choice, sample record, make_move.  Returns the square played or 255.
template <int RULES>
__device__ __noinline__ int play_game(const EngineView& ev, int g, float temperature) {
    Board b{ev.black[g], ev.white[g], ev.side[g], ev.flags[g]};
    if (!ev.live[g] || is_over(b)) return 255;
    const int4* hot = ev.hot + (size_t)g * ev.cap;
    const int4* cold = ev.cold + (size_t)g * ev.cap;
    double pi[65];
    for (int i = 0; i < 65; ++i) pi[i] = 0.0;
    const int4 c = cold[0];
    const int nc = c.z & 0xFF;
    long long total = 0;
    for (int i = 0; i < nc; ++i) total += hot[c.y + i].x;
    bool all_zero = true;
    if (total > 0) {
        for (int i = 0; i < nc; ++i) {
            const int n = hot[c.y + i].x;
            const int mv = (cold[c.y + i].z >> 8) & 0xFF;
            pi[mv] = __ddiv_rn((double)n, (double)total);  // count / total_visits (mcts.py:670)
            if (n) all_zero = false;
        }
    }
    if (temperature > 0.0f && !all_zero) {  // mcts.py:673-676
        const double e = __ddiv_rn(1.0, (double)temperature);
        if (e != 1.0)
            for (int i = 0; i < 65; ++i) pi[i] = pow(pi[i], e);
        const double s = np_sum65(pi);
        for (int i = 0; i < 65; ++i)
            if (pi[i] != 0.0) pi[i] = __ddiv_rn(pi[i], s);  // 0 / s = 0: skip the ~55 illegal squares (their division takes the slow path)
    }
    int mv = 0;
    const int ply = ev.ply[g];
    if (temperature == 0.0f || all_zero) {  // np.argmax: first maximum (mcts.py:679-681)
        double best = pi[0];
        for (int i = 1; i < 65; ++i) if (pi[i] > best) { best = pi[i]; mv = i; }
    } else {  // np.random.choice: cumsum, normalise by cdf[-1], searchsorted(side='right')
        uint64_t st = stream_seed(ev.seed, ev.game_id[g], 0x80000000ULL + (uint64_t)ply);
        const double u = (double)(rng_next(st) >> 11) * (1.0 / 9007199254740992.0);
        double acc = 0.0;
        for (int i = 0; i < 65; ++i) acc = __dadd_rn(acc, pi[i]);
        double run = 0.0;
        mv = 64;
        for (int i = 0; i < 65; ++i) {
            if (pi[i] == 0.0) continue;  // run does not change, and run / acc > u was false for the previous entry (0 > u at the start)
            run = __dadd_rn(run, pi[i]);
            if (__ddiv_rn(run, acc) > u) { mv = i; break; }
        }
    }
    // record the sample BEFORE the move (self_play.py:87-94)
    if (ply < 64) {
        const size_t o = (size_t)g * 64 + ply;
        ev.s_black[o] = b.black; ev.s_white[o] = b.white; ev.s_side[o] = b.side;
        float* dst = ev.s_pi + o * 65;
        for (int i = 0; i < 65; ++i) dst[i] = (float)pi[i];
    }
    uint64_t nl;
    if (!try_move<RULES>(b, mv, nl)) {
        // the reference would spin forever here (SURVEY.md 8(a) A7 hazard); park the slot instead
        ev.live[g] = 0;
        atomicAdd(stat_at(ev, ST_STALLED), 1ULL);
        return 255;
    }
    atomicAdd(stat_at(ev, ST_STEPS), 1ULL);
    ev.black[g] = b.black; ev.white[g] = b.white; ev.side[g] = b.side; ev.flags[g] = b.flags;
    ev.ply[g] = ply + 1;
    if (is_over(b)) ev.finished[g] = 1;
    return mv;
}

template <int RULES>
__global__ void __launch_bounds__(128) play_kernel(EngineView ev, float temperature, uint8_t* __restrict__ out_moves) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ev.G) return;
    const int mv = play_game<RULES>(ev, g, temperature);
    if (out_moves) out_moves[g] = (uint8_t)mv;
}

// Finished game of slot g, by its warp: z back-fill (self_play.py:117-126), samples to the ring,
// recycle or park the slot.
// `lane` counts inside the owning group of `width` lanes (a whole warp, or 8 lanes in the 4-games-per-warp kernels)
__device__ __forceinline__ void finalize_game(const EngineView& ev, int g, int lane, int recycle, unsigned mask = kFull,
                                              int width = 32) {
    if (!ev.finished[g]) return;
    const int n = ev.ply[g] < 64 ? ev.ply[g] : 64;
    const int w = (ev.flags[g] & F_WIN_MASK) >> F_WIN_SHIFT;
    __syncwarp(mask);
    unsigned long long at = 0;
    if (lane == 0) at = atomicAdd(ev.ring_count, (unsigned long long)n);
    at = __shfl_sync(mask, at, 0, width);
    int stored = 0;
    for (int p = 0; p < n; ++p) {
        const unsigned long long dst = at + p;
        if (dst >= (unsigned long long)ev.ring_cap) break;
        const size_t o = (size_t)g * 64 + p;
        if (lane == 0) {
            const int s = ev.s_side[o];
            ev.r_black[dst] = ev.s_black[o]; ev.r_white[dst] = ev.s_white[o]; ev.r_side[dst] = (uint8_t)s;
            ev.r_z[dst] = (int8_t)(w == 0 ? 0 : (s == w ? 1 : -1));
        }
        for (int i = lane; i < 65; i += width) ev.r_pi[dst * 65 + i] = ev.s_pi[o * 65 + i];
        ++stored;
    }
    if (lane == 0) {
        atomicAdd(stat_at(ev, ST_FINISHED), 1ULL);
        atomicAdd(stat_at(ev, ST_SAMPLES), (unsigned long long)stored);
        if (stored < n) atomicAdd(stat_at(ev, ST_DROPPED), (unsigned long long)(n - stored));
        ev.finished[g] = 0;
        if (recycle && (ev.game_limit == 0 || ev.game_id[g] + (uint64_t)ev.G < ev.game_limit)) {
            ev.black[g] = kStartBlack; ev.white[g] = kStartWhite; ev.side[g] = 1; ev.flags[g] = 0;
            ev.ply[g] = 0;
            ev.game_id[g] += (uint64_t)ev.G;
        } else {
            ev.live[g] = 0;
        }
    }
    __syncwarp(mask);
}

__global__ void __launch_bounds__(kBlock) finalize_kernel(EngineView ev, int recycle) {
    const int g = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (g >= ev.G) return;
    finalize_game(ev, g, threadIdx.x & 31, recycle);
}

// ---- wave 1, several games per warp (rvs_treeg.cuh) -------------------------------------------
constexpr int kBlockG = 32;  // one warp per CTA: the finest grain for the block scheduler

template <int LPG>
__device__ __forceinline__ void flush_stats_g(const EngineView& ev, const TreeCtxG<LPG>& cx, bool act) {
    if (act && cx.g.lane == 0) {
        atomicAdd(stat_at(ev, ST_SIMS), (unsigned long long)cx.sims);
        atomicAdd(stat_at(ev, ST_EVALS), (unsigned long long)cx.evals);
        atomicAdd(stat_at(ev, ST_STEPS), (unsigned long long)cx.steps);
        atomicAdd(stat_at(ev, ST_BYTES), (unsigned long long)cx.bytes);
        atomicAdd(stat_at(ev, ST_NODES), (unsigned long long)cx.created);
        if (cx.overflow) atomicAdd(stat_at(ev, ST_OVERFLOW), 1ULL);
    }
}

template <int LPG>
__device__ __forceinline__ TreeCtxG<LPG> make_ctx_g(const EngineView& ev, int g, const Grp<LPG>& grp) {
    return TreeCtxG<LPG>{ev.hot + (size_t)g * ev.cap, ev.cold + (size_t)g * ev.cap, ev.brd + (size_t)g * ev.cap, ev.cap, 1, ev.c_puct, 0, 0, 0, 0, 0, 0, grp, nullptr, 0, 0u};
}

template <int LPG>
__device__ __forceinline__ void init_root_g(TreeCtxG<LPG>& cx, int side, bool act) {  // mcts.py:334-341
    if (act && cx.g.lane == 0) {
        cx.hot[0] = make_int4(0, 0, 0, 0);
        cx.cold[0] = make_int4(__float_as_int(1.0f), -1, (255 << 8) | (side << 16), 0);
    }
    cx.n_nodes = 1;
    __syncwarp();
}

// finalize_game() for the LPG-lane groups of a converged warp
template <int LPG>
__device__ __forceinline__ void finalize_game_g(const EngineView& ev, int g, int lane, int recycle, bool act) {
    const bool fin = act && ev.finished[g];
    const int n = fin ? (ev.ply[g] < 64 ? ev.ply[g] : 64) : 0;
    const int w = fin ? (ev.flags[g] & F_WIN_MASK) >> F_WIN_SHIFT : 0;
    __syncwarp();
    unsigned long long at = 0;
    if (fin && lane == 0) at = atomicAdd(ev.ring_count, (unsigned long long)n);
    at = __shfl_sync(kFull, at, 0, LPG);
    int stored = 0;
    for (int p = 0; p < n; ++p) {
        const unsigned long long dst = at + p;
        if (dst >= (unsigned long long)ev.ring_cap) break;
        const size_t o = (size_t)g * 64 + p;
        if (lane == 0) {
            const int s = ev.s_side[o];
            ev.r_black[dst] = ev.s_black[o]; ev.r_white[dst] = ev.s_white[o]; ev.r_side[dst] = (uint8_t)s;
            ev.r_z[dst] = (int8_t)(w == 0 ? 0 : (s == w ? 1 : -1));
        }
        for (int i = lane; i < 65; i += LPG) ev.r_pi[dst * 65 + i] = ev.s_pi[o * 65 + i];
        ++stored;
    }
    if (fin && lane == 0) {
        atomicAdd(stat_at(ev, ST_FINISHED), 1ULL);
        atomicAdd(stat_at(ev, ST_SAMPLES), (unsigned long long)stored);
        if (stored < n) atomicAdd(stat_at(ev, ST_DROPPED), (unsigned long long)(n - stored));
        ev.finished[g] = 0;
        if (recycle && (ev.game_limit == 0 || ev.game_id[g] + (uint64_t)ev.G < ev.game_limit)) {
            ev.black[g] = kStartBlack; ev.white[g] = kStartWhite; ev.side[g] = 1; ev.flags[g] = 0;
            ev.ply[g] = 0;
            ev.game_id[g] += (uint64_t)ev.G;
        } else {
            ev.live[g] = 0;
        }
    }
    __syncwarp();
}

// Slots ordered by game phase (disc count; parked / finished slots last): the four games that share
// a warp then have rollouts of similar length, so the groups of a warp stay converged.
__global__ void __launch_bounds__(1024) phase_order_kernel(EngineView ev) {
    __shared__ int hist[66];
    __shared__ int start[66];
    for (int i = threadIdx.x; i < 66; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int g = threadIdx.x; g < ev.G; g += blockDim.x) {
        const int key = (ev.live[g] && !(ev.flags[g] & F_OVER)) ? popc64(ev.black[g] | ev.white[g]) : 65;
        atomicAdd(&hist[key], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = 0;
        for (int k = 0; k < 66; ++k) { start[k] = a; a += hist[k]; }
    }
    __syncthreads();
    for (int g = threadIdx.x; g < ev.G; g += blockDim.x) {
        const int key = (ev.live[g] && !(ev.flags[g] & F_OVER)) ? popc64(ev.black[g] | ev.white[g]) : 65;
        ev.order[atomicAdd(&start[key], 1)] = g;
    }
}

// MCTS.search with batch_size 1 and a built-in evaluator, an LPG-lane group per game: every simulation stays
// in registers / shared memory (path, position, rollout) except for the tree rows.
// The warp stays converged (rvs_treeg.cuh); groups without a game are predicated off.  A group
// owns the slots gidx, gidx + n_groups, ... (phase order), so any G runs on a resident grid.
constexpr int kMaxWarpsG = kNumSMs * 16;  // 16 one-warp CTAs per SM (__launch_bounds__(32, 16): <= 128 registers)

// MINB = resident one-warp CTAs per SM the launch is compiled for: 16 (<= 128 registers) when the grid needs them all,
// 8 (<= 255 registers: no spills in the simulation loop) when the games fit in 148 x 8 warps (e.g. 4096 games x 8 lanes)
template <int RULES, int EVAL, int LPG, int MINB>
__global__ void __launch_bounds__(kBlockG, MINB) search_k1g_kernel(EngineView ev, int S) {
    constexpr int GPB = kBlockG / LPG;
    __shared__ __align__(16) uint8_t lut[LutCfg<LPG>::kBytes];
    __shared__ int spath[GPB][kMaxPath + 1];
    __shared__ __align__(16) unsigned sgather[LPG == 8 ? kGatherWords : 4];  // 3 x 64-byte exchange buffers per 8-lane group (gather_off)
    __shared__ __align__(16) uint64_t srays[RayCfg<LPG>::kWords64];  // flip-ray table (ray_init)
    lut_init<LPG>(lut, threadIdx.x, kBlockG);
    ray_init<RULES, LPG>(srays, threadIdx.x, kBlockG);
    __syncthreads();
    const int n_groups = gridDim.x * GPB;
    const Grp<LPG> grp = make_grp<RULES, LPG>(threadIdx.x & 31, lut, spath[threadIdx.x / LPG], LPG == 8 ? sgather : nullptr, srays);
    __shared__ int4 sstage[GPB][2 * (1 + StageCfg<LPG>::kRows) + 2];  // hot node rows of each group's search
    for (int slot = blockIdx.x * GPB + (int)threadIdx.x / LPG; __any_sync(kFull, slot < ev.G); slot += n_groups) {
        const bool act = slot < ev.G;
        const int g = act ? ev.order[slot] : 0;
        TreeCtxG<LPG> cx = make_ctx_g(ev, g, grp);
        cx.stage = sstage[threadIdx.x / LPG];
        const Board root{ev.black[g], ev.white[g], ev.side[g], ev.flags[g]};
        const uint64_t game_id = ev.game_id[g];
        const uint64_t search_id = (uint64_t)ev.ply[g];
        init_root_g(cx, root.side, act);
        const GBoard root_g = gboard_load(grp, root);
        for (int sim = 0; sim < S; ++sim) {
            const uint64_t st = EVAL == RVS_EVAL_ROLLOUT ? stream_seed(ev.seed, game_id, (search_id << 16) | (uint64_t)sim) : 0ULL;
            simulate_one_g<EVAL>(cx, root_g, st, act);
            if (sim == 0) {
                if (ev.noise_eps > 0.0f) {
                    if (act && grp.lane == 0) root_noise_apply(ev, g, game_id, search_id);
                    __syncwarp();
                }
                stage_root(cx, act);  // the root was expanded by simulation 0: its rows move to shared memory
            }
        }
        unstage_root(cx);
        if (act && grp.lane == 0) ev.n_nodes[g] = cx.n_nodes;
        flush_stats_g(ev, cx, act);
    }
}

// move choice + sample record + make_move + game end of one ply, out of line so that its registers (f64 pi
// arithmetic) do not weigh on the allocation of the simulation loop
template <int RULES, int LPG>
__device__ __noinline__ void end_of_ply(const EngineView& ev, int g, int lane, float temperature, int recycle, bool alive) {
    __syncwarp();
    if (alive && lane == 0) play_game<RULES>(ev, g, temperature);
    __syncwarp();
    finalize_game_g<LPG>(ev, g, lane, recycle, alive);
}

// Persistent self-play (SelfPlay.generate_games, self_play.py:66-131, with MCTS batch_size 1), an LPG-lane
// group per game, work-conserving: every group
// keeps playing plies of its slots, round-robin, until the launch-wide budget of game-plies is used up
template <int RULES, int EVAL, int LPG, int MINB>
__global__ void __launch_bounds__(kBlockG, MINB) selfplay_k1g_kernel(EngineView ev, int S, float temperature,
                                                                   unsigned long long budget, int recycle) {
    constexpr int GPB = kBlockG / LPG;
    __shared__ __align__(16) uint8_t lut[LutCfg<LPG>::kBytes];
    __shared__ int spath[GPB][kMaxPath + 1];
    __shared__ __align__(16) unsigned sgather[LPG == 8 ? kGatherWords : 4];  // 3 x 64-byte exchange buffers per 8-lane group (gather_off)
    __shared__ __align__(16) uint64_t srays[RayCfg<LPG>::kWords64];  // flip-ray table (ray_init)
    lut_init<LPG>(lut, threadIdx.x, kBlockG);
    ray_init<RULES, LPG>(srays, threadIdx.x, kBlockG);
    __syncthreads();
    const int n_groups = gridDim.x * GPB;
    const int slot0 = blockIdx.x * GPB + (int)threadIdx.x / LPG;
    const int n_mine = slot0 < ev.G ? (ev.G - slot0 + n_groups - 1) / n_groups : 0;  // slots of this group
    const Grp<LPG> grp = make_grp<RULES, LPG>(threadIdx.x & 31, lut, spath[threadIdx.x / LPG], LPG == 8 ? sgather : nullptr, srays);
    __shared__ int4 sstage[GPB][2 * (1 + StageCfg<LPG>::kRows) + 2];  // hot node rows of each group's search
    bool quit = n_mine == 0;
    int k = 0, idle = 0;  // current slot of the round-robin; consecutive slots found parked / finished
    while (true) {
        if (!__any_sync(kFull, !quit)) break;
        const int g = quit ? 0 : ev.order[slot0 + k * n_groups];
        bool alive = !quit && ev.live[g];
        const Board root{ev.black[g], ev.white[g], ev.side[g], ev.flags[g]};
        alive = alive && !is_over(root);
        unsigned long long t = 0;
        if (alive && grp.lane == 0) t = atomicAdd(ev.ply_counter, 1ULL);
        t = __shfl_sync(kFull, t, 0, LPG);
        if (alive && t >= budget) { alive = false; quit = true; }
        TreeCtxG<LPG> cx = make_ctx_g(ev, g, grp);
        cx.stage = sstage[threadIdx.x / LPG];
        const uint64_t game_id = ev.game_id[g];
        const uint64_t search_id = (uint64_t)ev.ply[g];
        init_root_g(cx, root.side, alive);
        const GBoard root_g = gboard_load(grp, root);
        if (__any_sync(kFull, alive)) {
            for (int sim = 0; sim < S; ++sim) {
                const uint64_t st = EVAL == RVS_EVAL_ROLLOUT ? stream_seed(ev.seed, game_id, (search_id << 16) | (uint64_t)sim) : 0ULL;
                simulate_one_g<EVAL>(cx, root_g, st, alive);
                if (sim == 0) {
                    if (ev.noise_eps > 0.0f) {
                        if (alive && grp.lane == 0) root_noise_apply(ev, g, game_id, search_id);
                        __syncwarp();
                    }
                    stage_root(cx, alive);  // the root was expanded by simulation 0: its rows move to shared memory
                }
            }
            unstage_root(cx);
        }
        if (alive && grp.lane == 0) ev.n_nodes[g] = cx.n_nodes;
        flush_stats_g(ev, cx, alive);
        end_of_ply<RULES, LPG>(ev, g, grp.lane, temperature, recycle, alive);
        idle = alive ? 0 : idle + 1;
        if (idle >= n_mine) quit = true;  // every slot of this group is parked or over
        k = k + 1 < n_mine ? k + 1 : 0;
    }
}

// ring -> trainer format (states [n,3,8,8] f32, pi [n,65] f32, z [n] f32)
template <int RULES>
__global__ void __launch_bounds__(256) drain_kernel(EngineView ev, int64_t n, float4* __restrict__ states,
                                                     float* __restrict__ pi, float* __restrict__ z) {
    const int64_t total = n * 48;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / 48;
        const int r = (int)(t - i * 48);
        const bool blk = ev.r_side[i] == 1;
        const uint64_t P = blk ? ev.r_black[i] : ev.r_white[i], O = blk ? ev.r_white[i] : ev.r_black[i];
        const int plane = r >> 4, q = r & 15;
        const uint64_t src = plane == 0 ? P : (plane == 1 ? O : legal_moves<RULES>(P, O));
        const uint32_t nib = (uint32_t)(src >> (4 * q)) & 15u;
        states[t] = make_float4((float)(nib & 1), (float)((nib >> 1) & 1), (float)((nib >> 2) & 1), (float)((nib >> 3) & 1));
        if (r == 0) z[i] = (float)ev.r_z[i];
    }
    const int64_t np = n * 65;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < np; t += (int64_t)gridDim.x * blockDim.x)
        pi[t] = ev.r_pi[t];
}

// rvs_engine_drain_packed_async: min(ring_count, capacity) oldest samples -> caller buffers, count read on the device
__global__ void __launch_bounds__(256) drain_packed_kernel(EngineView ev, int64_t capacity, uint64_t* __restrict__ black,
                                                            uint64_t* __restrict__ white, uint8_t* __restrict__ side,
                                                            int8_t* __restrict__ z, float* __restrict__ pi) {
    unsigned long long cnt = *ev.ring_count;
    if (cnt > (unsigned long long)ev.ring_cap) cnt = (unsigned long long)ev.ring_cap;
    const int64_t n = (int64_t)cnt < capacity ? (int64_t)cnt : capacity;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * 65; t += stride) {
        pi[t] = ev.r_pi[t];
        if (t < n) { black[t] = ev.r_black[t]; white[t] = ev.r_white[t]; side[t] = ev.r_side[t]; z[t] = ev.r_z[t]; }
    }
}
// ... then the samples beyond `capacity` (rare: the caller sizes its buffers for a generation) move to the front of
// the ring, chunk by chunk through one CTA so that overlapping ranges are safe, and the count is published
__global__ void __launch_bounds__(1024) drain_commit_kernel(EngineView ev, int64_t capacity, int64_t* __restrict__ out_count) {
    unsigned long long cnt = *ev.ring_count;
    if (cnt > (unsigned long long)ev.ring_cap) cnt = (unsigned long long)ev.ring_cap;
    const int64_t n = (int64_t)cnt < capacity ? (int64_t)cnt : capacity;
    const int64_t rest = (int64_t)cnt - n;
    for (int64_t base = 0; base < rest; base += blockDim.x) {  // forward move in chunks: dst < src, chunk-synchronous
        const int64_t i = base + threadIdx.x;
        uint64_t b = 0, w = 0; uint8_t sd = 0; int8_t zz = 0;
        if (i < rest) { b = ev.r_black[n + i]; w = ev.r_white[n + i]; sd = ev.r_side[n + i]; zz = ev.r_z[n + i]; }
        __syncthreads();
        if (i < rest) { ev.r_black[i] = b; ev.r_white[i] = w; ev.r_side[i] = sd; ev.r_z[i] = zz; }
        __syncthreads();
    }
    for (int64_t base = 0; base < rest * 65; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        float v = 0.f;
        if (i < rest * 65) v = ev.r_pi[n * 65 + i];
        __syncthreads();
        if (i < rest * 65) ev.r_pi[i] = v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *ev.ring_count = (unsigned long long)rest;
        *out_count = n;
        __threadfence_system();
    }
}

__global__ void reset_games_kernel(EngineView ev) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ev.G) return;
    ev.black[g] = kStartBlack; ev.white[g] = kStartWhite; ev.side[g] = 1; ev.flags[g] = 0;
    ev.game_id[g] = (uint64_t)g; ev.ply[g] = 0; ev.finished[g] = 0; ev.n_nodes[g] = 0;
    ev.live[g] = (ev.game_limit == 0 || (uint64_t)g < ev.game_limit) ? 1 : 0;
}

// positions handed in by the caller: derive game_over / winner the way Board would have when the
// game ended (neither side can move), so that root terminal handling matches mcts.py:567-575.
template <int RULES>
__global__ void set_positions_kernel(EngineView ev, const uint64_t* __restrict__ black, const uint64_t* __restrict__ white,
                                     const uint8_t* __restrict__ side, int n, uint64_t epoch) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    Board b{black[g], white[g], side[g], 0};
    ev.game_id[g] = (uint64_t)g + epoch * (uint64_t)ev.G;
    if ((b.side != 1 && b.side != 2) || (b.black & b.white)) {
        // not a position: the slot is parked on an empty finished board (a search sees a terminal root: no visits)
        ev.black[g] = 0; ev.white[g] = 0; ev.side[g] = 1; ev.flags[g] = F_OVER;
        ev.ply[g] = 0; ev.live[g] = 0; ev.finished[g] = 0; ev.n_nodes[g] = 0;
        atomicAdd(stat_at(ev, ST_BADPOS), 1ULL);
        return;
    }
    if (board_legal<RULES>(b) == 0) {
        Board o = b; o.side = (uint8_t)(3 - b.side);
        if (board_legal<RULES>(o) == 0) {
            const int nb = popc64(b.black), nw = popc64(b.white);
            const int w = nb > nw ? 1 : (nw > nb ? 2 : 0);
            b.flags = (uint8_t)(F_OVER | (w << F_WIN_SHIFT));
        }
    }
    ev.black[g] = b.black; ev.white[g] = b.white; ev.side[g] = b.side; ev.flags[g] = b.flags;
    ev.ply[g] = 0; ev.live[g] = 1; ev.finished[g] = 0; ev.n_nodes[g] = 0;
}

}  // namespace rvs

// ------------------------------------------------------------------------------ C ABI
using namespace rvs;

namespace {

template <typename T>
int dalloc(rvs_engine* h, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
    if (e != cudaSuccess) return fail(-100 - (int)e, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    cudaMemset(q, 0, count * sizeof(T));
    h->allocs[h->n_allocs++] = q;
    *p = (T*)q;
    return 0;
}

int io_stage(rvs_engine* h, size_t bytes, void** out) {
    if (h->io_stage_cap < bytes) {
        if (h->io_stage) cudaFree(h->io_stage);
        h->io_stage = nullptr; h->io_stage_cap = 0;
        RVS_CUDA(cudaMalloc(&h->io_stage, bytes));
        h->io_stage_cap = bytes;
    }
    *out = h->io_stage;
    return 0;
}

inline int games_grid(int G) { return (G + kWarpsPerBlock - 1) / kWarpsPerBlock; }

// lanes per game of the wave-1 kernels (rvs_treeg.cuh)
inline int lanes_per_game(const rvs_engine* h) {
    const int G = h->v.G;
    if (h->lanes_per_game) return h->lanes_per_game;  // rvs_engine_set_lanes_per_game
    // measured on B200: with few games the kernel is bound by the latency of one ply's
    // dependency chain, so more lanes per game (shorter per-lane chains, more warps) win; with many
    // games it is issue bound and fewer lanes per game (fewer instructions per game-ply) win
    // 16384 games: 3.6e8 / 4.9e8 / 4.0e8 sims/s with 8 / 4 / 2 lanes per game; 32768: - / 4.5e8 / 4.9e8; 65536: - / 4.8e8 / 5.6e8
    return G <= 6144 ? 8 : (G <= 24576 ? 4 : 2);
}


#define RVS_ENGINE_LAUNCH(h, ...)          \
    do {                                   \
        RVS_LAUNCH(__VA_ARGS__);           \
        (h)->launches++;                   \
    } while (0)

}  // namespace

extern "C" {

int rvs_engine_create(const rvs_engine_config* cfg, rvs_engine** out) {
    if (!cfg || !out) return fail(-1, "rvs_engine_create: null argument");
    if (cfg->struct_size != (int)sizeof(rvs_engine_config)) return fail(-1, "rvs_engine_create: struct_size mismatch (%d != %zu)", cfg->struct_size, sizeof(rvs_engine_config));
    if (cfg->n_games < 1 || cfg->max_sims < 1 || cfg->max_wave < 1 || cfg->max_sims > 65535)
        return fail(-1, "rvs_engine_create: n_games/max_sims/max_wave out of range");
    if (cfg->rules != RVS_RULES_REF && cfg->rules != RVS_RULES_STRICT) return fail(-1, "rvs_engine_create: bad rules");
    if (cfg->evaluator < RVS_EVAL_E0 || cfg->evaluator > RVS_EVAL_NN) return fail(-1, "rvs_engine_create: bad evaluator");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(-2, "rvs_engine_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(-1, "rvs_engine_create: bad device %d", cfg->device);
    DeviceGuard dg;
    if (int rc0 = dg.enter(cfg->device)) return rc0;
    rvs_engine* h = new (std::nothrow) rvs_engine();
    if (!h) return fail(-3, "out of host memory");
    h->cfg = *cfg;
    EngineView& v = h->v;
    v.G = cfg->n_games;
    v.kmax = cfg->max_wave;
    v.cap = cfg->nodes_per_game > 0 ? cfg->nodes_per_game : 2 + 34 * cfg->max_sims;
    v.c_puct = cfg->c_puct;
    v.noise_eps = 0.0f;
    v.noise_alpha = 0.0;
    v.seed = cfg->seed;
    v.game_limit = 0;
    v.mode = RVS_MODE_REF;
    v.ring_cap = cfg->sample_capacity > 0 ? cfg->sample_capacity : (int64_t)64 * v.G;
    const size_t G = v.G, GK = G * v.kmax, GN = G * (size_t)v.cap;
    int rc = 0;
    if ((rc = dalloc(h, &v.black, G)) || (rc = dalloc(h, &v.white, G)) || (rc = dalloc(h, &v.side, G)) ||
        (rc = dalloc(h, &v.flags, G)) || (rc = dalloc(h, &v.game_id, G)) || (rc = dalloc(h, &v.ply, G)) ||
        (rc = dalloc(h, &v.live, G)) || (rc = dalloc(h, &v.finished, G)) || (rc = dalloc(h, &v.hot, GN)) ||
        (rc = dalloc(h, &v.cold, GN)) ||
        (rc = dalloc(h, &v.brd, (cfg->evaluator == RVS_EVAL_E0 || cfg->evaluator == RVS_EVAL_ROLLOUT) ? GN : (size_t)1)) || (rc = dalloc(h, &v.n_nodes, G)) || (rc = dalloc(h, &v.order, G)) || (rc = dalloc(h, &v.w_node, GK)) ||
        (rc = dalloc(h, &v.w_plen, GK)) || (rc = dalloc(h, &v.w_path, GK * kMaxPath)) || (rc = dalloc(h, &v.w_black, GK)) ||
        (rc = dalloc(h, &v.w_white, GK)) || (rc = dalloc(h, &v.w_sf, GK)) || (rc = dalloc(h, &v.w_lm, GK)) ||
        (rc = dalloc(h, &v.w_val, GK)) || (rc = dalloc(h, &v.w_sides, GK)) || (rc = dalloc(h, &v.s_black, G * 64)) || (rc = dalloc(h, &v.s_white, G * 64)) ||
        (rc = dalloc(h, &v.s_side, G * 64)) || (rc = dalloc(h, &v.s_pi, G * 64 * 65)) ||
        (rc = dalloc(h, &v.r_black, (size_t)v.ring_cap)) || (rc = dalloc(h, &v.r_white, (size_t)v.ring_cap)) ||
        (rc = dalloc(h, &v.r_side, (size_t)v.ring_cap)) || (rc = dalloc(h, &v.r_z, (size_t)v.ring_cap)) ||
        (rc = dalloc(h, &v.r_pi, (size_t)v.ring_cap * 65)) || (rc = dalloc(h, &v.ring_count, 1)) || (rc = dalloc(h, &v.ply_counter, 1)) ||
        (rc = dalloc(h, &v.stats, (size_t)kStatStripes * kStatStride)) || (rc = dalloc(h, &h->visits, G * 65)) || (rc = dalloc(h, &h->moves, G))) {
        rvs_engine_destroy(h);
        return rc;
    }
    if (cudaHostAlloc((void**)&h->pinned_count, 64, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        rvs_engine_destroy(h);
        return fail(-3, "rvs_engine_create: cudaHostAlloc failed");
    }
    reset_games_kernel<<<(v.G + 127) / 128, 128>>>(v);
    g_launches.fetch_add(1);
    h->launches++;
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        rvs_engine_destroy(h);
        return fail(-100 - (int)e, "rvs_engine_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return 0;
}

int rvs_engine_destroy(rvs_engine* h) {
    if (!h) return 0;
    DeviceGuard dg;
    dg.enter(h->cfg.device);
    cudaDeviceSynchronize();
    for (int i = 0; i < h->n_allocs; ++i) cudaFree(h->allocs[i]);
    if (h->pinned_count) cudaFreeHost(h->pinned_count);
    if (h->io_stage) cudaFree(h->io_stage);
    if (h->ext_probs) cudaFree(h->ext_probs);
    if (h->ext_values) cudaFree(h->ext_values);
    if (h->ext_planes) cudaFree(h->ext_planes);
    if (h->ext_valid) cudaFree(h->ext_valid);
    if (h->net) rvs_net_destroy(h->net);
    delete h;
    return 0;
}

int rvs_engine_reset(rvs_engine* h, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    cudaStream_t s = (cudaStream_t)stream;
    RVS_ENGINE_LAUNCH(h, reset_games_kernel, (h->v.G + 127) / 128, 128, 0, s, h->v);
    RVS_CUDA(cudaMemsetAsync(h->v.ring_count, 0, 8, s));
    h->searching = false;
    h->epoch = 0;
    return 0;
}

int rvs_engine_set_positions(rvs_engine* h, const uint64_t* black, const uint64_t* white, const uint8_t* side, int32_t n,
                             int mem, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    if (n < 0 || n > h->v.G || (n > 0 && (!black || !white || !side))) return fail(-1, "rvs_engine_set_positions: bad arguments");
    if (mem < RVS_MEM_DEVICE || mem > RVS_MEM_HOST_ASYNC) return fail(-1, "rvs_engine_set_positions: bad mem %d", mem);
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t *db = black, *dw = white;
    const uint8_t* ds = side;
    if (mem != RVS_MEM_DEVICE) {
        for (int32_t i = 0; i < n; ++i)  // host inputs are validated before anything is uploaded
            if ((side[i] != 1 && side[i] != 2) || (black[i] & white[i]))
                return fail(-1, "rvs_engine_set_positions: position %d is invalid (side %d, %s)", i, (int)side[i],
                            (black[i] & white[i]) ? "black and white discs overlap" : "side must be 1 or 2");
        void* st = nullptr;
        const size_t nb = (size_t)n * 8;
        if ((rc = io_stage(h, 2 * nb + n, &st))) return rc;
        RVS_CUDA(cudaMemcpyAsync(st, black, nb, cudaMemcpyHostToDevice, s));
        RVS_CUDA(cudaMemcpyAsync((char*)st + nb, white, nb, cudaMemcpyHostToDevice, s));
        RVS_CUDA(cudaMemcpyAsync((char*)st + 2 * nb, side, n, cudaMemcpyHostToDevice, s));
        db = (const uint64_t*)st; dw = (const uint64_t*)((char*)st + nb); ds = (const uint8_t*)((char*)st + 2 * nb);
    }
    const uint64_t epoch = h->epoch++;
    if (h->cfg.rules == RVS_RULES_STRICT)
        RVS_ENGINE_LAUNCH(h, set_positions_kernel<RULES_STRICT>, (n + 127) / 128, 128, 0, s, h->v, db, dw, ds, n, epoch);
    else
        RVS_ENGINE_LAUNCH(h, set_positions_kernel<RULES_REF>, (n + 127) / 128, 128, 0, s, h->v, db, dw, ds, n, epoch);
    // pageable host memory: the staged copies above may still be reading the caller's arrays
    if (mem == RVS_MEM_HOST) RVS_CUDA(cudaStreamSynchronize(s));
    h->searching = false;
    return 0;
}

int rvs_engine_get_positions(rvs_engine* h, uint64_t* black, uint64_t* white, uint8_t* side, uint8_t* flags, int32_t n,
                             int mem, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (n < 0 || n > h->v.G) return fail(-1, "rvs_engine_get_positions: bad n");
    cudaStream_t s = (cudaStream_t)stream;
    const bool host = mem != RVS_MEM_DEVICE;
    const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (black) RVS_CUDA(cudaMemcpyAsync(black, h->v.black, (size_t)n * 8, kind, s));
    if (white) RVS_CUDA(cudaMemcpyAsync(white, h->v.white, (size_t)n * 8, kind, s));
    if (side) RVS_CUDA(cudaMemcpyAsync(side, h->v.side, n, kind, s));
    if (flags) RVS_CUDA(cudaMemcpyAsync(flags, h->v.flags, n, kind, s));
    if (host) RVS_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int rvs_engine_search(rvs_engine* h, int32_t num_sims, int32_t wave, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (num_sims < 1 || num_sims > h->cfg.max_sims) return fail(-1, "rvs_engine_search: num_sims %d outside [1,%d]", num_sims, h->cfg.max_sims);
    if (wave < 1 || wave > h->cfg.max_wave) return fail(-1, "rvs_engine_search: wave %d outside [1,%d]", wave, h->cfg.max_wave);
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = games_grid(h->v.G);
    const bool strict = h->cfg.rules == RVS_RULES_STRICT;
    if (wave == 1 && h->v.mode == RVS_MODE_REF && (h->cfg.evaluator == RVS_EVAL_E0 || h->cfg.evaluator == RVS_EVAL_ROLLOUT)) {
        const bool e0 = h->cfg.evaluator == RVS_EVAL_E0;
        {
            RVS_ENGINE_LAUNCH(h, phase_order_kernel, 1, 1024, 0, s, h->v);
            const int lpg = lanes_per_game(h);
#define RVS_SEARCH_GM(LPG, MINB)                                                                                                  \
    do {                                                                                                                          \
        if (strict && e0) RVS_ENGINE_LAUNCH(h, (search_k1g_kernel<RULES_STRICT, RVS_EVAL_E0, LPG, MINB>), gridg, kBlockG, 0, s, h->v, num_sims); \
        else if (strict) RVS_ENGINE_LAUNCH(h, (search_k1g_kernel<RULES_STRICT, RVS_EVAL_ROLLOUT, LPG, MINB>), gridg, kBlockG, 0, s, h->v, num_sims); \
        else if (e0) RVS_ENGINE_LAUNCH(h, (search_k1g_kernel<RULES_REF, RVS_EVAL_E0, LPG, MINB>), gridg, kBlockG, 0, s, h->v, num_sims); \
        else RVS_ENGINE_LAUNCH(h, (search_k1g_kernel<RULES_REF, RVS_EVAL_ROLLOUT, LPG, MINB>), gridg, kBlockG, 0, s, h->v, num_sims);  \
    } while (0)
#define RVS_SEARCH_G(LPG)                                                                                                         \
    do {                                                                                                                          \
        const int need = (h->v.G * LPG + kBlockG - 1) / kBlockG;                                                                  \
        const int gridg = need < kMaxWarpsG ? need : kMaxWarpsG;                                                                  \
        if (need <= kNumSMs * 8) RVS_SEARCH_GM(LPG, 8);                                                                           \
        else RVS_SEARCH_GM(LPG, 16);                                                                                              \
    } while (0)
            if (lpg == 8) RVS_SEARCH_G(8);
            else if (lpg == 2) RVS_SEARCH_G(2);
            else RVS_SEARCH_G(4);
#undef RVS_SEARCH_G
#undef RVS_SEARCH_GM
        }
        h->searching = false;
        return 0;
    }
    switch (h->cfg.evaluator) {
    case RVS_EVAL_E0:
        if (strict) RVS_ENGINE_LAUNCH(h, (search_fused_kernel<RULES_STRICT, RVS_EVAL_E0>), grid, kBlock, 0, s, h->v, num_sims, wave);
        else RVS_ENGINE_LAUNCH(h, (search_fused_kernel<RULES_REF, RVS_EVAL_E0>), grid, kBlock, 0, s, h->v, num_sims, wave);
        break;
    case RVS_EVAL_ROLLOUT:
        if (strict) RVS_ENGINE_LAUNCH(h, (search_fused_kernel<RULES_STRICT, RVS_EVAL_ROLLOUT>), grid, kBlock, 0, s, h->v, num_sims, wave);
        else RVS_ENGINE_LAUNCH(h, (search_fused_kernel<RULES_REF, RVS_EVAL_ROLLOUT>), grid, kBlock, 0, s, h->v, num_sims, wave);
        break;
    case RVS_EVAL_NN:
        return rvs_net_search(h, num_sims, wave, s);
    default:
        return fail(-1, "rvs_engine_search: evaluator EXTERNAL needs begin_search/select/process");
    }
    h->searching = false;
    return 0;
}

int rvs_engine_begin_search(rvs_engine* h, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    RVS_ENGINE_LAUNCH(h, begin_search_kernel, games_grid(h->v.G), kBlock, 0, (cudaStream_t)stream, h->v);
    h->searching = true;
    h->cur_k = 0;
    h->waves_done = 0;
    return 0;
}

int rvs_engine_select(rvs_engine* h, int32_t k, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (!h->searching) return fail(-1, "rvs_engine_select: call rvs_engine_begin_search first");
    if (k < 1 || k > h->cfg.max_wave) return fail(-1, "rvs_engine_select: k %d outside [1,%d]", k, h->cfg.max_wave);
    if (h->cur_k != 0) return fail(-1, "rvs_engine_select: previous wave not processed");
    cudaStream_t s = (cudaStream_t)stream;
    if (h->cfg.rules == RVS_RULES_STRICT) RVS_ENGINE_LAUNCH(h, select_kernel<RULES_STRICT>, games_grid(h->v.G), kBlock, 0, s, h->v, k);
    else RVS_ENGINE_LAUNCH(h, select_kernel<RULES_REF>, games_grid(h->v.G), kBlock, 0, s, h->v, k);
    h->cur_k = k;
    return 0;
}

int rvs_engine_leaf_planes(rvs_engine* h, float* out_planes, uint8_t* out_valid, int mem, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (h->cur_k == 0) return fail(-1, "rvs_engine_leaf_planes: no selected wave");
    if (!out_planes) return fail(-1, "rvs_engine_leaf_planes: null output");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t slots = (int64_t)h->v.G * h->cur_k;
    float* dp = out_planes;
    uint8_t* dv = out_valid;
    if (mem != RVS_MEM_DEVICE) {
        const size_t need = (size_t)h->v.G * h->cfg.max_wave;
        if (!h->ext_planes) RVS_CUDA(cudaMalloc(&h->ext_planes, need * 192 * sizeof(float)));
        if (!h->ext_valid) RVS_CUDA(cudaMalloc(&h->ext_valid, need));
        dp = h->ext_planes;
        dv = h->ext_valid;
    }
    RVS_ENGINE_LAUNCH(h, leaf_planes_kernel, grid_for(slots * 48, 256), 256, 0, s, h->v, h->cur_k, (float4*)dp, dv);
    if (mem != RVS_MEM_DEVICE) {
        RVS_CUDA(cudaMemcpyAsync(out_planes, dp, slots * 192 * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (out_valid) RVS_CUDA(cudaMemcpyAsync(out_valid, dv, slots, cudaMemcpyDeviceToHost, s));
        RVS_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int rvs_engine_process(rvs_engine* h, const float* probs, const float* values, int mem, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (h->cur_k == 0) return fail(-1, "rvs_engine_process: no selected wave");
    if (!probs || !values) return fail(-1, "rvs_engine_process: null input");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t slots = (int64_t)h->v.G * h->cur_k;
    const float *dp = probs, *dv = values;
    if (mem != RVS_MEM_DEVICE) {
        const size_t need = (size_t)h->v.G * h->cfg.max_wave;
        if (!h->ext_probs) RVS_CUDA(cudaMalloc(&h->ext_probs, need * 65 * sizeof(float)));
        if (!h->ext_values) RVS_CUDA(cudaMalloc(&h->ext_values, need * sizeof(float)));
        RVS_CUDA(cudaMemcpyAsync(h->ext_probs, probs, slots * 65 * sizeof(float), cudaMemcpyHostToDevice, s));
        RVS_CUDA(cudaMemcpyAsync(h->ext_values, values, slots * sizeof(float), cudaMemcpyHostToDevice, s));
        dp = h->ext_probs;
        dv = h->ext_values;
    }
    RVS_ENGINE_LAUNCH(h, process_probs_kernel, games_grid(h->v.G), kBlock, 0, s, h->v, h->cur_k, dp, dv, h->waves_done == 0 ? 1 : 0, (const int*)nullptr);
    h->waves_done++;
    h->cur_k = 0;
    return 0;
}

int rvs_engine_root_visits(rvs_engine* h, int32_t* out, int32_t n, int mem, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (!out || n < 0 || n > h->v.G) return fail(-1, "rvs_engine_root_visits: bad arguments");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const bool host = mem == RVS_MEM_HOST || mem == RVS_MEM_HOST_ASYNC;
    int32_t* d = host ? h->visits : out;
    RVS_ENGINE_LAUNCH(h, root_visits_kernel, games_grid(n), kBlock, 0, s, h->v, d, n);
    if (host) {
        RVS_CUDA(cudaMemcpyAsync(out, d, (size_t)n * 65 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        if (mem == RVS_MEM_HOST) RVS_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int rvs_engine_play(rvs_engine* h, float temperature, int recycle, uint8_t* out_moves, int mem, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (temperature < 0.0f) return fail(-1, "rvs_engine_play: negative temperature");
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* dm = out_moves ? (mem != RVS_MEM_DEVICE ? h->moves : out_moves) : nullptr;
    if (h->cfg.rules == RVS_RULES_STRICT) RVS_ENGINE_LAUNCH(h, play_kernel<RULES_STRICT>, (h->v.G + 127) / 128, 128, 0, s, h->v, temperature, dm);
    else RVS_ENGINE_LAUNCH(h, play_kernel<RULES_REF>, (h->v.G + 127) / 128, 128, 0, s, h->v, temperature, dm);
    RVS_ENGINE_LAUNCH(h, finalize_kernel, games_grid(h->v.G), kBlock, 0, s, h->v, recycle);
    if (out_moves && mem != RVS_MEM_DEVICE) {
        RVS_CUDA(cudaMemcpyAsync(out_moves, dm, h->v.G, cudaMemcpyDeviceToHost, s));
        RVS_CUDA(cudaStreamSynchronize(s));
    }
    h->searching = false;
    return 0;
}

int rvs_engine_selfplay(rvs_engine* h, int32_t num_sims, float temperature, int64_t plies, int recycle, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (num_sims < 1 || num_sims > h->cfg.max_sims) return fail(-1, "rvs_engine_selfplay: num_sims %d outside [1,%d]", num_sims, h->cfg.max_sims);
    if (plies < 0 || temperature < 0.0f) return fail(-1, "rvs_engine_selfplay: bad arguments");
    if (h->cfg.evaluator == RVS_EVAL_EXTERNAL)
        return fail(-1, "rvs_engine_selfplay: the external evaluator is driven by the caller (begin_search / select / process + play)");
    if (h->cfg.evaluator == RVS_EVAL_NN || h->v.mode == RVS_MODE_FAST) {
        // the network evaluates all leaves of a wave in one batch (and FAST waves are lockstep by construction), so
        // this self-play advances in lockstep: ceil(plies / n_games) rounds of (search, play) on `stream`.
        // REF mode: wave 1 (= MCTS(batch_size=1)); FAST mode: waves of max_wave
        const int64_t rounds = (plies + h->v.G - 1) / h->v.G;
        const int wave = h->v.mode == RVS_MODE_FAST ? h->cfg.max_wave : 1;
        for (int64_t r = 0; r < rounds; ++r) {
            if ((rc = rvs_engine_search(h, num_sims, wave, stream))) return rc;
            if ((rc = rvs_engine_play(h, temperature, recycle, nullptr, RVS_MEM_DEVICE, stream))) return rc;
        }
        return 0;
    }
    if (h->cfg.evaluator != RVS_EVAL_E0 && h->cfg.evaluator != RVS_EVAL_ROLLOUT)
        return fail(-1, "rvs_engine_selfplay: the external evaluator is driven by the caller (begin_search / select / process + play)");
    cudaStream_t s = (cudaStream_t)stream;
    RVS_CUDA(cudaMemsetAsync(h->v.ply_counter, 0, 8, s));
    const bool strict = h->cfg.rules == RVS_RULES_STRICT, e0 = h->cfg.evaluator == RVS_EVAL_E0;
    const unsigned long long budget = (unsigned long long)plies;
    {
        RVS_ENGINE_LAUNCH(h, phase_order_kernel, 1, 1024, 0, s, h->v);
        const int lpg = lanes_per_game(h);
#define RVS_SELFPLAY_GM(LPG, MINB)                                                                                                \
    do {                                                                                                                          \
        if (strict && e0) RVS_ENGINE_LAUNCH(h, (selfplay_k1g_kernel<RULES_STRICT, RVS_EVAL_E0, LPG, MINB>), gridg, kBlockG, 0, s, h->v, num_sims, temperature, budget, recycle); \
        else if (strict) RVS_ENGINE_LAUNCH(h, (selfplay_k1g_kernel<RULES_STRICT, RVS_EVAL_ROLLOUT, LPG, MINB>), gridg, kBlockG, 0, s, h->v, num_sims, temperature, budget, recycle); \
        else if (e0) RVS_ENGINE_LAUNCH(h, (selfplay_k1g_kernel<RULES_REF, RVS_EVAL_E0, LPG, MINB>), gridg, kBlockG, 0, s, h->v, num_sims, temperature, budget, recycle); \
        else RVS_ENGINE_LAUNCH(h, (selfplay_k1g_kernel<RULES_REF, RVS_EVAL_ROLLOUT, LPG, MINB>), gridg, kBlockG, 0, s, h->v, num_sims, temperature, budget, recycle); \
    } while (0)
#define RVS_SELFPLAY_G(LPG)                                                                                                       \
    do {                                                                                                                          \
        const int need = (h->v.G * LPG + kBlockG - 1) / kBlockG;                                                                  \
        const int gridg = need < kMaxWarpsG ? need : kMaxWarpsG;                                                                  \
        if (need <= kNumSMs * 8) RVS_SELFPLAY_GM(LPG, 8);                                                                         \
        else RVS_SELFPLAY_GM(LPG, 16);                                                                                            \
    } while (0)
        if (lpg == 8) RVS_SELFPLAY_G(8);
        else if (lpg == 2) RVS_SELFPLAY_G(2);
        else RVS_SELFPLAY_G(4);
#undef RVS_SELFPLAY_G
#undef RVS_SELFPLAY_GM
    }
    h->searching = false;
    return 0;
}

// number of completed samples waiting in the ring, through the handle's pinned word (the copy is ordered after
// everything enqueued on `s`, so the host waits for exactly the work the samples depend on and nothing else)
static int ring_pending(rvs_engine* h, cudaStream_t s, int64_t* n) {
    RVS_CUDA(cudaMemcpyAsync(h->pinned_count, h->v.ring_count, 8, cudaMemcpyDeviceToHost, s));
    RVS_CUDA(cudaStreamSynchronize(s));
    const unsigned long long cnt = *h->pinned_count;
    *n = (int64_t)cnt < h->v.ring_cap ? (int64_t)cnt : h->v.ring_cap;
    return 0;
}

int rvs_engine_drain_samples(rvs_engine* h, float* states, float* pi, float* z, int64_t capacity, int64_t* out_count,
                             int mem, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    if (!out_count) return fail(-1, "rvs_engine_drain_samples: null out_count");
    cudaStream_t s = (cudaStream_t)stream;
    int64_t n = 0;
    if ((rc = ring_pending(h, s, &n))) return rc;
    *out_count = n;
    if (n == 0) return 0;
    if (!states || !pi || !z) return fail(-1, "rvs_engine_drain_samples: null output");
    if (capacity < n) return fail(-4, "rvs_engine_drain_samples: capacity %lld < %lld samples pending", (long long)capacity, (long long)n);
    const bool host = mem != RVS_MEM_DEVICE;
    float *ds = states, *dp = pi, *dz = z;
    const size_t bs = (size_t)n * 192 * 4, bp = (size_t)n * 65 * 4, bz = (size_t)n * 4;
    if (host) {
        void* st = nullptr;
        if ((rc = io_stage(h, bs + bp + bz, &st))) return rc;
        ds = (float*)st; dp = (float*)((char*)st + bs); dz = (float*)((char*)st + bs + bp);
    }
    const int grid = grid_for(n * 65, 256);
    if (h->cfg.rules == RVS_RULES_STRICT) RVS_ENGINE_LAUNCH(h, drain_kernel<RULES_STRICT>, grid, 256, 0, s, h->v, n, (float4*)ds, dp, dz);
    else RVS_ENGINE_LAUNCH(h, drain_kernel<RULES_REF>, grid, 256, 0, s, h->v, n, (float4*)ds, dp, dz);
    RVS_CUDA(cudaMemsetAsync(h->v.ring_count, 0, 8, s));
    if (host) {
        RVS_CUDA(cudaMemcpyAsync(states, ds, bs, cudaMemcpyDeviceToHost, s));
        RVS_CUDA(cudaMemcpyAsync(pi, dp, bp, cudaMemcpyDeviceToHost, s));
        RVS_CUDA(cudaMemcpyAsync(z, dz, bz, cudaMemcpyDeviceToHost, s));
        RVS_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int rvs_engine_drain_packed(rvs_engine* h, uint64_t* black, uint64_t* white, uint8_t* side, int8_t* z, float* pi,
                            int64_t capacity, int64_t* out_count, int mem, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    if (!out_count) return fail(-1, "rvs_engine_drain_packed: null out_count");
    cudaStream_t s = (cudaStream_t)stream;
    int64_t n = 0;
    if ((rc = ring_pending(h, s, &n))) return rc;
    *out_count = n;
    if (n == 0) return 0;
    if (!black || !white || !side || !z || !pi) return fail(-1, "rvs_engine_drain_packed: null output");
    if (capacity < n) return fail(-4, "rvs_engine_drain_packed: capacity %lld < %lld samples pending", (long long)capacity, (long long)n);
    const cudaMemcpyKind kind = mem == RVS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    RVS_CUDA(cudaMemcpyAsync(black, h->v.r_black, (size_t)n * 8, kind, s));
    RVS_CUDA(cudaMemcpyAsync(white, h->v.r_white, (size_t)n * 8, kind, s));
    RVS_CUDA(cudaMemcpyAsync(side, h->v.r_side, (size_t)n, kind, s));
    RVS_CUDA(cudaMemcpyAsync(z, h->v.r_z, (size_t)n, kind, s));
    RVS_CUDA(cudaMemcpyAsync(pi, h->v.r_pi, (size_t)n * 65 * sizeof(float), kind, s));
    RVS_CUDA(cudaMemsetAsync(h->v.ring_count, 0, 8, s));
    if (mem != RVS_MEM_DEVICE) RVS_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int rvs_engine_drain_packed_async(rvs_engine* h, uint64_t* black, uint64_t* white, uint8_t* side, int8_t* z, float* pi,
                                  int64_t capacity, int64_t* out_count_dev, void* stream) {
    RVS_ENTER(h);
    if (!out_count_dev || !black || !white || !side || !z || !pi || capacity < 0)
        return fail(-1, "rvs_engine_drain_packed_async: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    // two launches: the copy reads the count, the second kernel moves the remainder to the front and
    // publishes the count -- nothing visits the host
    RVS_ENGINE_LAUNCH(h, drain_packed_kernel, grid_for(capacity * 17, 256), 256, 0, s, h->v, capacity, black, white, side, z, pi);
    RVS_ENGINE_LAUNCH(h, drain_commit_kernel, 1, 1024, 0, s, h->v, capacity, out_count_dev);
    return 0;
}

int rvs_engine_set_lanes_per_game(rvs_engine* h, int32_t lanes) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (lanes != 0 && lanes != 2 && lanes != 4 && lanes != 8) return fail(-1, "rvs_engine_set_lanes_per_game: %d not in {0, 2, 4, 8}", lanes);
    h->lanes_per_game = lanes;
    return 0;
}

int rvs_engine_set_option(rvs_engine* h, int32_t option, int64_t value) {
    RVS_ENTER(h);
    switch (option) {
    case RVS_OPT_LANES_PER_GAME:
        return rvs_engine_set_lanes_per_game(h, (int32_t)value);
    case RVS_OPT_NET_GRAPH:
        h->net_graph = value != 0;
        return 0;
    case RVS_OPT_SEARCH_MODE:
        if (value != RVS_MODE_REF && value != RVS_MODE_FAST) return fail(-1, "rvs_engine_set_option: search mode %lld not in {REF, FAST}", (long long)value);
        h->v.mode = (int)value;
        return 0;
    case RVS_OPT_GAME_LIMIT:
        if (value < 0 || (value > 0 && value < h->v.G))
            return fail(-1, "rvs_engine_set_option: game limit %lld must be 0 or >= n_games (%d)", (long long)value, h->v.G);
        h->v.game_limit = (uint64_t)value;
        return 0;
    case RVS_OPT_NET_MAX_CTAS:
        if (value < 0 || value > kNumSMs) return fail(-1, "rvs_engine_set_option: net_max_ctas %lld outside [0,%d]", (long long)value, kNumSMs);
        h->net_max_ctas = (int)value;
        return 0;
    case RVS_OPT_NET_PIPELINE:
        h->net_pipeline = value != 0;
        return 0;
    case RVS_OPT_NET_TOWER:
        h->net_tower = value != 0;
        return 0;
    default:
        return fail(-1, "rvs_engine_set_option: unknown option %d", option);
    }
}

int rvs_engine_set_root_noise(rvs_engine* h, double alpha, float epsilon) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (!(epsilon >= 0.0f && epsilon <= 1.0f)) return fail(-1, "rvs_engine_set_root_noise: epsilon %g outside [0,1]", (double)epsilon);
    if (epsilon > 0.0f && !(alpha >= 1e-3 && alpha <= 1e3)) return fail(-1, "rvs_engine_set_root_noise: alpha %g outside [1e-3,1e3]", alpha);
    h->v.noise_alpha = alpha;
    h->v.noise_eps = epsilon;
    return 0;
}

int rvs_engine_stats_get(rvs_engine* h, rvs_engine_stats* out, void* stream) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (!out) return fail(-1, "rvs_engine_stats_get: null output");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long raw[kStatStripes * kStatStride], st[ST_COUNT] = {};
    RVS_CUDA(cudaMemcpyAsync(raw, h->v.stats, sizeof(raw), cudaMemcpyDeviceToHost, s));
    RVS_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < kStatStripes; ++i)
        for (int k = 0; k < ST_COUNT; ++k) st[k] += raw[i * kStatStride + k];
    out->sims = (int64_t)st[ST_SIMS];
    out->evals = (int64_t)st[ST_EVALS];
    out->board_steps = (int64_t)st[ST_STEPS];
    out->nodes = (int64_t)st[ST_NODES];
    out->tree_bytes = (int64_t)st[ST_BYTES];
    out->games_finished = (int64_t)st[ST_FINISHED];
    out->samples = (int64_t)st[ST_SAMPLES];
    out->launches = h->launches;
    out->overflow = (int64_t)st[ST_OVERFLOW];
    out->samples_dropped = (int64_t)st[ST_DROPPED];
    out->stalled = (int64_t)st[ST_STALLED];
    out->nn_evals = (int64_t)st[ST_NNEVALS];
    out->bad_positions = (int64_t)st[ST_BADPOS];
    return 0;
}

}  // extern "C"

int rvs_engine_process_mapped(rvs_engine* h, const float* probs, const float* values, const int* inv, cudaStream_t s) {
    RVS_ENTER(h);
    int rc = 0;
    (void)rc;
    if (h->cur_k == 0) return fail(-1, "rvs_engine_process: no selected wave");
    RVS_ENGINE_LAUNCH(h, process_probs_kernel, games_grid(h->v.G), kBlock, 0, s, h->v, h->cur_k, probs, values, h->waves_done == 0 ? 1 : 0, inv);
    h->waves_done++;
    h->cur_k = 0;
    return 0;
}

int rvs_engine_nn_step(rvs_engine* h, int g0, int g1, int flags, const float* probs, const float* values, int* rows,
                       uint64_t* bits_out, int* n_cur, int* n_next, void* tiles_out, cudaStream_t s, bool pdl) {
    if (g1 <= g0) return 0;
    const int grid = (g1 - g0 + kWarpsPerBlock - 1) / kWarpsPerBlock;
    // pdl = false (pipelined half-batches): a plain launch.  A dependent launched early becomes RESIDENT and waits; beside
    // the other half's whole-network CTAs only one small CTA fits per SM, and a squatting one keeps the other half's tree
    // step / heads out.
    if (!pdl) {
        if (h->cfg.rules == RVS_RULES_STRICT) RVS_LAUNCH(nn_step_kernel<RULES_STRICT>, grid, kBlock, 0, s, h->v, g0, g1, flags, probs, values, rows, bits_out, n_cur, n_next, (uint4*)tiles_out);
        else RVS_LAUNCH(nn_step_kernel<RULES_REF>, grid, kBlock, 0, s, h->v, g0, g1, flags, probs, values, rows, bits_out, n_cur, n_next, (uint4*)tiles_out);
    } else if (h->cfg.rules == RVS_RULES_STRICT) RVS_LAUNCH_PDL(nn_step_kernel<RULES_STRICT>, grid, kBlock, 0, s, h->v, g0, g1, flags, probs, values, rows, bits_out, n_cur, n_next, (uint4*)tiles_out);
    else RVS_LAUNCH_PDL(nn_step_kernel<RULES_REF>, grid, kBlock, 0, s, h->v, g0, g1, flags, probs, values, rows, bits_out, n_cur, n_next, (uint4*)tiles_out);
    h->launches++;
    return 0;
}
