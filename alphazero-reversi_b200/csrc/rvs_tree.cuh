// rvs_tree.cuh -- K2: lockstep batched MCTS, one warp per game, trees as flat rows in HBM.
//
// Reference semantics restated (src/mcts/mcts.py; verified against the live reference
// through oracle/rvs_oracle.c):
//   MCTSNode.ucb_score            mcts.py:84-114   -> score_child()
//   MCTS._traverse                mcts.py:409-444  -> select_one()
//   MCTS._process_batch           mcts.py:544-623  -> process_wave()
//   MCTS._backpropagate_path      mcts.py:625-640  -> backup_path()
//
// Tree layout (per game, `cap` rows each, two arrays of 16-byte rows = SoA of packed rows):
//   hot [node] = { int N; float W; int vlf; float cache }      selection + backup touch only this
//       vlf: bits 0-15 virtual loss, bit 16 cache valid, bit 17 terminal, bits 18-19 terminal
//            value code (0 -> 0.0, 1 -> +1.0, 2 -> -1.0)
//   cold[node] = { float P; int first_child; int meta; int pad }
//       meta: bits 0-7 nchild, 8-15 move square (255 root), 16-17 turn label
// Children of a node are contiguous (first_child .. first_child+nchild), created in ascending
// square order = the reference's dict insertion order (row-major legal squares,
// mcts.py:608-611), so "lowest lane wins ties" == the reference's first-max rule
// (mcts.py:425 strict '>').
//
// Float rules: every tree operation is a single IEEE f32 op issued through __f*_rn
// intrinsics (never contracted to FMA) in the reference's operation order.
#pragma once
#include "rvs_board.cuh"

namespace rvs {

constexpr int kVLMask = 0xFFFF;
constexpr int kCacheValid = 1 << 16;
constexpr int kTerminal = 1 << 17;
constexpr int kTermShift = 18;
constexpr int kMaxPath = 64;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned ordered_key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ int4 shfl4(const int4& v, int src) {
    return make_int4(__shfl_sync(kFull, v.x, src), __shfl_sync(kFull, v.y, src), __shfl_sync(kFull, v.z, src),
                     __shfl_sync(kFull, v.w, src));
}
__device__ __forceinline__ float term_value_of(int vlf) {
    const int code = (vlf >> kTermShift) & 3;
    return code == 1 ? 1.0f : (code == 2 ? -1.0f : 0.0f);
}

// per-warp view of one game's tree + running counters
struct TreeCtx {
    int4* hot;
    int4* cold;
    int cap;
    int n_nodes;
    float c_puct;
    int lane;
    int overflow;
    // per-launch, per-game counters (32 bit is ample: <= 65535 sims x 64 plies)
    unsigned steps, sims, evals;
    unsigned bytes;    // algorithmic HBM bytes (DESIGN.md 'K2 roofline'): 32 B per row touched
    unsigned created;  // child nodes created
    DirLane dir;                 // this lane's direction slice for the warp-cooperative board ops
};

// per-game wave scratch (global memory), slot j in [0, wave)
struct WaveScratch {
    int* node;          // leaf node of sim j, -1 when the sim ended on a terminal node
    int* plen;          // path length
    int* path;          // [wave][kMaxPath]
    uint64_t* black;    // leaf position
    uint64_t* white;
    uint16_t* sf;       // side | flags << 8
    uint64_t* lm;       // legal mask of the leaf (filled by precompute)
    float* val;         // evaluator value
    uint64_t* sides;    // FAST mode: bit d = the side to move at path node d is WHITE
};

// MCTSNode.ucb_score (mcts.py:96-114) for a visited child without a valid cache
__device__ __forceinline__ float score_child(int N, float W, int VL, float P, int turn, float c_puct, float sq) {
    // W == 0 (draws, or wins and losses that cancel: ~19 % of the evaluations, ncu) would send the IEEE division down
    // its slow path (a call); +-0 / n = +-0, so the quotient is W itself
    const bool wz = W == 0.0f;
    float q = __fdiv_rn(wz ? 1.0f : W, (float)(N > 1 ? N : 1));
    q = wz ? W : q;
    float u = __fmul_rn(c_puct, P);
    u = __fmul_rn(u, sq);
    u = __fdiv_rn(u, (float)(1 + N + VL));
    if (turn != 1) q = -q;
    return __fadd_rn(__fadd_rn(q, u), 0.0f);  // + 0.0f canonicalises -0 (cannot change any other value)
}

// MCTS._backpropagate_path (mcts.py:625-640): lane d owns path node d; nodes of one path are
// distinct, so the read-modify-writes need no atomics.
__device__ __forceinline__ void backup_path(TreeCtx& cx, int p0, int p1, int plen, float v) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int d = cx.lane + 32 * half;
        if (d < plen) {
            const int n = half ? p1 : p0;
            int4 h = cx.hot[n];
            const float sv = ((plen - 1 - d) & 1) ? -v : v;
            h.x += 1;
            h.y = __float_as_int(__fadd_rn(__int_as_float(h.y), sv));
            int vl = h.z & kVLMask;
            if (vl > 0) --vl;
            h.z = (h.z & ~(kVLMask | kCacheValid)) | vl;
            cx.hot[n] = h;
        }
    }
    cx.bytes += 32u * (unsigned)plen;  // hot row read + write per path node
    __syncwarp();
}

// MCTS._traverse (mcts.py:409-444).  Returns the leaf node; lane d holds path node d in
// p0 (d<32) / p1 (d>=32); `b` is advanced to the leaf position; `leaf_vlf` = leaf's vlf word.
__device__ __forceinline__ int select_one(TreeCtx& cx, CoopBoard& b, int& p0, int& p1, int& plen, int& leaf_vlf) {
    int node = 0;
    p0 = 0; p1 = 0;  // path[0] = root (node 0) for lane 0; other lanes overwritten as we descend
    plen = 1;
    int4 h = cx.hot[0];
    int4 c = cx.cold[0];
    cx.bytes += 32;  // root rows
    const unsigned key_floor = ordered_key(-INFINITY);
    while (true) {
        const int nchild = c.z & 0xFF;
        if (nchild == 0 || (h.z & kTerminal)) break;
        h.z += 1;  // node.virtual_loss += 1 (mcts.py:416)
        if (cx.lane == 0) reinterpret_cast<int*>(&cx.hot[node])[2] = h.z;
        const float sq = __fsqrt_rn((float)h.x);  // == f32(math.sqrt(N)) (double rounding is innocuous)
        const int fc = c.y;
        cx.bytes += 32u * (unsigned)nchild;  // hot + cold row of every child scanned
        unsigned best_key = key_floor;
        int best_i = -1;
        int4 bh = h, bc = c;
        for (int base = 0; base < nchild; base += 32) {
            const int i = base + cx.lane;
            int4 ch = make_int4(0, 0, 0, 0), cc = make_int4(0, 0, 0, 0);
            unsigned key = 0;
            if (i < nchild) {
                ch = cx.hot[fc + i];
                cc = cx.cold[fc + i];
                float score;
                if (ch.x == 0) {
                    score = INFINITY;  // mcts.py:96-97
                } else if (ch.z & kCacheValid) {
                    score = __int_as_float(ch.w);  // mcts.py:99-100 (stale until the child is backed up)
                } else {
                    score = score_child(ch.x, __int_as_float(ch.y), ch.z & kVLMask, __int_as_float(cc.x),
                                        (cc.z >> 16) & 3, cx.c_puct, sq);
                    ch.w = __float_as_int(score);
                    ch.z |= kCacheValid;
                    reinterpret_cast<int2*>(&cx.hot[fc + i])[1] = make_int2(ch.z, ch.w);
                }
                key = (score == score) ? ordered_key(score) : 0u;
            }
            const unsigned mx = __reduce_max_sync(kFull, key);
            if (mx > best_key) {  // strict: an earlier chunk keeps ties
                const int src = __ffs(__ballot_sync(kFull, key == mx)) - 1;
                best_key = mx;
                best_i = base + src;
                bh = shfl4(ch, src);
                bc = shfl4(cc, src);
            }
        }
        if (best_i < 0) { cx.overflow |= 2; break; }  // reference would raise (next_node is None)
        coop_apply_move(cx.dir, b, (bc.z >> 8) & 0xFF);  // game.make_move(*next_move) (mcts.py:439)
        ++cx.steps;
        node = fc + best_i;
        h = bh;
        c = bc;
        if (plen < kMaxPath) {
            if (cx.lane == (plen & 31)) { if (plen < 32) p0 = node; else p1 = node; }
            ++plen;
        } else {
            cx.overflow |= 4;
            break;
        }
    }
    leaf_vlf = h.z;
    __syncwarp();
    return node;
}

// node.expand (mcts.py:141-161, 605-618): children for the legal squares in ascending order.
// prior(sq) is supplied by the caller (uniform constant or a probs row).
template <typename PriorFn>
__device__ __forceinline__ void expand_node(TreeCtx& cx, int node, uint64_t lm, PriorFn prior) {
    int4 c = cx.cold[node];
    if ((c.z & 0xFF) != 0) return;  // 'if action not in self.children' (mcts.py:154)
    const int nc = popc64(lm);
    if (cx.n_nodes + nc > cx.cap) { cx.overflow |= 1; return; }
    const int fc = cx.n_nodes;
    const int turn = 3 - ((c.z >> 16) & 3);  // mcts.py:618 (flips even after an auto-pass)
    for (int i = cx.lane; i < nc; i += 32) {
        const int sq = nth_set_bit(lm, i);
        cx.hot[fc + i] = make_int4(0, 0, 0, 0);
        cx.cold[fc + i] = make_int4(__float_as_int(prior(sq)), -1, (sq << 8) | (turn << 16), 0);
    }
    if (cx.lane == 0) {
        c.y = fc;
        c.z = (c.z & ~0xFF) | nc;
        cx.cold[node] = c;
    }
    cx.n_nodes += nc;
    cx.created += (unsigned)nc;
    cx.bytes += 32u * (unsigned)nc;  // rows written
    __syncwarp();
}

__device__ __forceinline__ void load_path(const WaveScratch& ws, int j, int lane, int& p0, int& p1, int& plen) {
    plen = ws.plen[j];
    p0 = ws.path[j * kMaxPath + lane];
    p1 = ws.path[j * kMaxPath + 32 + lane];
}
__device__ __forceinline__ void store_leaf(const WaveScratch& ws, int j, int lane, int node, int p0, int p1, int plen,
                                           const Board& b) {
    ws.path[j * kMaxPath + lane] = p0;
    ws.path[j * kMaxPath + 32 + lane] = p1;
    if (lane == 0) {
        ws.node[j] = node;
        ws.plen[j] = plen;
        ws.black[j] = b.black;
        ws.white[j] = b.white;
        ws.sf[j] = (uint16_t)(b.side | (b.flags << 8));
    }
}

// One wave of selections (mcts.py:355-386): k traversals from the root; a traversal that ends
// on a terminal-flagged node is backed up at once (mcts.py:364-366) and leaves slot j empty.
__device__ __forceinline__ void select_wave(TreeCtx& cx, const Board& root, const WaveScratch& ws, int k) {
    const CoopBoard root_c = coop_load(cx.dir, root);
    for (int j = 0; j < k; ++j) {
        CoopBoard b = root_c;
        int p0, p1, plen, vlf;
        const int node = select_one(cx, b, p0, p1, plen, vlf);
        ++cx.sims;
        if (vlf & kTerminal) {
            backup_path(cx, p0, p1, plen, term_value_of(vlf));
            if (cx.lane == 0) ws.node[j] = -1;
        } else {
            store_leaf(ws, j, cx.lane, node, p0, p1, plen, coop_store(cx.dir, b));
        }
    }
    __syncwarp();
}

// terminal leaf found at evaluation time (mcts.py:567-579): flag it with its ABSOLUTE value
// (+1 black won, -1 white won, 0 draw / not over) and back that value up
__device__ __forceinline__ void mark_terminal_and_backup(TreeCtx& cx, int node, int flags, int p0, int p1, int plen) {
    const int w = (flags & F_WIN_MASK) >> F_WIN_SHIFT;
    const int code = !(flags & F_OVER) ? 0 : (w == 1 ? 1 : (w == 2 ? 2 : 0));
    if (cx.lane == 0) {
        int* z = &reinterpret_cast<int*>(&cx.hot[node])[2];
        *z = (*z & ~(3 << kTermShift)) | kTerminal | (code << kTermShift);
    }
    __syncwarp();
    backup_path(cx, p0, p1, plen, code == 1 ? 1.0f : (code == 2 ? -1.0f : 0.0f));
}

// MCTS._process_batch (mcts.py:544-623) for one game's wave.  ws.lm / ws.val must be filled
// for every slot with node >= 0.  Pass 1 marks and backs up terminal leaves (ABSOLUTE value,
// mcts.py:567-579), pass 2 expands and backs up the evaluated ones, both in slot order.
template <typename PriorFn>
__device__ __forceinline__ void process_wave(TreeCtx& cx, const WaveScratch& ws, int k, PriorFn prior_for_slot) {
    for (int j = 0; j < k; ++j) {
        const int node = ws.node[j];
        if (node < 0 || ws.lm[j] != 0) continue;
        int p0, p1, plen;
        load_path(ws, j, cx.lane, p0, p1, plen);
        mark_terminal_and_backup(cx, node, ws.sf[j] >> 8, p0, p1, plen);
    }
    for (int j = 0; j < k; ++j) {
        const int node = ws.node[j];
        if (node < 0) continue;
        const uint64_t lm = ws.lm[j];
        if (lm == 0) continue;
        ++cx.evals;
        expand_node(cx, node, lm, [&](int sq) { return prior_for_slot(j, sq); });
        int p0, p1, plen;
        load_path(ws, j, cx.lane, p0, p1, plen);
        backup_path(cx, p0, p1, plen, ws.val[j]);
    }
}

// ================================================================================================
// RVS_MODE_FAST: virtual-loss PUCT with effective per-wave leaf batching.  NOT reference behaviour (the
// reference sends all simulations of a wave down one path: src/mcts/mcts.py:96-100,113,355-392); the
// specification is restated independently in oracle/rvs_oracle.c (mcts_search_fast) and these kernels
// match it bit for bit.  Differences from the functions above: no +inf for unvisited children, no score
// cache, virtual losses count as losses in q and as visits in u, the LEAF gets a virtual loss too, W is
// kept from the perspective of the player who moved into the node (actual movers, auto-passes respected),
// priors are bf16-rounded, and the first wave of a search is a single simulation (root expansion).
// ================================================================================================
__device__ __forceinline__ float score_child_fast(int N, float W, int VL, float P, float c_puct, float sq) {
    const int n = N + VL;
    const float q = n > 0 ? __fdiv_rn(__fsub_rn(W, (float)VL), (float)n) : 0.0f;
    float u = __fmul_rn(c_puct, P);
    u = __fmul_rn(u, sq);
    u = __fdiv_rn(u, (float)(1 + n));
    return __fadd_rn(q, u);
}

// v_black: the value from BLACK's perspective; path node d receives it from the perspective of the player
// who moved into it = the side to move at node d-1 (the root: its own side to move)
__device__ __forceinline__ void backup_path_fast(TreeCtx& cx, int p0, int p1, int plen, uint64_t sides, float v_black) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int d = cx.lane + 32 * half;
        if (d < plen) {
            const int n = half ? p1 : p0;
            int4 h = cx.hot[n];
            const bool white = (sides >> (d > 0 ? d - 1 : 0)) & 1ULL;
            h.x += 1;
            h.y = __float_as_int(__fadd_rn(__int_as_float(h.y), white ? -v_black : v_black));
            int vl = h.z & kVLMask;
            if (vl > 0) --vl;
            h.z = (h.z & ~kVLMask) | vl;
            cx.hot[n] = h;
        }
    }
    cx.bytes += 32u * (unsigned)plen;
    __syncwarp();
}

__device__ __forceinline__ int select_one_fast(TreeCtx& cx, CoopBoard& b, int& p0, int& p1, int& plen, int& leaf_vlf,
                                               uint64_t& sides) {
    int node = 0;
    p0 = 0; p1 = 0;
    plen = 1;
    sides = b.side == 2 ? 1ULL : 0ULL;
    int4 h = cx.hot[0];
    int4 c = cx.cold[0];
    cx.bytes += 32;
    const unsigned key_floor = ordered_key(-INFINITY);
    while (true) {
        const int nchild = c.z & 0xFF;
        if (nchild == 0 || (h.z & kTerminal)) break;
        h.z += 1;  // VL[node] += 1
        if (cx.lane == 0) reinterpret_cast<int*>(&cx.hot[node])[2] = h.z;
        const float sq = __fsqrt_rn((float)(h.x + (h.z & kVLMask)));
        const int fc = c.y;
        cx.bytes += 32u * (unsigned)nchild;
        unsigned best_key = key_floor;
        int best_i = -1;
        int4 bh = h, bc = c;
        for (int base = 0; base < nchild; base += 32) {
            const int i = base + cx.lane;
            int4 ch = make_int4(0, 0, 0, 0), cc = make_int4(0, 0, 0, 0);
            unsigned key = 0;
            if (i < nchild) {
                ch = cx.hot[fc + i];
                cc = cx.cold[fc + i];
                const float score = score_child_fast(ch.x, __int_as_float(ch.y), ch.z & kVLMask, __int_as_float(cc.x), cx.c_puct, sq);
                key = (score == score) ? ordered_key(__fadd_rn(score, 0.0f)) : 0u;
            }
            const unsigned mx = __reduce_max_sync(kFull, key);
            if (mx > best_key) {
                const int src = __ffs(__ballot_sync(kFull, key == mx)) - 1;
                best_key = mx;
                best_i = base + src;
                bh = shfl4(ch, src);
                bc = shfl4(cc, src);
            }
        }
        if (best_i < 0) { cx.overflow |= 2; break; }
        coop_apply_move(cx.dir, b, (bc.z >> 8) & 0xFF);
        ++cx.steps;
        node = fc + best_i;
        h = bh;
        c = bc;
        if (plen < kMaxPath) {
            if (cx.lane == (plen & 31)) { if (plen < 32) p0 = node; else p1 = node; }
            if (b.side == 2) sides |= 1ULL << plen;
            ++plen;
        } else {
            cx.overflow |= 4;
            break;
        }
    }
    h.z += 1;  // the leaf carries a virtual loss too: later simulations of the wave avoid it
    if (cx.lane == 0) reinterpret_cast<int*>(&cx.hot[node])[2] = h.z;
    leaf_vlf = h.z;
    __syncwarp();
    return node;
}

__device__ __forceinline__ float term_black_value(int vlf) { return term_value_of(vlf); }  // codes are absolute: +1 black, -1 white

__device__ __forceinline__ void select_wave_fast(TreeCtx& cx, const Board& root, const WaveScratch& ws, int k) {
    const CoopBoard root_c = coop_load(cx.dir, root);
    for (int j = 0; j < k; ++j) {
        CoopBoard b = root_c;
        int p0, p1, plen, vlf;
        uint64_t sides;
        const int node = select_one_fast(cx, b, p0, p1, plen, vlf, sides);
        ++cx.sims;
        if (vlf & kTerminal) {
            backup_path_fast(cx, p0, p1, plen, sides, term_black_value(vlf));
            if (cx.lane == 0) ws.node[j] = -1;
        } else {
            store_leaf(ws, j, cx.lane, node, p0, p1, plen, coop_store(cx.dir, b));
            if (cx.lane == 0) ws.sides[j] = sides;
        }
    }
    __syncwarp();
}

__device__ __forceinline__ float bf16_round_f32(float x) {  // round-to-nearest-even to bf16 precision (finite inputs)
    unsigned u = __float_as_uint(x);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return __uint_as_float(u & 0xFFFF0000u);
}

// the FAST counterpart of process_wave(): ws.val[j] is the evaluator's value from the LEAF MOVER's perspective
template <typename PriorFn>
__device__ __forceinline__ void process_wave_fast(TreeCtx& cx, const WaveScratch& ws, int k, PriorFn prior_for_slot) {
    for (int j = 0; j < k; ++j) {
        const int node = ws.node[j];
        if (node < 0 || ws.lm[j] != 0) continue;
        int p0, p1, plen;
        load_path(ws, j, cx.lane, p0, p1, plen);
        const int flags = ws.sf[j] >> 8;
        const int w = (flags & F_WIN_MASK) >> F_WIN_SHIFT;
        const int code = !(flags & F_OVER) ? 0 : (w == 1 ? 1 : (w == 2 ? 2 : 0));
        if (cx.lane == 0) {
            int* z = &reinterpret_cast<int*>(&cx.hot[node])[2];
            *z = (*z & ~(3 << kTermShift)) | kTerminal | (code << kTermShift);
        }
        __syncwarp();
        backup_path_fast(cx, p0, p1, plen, ws.sides[j], code == 1 ? 1.0f : (code == 2 ? -1.0f : 0.0f));
    }
    for (int j = 0; j < k; ++j) {
        const int node = ws.node[j];
        if (node < 0) continue;
        const uint64_t lm = ws.lm[j];
        if (lm == 0) continue;
        ++cx.evals;
        expand_node(cx, node, lm, [&](int sq) { return bf16_round_f32(prior_for_slot(j, sq)); });
        int p0, p1, plen;
        load_path(ws, j, cx.lane, p0, p1, plen);
        const float v = ws.val[j];
        backup_path_fast(cx, p0, p1, plen, ws.sides[j], (ws.sf[j] & 0xFF) == 1 ? v : -v);
    }
}

}  // namespace rvs
