/*
 * rvs_oracle.c -- TEST INFRASTRUCTURE ONLY (see rvs_oracle.h).
 *
 * Plain-C restatement of the reference's board, MCTS and self-play semantics.
 * Compile with -ffp-contract=off: the tree arithmetic must be IEEE f32 with no FMA
 * contraction (numpy>=2 keeps python_float (op) np.float32 in float32).
 *
 * Pinned against the live reference through tests/golden/ (oracle/gen_golden.py).
 */
#include "rvs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define M64 0xFFFFFFFFFFFFFFFFULL
#define NOT_A 0xFEFEFEFEFEFEFEFEULL /* bit (r*8+c) with c != 0 */
#define NOT_H 0x7F7F7F7F7F7F7F7FULL /* c != 7 */

static __thread int64_t g_steps = 0; /* moves applied by this thread (bench accounting) */

static inline uint64_t sh(uint64_t x, int s) { return s > 0 ? (x << s) : (x >> (-s)); }
static inline int popc(uint64_t x) { return __builtin_popcountll(x); }

/* Board.__init__ (src/game/board.py:25-39) */
void orc_board_init(orc_board *b) {
    b->black = 0x0000000810000000ULL;
    b->white = 0x0000001008000000ULL;
    b->side = 1;
    b->over = 0;
    b->winner = 0;
    b->passes = 0;
}

/* Board.get_valid_moves (src/game/board.py:70-133): shifts in the reference's order
 * E,W,S,N,SE,NW,SW,NE = +1,-1,+8,-8,+9,-9,+7,-7; NO file masks in REF rules
 * (board.py:102-124), 1 seed step + 5 propagation steps + 1 landing step. */
static uint64_t legal_ref(uint64_t P, uint64_t O) {
    static const int S[8] = {1, -1, 8, -8, 9, -9, 7, -7};
    uint64_t E = ~(P | O) & M64, v = 0;
    for (int i = 0; i < 8; i++) {
        int s = S[i];
        uint64_t c = sh(P, s) & O;
        for (int k = 0; k < 5; k++) c |= sh(c, s) & O;
        v |= sh(c, s) & E;
    }
    return v;
}

/* true Othello: the source of every horizontal-component shift is masked */
static inline uint64_t sh_strict(uint64_t x, int s) {
    switch (s) {
    case 1: case 9: case -7: return sh(x & NOT_H, s);
    case -1: case -9: case 7: return sh(x & NOT_A, s);
    default: return sh(x, s);
    }
}
static uint64_t legal_strict(uint64_t P, uint64_t O) {
    static const int S[8] = {1, -1, 8, -8, 9, -9, 7, -7};
    uint64_t E = ~(P | O) & M64, v = 0;
    for (int i = 0; i < 8; i++) {
        int s = S[i];
        uint64_t c = sh_strict(P, s) & O;
        for (int k = 0; k < 5; k++) c |= sh_strict(c, s) & O;
        v |= sh_strict(c, s) & E;
    }
    return v;
}

uint64_t orc_legal(uint64_t P, uint64_t O, int rules) {
    return rules == ORC_RULES_STRICT ? legal_strict(P, O) : legal_ref(P, O);
}

/* Board.make_move flip scan (src/game/board.py:190-219): directions
 * [1,-1,8,-8,7,-7,9,-9]; mask looked up by abs(d) (board.py:208) so -1,-7,-9 get the
 * masks of +1,+7,+9; up to size-1 = 7 steps. */
static uint64_t flips_ref(uint64_t P, uint64_t O, uint64_t mv) {
    static const int D[8] = {1, -1, 8, -8, 7, -7, 9, -9};
    uint64_t f = 0;
    for (int i = 0; i < 8; i++) {
        int d = D[i], a = d < 0 ? -d : d;
        uint64_t m = a == 1 ? NOT_A : a == 7 ? NOT_A : a == 9 ? NOT_H : M64;
        uint64_t cur = mv, line = 0;
        for (int k = 0; k < 7; k++) {
            cur = sh(cur, d);
            if ((cur & O & m) == 0) break;
            line |= cur;
        }
        if (cur & P & m) f |= line;
    }
    return f;
}
static uint64_t flips_strict(uint64_t P, uint64_t O, uint64_t mv) {
    static const int D[8] = {1, -1, 8, -8, 7, -7, 9, -9};
    uint64_t f = 0;
    for (int i = 0; i < 8; i++) {
        int d = D[i];
        uint64_t cur = mv, line = 0;
        for (int k = 0; k < 7; k++) {
            cur = sh_strict(cur, d);
            if ((cur & O) == 0) break;
            line |= cur;
        }
        if (cur & P) f |= line;
    }
    return f;
}
uint64_t orc_flips(uint64_t P, uint64_t O, int idx, int rules) {
    uint64_t mv = 1ULL << idx;
    return rules == ORC_RULES_STRICT ? flips_strict(P, O, mv) : flips_ref(P, O, mv);
}

uint64_t orc_board_legal(const orc_board *b, int rules) {
    return b->side == 1 ? orc_legal(b->black, b->white, rules)
                        : orc_legal(b->white, b->black, rules);
}

/* ReversiGame.make_move -> Board.make_move (src/game/game.py:36-70,
 * src/game/board.py:135-251; winner board.py:363-373).  Legality is membership in the
 * legal mask (board.py:173-179), so zero-flip phantom moves are accepted in REF rules. */
int orc_apply(orc_board *b, int idx, int rules) {
    if (b->over) return 0; /* game.py:47-48 */
    if (idx < 0 || idx > 63) return 0;
    uint64_t P = b->side == 1 ? b->black : b->white;
    uint64_t O = b->side == 1 ? b->white : b->black;
    uint64_t mv = 1ULL << idx;
    if (!(orc_legal(P, O, rules) & mv)) return 0;
    uint64_t f = orc_flips(P, O, idx, rules);
    P ^= mv | f;
    O ^= f;
    g_steps++;
    if (b->side == 1) { b->black = P; b->white = O; } else { b->white = P; b->black = O; }
    b->side = (uint8_t)(3 - b->side);
    b->passes = 0;
    if (orc_board_legal(b, rules) == 0) { /* board.py:242-249 auto-pass */
        b->side = (uint8_t)(3 - b->side);
        b->passes = 1;
        if (orc_board_legal(b, rules) == 0) {
            int nb = popc(b->black), nw = popc(b->white);
            b->over = 1;
            b->winner = nb > nw ? 1 : nw > nb ? 2 : 0;
        }
    }
    return 1;
}

/* perft under the reference's move/auto-pass semantics: leaf = depth 0 or game over
 * (4, 12, 56, 244, 1396, 8200, 55134, 391210 in REF rules). */
uint64_t orc_perft(const orc_board *b, int depth, int rules) {
    if (depth == 0 || b->over) return 1;
    uint64_t lm = orc_board_legal(b, rules), n = 0;
    if (lm == 0) return 1;
    while (lm) {
        int idx = __builtin_ctzll(lm);
        lm &= lm - 1;
        orc_board c = *b;
        orc_apply(&c, idx, rules);
        n += orc_perft(&c, depth - 1, rules);
    }
    return n;
}

/* ReversiGame.get_canonical_state (src/game/game.py:131-162): plane 0 side-to-move
 * discs, plane 1 opponent discs, plane 2 legal mask; [row][col], bit = row*8+col. */
void orc_planes(const orc_board *b, int rules, float *out) {
    uint64_t P = b->side == 1 ? b->black : b->white;
    uint64_t O = b->side == 1 ? b->white : b->black;
    uint64_t L = orc_legal(P, O, rules);
    for (int i = 0; i < 64; i++) {
        out[i] = (float)((P >> i) & 1);
        out[64 + i] = (float)((O >> i) & 1);
        out[128 + i] = (float)((L >> i) & 1);
    }
}

/* ---- shared counter RNG (new-engine spec; no reference behaviour) ---- */
uint64_t orc_mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}
uint64_t orc_stream_seed(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t x = orc_mix64(seed + 0x9E3779B97F4A7C15ULL * (a + 1));
    return orc_mix64(x ^ (0xD1B54A32D192ED03ULL * (b + 1)));
}
static inline uint64_t rng_next(uint64_t *s) {
    *s += 0x9E3779B97F4A7C15ULL;
    return orc_mix64(*s);
}
static inline int rng_pick(uint64_t r, int n) { return (int)(((r >> 32) * (uint64_t)n) >> 32); }
static inline int nth_set_bit(uint64_t m, int k) {
    while (k--) m &= m - 1;
    return __builtin_ctzll(m);
}

/* playout move picker: PCG-RXS-M-XS-32 seeded from the 64-bit stream id (DESIGN.md "RNG") */
static inline uint32_t roll_next(uint32_t *s) {
    *s = *s * 747796405u + 2891336453u;
    uint32_t w = ((*s >> ((*s >> 28u) + 4u)) ^ *s) * 277803737u;
    return (w >> 22u) ^ w;
}

int orc_random_playout(orc_board *b, uint64_t stream, int rules) {
    int plies = 0;
    uint32_t rs = (uint32_t)(stream ^ (stream >> 32));
    while (!b->over) {
        uint64_t lm = orc_board_legal(b, rules);
        if (lm == 0) break; /* only for hand-made positions that never auto-passed */
        int k = (int)(((uint64_t)roll_next(&rs) * (uint64_t)popc(lm)) >> 32);
        orc_apply(b, nth_set_bit(lm, k), rules);
        plies++;
    }
    return plies;
}

void orc_random_playouts(int64_t n, uint64_t seed, int rules, uint64_t *black,
                         uint64_t *white, uint8_t *winner, uint8_t *plies) {
    for (int64_t g = 0; g < n; g++) {
        orc_board b;
        orc_board_init(&b);
        int p = orc_random_playout(&b, orc_stream_seed(seed, (uint64_t)g, 0), rules);
        black[g] = b.black; white[g] = b.white;
        winner[g] = b.winner; plies[g] = (uint8_t)p;
    }
}


/* ---------------- Dirichlet root noise (new-engine feature, no reference behaviour) ----------------
 * The reference threads dirichlet_alpha / dirichlet_epsilon through its config (src/config.py:25-26,
 * src/self_play/self_play.py:18-47) and never uses them.  This restates the ENGINE's
 * specification (alphazero-reversi_b200/csrc/rvs_noise.cuh header comment), written independently:
 * fixed-polynomial log / exp, every step one IEEE f64 operation (this file is built with
 * -ffp-contract=off), so the CUDA path can be compared bit for bit. */
static double g_noise_alpha = 0.0;
static float g_noise_eps = 0.0f;
void orc_set_root_noise(double alpha, float eps) { g_noise_alpha = alpha; g_noise_eps = eps; }

static double o_log(double x) {
    uint64_t u; memcpy(&u, &x, 8);
    int e = (int)((u >> 52) & 0x7FF) - 1023;
    u = (u & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL;
    double m; memcpy(&m, &u, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    static const double k[12] = {1.0 / 23.0, 1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0,
                                 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0, 1.0};
    double p = k[0];
    for (int i = 1; i < 12; i++) { p = p * s2; p = p + k[i]; }
    double lm = (2.0 * s) * p;
    return (double)e * 0.6931471805599453 + lm;
}
static double o_exp(double x) {
    if (!(x > -690.0)) return 0.0;
    double kf = floor(x * 1.4426950408889634 + 0.5);
    double r = x - kf * 0.6931471805599453;
    static const double c[14] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0,
                                 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0,
                                 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};
    double p = c[0];
    for (int i = 1; i < 14; i++) { p = p * r; p = p + c[i]; }
    uint64_t sb = (uint64_t)((int)kf + 1023) << 52;
    double sc; memcpy(&sc, &sb, 8);
    return p * sc;
}
static double o_unif(uint64_t *s) { return (double)((rng_next(s) >> 11) + 1ULL) * (1.0 / 9007199254740992.0); }
static double o_normal(uint64_t *s) {
    for (;;) {
        double v1 = 2.0 * o_unif(s) - 1.0;
        double v2 = 2.0 * o_unif(s) - 1.0;
        double q = v1 * v1 + v2 * v2;
        if (q >= 1.0 || q == 0.0) continue;
        return v1 * sqrt((-2.0 * o_log(q)) / q);
    }
}
static double o_log_gamma_dev(double a, uint64_t *s) {
    double a1 = a < 1.0 ? a + 1.0 : a;
    double d = a1 + (-1.0 / 3.0);
    double c = 1.0 / sqrt(9.0 * d);
    double lg;
    for (;;) {
        double z = o_normal(s);
        double t = 1.0 + c * z;
        if (t <= 0.0) continue;
        double v = (t * t) * t;
        double lu = o_log(o_unif(s));
        double lv = o_log(v);
        double zz = z * z;
        double rhs = 0.5 * zz;
        rhs = rhs + d;
        double dv = d * v;
        rhs = rhs - dv;
        double dl = d * lv;
        rhs = rhs + dl;
        if (lu < rhs) { lg = o_log(d) + lv; break; }
    }
    if (a < 1.0) lg = lg + o_log(o_unif(s)) / a;
    return lg;
}
/* eta[k] ~ Dirichlet(alpha), exported for the tests */
void orc_dirichlet(double alpha, int k, uint64_t stream, float *eta) {
    double lg[64];
    uint64_t s = stream;
    double mx = -1e300;
    for (int i = 0; i < k; i++) { lg[i] = o_log_gamma_dev(alpha, &s); if (lg[i] > mx) mx = lg[i]; }
    double sum = 0.0;
    for (int i = 0; i < k; i++) { lg[i] = o_exp(lg[i] - mx); sum = sum + lg[i]; }
    for (int i = 0; i < k; i++) eta[i] = (float)(lg[i] / sum);
}
double orc_det_log(double x) { return o_log(x); }
double orc_det_exp(double x) { return o_exp(x); }

/* ---------------- MCTS (src/mcts/mcts.py) ---------------- */
typedef struct {
    int32_t N;          /* visit_count      mcts.py:57 */
    float W;            /* value_sum        mcts.py:58 (f32 under numpy>=2) */
    float P;            /* prior            mcts.py:59 */
    float cache;        /* cached_ucb       mcts.py:69,113 */
    int32_t first_child, nchild;
    int32_t VL;         /* virtual_loss     mcts.py:65 */
    uint8_t move, turn, terminal, cache_valid;
    float term_value;   /* terminal_value   mcts.py:68 */
} onode;

typedef struct {
    onode *nodes;
    int n, cap;
    float c_puct;
} otree;

static int tree_new(otree *t, float prior, int turn, int move) {
    if (t->n >= t->cap) return -1;
    onode *x = &t->nodes[t->n];
    memset(x, 0, sizeof(*x));
    x->P = prior; x->turn = (uint8_t)turn; x->move = (uint8_t)move;
    x->first_child = -1;
    /* MCTSNode.__init__ sets cached_ucb=-inf (mcts.py:69) but ucb_score returns +inf
     * while N==0 and the attribute is deleted on the first backup (mcts.py:639-640),
     * so the initial value is never observed: model it as "no cache". */
    x->cache_valid = 0;
    return t->n++;
}

/* MCTSNode.ucb_score (mcts.py:84-114), f32 op order, no FMA */
static float ucb(otree *t, int ci, int parentN) {
    onode *c = &t->nodes[ci];
    if (c->N == 0) return INFINITY;
    if (c->cache_valid) return c->cache;
    int visits = c->N + c->VL;
    float q = c->W / (float)(c->N > 1 ? c->N : 1);
    float sq = (float)sqrt((double)parentN);
    float u = t->c_puct * c->P;
    u = u * sq;
    u = u / (float)(1 + visits);
    if (c->turn != 1) q = -q;
    c->cache = q + u;
    c->cache_valid = 1;
    return c->cache;
}

/* MCTS._backpropagate_path (mcts.py:625-640) */
static void backprop(otree *t, const int *path, int len, float v) {
    float sv = v;
    for (int i = len - 1; i >= 0; i--) {
        onode *x = &t->nodes[path[i]];
        if (x->VL > 0) x->VL--;
        x->N++;
        x->W = x->W + sv;
        sv = -sv;
        x->cache_valid = 0;
    }
}

static void eval_e0(const orc_board *lv, int n, float *probs, float *values) {
    /* logits==0 -> F.softmax gives f32(1/65) = 0x3C7C0FC1 for all 65 entries;
     * value = (own-opp)/64 decoded from planes 0/1 (SURVEY.md 8(c) evaluator E0) */
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 65; k++) probs[i * 65 + k] = 1.0f / 65.0f;
        int own = popc(lv[i].side == 1 ? lv[i].black : lv[i].white);
        int opp = popc(lv[i].side == 1 ? lv[i].white : lv[i].black);
        values[i] = (float)(own - opp) / 64.0f;
    }
}

/* ---------------- FAST search mode (engine feature; NOT reference behaviour) ----------------
 * Parity status: UNPINNED BY THE REFERENCE.  The reference's "batched" search sends every simulation of a
 * wave down one path (+inf for unvisited children whatever the virtual loss, stale cached UCB, backups only
 * after the wave: src/mcts/mcts.py:96-100,113,355-392), so a wave of K costs K evaluations of
 * about one unique leaf.  RVS_MODE_FAST is this engine's own specification of the textbook alternative --
 * virtual-loss PUCT -- and this function is its independent restatement; the CUDA kernels must match it bit
 * for bit (tests/test_gpu_fast.py).  Specification:
 *   - W of a node is the value sum from the perspective of the player who moved INTO it (the side to move at
 *     its parent, as actually played: auto-passes are respected); the root's W is from the root mover's side.
 *   - wave schedule: the first wave is ONE simulation (it expands the root); then waves of min(K, remaining).
 *   - descent while the node is expanded and not terminal: VL[node] += 1; Nt = N + VL of the node;
 *       score(child) = q + u,  n = N_c + VL_c,  q = n > 0 ? (W_c - f32(VL_c)) / f32(n) : 0,
 *       u = ((c_puct * P_c) * f32(sqrt(Nt))) / f32(1 + n)      (IEEE f32, no FMA, this operation order)
 *     first maximum in creation order wins; no score cache, no +inf.  The LEAF gets VL += 1 as well, so the
 *     following simulations of the wave are steered to other leaves.
 *   - a terminal-flagged leaf is backed up at once; after the wave, in selection order: a leaf without legal
 *     moves is flagged terminal with its outcome; the others are evaluated (one evaluation per distinct node),
 *     expanded if still unexpanded with priors rounded to bf16 (round-to-nearest-even), and backed up.
 *   - backup, leaf to root: VL -= 1 (if > 0), N += 1, W += value for the player who moved into the node
 *     (evaluator value v is from the leaf mover's perspective; terminal: +1 / -1 / 0 by the winner).
 *   - Dirichlet root noise as in the reference-compatible mode (after the root expansion).
 */
static int g_search_mode = 0;
void orc_set_search_mode(int mode) { g_search_mode = mode; }

static float bf16_round(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    u &= 0xFFFF0000u;
    memcpy(&x, &u, 4);
    return x;
}

typedef struct { int node; int plen; int path[64]; uint8_t side[64]; orc_board b; int sim; } fleaf;

static void backprop_fast(otree *t, const fleaf *L, float v_black) {
    for (int i = L->plen - 1; i >= 0; i--) {
        onode *x = &t->nodes[L->path[i]];
        int mover = i == 0 ? L->side[0] : L->side[i - 1]; /* who moved into path[i] */
        if (x->VL > 0) x->VL--;
        x->N++;
        x->W = x->W + (mover == 1 ? v_black : -v_black);
    }
}

static int mcts_search_fast(const orc_board *root, int num_sims, int wave, float c_puct, int rules,
                            int evaluator, orc_eval_fn fn, void *ctx, uint64_t seed, uint64_t game_id,
                            uint64_t search_id, int32_t *visits, int32_t *root_n, float *root_w,
                            int64_t *n_evals, int64_t *n_unique) {
    otree t;
    t.cap = 2 + num_sims * 34;
    t.nodes = (onode *)malloc(sizeof(onode) * (size_t)t.cap);
    t.n = 0;
    t.c_puct = c_puct;
    fleaf *leaves = (fleaf *)malloc(sizeof(fleaf) * (size_t)wave);
    orc_board *lb = (orc_board *)malloc(sizeof(orc_board) * (size_t)wave);
    float *probs = (float *)malloc(sizeof(float) * 65 * (size_t)wave);
    float *values = (float *)malloc(sizeof(float) * (size_t)wave);
    int *lidx = (int *)malloc(sizeof(int) * (size_t)wave);
    int *rep = (int *)malloc(sizeof(int) * (size_t)wave);
    int64_t evals = 0, unique = 0;
    int rc = 0;
    tree_new(&t, 1.0f, root->side, 255);
    for (int start = 0; start < num_sims && rc >= 0;) {
        int k = start == 0 ? 1 : (num_sims - start < wave ? num_sims - start : wave);
        int nleaf = 0;
        for (int j = 0; j < k; j++) {
            fleaf *L = &leaves[nleaf];
            L->b = *root;
            L->plen = 0;
            L->sim = start + j;
            int node = 0;
            L->side[0] = L->b.side;
            L->path[L->plen++] = 0;
            while (t.nodes[node].nchild > 0 && !t.nodes[node].terminal) {
                t.nodes[node].VL++;
                int Nt = t.nodes[node].N + t.nodes[node].VL;
                float sq = (float)sqrt((double)Nt);
                float best = -INFINITY;
                int next = -1;
                int fc = t.nodes[node].first_child, nc = t.nodes[node].nchild;
                for (int c = fc; c < fc + nc; c++) {
                    onode *x = &t.nodes[c];
                    int n = x->N + x->VL;
                    float q = n > 0 ? (x->W - (float)x->VL) / (float)n : 0.0f;
                    float u = t.c_puct * x->P;
                    u = u * sq;
                    u = u / (float)(1 + n);
                    float sc = q + u;
                    if (sc > best) { best = sc; next = c; }
                }
                if (next < 0) { rc = -2; break; }
                orc_apply(&L->b, t.nodes[next].move, rules);
                node = next;
                if (L->plen >= 64) { rc = -3; break; }
                L->side[L->plen] = L->b.side;
                L->path[L->plen++] = node;
            }
            if (rc < 0) break;
            L->node = node;
            t.nodes[node].VL++; /* the leaf carries a virtual loss too */
            if (t.nodes[node].terminal) {
                backprop_fast(&t, L, t.nodes[node].term_value);
                continue;
            }
            nleaf++;
        }
        if (rc < 0) break;
        int ne = 0;
        for (int i = 0; i < nleaf; i++) {
            fleaf *L = &leaves[i];
            uint64_t lm = orc_board_legal(&L->b, rules);
            if (lm == 0) { /* flagged with the outcome from BLACK's perspective (0 when not over) */
                onode *x = &t.nodes[L->node];
                x->terminal = 1;
                x->term_value = !L->b.over ? 0.0f : L->b.winner == 1 ? 1.0f : L->b.winner == 2 ? -1.0f : 0.0f;
                backprop_fast(&t, L, x->term_value);
                continue;
            }
            rep[ne] = -1;
            for (int q = 0; q < ne; q++)
                if (leaves[lidx[q]].node == L->node) { rep[ne] = q; break; }
            lb[ne] = L->b;
            lidx[ne] = i;
            ne++;
        }
        if (ne > 0) {
            if (evaluator == ORC_EVAL_E0) {
                eval_e0(lb, ne, probs, values);
            } else if (evaluator == ORC_EVAL_ROLLOUT) {
                for (int i = 0; i < ne; i++) {
                    for (int q = 0; q < 65; q++) probs[i * 65 + q] = 1.0f / 65.0f;
                    orc_board c = lb[i];
                    uint64_t st = orc_stream_seed(seed, game_id, (search_id << 16) | (uint64_t)leaves[lidx[i]].sim);
                    orc_random_playout(&c, st, rules);
                    values[i] = (!c.over || c.winner == 0) ? 0.0f : (c.winner == lb[i].side ? 1.0f : -1.0f);
                }
            } else {
                fn(ctx, lb, ne, probs, values);
            }
            evals += ne;
            for (int i = 0; i < ne; i++) {
                fleaf *L = &leaves[lidx[i]];
                onode *x = &t.nodes[L->node];
                if (rep[i] < 0) unique++;
                if (x->nchild == 0) {
                    uint64_t lm = orc_board_legal(&L->b, rules);
                    int nc = popc(lm);
                    if (t.n + nc > t.cap) { rc = -1; break; }
                    int fc = t.n;
                    while (lm) {
                        int sq = __builtin_ctzll(lm);
                        lm &= lm - 1;
                        tree_new(&t, bf16_round(probs[i * 65 + sq]), 3 - x->turn, sq);
                    }
                    x = &t.nodes[L->node];
                    x->first_child = fc;
                    x->nchild = nc;
                }
                /* a duplicate leaf of the wave shares the evaluation of its first occurrence (rollout
                 * evaluators draw per simulation, so each copy keeps its own value there) */
                float v = (rep[i] >= 0 && evaluator != ORC_EVAL_ROLLOUT) ? values[rep[i]] : values[i];
                backprop_fast(&t, L, lb[i].side == 1 ? v : -v);
            }
        }
        if (start == 0 && g_noise_eps > 0.0f && t.nodes[0].nchild > 0 && t.nodes[0].nchild <= 64) {
            float eta[64];
            int fc = t.nodes[0].first_child, nc = t.nodes[0].nchild;
            orc_dirichlet(g_noise_alpha, nc, orc_stream_seed(seed, game_id, 0xD1000000ULL + search_id), eta);
            for (int c = 0; c < nc; c++) {
                float keep = (1.0f + (-g_noise_eps)) * t.nodes[fc + c].P;
                float add = g_noise_eps * eta[c];
                t.nodes[fc + c].P = keep + add;
            }
        }
        start += k;
    }
    for (int i = 0; i < 65; i++) visits[i] = 0;
    if (rc >= 0) {
        onode *r = &t.nodes[0];
        for (int c = r->first_child; c >= 0 && c < r->first_child + r->nchild; c++) visits[t.nodes[c].move] = t.nodes[c].N;
        if (root_n) *root_n = r->N;
        if (root_w) *root_w = r->W;
        rc = t.n;
    }
    if (n_evals) *n_evals = evals;
    if (n_unique) *n_unique = unique;
    free(t.nodes); free(leaves); free(lb); free(probs); free(values); free(lidx); free(rep);
    return rc;
}

int orc_mcts_search_fast(const orc_board *root, int num_sims, int wave, float c_puct, int rules, int evaluator,
                         orc_eval_fn fn, void *ctx, uint64_t seed, uint64_t game_id, uint64_t search_id,
                         int32_t *visits, int32_t *root_n, float *root_w, int64_t *n_evals, int64_t *n_unique) {
    if (wave < 1) wave = 1;
    return mcts_search_fast(root, num_sims, wave, c_puct, rules, evaluator, fn, ctx, seed, game_id, search_id, visits,
                            root_n, root_w, n_evals, n_unique);
}

typedef struct { int node; int plen; int path[64]; orc_board b; int sim; } oleaf;

int orc_mcts_search(const orc_board *root, int num_sims, int wave, float c_puct, int rules,
                    int evaluator, orc_eval_fn fn, void *ctx, uint64_t seed,
                    uint64_t game_id, uint64_t search_id, int32_t *visits, int32_t *root_n,
                    float *root_w, int64_t *n_evals) {
    if (g_search_mode == 1) /* orc_set_search_mode(1): the engine's FAST mode, see above */
        return orc_mcts_search_fast(root, num_sims, wave, c_puct, rules, evaluator, fn, ctx, seed, game_id, search_id,
                                    visits, root_n, root_w, n_evals, 0);
    otree t;
    t.cap = 2 + num_sims * 34;
    if (wave < 1) wave = 1;
    /* per-thread grow-only scratch (a malloc/free pair per search serialises host threads) */
    static __thread onode *tl_nodes = 0; static __thread int tl_cap = 0;
    static __thread oleaf *tl_leaves = 0; static __thread orc_board *tl_lb = 0;
    static __thread float *tl_probs = 0, *tl_values = 0; static __thread int *tl_lidx = 0;
    static __thread int tl_wave = 0;
    if (tl_cap < t.cap) { free(tl_nodes); tl_nodes = (onode *)malloc(sizeof(onode) * (size_t)t.cap); tl_cap = t.cap; }
    if (tl_wave < wave) {
        free(tl_leaves); free(tl_lb); free(tl_probs); free(tl_values); free(tl_lidx);
        tl_leaves = (oleaf *)malloc(sizeof(oleaf) * (size_t)wave);
        tl_lb = (orc_board *)malloc(sizeof(orc_board) * (size_t)wave);
        tl_probs = (float *)malloc(sizeof(float) * 65 * (size_t)wave);
        tl_values = (float *)malloc(sizeof(float) * (size_t)wave);
        tl_lidx = (int *)malloc(sizeof(int) * (size_t)wave);
        tl_wave = wave;
    }
    t.nodes = tl_nodes;
    t.n = 0;
    t.c_puct = c_puct;
    oleaf *leaves = tl_leaves;
    orc_board *lb = tl_lb;
    float *probs = tl_probs;
    float *values = tl_values;
    int *lidx = tl_lidx;
    int64_t evals = 0;
    int rc = 0;

    tree_new(&t, 1.0f, root->side, 255); /* mcts.py:334-341 */

    for (int start = 0; start < num_sims && rc >= 0; start += wave) { /* mcts.py:348-349 */
        int k = num_sims - start < wave ? num_sims - start : wave;
        int nleaf = 0;
        for (int j = 0; j < k; j++) { /* mcts.py:355-386 */
            oleaf *L = &leaves[nleaf];
            L->b = *root;
            L->plen = 0;
            L->sim = start + j;
            int node = 0;
            L->path[L->plen++] = 0;
            /* _traverse (mcts.py:409-444) */
            while (t.nodes[node].nchild > 0 && !t.nodes[node].terminal) {
                t.nodes[node].VL++;
                float best = -INFINITY;
                int next = -1;
                int fc = t.nodes[node].first_child, nc = t.nodes[node].nchild;
                int pn = t.nodes[node].N;
                for (int c = fc; c < fc + nc; c++) {
                    float s = ucb(&t, c, pn);
                    if (s > best) { best = s; next = c; }
                }
                if (next < 0) { rc = -2; break; } /* reference would raise here */
                orc_apply(&L->b, t.nodes[next].move, rules);
                node = next;
                if (L->plen >= 64) { rc = -3; break; }
                L->path[L->plen++] = node;
            }
            if (rc < 0) break;
            L->node = node;
            if (t.nodes[node].terminal) { /* mcts.py:364-366 */
                backprop(&t, L->path, L->plen, t.nodes[node].term_value);
                continue;
            }
            nleaf++;
        }
        if (rc < 0) break;
        /* _process_batch (mcts.py:544-623): pass 1 terminal detection + backup */
        int ne = 0;
        for (int i = 0; i < nleaf; i++) {
            oleaf *L = &leaves[i];
            uint64_t lm = orc_board_legal(&L->b, rules);
            if (lm == 0) { /* mcts.py:567-579: ABSOLUTE value, 0 if not over */
                onode *x = &t.nodes[L->node];
                x->terminal = 1;
                x->term_value = !L->b.over ? 0.0f
                                : L->b.winner == 1 ? 1.0f
                                : L->b.winner == 2 ? -1.0f : 0.0f;
                backprop(&t, L->path, L->plen, x->term_value);
                continue;
            }
            lb[ne] = L->b;
            lidx[ne] = i;
            ne++;
        }
        if (ne == 0) continue;
        if (evaluator == ORC_EVAL_E0) {
            eval_e0(lb, ne, probs, values);
        } else if (evaluator == ORC_EVAL_ROLLOUT) {
            for (int i = 0; i < ne; i++) {
                for (int q = 0; q < 65; q++) probs[i * 65 + q] = 1.0f / 65.0f;
                orc_board c = lb[i];
                uint64_t st = orc_stream_seed(seed, game_id,
                                              (search_id << 16) | (uint64_t)leaves[lidx[i]].sim);
                orc_random_playout(&c, st, rules);
                values[i] = (!c.over || c.winner == 0) ? 0.0f
                            : (c.winner == lb[i].side ? 1.0f : -1.0f);
            }
        } else {
            fn(ctx, lb, ne, probs, values);
        }
        evals += ne;
        /* pass 2: expand (mcts.py:600-618) + backup (mcts.py:623) */
        for (int i = 0; i < ne; i++) {
            oleaf *L = &leaves[lidx[i]];
            onode *x = &t.nodes[L->node];
            if (x->terminal) continue;
            if (x->nchild == 0) {
                uint64_t lm = orc_board_legal(&L->b, rules);
                int nc = popc(lm);
                if (t.n + nc > t.cap) { rc = -1; break; }
                int fc = t.n;
                int turn = 3 - x->turn; /* mcts.py:618: flips even after an auto-pass */
                while (lm) {
                    int sq = __builtin_ctzll(lm);
                    lm &= lm - 1;
                    tree_new(&t, probs[i * 65 + sq], turn, sq);
                }
                x = &t.nodes[L->node];
                x->first_child = fc;
                x->nchild = nc;
            }
            backprop(&t, L->path, L->plen, values[i]);
        }
        if (start == 0 && g_noise_eps > 0.0f && t.nodes[0].nchild > 0 && t.nodes[0].nchild <= 64) {
            /* engine feature: mix Dirichlet noise into the root priors right after the root expansion */
            float eta[64];
            int fc = t.nodes[0].first_child, nc = t.nodes[0].nchild;
            orc_dirichlet(g_noise_alpha, nc, orc_stream_seed(seed, game_id, 0xD1000000ULL + search_id), eta);
            for (int c = 0; c < nc; c++) {
                float keep = (1.0f + (-g_noise_eps)) * t.nodes[fc + c].P;
                float add = g_noise_eps * eta[c];
                t.nodes[fc + c].P = keep + add;
            }
        }
    }

    for (int i = 0; i < 65; i++) visits[i] = 0;
    if (rc >= 0) {
        onode *r = &t.nodes[0];
        for (int c = r->first_child; c >= 0 && c < r->first_child + r->nchild; c++)
            visits[t.nodes[c].move] = t.nodes[c].N;
        if (root_n) *root_n = r->N;
        if (root_w) *root_w = r->W;
        rc = t.n;
    }
    if (n_evals) *n_evals = evals;
    return rc;
}

/* numpy pairwise sum of 65 doubles (np.sum on a contiguous f64 array, n < 128:
 * 8 running accumulators, combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail) */
static double np_sum65(const double *a) {
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    for (int i = 8; i < 64; i += 8)
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    res += a[64];
    return res;
}

/* MCTS.get_action_probs (mcts.py:660-676) */
void orc_action_probs(const int32_t *visits, double temperature, double *pi) {
    long total = 0;
    for (int i = 0; i < 65; i++) { total += visits[i]; pi[i] = 0.0; }
    if (total > 0)
        for (int i = 0; i < 65; i++) pi[i] = (double)visits[i] / (double)total;
    int all_zero = 1;
    for (int i = 0; i < 65; i++) if (pi[i] != 0.0) all_zero = 0;
    if (temperature > 0 && !all_zero) {
        double tmp[65];
        double e = 1.0 / temperature;
        for (int i = 0; i < 65; i++) tmp[i] = e == 1.0 ? pi[i] : pow(pi[i], e);
        double s = np_sum65(tmp);
        for (int i = 0; i < 65; i++) pi[i] = tmp[i] / s;
    }
}

/* SelfPlay.generate_games inner loop (self_play.py:80-126) with the device move
 * sampler: T==0 or all-zero pi -> first argmax (mcts.py:679-681); otherwise inverse CDF
 * like np.random.choice (cumsum, normalise by cdf[-1], searchsorted side='right') on
 * u = (rng >> 11) * 2^-53 from stream_seed(seed, game_id, 0x80000000 + ply). */
int orc_self_play_game(int num_sims, int wave, float c_puct, int rules, int evaluator,
                       orc_eval_fn fn, void *ctx, uint64_t seed, uint64_t game_id,
                       double temperature, orc_sample *out, int max_plies, uint8_t *winner) {
    orc_board b;
    orc_board_init(&b);
    int ply = 0;
    while (!b.over && ply < max_plies) {
        orc_sample *s = &out[ply];
        /* rollout streams are keyed by (game, (ply << 16) | sim) */
        int rc = orc_mcts_search(&b, num_sims, wave, c_puct, rules, evaluator, fn, ctx, seed,
                                 game_id, (uint64_t)ply, s->visits, 0, 0, 0);
        if (rc < 0) return rc;
        double pi[65];
        orc_action_probs(s->visits, temperature, pi);
        int all_zero = 1, mv = 0;
        for (int i = 0; i < 65; i++) if (pi[i] != 0.0) all_zero = 0;
        if (temperature == 0.0 || all_zero) {
            double best = pi[0];
            for (int i = 1; i < 65; i++) if (pi[i] > best) { best = pi[i]; mv = i; }
        } else {
            uint64_t st = orc_stream_seed(seed, game_id, 0x80000000ULL + (uint64_t)ply);
            double u = (double)(rng_next(&st) >> 11) * (1.0 / 9007199254740992.0);
            double cdf[65], acc = 0.0;
            for (int i = 0; i < 65; i++) { acc += pi[i]; cdf[i] = acc; }
            for (int i = 0; i < 65; i++) cdf[i] /= acc;
            mv = 64;
            for (int i = 0; i < 65; i++) if (cdf[i] > u) { mv = i; break; }
        }
        s->black = b.black; s->white = b.white; s->side = b.side;
        s->move = (uint8_t)mv; s->z = 0; s->pad = 0;
        if (!orc_apply(&b, mv, rules)) return -10; /* reference would loop forever */
        ply++;
    }
    if (!b.over) return -11;
    for (int i = 0; i < ply; i++) /* self_play.py:117-126 */
        out[i].z = b.winner == 0 ? 0 : (out[i].side == b.winner ? 1 : -1);
    if (winner) *winner = b.winner;
    return ply;
}

/* bench helper: the same search over n root positions on the calling thread (the benchmark
 * splits a batch over host threads).  game ids game0..game0+n-1; *steps += moves applied. */
int orc_search_batch(const uint64_t *black, const uint64_t *white, const uint8_t *side, int n,
                     int num_sims, int wave, float c_puct, int rules, int evaluator, uint64_t seed,
                     uint64_t game0, uint64_t search_id, int32_t *visits, int64_t *evals,
                     int64_t *steps) {
    int64_t ev = 0, e1 = 0;
    g_steps = 0;
    for (int i = 0; i < n; i++) {
        orc_board b = {black[i], white[i], side[i], 0, 0, 0};
        if (orc_board_legal(&b, rules) == 0) {
            orc_board o = b;
            o.side = (uint8_t)(3 - b.side);
            if (orc_board_legal(&o, rules) == 0) {
                int nb = popc(b.black), nw = popc(b.white);
                b.over = 1;
                b.winner = nb > nw ? 1 : nw > nb ? 2 : 0;
            }
        }
        int rc = orc_mcts_search(&b, num_sims, wave, c_puct, rules, evaluator, 0, 0, seed,
                                 game0 + (uint64_t)i, search_id, visits + (size_t)i * 65, 0, 0, &e1);
        if (rc < 0) return rc;
        ev += e1;
    }
    if (evals) *evals = ev;
    if (steps) *steps = g_steps;
    return 0;
}
