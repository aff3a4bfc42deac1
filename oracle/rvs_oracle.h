/*
 * rvs_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's self-play hot path, used solely as the
 * parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing in the product package may include, link or call it.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  The oracle is pinned against the live Python reference by
 * oracle/gen_golden.py -> tests/golden/ (see tests/test_oracle_*.py).
 */
#ifndef RVS_ORACLE_H
#define RVS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_RULES_REF 0    /* bug-compatible with src/game/board.py (graded mode) */
#define ORC_RULES_STRICT 1 /* true Othello (file masks on every +-1/+-7/+-9 shift) */

#define ORC_EVAL_E0 0       /* logits==0 -> prior 1/65 (f32), value=(own-opp)/64 */
#define ORC_EVAL_ROLLOUT 1  /* uniform prior, value = one uniform random playout */
#define ORC_EVAL_CALLBACK 2 /* caller-supplied priors/values (table evaluators, torch models) */

typedef struct orc_board {
    uint64_t black, white;
    uint8_t side;   /* 1 = BLACK, 2 = WHITE (board.py:22-23) */
    uint8_t over;   /* Board.game_over */
    uint8_t winner; /* 0 draw, 1, 2 ; only meaningful when over */
    uint8_t passes; /* Board.passed_moves_in_a_row */
} orc_board;

/* evaluator callback: n leaves -> probs[n*65] (softmax output), values[n] */
typedef void (*orc_eval_fn)(void *ctx, const orc_board *leaves, int n, float *probs,
                            float *values);

void orc_board_init(orc_board *b);
uint64_t orc_legal(uint64_t P, uint64_t O, int rules);
uint64_t orc_flips(uint64_t P, uint64_t O, int idx, int rules);
/* returns 1 when the move was legal and applied, 0 otherwise (game.py:47-48,70) */
int orc_apply(orc_board *b, int idx, int rules);
uint64_t orc_board_legal(const orc_board *b, int rules);
uint64_t orc_perft(const orc_board *b, int depth, int rules);
/* canonical planes f32[3][8][8] (game.py:131-162) */
void orc_planes(const orc_board *b, int rules, float *out192);

/* shared counter RNG (specified in DESIGN.md "RNG") */
uint64_t orc_mix64(uint64_t x);
uint64_t orc_stream_seed(uint64_t seed, uint64_t a, uint64_t b);
/* one uniform random playout from b to the end; returns plies played.  */
int orc_random_playout(orc_board *b, uint64_t stream, int rules);
/* n games from the start position, stream = orc_stream_seed(seed, game, 0) */
void orc_random_playouts(int64_t n, uint64_t seed, int rules, uint64_t *black,
                         uint64_t *white, uint8_t *winner, uint8_t *plies);

/* MCTS.search (mcts.py:322-407).  visits[65] <- root child visit counts by square.
 * root_n / root_w (optional) <- root visit count / value sum.  Returns number of
 * nodes allocated, or <0 on pool overflow. */
int orc_mcts_search(const orc_board *root, int num_sims, int wave, float c_puct, int rules,
                    int evaluator, orc_eval_fn fn, void *ctx, uint64_t seed,
                    uint64_t game_id, uint64_t search_id, int32_t *visits, int32_t *root_n,
                    float *root_w, int64_t *n_evals);

/* RVS_MODE_FAST (engine feature, virtual-loss PUCT; specification in rvs_oracle.c).  PARITY UNPINNED BY THE
 * REFERENCE: the reference has no such mode, the oracle is the specification the CUDA kernels are held to.
 * orc_set_search_mode(1) routes orc_mcts_search / orc_self_play_game / orc_search_batch through it. */
void orc_set_search_mode(int mode);
int orc_mcts_search_fast(const orc_board *root, int num_sims, int wave, float c_puct, int rules, int evaluator,
                         orc_eval_fn fn, void *ctx, uint64_t seed, uint64_t game_id, uint64_t search_id,
                         int32_t *visits, int32_t *root_n, float *root_w, int64_t *n_evals, int64_t *n_unique);

/* Dirichlet noise on the root priors (engine feature, see rvs_oracle.c): applied by every following
 * orc_mcts_search / orc_self_play_game right after the root expansion; eps == 0 switches it off */
void orc_set_root_noise(double alpha, float eps);
void orc_dirichlet(double alpha, int k, uint64_t stream, float *eta);
double orc_det_log(double x);
double orc_det_exp(double x);

/* bench helper: orc_mcts_search over n roots on the calling thread */
int orc_search_batch(const uint64_t *black, const uint64_t *white, const uint8_t *side, int n,
                     int num_sims, int wave, float c_puct, int rules, int evaluator, uint64_t seed,
                     uint64_t game0, uint64_t search_id, int32_t *visits, int64_t *evals,
                     int64_t *steps);

/* MCTS.get_action_probs pi (mcts.py:660-676): f64 pi[65] from visit counts */
void orc_action_probs(const int32_t *visits, double temperature, double *pi65);

/* Full self-play of one game with deterministic argmax/sampled moves (self_play.py:51-145).
 * move choice: temperature==0 -> argmax(pi) first max; else inverse-CDF with the
 * shared RNG (device self-play spec).  Outputs per ply records. */
typedef struct orc_sample {
    uint64_t black, white;
    uint8_t side;
    int8_t z;
    uint8_t move;
    uint8_t pad;
    int32_t visits[65];
} orc_sample;
int orc_self_play_game(int num_sims, int wave, float c_puct, int rules, int evaluator,
                       orc_eval_fn fn, void *ctx, uint64_t seed, uint64_t game_id,
                       double temperature, orc_sample *out, int max_plies, uint8_t *winner);

#ifdef __cplusplus
}
#endif
#endif
