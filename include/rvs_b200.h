/*
 * rvs_b200.h -- C ABI of the B200-native Reversi self-play hot path (librvs_b200.so).
 *
 * The reference (RandomMike1280/AlphaZero-Reversi) has no plugin / FFI interface on this
 * path: the boundary is its Python class API (ReversiGame, MCTS, SelfPlay).  Each entry
 * point below names the reference method(s) it replaces (paths relative to the reference
 * root); the Python mirror classes in alphazero-reversi_b200/ bind them with ctypes, and
 * INTEGRATION.md shows the same binding added to the reference itself.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; rvs_last_error() gives a
 *     thread-local message.  No C++ exception crosses the boundary.
 *   - `mem` selects where the caller's bulk pointers live: RVS_MEM_DEVICE (CUDA device
 *     pointers, e.g. torch.Tensor.data_ptr()) or RVS_MEM_HOST (plain host memory; the
 *     library stages through pinned buffers and the copies run on `stream`).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls with
 *     RVS_MEM_HOST outputs synchronise the stream before returning; device-only calls
 *     are asynchronous.
 *   - the caller owns every buffer it passes; engine handles own their internal HBM
 *     pools.  A handle is bound to one device and is not thread-safe.
 *   - squares are indices row*8+col (bit index of the reference bitboards,
 *     src/game/board.py:49,170); side 1 = BLACK, 2 = WHITE (board.py:22-23).
 */
#ifndef RVS_B200_H
#define RVS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVS_MEM_DEVICE 0
#define RVS_MEM_HOST 1
/* host pointers in PINNED memory, copies only enqueued on `stream`: the call returns without
 * synchronising, so several engine handles can be pipelined on several streams; the caller
 * synchronises the stream before reading outputs / reusing inputs (honoured by
 * rvs_engine_set_positions and rvs_engine_root_visits; every other entry point treats it exactly
 * like RVS_MEM_HOST: staged copies, outputs complete on return) */
#define RVS_MEM_HOST_ASYNC 2

#define RVS_RULES_REF 0    /* bug-compatible with src/game/board.py (graded) */
#define RVS_RULES_STRICT 1 /* true Othello */

/* flags byte of a position: bit0 game over, bits1-2 winner (0 draw, 1 black, 2 white),
 * bit3 the last move was followed by an auto-pass (board.py:242-249) */
#define RVS_FLAG_OVER 1
#define RVS_FLAG_WINNER_SHIFT 1
#define RVS_FLAG_WINNER_MASK 6
#define RVS_FLAG_PASSED 8

#define RVS_PLANES_F32_NCHW 0  /* [n,3,8,8] float32: ReversiGame.get_canonical_state */
#define RVS_PLANES_BF16_NHWC 1 /* [n,8,8,16] bf16, channels 3..15 zero: K4 network input */

#define RVS_EVAL_E0 0       /* deterministic: prior f32(1/65), value (own-opp)/64 */
#define RVS_EVAL_ROLLOUT 1  /* uniform prior, value = one uniform random playout */
#define RVS_EVAL_EXTERNAL 2 /* caller evaluates leaves (any model.predict duck type) */
#define RVS_EVAL_NN 3       /* built-in bf16 ResNet on tcgen05 (K4) */

/* wave semantics of the search (RVS_OPT_SEARCH_MODE) */
#define RVS_MODE_REF 0  /* the reference's waves, bugs included (src/mcts/mcts.py:96-100,113,355-392): graded */
#define RVS_MODE_FAST 1 /* effective virtual-loss leaf batching: the simulations of a wave spread over distinct leaves */

/* rvs_engine_set_option keys.  Only RVS_OPT_SEARCH_MODE changes results. */
#define RVS_OPT_LANES_PER_GAME 1 /* = rvs_engine_set_lanes_per_game */
#define RVS_OPT_NET_GRAPH 2      /* RVS_EVAL_NN: waves 2.. of a search replay one captured CUDA graph (0 / 1, default 0) */
#define RVS_OPT_SEARCH_MODE 3    /* RVS_MODE_REF (default) / RVS_MODE_FAST */
#define RVS_OPT_GAME_LIMIT 4     /* self-play: a finished slot restarts only while its next game id stays below this
                                    value (0 = no limit); with n slots, game ids 0 .. limit-1 are each played exactly once */
#define RVS_OPT_NET_MAX_CTAS 5   /* RVS_EVAL_NN: cap of the persistent tcgen05 grids (0 = all SMs); leaves SMs to the tree
                                    kernels of the other half-batch when a search is pipelined */
#define RVS_OPT_NET_PIPELINE 6   /* RVS_EVAL_NN, wave 1: 1 (default) = two half-batches ping-pong on two streams, so that the
                                    tree step and the heads of one half run beside the whole-network kernel of the other
                                    (128 filters, >= 2048 games; the tensor-core grid then leaves 12 SMs free unless
                                    RVS_OPT_NET_MAX_CTAS says otherwise).  0 = one lockstep batch per wave.  Identical results. */
#define RVS_OPT_NET_TOWER 7      /* RVS_EVAL_NN, 128 filters: 1 = first layer + residual tower + head planes run as ONE persistent
                                    launch (default), 0 = one launch per layer.  Bit-identical results. */

const char *rvs_last_error(void);
int rvs_version(void);
/* number of kernels this library has launched since load (bench.py "gpu_launches") */
int64_t rvs_launch_count(void);

/* ---- K1: stateless board operations ------------------------------------------------ */

/* Board.get_valid_moves (src/game/board.py:70-133) for n positions -> legal bit masks */
int rvs_legal_masks(const uint64_t *black, const uint64_t *white, const uint8_t *side,
                    uint64_t *out_mask, int64_t n, int rules, int mem, void *stream);

/* flip scan of Board.make_move / _get_flipped_pieces (board.py:190-219, 295-348) */
int rvs_flip_masks(const uint64_t *black, const uint64_t *white, const uint8_t *side,
                   const uint8_t *move, uint64_t *out_flip, int64_t n, int rules, int mem,
                   void *stream);

/* ReversiGame.make_move (src/game/game.py:36-70 -> board.py:135-251) in place on n
 * positions: ok[i]=0 and the position is untouched when the game is over or the square is
 * not in the legal mask.  out_next_legal (optional) = legal mask of the side to move next. */
int rvs_apply_moves(uint64_t *black, uint64_t *white, uint8_t *side, uint8_t *flags,
                    const uint8_t *move, uint8_t *ok, uint64_t *out_next_legal, int64_t n,
                    int rules, int mem, void *stream);

/* n uniform-random games from the start position (BASELINE config 1); game g uses RNG
 * stream (seed, first_game+g).  Any output pointer may be NULL.  *out_total_plies (host)
 * receives the number of board-steps played. */
int rvs_random_playouts(int64_t n_games, uint64_t seed, uint64_t first_game, int rules,
                        uint64_t *out_black, uint64_t *out_white, uint8_t *out_winner,
                        uint8_t *out_plies, int64_t *out_total_plies, int mem, void *stream);

/* perft under the reference's auto-pass semantics; *out_count is a host pointer */
int rvs_perft(uint64_t black, uint64_t white, int side, int depth, int rules,
              uint64_t *out_count, void *stream);

/* ---- K3: leaf encoding ------------------------------------------------------------- */

/* ReversiGame.get_canonical_state (game.py:131-162) */
int rvs_encode_planes(const uint64_t *black, const uint64_t *white, const uint8_t *side,
                      void *out, int64_t n, int layout, int rules, int mem, void *stream);

/* ---- K2(+K3+K4): lockstep batched MCTS engine --------------------------------------- */

typedef struct rvs_engine rvs_engine;

typedef struct rvs_engine_config {
    int32_t struct_size; /* sizeof(rvs_engine_config) */
    int32_t device;
    int32_t n_games;        /* concurrent game slots on this device */
    int32_t max_sims;       /* upper bound of num_simulations (sizes the node pools) */
    int32_t max_wave;       /* upper bound of the per-game wave (MCTS batch_size) */
    int32_t rules;
    int32_t evaluator;
    float c_puct;           /* MCTS(c_puct=...)  src/mcts/mcts.py:197 */
    uint64_t seed;
    int32_t nodes_per_game; /* 0 = worst case 2 + 34*max_sims */
    int32_t net_blocks;     /* RVS_EVAL_NN: AlphaZeroNetwork(num_res_blocks, num_filters) */
    int32_t net_filters;
    int32_t sample_capacity; /* self-play sample ring, 0 = 64*n_games */
} rvs_engine_config;

typedef struct rvs_engine_stats {
    int64_t sims;        /* simulations run (root-to-leaf traversals) */
    int64_t evals;       /* leaf evaluations requested from the evaluator */
    int64_t board_steps; /* moves applied (tree descent + rollouts + played moves) */
    int64_t nodes;       /* child nodes created by expansions */
    int64_t games_finished;
    int64_t samples;     /* samples recorded */
    int64_t launches;    /* kernels launched by this engine */
    int64_t overflow;    /* node-pool / path overflows (must stay 0) */
    int64_t tree_bytes;  /* algorithmic HBM bytes of the tree kernels: 32 B per node row touched */
    int64_t samples_dropped; /* samples lost because the ring was full (drain more often) */
    int64_t stalled;     /* slots parked after an illegal move choice (num_sims <= wave hazard) */
    int64_t nn_evals;    /* RVS_EVAL_NN: boards actually run through the network (terminal / duplicate leaves of a wave are compacted away) */
    int64_t bad_positions; /* positions rejected by rvs_engine_set_positions from DEVICE memory (side not 1/2 or overlapping discs): their slots are parked */
} rvs_engine_stats;

int rvs_engine_create(const rvs_engine_config *cfg, rvs_engine **out);
int rvs_engine_destroy(rvs_engine *h);

/* every slot back to the start position with a fresh game id (ReversiGame(), game.py:14-26) */
int rvs_engine_reset(rvs_engine *h, void *stream);
/* root positions for slots [0,n): MCTS.search(game) takes the caller's game (mcts.py:322).
 * Every position must have side in {1,2} and disjoint disc sets: host inputs are validated before the upload
 * (error -1), device inputs on the device (bad slots are parked and counted in stats.bad_positions).
 * RNG streams: slot g of the e-th set_positions call on a handle (e = 0 after create / reset) plays game id
 * g + e * n_games, so successive searches through this entry point draw independent rollout / noise streams. */
int rvs_engine_set_positions(rvs_engine *h, const uint64_t *black, const uint64_t *white,
                             const uint8_t *side, int32_t n, int mem, void *stream);
int rvs_engine_get_positions(rvs_engine *h, uint64_t *black, uint64_t *white, uint8_t *side,
                             uint8_t *flags, int32_t n, int mem, void *stream);

/* MCTS.search (src/mcts/mcts.py:322-407) for all slots with a built-in evaluator:
 * fresh root, num_sims simulations in waves of `wave` (= MCTS batch_size). */
int rvs_engine_search(rvs_engine *h, int32_t num_sims, int32_t wave, void *stream);

/* The same search split at the evaluator (RVS_EVAL_EXTERNAL): begin -> repeat { select k
 * leaves per game; caller evaluates rvs_engine_leaf_planes(); process } .  This is
 * MCTS._traverse (mcts.py:409-444) / MCTS._process_batch (mcts.py:544-623). */
int rvs_engine_begin_search(rvs_engine *h, void *stream);
int rvs_engine_select(rvs_engine *h, int32_t k, void *stream);
/* canonical planes of the leaves selected last, [n_games*k, 3,8,8] f32 (slot-major); leaves
 * that need no evaluation (terminal hits) are all-zero.  out_valid (optional, uint8
 * [n_games*k]) marks leaves whose evaluation will be consumed. */
int rvs_engine_leaf_planes(rvs_engine *h, float *out_planes, uint8_t *out_valid, int mem,
                           void *stream);
/* probs [n_games*k,65] f32 = softmax(logits) (mcts.py:596), values [n_games*k] f32 */
int rvs_engine_process(rvs_engine *h, const float *probs, const float *values, int mem,
                       void *stream);

/* root child visit counts by square, [n,65] int32 (index 64 = pass, always 0): the dict
 * MCTS.search returns (mcts.py:406-407) */
int rvs_engine_root_visits(rvs_engine *h, int32_t *out, int32_t n, int mem, void *stream);

/* One self-play ply for every live slot (src/self_play/self_play.py:80-101): pi from the
 * root visits (MCTS.get_action_probs, mcts.py:660-676), move choice (argmax if
 * temperature==0 else inverse-CDF sampling), sample record, make_move; finished games get
 * z back-filled (self_play.py:117-126) and, if recycle!=0, the slot restarts.
 * out_moves (optional, uint8 [n_games]) receives the squares played (255 = idle slot). */
int rvs_engine_play(rvs_engine *h, float temperature, int recycle, uint8_t *out_moves, int mem,
                    void *stream);

/* Persistent self-play (SelfPlay.generate_games, self_play.py:66-131, with MCTS batch_size 1 and a
 * built-in evaluator): ONE launch in which every slot keeps playing plies -- search, move choice,
 * sample record, make_move, game end, recycling -- until `plies` game-plies have been played in
 * total.  Work conserving: slots do not wait for each other between plies.  Per-game results are
 * identical to repeating rvs_engine_search(num_sims, 1) + rvs_engine_play.  With RVS_EVAL_NN the network
 * evaluates a whole wave in one batch, so the call runs ceil(plies / n_games) lockstep rounds of exactly that. */
int rvs_engine_selfplay(rvs_engine *h, int32_t num_sims, float temperature, int64_t plies,
                        int recycle, void *stream);

/* completed-game samples in the trainer's format (self_play.py:72-77, pipeline.py:226-228):
 * states [cap,3,8,8] f32, pi [cap,65] f32, z [cap] f32.  Returns up to `capacity` samples
 * and removes them from the ring. */
int rvs_engine_drain_samples(rvs_engine *h, float *states, float *pi, float *z, int64_t capacity,
                             int64_t *out_count, int mem, void *stream);

/* The same samples PACKED, as the engine keeps them (277 B instead of 1032 B per sample): position
 * before the move as black / white bitboards + side to move (the reference's game_data
 * 'current_players', self_play.py:91), z (self_play.py:117-126) as int8 and pi [65] f32.  The
 * canonical planes are a pure function of (black, white, side): rvs_encode_planes.  This is the
 * replay-file / NCCL-gather format (alphazero-reversi_b200/replay.py). */
int rvs_engine_drain_packed(rvs_engine *h, uint64_t *black, uint64_t *white, uint8_t *side, int8_t *z,
                            float *pi, int64_t capacity, int64_t *out_count, int mem, void *stream);

/* rvs_engine_drain_packed without any host synchronisation: DEVICE output buffers, and the number of samples
 * is written ON `stream` to *out_count_dev (device memory, or pinned host memory mapped into the device's address
 * space).  At most `capacity` samples are taken (the oldest first); the rest stay in the ring.  The caller orders
 * its consumers after this call on `stream` (or on an event recorded there): used to gather a generation's samples
 * on a side stream while the next generation is already searching. */
int rvs_engine_drain_packed_async(rvs_engine *h, uint64_t *black, uint64_t *white, uint8_t *side, int8_t *z,
                                  float *pi, int64_t capacity, int64_t *out_count_dev, void *stream);

/* Dirichlet noise on the root priors: P' = (1-eps) P + eps Dir(alpha), mixed in right after the root
 * is expanded by every following search (BASELINE config 4).  The reference only CONFIGURES this
 * (dirichlet_alpha / dirichlet_epsilon, src/config.py:25-26, src/self_play/self_play.py:18-47) and
 * never applies it, so it is off by default (epsilon = 0) and the sampling algorithm is this engine's
 * own (csrc/rvs_noise.cuh), restated by the oracle. */
int rvs_engine_set_root_noise(rvs_engine *h, double alpha, float epsilon);

/* Tuning knob of the wave-1 kernels (rvs_engine_search / rvs_engine_selfplay with a built-in evaluator):
 * lanes of a warp that cooperate on one game, 8 / 4 / 2 (0 = automatic: 8 up to 6144 games per handle,
 * 4 up to 24576, else 2).  Results never depend on it.  Worth setting to 4 when several handles are pipelined on one
 * GPU, i.e. when far more games are in flight than one handle holds. */
int rvs_engine_set_lanes_per_game(rvs_engine *h, int32_t lanes);

/* Engine options, see the RVS_OPT_* keys above. */
int rvs_engine_set_option(rvs_engine *h, int32_t option, int64_t value);

int rvs_engine_stats_get(rvs_engine *h, rvs_engine_stats *out, void *stream);

/* ---- K4: network -------------------------------------------------------------------- */

/* Loads AlphaZeroNetwork weights (src/model/network.py:33-69) given as the flat f32
 * concatenation of the state_dict tensors in canonical key order (see
 * alphazero-reversi_b200/network.py: pack_state_dict); BN is folded here. */
int rvs_engine_load_weights(rvs_engine *h, const float *flat, int64_t n_floats, int mem,
                            void *stream);
/* AlphaZeroNetwork.predict (network.py:136-158) on packed positions: logits [n,65] f32,
 * value [n] f32 (bf16 tensor-core compute, f32 accumulate) */
int rvs_engine_predict(rvs_engine *h, const uint64_t *black, const uint64_t *white,
                       const uint8_t *side, int64_t n, float *out_logits, float *out_value,
                       int mem, void *stream);
/* The same forward pass, returning what the built-in NN search consumes: probs [n,65] f32 =
 * softmax(logits) as computed by the engine's own head kernel (the reference applies F.softmax to the
 * logits, src/mcts/mcts.py:596) and value [n].  Feeding these to rvs_engine_process reproduces
 * rvs_engine_search(RVS_EVAL_NN) visit for visit. */
int rvs_engine_predict_probs(rvs_engine *h, const uint64_t *black, const uint64_t *white,
                             const uint8_t *side, int64_t n, float *out_probs, float *out_value,
                             int mem, void *stream);

#ifdef __cplusplus
}
#endif
#endif
