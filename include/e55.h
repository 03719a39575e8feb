/*
 * e55.h -- C ABI of the B200-native policy/value network evaluator (libe55.so).
 *
 * Drop-in boundary for ONE path of nyashiki/erweitern_55: the batched forward pass behind
 * `Network.predict` (reference erweitern_55/network.py:201-216), i.e. the graph built by
 * `_alphazero_network` / `_residual_block` (network.py:75-146), with weights exchanged in the Keras
 * `get_weights()` list format (network.py:228-243).  In the reference this boundary is
 * `keras.Model.predict` -> `tf.Session.run` (TensorFlow 1.15.2, not vendored); a maintainer binds
 * these symbols with ctypes from network.py.
 *
 * Conventions: plain pointers and sizes only; the caller owns every in/out buffer; the library owns
 * the context, device weights, staging buffers, streams and graphs.  Every function returns E55_OK (0)
 * or a negative error code; the message is available through e55_last_error().  No C++ exception
 * crosses the ABI.  One context = one GPU + one stream; calls on one context are serialised
 * internally (any host thread may call); distinct contexts are independent (one per GPU for sharding).
 * There is no CPU fallback: without a usable CUDA device e55_create() fails.
 */
#ifndef E55_H_
#define E55_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct e55_ctx e55_ctx;

enum {
  E55_PENDING = 1,           /* asynchronous service calls only: no room right now / result not ready yet */
  E55_OK = 0,
  E55_ERR_INVALID = -1,      /* bad argument / shape */
  E55_ERR_CUDA = -2,         /* CUDA runtime error (message has the CUDA string) */
  E55_ERR_NO_WEIGHTS = -3,   /* predict before set_weights */
  E55_ERR_UNSUPPORTED = -4,  /* architecture not supported by the requested precision path */
  E55_ERR_NO_DEVICE = -5     /* no CUDA device / device is not sm_100 */
};

enum {
  E55_PRECISION_FP32 = 0,    /* CUDA-core fp32 path (same arithmetic type as the reference) */
  E55_PRECISION_BF16_TC = 1  /* bf16 operands on tcgen05 tensor cores, fp32 accumulation */
};

/* Library ABI version (bumped on incompatible change). */
int e55_version(void);

/* Replaces Network.__init__ session/graph construction (network.py:17-69).  `blocks`, `filters`,
 * `in_planes` describe the tower (reference defaults 7, 128, 266); `max_batch` is a sizing hint,
 * larger batches grow the buffers on demand. */
int e55_create(e55_ctx** out, int device, int blocks, int filters, int in_planes, int max_batch);
void e55_destroy(e55_ctx* ctx);

/* Message of the last failing call on `ctx` (or of the last failing e55_create on this thread when
 * ctx is NULL).  Valid until the next call on the same ctx/thread. */
const char* e55_last_error(const e55_ctx* ctx);

/* Replaces Network.set_weights / model.set_weights (network.py:238-243; client.py:51).  Takes the
 * Keras get_weights() list: n_tensors float32 arrays, shapes[i][0..ndims[i]).  Folds inference
 * BatchNorm (eps 1e-3) into the preceding convolution, builds the fp32 and the bf16/UMMA weight
 * images and uploads them. */
int e55_set_weights(e55_ctx* ctx, int n_tensors, const float* const* data,
                    const int64_t* const* shapes, const int* ndims);

/* Select the arithmetic path used by the predict calls (default E55_PRECISION_BF16_TC). */
int e55_set_precision(e55_ctx* ctx, int precision);
int e55_get_precision(const e55_ctx* ctx, int* precision);

/* Replaces Network.predict (network.py:201-216) for HOST buffers: planes float32 [n,in_planes,5,5]
 * (NCHW) -> policy float32 [n,1725] raw logits (index c*25+y*5+x), value float32 [n,1] (tanh).
 * The whole input is one batch.  Includes H2D and D2H copies; returns after the outputs are in the
 * caller's buffers, which may be pinned (e55_host_alloc, cudaHostAlloc) or ordinary pageable memory.  Batches of 64 and
 * more are pipelined in chunks (host-side bit-packing of the planes by a small thread pool -> copy -> kernel ->
 * read-back -> copy-out overlap across chunks; see e55_set_host_threads); smaller ones are one copy in, one kernel,
 * one copy out, a pageable input being bit-packed by the calling thread on the way. */
int e55_predict(e55_ctx* ctx, const float* planes, int n, float* policy, float* value);

/* Host threads e55_predict may use to bit-pack large float-plane batches before the copy (exact: a chunk whose
 * planes are not "one value or zero" travels as floats; results are identical either way).  -1 (default) = the
 * library's choice: cores - 1 threads, pinned to the process's cores (at most 31; E55_HOST_THREADS overrides), and during the first calls it
 * measures how much of the batch to pack and how much to send as floats meanwhile (e55_get_host_split) and keeps
 * the fastest; n > 0 = always pack with n threads; 0 = off: the planes cross PCIe as they
 * are. */
int e55_set_host_threads(e55_ctx* ctx, int threads);
/* Measurement: the rate at which this context's host thread pool can merely READ `bytes` of host memory at `buf` (best of
 * `passes` passes, GB/s).  Whatever e55_predict does with float32 planes, 26 600 bytes per position have to be read from
 * host memory once, by the cores (packing) or by the copy engine: this is the roofline of the end-to-end number, and
 * bench.py reports it beside it (summed over the ranks of a box, which share the host's memory system). */
int e55_host_read_bandwidth(e55_ctx* ctx, const void* buf, size_t bytes, int passes, double* gbs_out);
/* What e55_predict measured: packing threads in use, and the running mean of microseconds per position of the
 * compacting and of the float pipeline on large batches (0 = not tried yet). */
int e55_get_host_pipeline(e55_ctx* ctx, int* threads_out, double* compact_us_out, double* float_us_out);
/* The compacting pipeline may leave a share of a large batch unpacked and send it as float planes in between the
 * packed chunks (link and cores both busy).  float_eighths_out: that share, in eighths of the batch, at the latest
 * large call (0 = all packed, 8 = the float pipeline); us_per_position_out (8 doubles, may be null): the running
 * mean at the shares 0, 1, 2, 3, 4, 5, 6 and 8 eighths (0 = not tried).  Pageable inputs are always packed whole (the
 * cores have to read them either way). */
int e55_get_host_split(e55_ctx* ctx, int* float_eighths_out, double* us_per_position_out);

/* Page-locked host memory for the arrays handed to the host-buffer calls (any thread, any device: portable).  Buffers
 * from here are read and written by the copy engines directly: no staging copy on either side of e55_predict.
 * Network.predict (erweitern_55_b200/network.py) returns its numpy results in such blocks and recycles them when the
 * arrays are garbage-collected.  Ordinary (pageable) buffers work everywhere too: the library stages them itself. */
int e55_host_alloc(size_t bytes, void** out);
int e55_host_free(void* p);

/* Same computation on DEVICE buffers, enqueued on the context stream; returns without
 * synchronising (call e55_sync).  Buffers must stay valid until then. */
int e55_predict_device(e55_ctx* ctx, const float* d_planes, int n, float* d_policy, float* d_value);

/* Compact input: the same evaluation from bit-packed planes, 8 bytes per plane
 * instead of 100.  bits uint32 [n,in_planes]: bit sq (= y*5+x) set where the plane is non-zero; scale float32
 * [n,in_planes]: the value those squares hold (1 for occupancy planes, the constant for count/ply planes), i.e.
 * planes[n][c][sq] = (bits[n][c] >> sq) & 1 ? scale[n][c] : 0 of what Network.get_inputs builds
 * (network.py:255-274).  Results are identical to e55_predict on the expanded planes. */
int e55_predict_packed(e55_ctx* ctx, const uint32_t* bits, const float* scale, int n, float* policy, float* value);
int e55_predict_packed_device(e55_ctx* ctx, const uint32_t* d_bits, const float* d_scale, int n, float* d_policy,
                              float* d_value);

/* Compact legal-move policy tail (the north star's fused legal-move-masked softmax, in the form a search wants it):
 * for position i the caller lists the policy indices of its legal moves in move_index[move_offset[i] ..
 * move_offset[i+1]) (move_offset int32 [n+1], starting at 0); priors[j] receives softmax over exactly those logits
 * -- what minishogilib.MCTS.evaluate computes from a policy row (mcts.py:60,105-106) -- so only a few dozen floats
 * per position cross PCIe instead of 1725.  value float32 [n] as in e55_predict. */
int e55_predict_packed_moves(e55_ctx* ctx, const uint32_t* bits, const float* scale, int n, const int32_t* move_offset,
                             const uint16_t* move_index, float* priors, float* value);

/* Fused policy tail for the caller-side post-processing the reference leaves to
 * minishogilib.MCTS.evaluate (mcts.py:59-60,103-106): probs = softmax over entries with
 * legal_mask != 0 (others 0), legal_mask uint8 [n,1725]; value as in e55_predict. */
int e55_predict_masked(e55_ctx* ctx, const float* planes, const uint8_t* legal_mask, int n,
                       float* probs, float* value);
int e55_predict_masked_device(e55_ctx* ctx, const float* d_planes, const uint8_t* d_legal_mask, int n,
                              float* d_probs, float* d_value);

/* Cross-game leaf aggregation (the per-GPU "inference stream" of a self-play shard): after
 * e55_service_start, any number of host threads may call the blocking e55_service_predict[_packed] with
 * the small batches mcts.py produces (root 1, leaves 16; mcts.py:57-58,91-106); a dispatcher thread coalesces
 * them into forwards of up to max_batch positions (a batch is dispatched when full, when a request does not
 * fit, or max_wait_us after its first request).  Results equal e55_predict[_packed] on the same positions. */
int e55_service_start(e55_ctx* ctx, int max_batch, int max_wait_us);
int e55_service_stop(e55_ctx* ctx);
int e55_service_predict(e55_ctx* ctx, const float* planes, int n, float* policy, float* value);
int e55_service_predict_packed(e55_ctx* ctx, const uint32_t* bits, const float* scale, int n, float* policy,
                               float* value);
int e55_service_predict_packed_moves(e55_ctx* ctx, const uint32_t* bits, const float* scale, int n,
                                     const int32_t* move_offset, const uint16_t* move_index, float* priors, float* value);
/* Asynchronous form, so that ONE host thread can keep many games in flight: submit copies the request into the
 * open batch and returns a ticket at once (E55_PENDING if there is no room and may_block is 0); fetch copies the
 * priors (move_offset[n] floats) and values out and releases the ticket (E55_PENDING if the batch has not finished
 * and wait is 0).  Every ticket must be fetched exactly once, before e55_service_stop. */
typedef struct e55_ticket {
  int32_t batch, off, moff, n, total, kind;
  uint64_t gen;
} e55_ticket;
int e55_service_submit_packed_moves(e55_ctx* ctx, const uint32_t* bits, const float* scale, int n,
                                    const int32_t* move_offset, const uint16_t* move_index, int may_block,
                                    e55_ticket* ticket_out);
int e55_service_fetch(e55_ctx* ctx, const e55_ticket* ticket, float* priors, float* value, int wait);
int e55_service_stats(e55_ctx* ctx, int64_t* batches_out, int64_t* positions_out);

/* Synthetic self-play load on the running service, tree operations excluded (what the service itself sustains):
 * `workers` host threads each keep leaf_batch-position requests of the kind the search sends (bit-packed planes + 30
 * legal moves per position; mcts.py:91-106) in flight on 48 asynchronous tickets until simulations/leaf_batch requests
 * per move have been answered for `moves` moves. */
int e55_bench_selfplay(e55_ctx* ctx, int workers, int moves, int simulations, int leaf_batch,
                       double* moves_per_s_out, double* evals_per_s_out, double* mean_batch_out);

/* Tower shape the context was created with (e55_create arguments). */
int e55_get_arch(const e55_ctx* ctx, int* blocks_out, int* filters_out, int* in_planes_out);

/* ---- native minishogi engine, search tree and self-play -------------------
 * Counterpart of the external Rust library minishogilib v0.6.12 the reference drives from Python
 * (selfplay.py:11-36,78; mcts.py:27-205; usi.py:39-140; tests/test_checkmate.py, tests/test_repetition.py).
 * Moves and positions travel as SFEN/USI strings ("5e4d", "3d4e+", "G*1d"); move lists are space separated. */
typedef struct e55_pos e55_pos;
e55_pos* e55_pos_create(void);                                  /* Position(); starts at the initial position */
void e55_pos_destroy(e55_pos* pos);
e55_pos* e55_pos_copy(const e55_pos* pos);                      /* Position.copy(True) */
int e55_pos_set_sfen(e55_pos* pos, const char* sfen);           /* "board side hands n [moves m1 m2 ...]" */
int e55_pos_sfen(const e55_pos* pos, char* buf, int buflen);
int e55_pos_print(const e55_pos* pos, char* buf, int buflen);   /* Position.print(): board as text */
int e55_pos_side_to_move(const e55_pos* pos);                   /* 0 = first player (black), 1 = second */
int e55_pos_ply(const e55_pos* pos);
int e55_pos_generate_moves(e55_pos* pos, char* buf, int buflen); /* legal moves; returns count */
int e55_pos_do_move(e55_pos* pos, const char* move);            /* E55_ERR_INVALID if the move is not legal */
int e55_pos_is_repetition(const e55_pos* pos, int out3[3]);     /* (is_repetition, my_check_rep, op_check_rep) */
int e55_pos_solve_checkmate_dfs(e55_pos* pos, int depth, char* move_buf, int buflen);  /* 1 = mate found */
int e55_pos_encode_packed(const e55_pos* pos, uint32_t* bits, float* scale);   /* 266-plane input, packed */
int e55_pos_encode_packed_batch(const e55_pos* const* pos, int n, uint32_t* bits, float* scale);  /* Network.get_inputs */
int e55_pos_policy_index(e55_pos* pos, const char* move);       /* index of a legal move in the 1725 logits */

/* minishogilib.MCTS as mcts.py drives it (mcts.py:27-205).  Nodes are small integers valid until the next
 * set_root/clear; the root is node 0.  One tree serves one search at a time. */
typedef struct e55_mcts e55_mcts;
e55_mcts* e55_mcts_create(double memory_gb);                    /* minishogilib.MCTS(memory_size), mcts.py:27 */
void e55_mcts_destroy(e55_mcts* m);
int e55_mcts_clear(e55_mcts* m);
int e55_mcts_configure(e55_mcts* m, float c_puct, float dirichlet_alpha, float dirichlet_eps, uint64_t seed);
int e55_mcts_set_root(e55_mcts* m, const e55_pos* pos, int reuse_tree);                    /* -> root node, mcts.py:50 */
int e55_mcts_expanded(const e55_mcts* m, int node);                                        /* mcts.py:56 */
int e55_mcts_evaluate(e55_mcts* m, int node, e55_pos* pos, const float* policy, float value01);  /* mcts.py:60,105 */
int e55_mcts_add_noise(e55_mcts* m, int node);                                             /* mcts.py:64 */
int e55_mcts_select_leaf(e55_mcts* m, int node, e55_pos* pos, int forced_playouts);        /* -> leaf node; pos advanced, mcts.py:97 */
int e55_mcts_backpropagate(e55_mcts* m, int leaf);                                         /* mcts.py:110 */
double e55_mcts_get_usage(const e55_mcts* m);                                              /* mcts.py:72 */
int e55_mcts_get_nodes(const e55_mcts* m);                                                 /* mcts.py:116 */
int e55_mcts_get_playouts(const e55_mcts* m, int node, int child_sum);                     /* mcts.py:88 */
int e55_mcts_best_move(e55_mcts* m, int node, char* buf, int buflen);                      /* 1 = move written, 0 = no children */
int e55_mcts_softmax_sample(e55_mcts* m, int node, float temperature, char* buf, int buflen);
int e55_mcts_softmax_sample_among_top_moves(e55_mcts* m, int node, float away, float temperature, char* buf, int buflen);
int e55_mcts_info(const e55_mcts* m, int node, char* pv_buf, int buflen, float* q_out);    /* -> pv length, mcts.py:114 */
int e55_mcts_dump(const e55_mcts* m, int node, int target_pruning, int remove_zeros, int* sum_n_out, float* q_out,
                  char* moves_buf, int buflen, int* counts, int max_counts);               /* -> entries, mcts.py:187 */
int e55_mcts_describe(const e55_mcts* m, int node, char* buf, int buflen);                 /* MCTS.print / debug */
int e55_mcts_visualize(const e55_mcts* m, int node, int node_num, char* buf, int buflen);  /* dot, mcts.py:205 */
/* The two halves of MCTS.run done natively on a network context (packed inputs, no Python per leaf):
 * search_root = set the root, evaluate it if needed, add noise (mcts.py:49-64); returns 1 for a terminal root.
 * search_batches = n_batches iterations of select / ONE network call / expand / back up (mcts.py:91-110);
 * returns the simulations done.  use_service routes the calls through the evaluation service. */
int e55_mcts_search_root(e55_mcts* m, e55_ctx* ctx, const e55_pos* pos, int reuse_tree, int use_dirichlet, int use_service);
int e55_mcts_search_batches(e55_mcts* m, e55_ctx* ctx, const e55_pos* pos, int n_batches, int leaf_batch,
                            int forced_playouts, int use_service);

typedef struct e55_selfplay_stats {
  double seconds;          /* wall time of the run */
  long moves;              /* game plies played (each = 1 root evaluation + `simulations` simulations, or a mate move) */
  long evals;              /* positions sent through the network */
  long games;              /* games finished */
  double mean_game_plies;
  double mean_gpu_batch;   /* positions per GPU forward as coalesced by the evaluation service */
  /* GPU-resident self-play only (0 from the host driver): */
  long tree_overflows;      /* descents abandoned because a game's tree pool was full: must be 0 (searches are suspended at 90 %) */
  long searches_suspended;  /* searches ended early at 90 % pool usage, as mcts.py:71-73 does with its memory */
  long mate_searches;       /* mate searches answered by the host */
  long mate_exhausted;      /* ... of which gave up at their node budget ("no mate" not proven; the reference's search is unbounded) */
  long records_truncated;   /* finished games whose record did not fit its buffer (flagged in the record) */
} e55_selfplay_stats;

/* Receives one finished game as text: "winner W ply P\n" then one line per ply
 * "<move> <sum_N> <Q> <k> <m1> <n1> ... <mk> <nk>" (gamerecord.py:3-9); a ply that is not a learning target (a fast
 * search under playout-cap oscillation) has '~' in front of its move.  Calls are serialised. */
typedef void (*e55_game_sink)(void* user, const char* record);

typedef struct e55_selfplay_config {
  int threads;             /* host threads; 0 = cores - 2 (cores - 1 on hosts of 12 cores or fewer): the service's threads need the rest */
  int games_per_thread;    /* games per thread in the shared pool kept in flight on the asynchronous service; 0 = 12
                              (best with a service batch of about 74 positions per thread: two batches' worth of leaves in flight) */
  long total_moves;        /* stop after this many plies in total */
  int simulations;         /* per move (client.py:29: 800) */
  int leaf_batch;          /* leaves per network call (client.py:28: 16) */
  int mate_depth;          /* solve_checkmate_dfs depth before each search, 0 = off (selfplay.py:36: 7) */
  int reuse_tree;          /* client.py:32 */
  int use_dirichlet;       /* client.py:31 */
  int sampling_moves;      /* plies sampled from the visit counts (selfplay.py:10: 10) */
  int max_moves;           /* selfplay.py:10: 512 */
  uint64_t seed;
  e55_game_sink sink;      /* optional: game records */
  void* sink_user;
  /* Playout-cap oscillation (selfplay.py:45-61,90-91; off by default as in client.py): with probability cap_frac a move
   * gets a FULL search -- cap_full simulations, forced playouts, Dirichlet noise, a fresh tree, pruned visit counts as
   * the training target -- otherwise a FAST one -- cap_fast simulations counting the visits the reused tree already has
   * ("immediate"), no noise, no forced playouts -- whose ply is not a learning target (its record line starts with '~'). */
  int cap_enable;
  int cap_full;            /* N (selfplay.py:10: 800) */
  int cap_fast;            /* n (128) */
  float cap_frac;          /* 0.25 */
} e55_selfplay_config;
void e55_selfplay_default_config(e55_selfplay_config* cfg);

/* Self-play on the running evaluation service: `threads` host threads work through a pool of threads x
 * games_per_thread games, each played the way selfplay.run does (selfplay.py:10-113) with client.py's search
 * settings, until `total_moves` plies have been played. */
int e55_selfplay_run_ex(e55_ctx* ctx, const e55_selfplay_config* cfg, e55_selfplay_stats* stats_out);
int e55_selfplay_run(e55_ctx* ctx, int threads, long total_moves, int simulations, int leaf_batch, int mate_depth,
                     uint64_t seed, e55_selfplay_stats* stats_out);

/* GPU-resident self-play: the same games with the search itself on the GPU -- tree descent with virtual loss, move
 * generation, repetition test, input encoding, expansion, back-up and move choice run as kernels next to the
 * network, one thread per game, `games` games at once; nothing but counters crosses PCIe and the host cores do not
 * matter (csrc/e55_gpu_selfplay.cuh).  Same rules, encoding and search as the host driver (shared source; tree reuse
 * and Dirichlet noise as in client.py:27-35, mate search of mate_depth plies before every search as in
 * selfplay.py:35-42 -- that one, a divergent recursion, is answered by the host from the roots the games post),
 * except that a round selects until a game's 16 network slots are filled (at most 24 selections) and an
 * unevaluated root takes a round of its own.  From 2 048 games on, the games run as two cohorts on two streams so
 * that one cohort's search overlaps the other's network launch.  simulations: a multiple of 16, at most 1024.
 * records (optional): the first record_cap finished games, 16384 uint16 words each: plies, winner, words used, truncated (0/1),
 * then per ply what selfplay.run records (selfplay.py:80-92) -- move, sum_N, Q * 65535, k, and k (move, N) pairs of
 * the visited root children; moves are packed as from (31 = drop) | to << 5 | piece << 10 | promotes << 14 (bit 15 of a
 * ply's own move word: not a learning target -- a fast search under playout-cap oscillation).
 * stats_out->mean_gpu_batch is the positions per network launch (16 slots per game of a cohort). */
int e55_gpu_selfplay_run(e55_ctx* ctx, int games, long total_moves, int simulations, int mate_depth, int reuse_tree,
                         int use_dirichlet, int sampling_moves, int max_moves, uint64_t seed, e55_selfplay_stats* stats_out,
                         uint16_t* records, int record_cap,
                         int* n_records_out);
/* Debug/verification: ONE game of the GPU-resident search, a round at a time, with the caller in the network's place
 * (no Dirichlet noise, no sampling, no mate search): the moves of `prefix` (packed as in the game records) are played from
 * the start position, then every select call runs one select_kernel round and reports its selections (leaf node, network
 * slot, the path from the root as packed moves) and the legal moves of the slots to evaluate; every expand call takes the
 * priors (in the slots' move order) and values, runs one expand_kernel round and reports the game's record so far.
 * tests/test_gpu_parity.py walks oracle/mcts_oracle.py beside it: same path for every selection, same visit counts.
 * info (8 ints): plies played, simulations done (-1: root not evaluated), selections of the round, nodes, edges, games
 * finished, record words, 1 if `record` holds a finished game (with its 4-word header).  Array sizes: sel_* [24],
 * sel_path [24][96], slot_cnt [16], slot_moves / slot_idx [16][256], priors [16][256], value [16], record [16384]. */
typedef struct e55_gpu_search e55_gpu_search;
int e55_gpu_search_open(e55_ctx* ctx, const uint16_t* prefix, int n_prefix, int simulations, int reuse_tree, e55_gpu_search** out);
void e55_gpu_search_close(e55_gpu_search* h);
int e55_gpu_search_select(e55_gpu_search* h, int32_t* info, int32_t* sel_leaf, int32_t* sel_slot, int32_t* sel_path_len,
                          uint16_t* sel_path, int32_t* slot_cnt, uint16_t* slot_moves, uint16_t* slot_idx);
int e55_gpu_search_expand(e55_gpu_search* h, const float* priors, const float* value, int32_t* info, uint16_t* record);

/* Playout-cap oscillation for the GPU-resident self-play and its debug stepper (selfplay.py:45-61,90-91; off by default,
 * as in client.py): with probability frac a move gets a full search (full_simulations, forced playouts, Dirichlet noise, a
 * fresh tree, pruned visit counts in the record), otherwise a fast one (fast_simulations counting the reused tree's visits,
 * no noise) whose ply is not a learning target (bit 15 of its move word in the record). */
int e55_gpu_selfplay_set_cap(e55_ctx* ctx, int enable, int full_simulations, int fast_simulations, float frac);

/* Device build of the rules engine: perft from the start position computed on the GPU (14, 181, 2512, 35401, ...). */
int e55_selftest_engine(int device, int depth, int64_t* nodes_out);
/* Validation of the device-side game history: n threads each play a random game and a random path below it with the
 * kernels' bookkeeping and encode the position they reach; the host replays the moves on a Position and compares
 * the 266 packed planes, the repetition verdict and count and the number of legal moves.  mismatches_out must be 0;
 * repetitions_out = positions that had occurred before (the walks are steered into repetitions), path_plies_out =
 * plies walked below the game positions in total. */
int e55_selftest_encode(int device, int n, uint64_t seed, int64_t* mismatches_out, int64_t* repetitions_out, int64_t* path_plies_out);

/* Block until all work enqueued on the context stream has finished. */
int e55_sync(e55_ctx* ctx);

/* The cudaStream_t the context launches on (for CUDA-event timing by the caller). */
int e55_stream(e55_ctx* ctx, void** stream_out);

/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
int e55_launch_count(const e55_ctx* ctx, int64_t* count_out);

/* Measurement: when enabled, every tensor-core forward records a CUDA-event pair on the context
 * stream around its dominant kernel (the fused tower kernel); e55_get_kernel_timing returns the
 * accumulated device time and launch count since the last e55_set_kernel_timing call.  Serialises
 * consecutive launches on the host, so leave it off outside measurement runs. */
int e55_set_kernel_timing(e55_ctx* ctx, int enable);
int e55_get_kernel_timing(e55_ctx* ctx, double* total_ms_out, int64_t* launches_out);

/* Debug/verification: activations after the first `n_convs` 3x3 convolutions of the tower
 * (1 = stem ... 1+2*blocks = trunk output, the input of both heads) as float32 [n,filters,5,5] on the
 * host, computed by the currently selected precision path. */
int e55_debug_activations(e55_ctx* ctx, const float* planes, int n, int n_convs, float* act_out);

/* Debug/verification: the stand-alone legal-move softmax kernel (the fp32 path's policy tail) on caller-supplied logits
 * float32 [n,1725]; same arguments as e55_predict_packed_moves.  The tensor-core path computes the same priors inside
 * the tower kernel (the logits never leave shared memory); tests hold the two against each other bit for bit. */
int e55_debug_legal_priors(e55_ctx* ctx, const float* logits, int n, const int32_t* move_offset, const uint16_t* move_index,
                           float* priors);

/* Tuning: number of weight chunks the tower kernel keeps group A's MMA issuer ahead of group B's
 * (0 = free running; clamped to ring depth - 1). */
int e55_debug_set_lag(e55_ctx* ctx, int lag);

/* Debug/profiling: switch the tower kernel's in-kernel cycle counters (CTA 0 only) on or off; when
 * counters_out is non-NULL the 9232 words (16 counters + per-chunk / per-event timestamps of the first
 * round) accumulated so far are copied out first, then reset.
 * [0..3] group-A MMA issuer: total, issue, wait-for-weights, wait-for-activations cycles; [4..7] the
 * same for group B; [8] weight producer wait; [10],[11] epilogue warp 0 wait / work. */
int e55_debug_profile(e55_ctx* ctx, int enable, uint64_t* counters_out);

/* Hardware self-test of the tcgen05 operand addressing this library depends on: one 128x128x64 bf16
 * UMMA from no-swizzle K-major shared-memory operands whose A descriptor start is displaced by
 * `row_shift` rows (16-byte granularity), compared with a CPU product.  max_abs_err_out receives the
 * largest deviation. */
int e55_selftest_umma(int device, int row_shift, float* max_abs_err_out);

/* Host-only self-test (no device needed) of the tower kernel's input-slice bookkeeping (`e55::tc::SliceBook`,
 * csrc/e55_tc.cuh): `slices` slices are written to the four stem-input buffers in a pseudo-random order and the
 * barrier parity the kernel would wait on is compared with plain counters before each.  mismatches_out receives
 * the number of disagreements (0 = correct at any launch length). */
int e55_selftest_slice_book(int64_t slices, uint64_t seed, int64_t* mismatches_out);

/* Tensor-pipe micro-benchmark: every SM issues `iters` rounds of 72 back-to-back 128 x n_dim x 16
 * UMMAs (n_dim 128 or 256) from shared memory; tflops_out receives the achieved dense bf16 rate. */
int e55_bench_umma(int device, int n_dim, int iters, float* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* E55_H_ */
