/* kgb200.h - C ABI of the B200-native KataGo NN-evaluator backend (libkgb200.so).
 *
 * Drop-in boundary: the reference selects ONE translation unit defining `namespace NeuralNet`
 * (cpp/neuralnet/nninterface.h:32-182) at link time.  A maintainer adds `neuralnet/b200backend.cpp`
 * (integration/b200backend.cpp in this repository) whose functions forward 1:1 to the entry
 * points below.  Plain pointers and sizes only; no exceptions, no C++ or torch types cross this boundary; every
 * function returns KGB_OK (0) or a negative status with the message available from kgb_last_error().
 * Buffers passed to kgb_forward are HOST memory owned by the caller (the reference's NNResultBuf / NNOutput arrays);
 * host<->device copies happen inside the call, exactly like cudabackend.cpp:3714-3782.
 */
#ifndef KGB200_H_
#define KGB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define KGB_API __attribute__((visibility("default")))
#else
#define KGB_API
#endif

#define KGB_OK 0
#define KGB_ERR_INVALID (-1)   /* bad argument / unsupported model */
#define KGB_ERR_IO (-2)        /* model file could not be read / sha256 mismatch */
#define KGB_ERR_CUDA (-3)      /* CUDA runtime or driver failure, or no sm_100 device */

typedef struct kgb_model kgb_model;     /* replaces LoadedModel     (nninterface.h:27, eigenbackend.cpp:94-107) */
typedef struct kgb_context kgb_context; /* replaces ComputeContext  (nninterface.h:17) */
typedef struct kgb_handle kgb_handle;   /* replaces ComputeHandle   (nninterface.h:21) - one per server thread, not thread-safe */

/* Subset of ModelDesc (cpp/neuralnet/desc.h:508-571) that NNEvaluator reads through getModelDesc(). */
typedef struct kgb_model_info {
  char name[128];
  char sha256[65];
  int32_t model_version;
  int32_t num_input_channels;
  int32_t num_input_global_channels;
  int32_t num_policy_channels;
  int32_t num_value_channels;
  int32_t num_score_value_channels;
  int32_t num_ownership_channels;
  int32_t trunk_num_channels;
  int32_t num_blocks;
  int32_t prefer_pass_alive_under_suicide_rules;
  float td_score_multiplier;
  float score_mean_multiplier;
  float score_stdev_multiplier;
  float lead_multiplier;
  float variance_time_multiplier;
  float shortterm_value_error_multiplier;
  float shortterm_score_error_multiplier;
  int64_t conv_macs_per_position; /* direct-convolution MACs per board point (ModelDesc::iterConvLayers, desc.cpp:2643) */
} kgb_model_info;

/* NeuralNet::globalInitialize / globalCleanup (nninterface.h:34-36). */
KGB_API int kgb_global_init(void);
KGB_API int kgb_global_cleanup(void);
/* Message of the last failure on the calling thread ("" if none). */
KGB_API const char* kgb_last_error(void);
/* NeuralNet::printDevices (nninterface.h:39): number of CUDA devices and, per device, name/compute capability. */
KGB_API int kgb_device_count(int* count);
KGB_API int kgb_device_name(int device, char* buf, int buf_len, int* cc_major, int* cc_minor);

/* NeuralNet::loadModelFile / freeLoadedModel / getModelDesc (nninterface.h:43-46).
 * expected_sha256 may be NULL or "" to skip verification. */
KGB_API int kgb_model_load_file(const char* path, const char* expected_sha256, kgb_model** out);
KGB_API void kgb_model_free(kgb_model* model);
KGB_API int kgb_model_get_info(const kgb_model* model, kgb_model_info* out);

/* NeuralNet::createComputeContext / freeComputeContext (nninterface.h:50-65).
 * fp16_mode: 0 = "fp32-equivalent" (3-term split-fp16 on the tensor pipe, fp32 streams), 1 = fp16 operands with fp32
 * accumulation, -1 = auto (fp16). */
KGB_API int kgb_context_create(const int* gpu_idxs, int num_gpu_idxs, int nn_x_len, int nn_y_len, int fp16_mode,
                       const kgb_model* model, kgb_context** out);
KGB_API void kgb_context_free(kgb_context* ctx);

/* NeuralNet::createComputeHandle / freeComputeHandle / isUsingFP16 (nninterface.h:76-94).
 * gpu_idx == -1 selects the context's first device (or device 0).  inputs_nhwc selects the layout of `spatial`
 * in kgb_forward: 1 = [n][Y][X][C], 0 = [n][C][Y][X]. */
KGB_API int kgb_handle_create(kgb_context* ctx, const kgb_model* model, int max_batch_size, int require_exact_nn_len,
                      int inputs_nhwc, int gpu_idx, kgb_handle** out);
KGB_API void kgb_handle_free(kgb_handle* handle);
KGB_API int kgb_handle_is_fp16(const kgb_handle* handle);

/* NeuralNet::getOutput (nninterface.h:117-123).  For each of the n rows:
 *   spatial        [n][num_input_channels * X * Y]   NNResultBuf::rowSpatialBuf (fillRowV7 output, no symmetry applied)
 *   global         [n][num_input_global_channels]    NNResultBuf::rowGlobalBuf
 *   symmetry       [n] in 0..7                       NNResultBuf::symmetry (applied to inputs, inverted on outputs)
 *   policy_optimism[n]                               NNResultBuf::policyOptimism
 * Outputs (raw logits; NNEvaluator does the softmax/tanh post-processing, nneval.cpp:960-1249):
 *   policy         [n][X*Y + 1]  NNOutput::policyProbs (pass last)
 *   value          [n][3]        whiteWinProb, whiteLossProb, whiteNoResultProb
 *   score_value    [n][6]        whiteScoreMean, whiteScoreMeanSq, whiteLead, varTimeLeft, shorttermWinlossError,
 *                                shorttermScoreError (filled by model version as eigenbackend.cpp:2583-2626)
 *   ownership      [n][X*Y] or NULL (NNOutput::whiteOwnerMap)
 */
KGB_API int kgb_forward(kgb_handle* handle, int n, const float* spatial, const float* global, const int32_t* symmetry,
                const float* policy_optimism, float* policy, float* value, float* score_value, float* ownership);

/* Same computation with every buffer already resident in this handle's device memory (used by the device-resident
 * self-play loop and by bench.py's HBM-resident timing).  Pointers are device pointers; the call is asynchronous on the
 * handle's stream - use kgb_handle_sync() before reading results from another stream. */
KGB_API int kgb_forward_device(kgb_handle* handle, int n, const float* d_spatial, const float* d_global, const int32_t* d_symmetry,
                       const float* d_policy_optimism, float* d_policy, float* d_value, float* d_score_value,
                       float* d_ownership);
KGB_API int kgb_handle_sync(kgb_handle* handle);
/* cudaStream_t of the handle (as an integer) so a torch/CUDA caller can order its own work against it. */
KGB_API uint64_t kgb_handle_stream(kgb_handle* handle);
/* Number of kernel launches one kgb_forward of batch n issues (for bench.py's gpu_launches accounting). */
KGB_API int kgb_handle_launches_per_forward(const kgb_handle* handle);

/* New weights into a live handle.  The reference polls its models directory and builds a fresh NNEvaluator for every new net
 * (command/selfplay.cpp:142-231 loadLatestNeuralNetIntoManager, :336-352 the polling thread; dataio/loadmodel.cpp:58
 * findLatestModel); games switch over between moves (switchNetsMidGame, program/play.cpp maybeGetNewNet).  Here the handle keeps
 * its graphs and buffers and only the weight arena changes:
 *   kgb_handle_stage_weights   packs `model` (same architecture as the handle's, else an error) for this handle's kernels and
 *                              copies it to a shadow arena on a side stream - evaluation continues meanwhile
 *   kgb_handle_commit_weights  orders "shadow -> live" on the handle's stream: every forward pass / self-play wave enqueued
 *                              afterwards runs the new net, none sees a mixture
 * Between GPUs of one node (one process per GPU) the packed arena travels by ncclBroadcast straight between device memories,
 * so only one rank reads and packs the model file:
 *   kgb_nccl_unique_id                   on one rank; the 128 bytes reach the others by any side channel (torch.distributed, MPI, a file)
 *   kgb_handle_comm_init                 on every rank, once per handle
 *   kgb_handle_broadcast_staged_weights  on every rank; the root must have staged; *ms_out = the collective's device time
 * followed by kgb_handle_commit_weights on every rank.  NCCL is bound at run time (libnccl.so.2, or KGB_NCCL_LIB). */
KGB_API int kgb_handle_weights_bytes(const kgb_handle* handle, uint64_t* bytes);
KGB_API int kgb_handle_stage_weights(kgb_handle* handle, const kgb_model* model);
KGB_API int kgb_handle_commit_weights(kgb_handle* handle);
/* Blocks the host until the staging copy (or a received broadcast) has landed in the shadow arena; not needed before a commit
 * (ordered on the device), useful before a barrier that lines the ranks up for a timed broadcast. */
KGB_API int kgb_handle_wait_staged(kgb_handle* handle);
KGB_API int kgb_nccl_unique_id(void* id_out_128_bytes);
KGB_API int kgb_handle_comm_init(kgb_handle* handle, const void* id_128_bytes, int rank, int num_ranks);
KGB_API int kgb_handle_broadcast_staged_weights(kgb_handle* handle, int root, float* ms_out);

/* FOR TESTING: NeuralNet::testEvaluateConv (nninterface.h:134-143).  weights in the model-file order
 * [ky][kx][in_c][out_c]; input/output [n][Y][X][C] (NHWC) fp32.  use_fp16 selects the operand mode as in
 * kgb_context_create.  Returns KGB_OK and fills output[n*Y*X*out_c]. */
KGB_API int kgb_test_conv(int ky, int kx, int in_c, int out_c, const float* weights, int n, int nn_x_len, int nn_y_len, int use_fp16,
                  const float* input, float* output);

/* ---------------------------------------------------------------------------------------------------------------
 * Boundary 2: device-resident self-play slots.  Boards, MCTS node pools and NN rows stay in HBM; one
 * kgb_selfplay_run(steps) call performs `steps` playout waves (select+featurize -> NN forward -> backup, one visit per
 * game per wave) without returning to the host.  Replaces, for the supported rule/parameter subset,
 * Search::runWholeSearch/playoutDescend (search.cpp:473,1189), selectBestChildToDescend (searchexplorehelpers.cpp:324),
 * NNEvaluator::evaluate's featurize/post-process halves (nneval.cpp:861-1262) and the per-move core of Play::runGame
 * (play.cpp:1757-1936).  Parameters are the reference's SearchParams names (search/searchparams.h). */
typedef struct kgb_selfplay kgb_selfplay;

typedef struct kgb_selfplay_config {
  int32_t num_games;                 /* concurrent games (= NN batch per wave); <= the handle's max_batch_size */
  int32_t max_visits;                /* maxVisits: visits per move on a cleared tree */
  int32_t max_moves;                 /* game length cap (0 = 2*X*Y) */
  int32_t multi_stone_suicide_legal; /* Rules::multiStoneSuicideLegal */
  int32_t early_temperature_moves;   /* moves chosen proportionally to visits (chosenMoveTemperatureEarly = 1), then argmax */
  float komi;
  double cpuct_exploration;          /* cpuctExploration       (default 1.0) */
  double cpuct_exploration_log;      /* cpuctExplorationLog    (0.45 in selfplay cfgs) */
  double cpuct_exploration_base;     /* cpuctExplorationBase   (500) */
  double fpu_reduction_max;          /* fpuReductionMax        (0.2) */
  double root_fpu_reduction_max;     /* rootFpuReductionMax    (0.1 in selfplay cfgs) */
  double win_loss_utility_factor;    /* winLossUtilityFactor   (1.0) */
  double no_result_utility_for_white;
  uint64_t seed;
  int32_t debug_fake_nn;             /* TEST ONLY: replace the evaluator by the deterministic hash net of oracle/ref_driver.cpp */
  int32_t disable_ladder_features;   /* 1 = leave NN input planes 14-17 zero (timing experiments only) */
  int32_t ladder_nodes_per_wave;     /* > 0: each of a game's 8 ladder-reader warps plays at most this many search moves per wave;
                                        a game whose searches are unfinished skips the wave (no visit) and resumes in the next.
                                        0 = run every search to the end inside the wave.  Features are identical either way. */
  int32_t max_playouts_per_wave;     /* a playout that ends without needing the evaluator (evaluation-cache hit, graph-search edge catch-up,
                                        cycle) is backed up inside the select kernel and the game starts its next playout in the same wave,
                                        up to this many (0 = 16).  The launch lasts as long as its slowest game, so a small bound (2-3) caps
                                        the tail: a game that uses it up delivers no leaf this wave and carries on in the next.  The sequence
                                        of playouts of a game, hence every result, is the same for every bound. */
  /* Score utility (Search::getScoreUtility, searchhelpers.cpp:272-279; selfplay8mainb18.cfg: 0.05, 0.30, 0.25, 0.50).
   * Both factors 0 = win/loss utility only. */
  double static_score_utility_factor;
  double dynamic_score_utility_factor;
  double dynamic_score_center_zero_weight;
  double dynamic_score_center_scale;
  double draw_equivalent_wins_for_white;   /* 0.5 in every stock config; used for integer komi results */
  /* More SearchParams of the selection / backup formulas (search/searchparams.h), by their cfg names.  All 0 = the plain
   * PUCT + visit-weighted average of the first fixtures. */
  double value_weight_exponent;                   /* valueWeightExponent (0.5): t-CDF value weighting in the backup */
  int32_t fpu_parent_weight_by_visited_policy;    /* fpuParentWeightByVisitedPolicy */
  int32_t debug_fixed_symmetry_plus_one;          /* TEST ONLY: k + 1 = evaluate every row under symmetry k (the reference's nnRandomize = false with
                                                     nnForcedSymmetry / default symmetry k, nneval.cpp:698-707); 0 = a random symmetry per row */
  double fpu_parent_weight_by_visited_policy_pow; /* fpuParentWeightByVisitedPolicyPow */
  double fpu_parent_weight;                       /* fpuParentWeight */
  double fpu_loss_prop;                           /* fpuLossProp */
  double root_fpu_loss_prop;                      /* rootFpuLossProp */
  double cpuct_utility_stdev_prior;               /* cpuctUtilityStdevPrior (0.25) */
  double cpuct_utility_stdev_prior_weight;        /* cpuctUtilityStdevPriorWeight (1.0) */
  double cpuct_utility_stdev_scale;               /* cpuctUtilityStdevScale (0 = off) */
  double root_desired_per_child_visits_coeff;     /* rootDesiredPerChildVisitsCoeff */
  double subtree_value_bias_factor;               /* subtreeValueBiasFactor (0.30 in selfplay8mainb18.cfg; 0 = off) */
  double subtree_value_bias_weight_exponent;      /* subtreeValueBiasWeightExponent (0.8) */
  int32_t use_graph_search;                       /* useGraphSearch: transpositions share a node (search.cpp:875-936) */
  int32_t graph_search_rep_bound;                 /* graphSearchRepBound (11) */
  int32_t debug_hold_at_max_visits;               /* hold mode: a game whose root has max_visits visits idles until kgb_selfplay_release (tests; game recording) */
  int32_t root_noise_enabled;                     /* rootNoiseEnabled: Dirichlet noise on the root policy */
  double root_dirichlet_noise_total_concentration;/* rootDirichletNoiseTotalConcentration (10.83) */
  double root_dirichlet_noise_weight;             /* rootDirichletNoiseWeight (0.25) */
  double root_policy_temperature;                 /* rootPolicyTemperature (0 = unset = 1.0) */
  double root_policy_temperature_early;           /* rootPolicyTemperatureEarly (0 = unset = 1.0) */
  double chosen_move_temperature_halflife;        /* chosenMoveTemperatureHalflife (0 = unset = 19): also the half-life of the early root temperature */
  /* Root move choice like Search::getChosenMoveLoc (play selection values, LCB, temperature).  use_play_selection = 0 keeps the
   * simple rule: proportional to visits for the first early_temperature_moves moves, then the most visited. */
  int32_t use_play_selection;
  int32_t use_lcb_for_selection;                  /* useLcbForSelection */
  int32_t use_non_buggy_lcb;                      /* useNonBuggyLcb */
  int32_t root_prune_useless_moves;               /* rootPruneUselessMoves: when the opponent's last four moves were passes, the root never plays
                                                     inside either player's pass-alive area (Search::isAllowedRootMove, searchhelpers.cpp:310-342) */
  double lcb_stdevs;                              /* lcbStdevs (5.0) */
  double min_visit_prop_for_lcb;                  /* minVisitPropForLCB (0.15) */
  double chosen_move_temperature;                 /* chosenMoveTemperature (0.15) */
  double chosen_move_temperature_early;           /* chosenMoveTemperatureEarly (0.75) */
  double chosen_move_temperature_only_below_prob; /* chosenMoveTemperatureOnlyBelowProb (0 = unset = 1.0) */
  double chosen_move_subtract;                    /* chosenMoveSubtract (0) */
  double chosen_move_prune;                       /* chosenMovePrune (1) */
  int32_t nn_cache_size_power_of_two;             /* nnCacheSizePowerOfTwo: evaluation cache shared by the games of this GPU (0 = off) */
  int32_t root_num_symmetries_to_sample;          /* rootNumSymmetriesToSample (4): the root is evaluated under that many symmetries, one per wave */
  int32_t ko_rule;                                /* Rules::koRule: 0 simple, 1 positional superko, 2 situational superko, 3 spight (area scoring) */
  int32_t full_history_rules;                     /* 1 = BoardHistory's game-end rules also under simple ko: a pass in a situation the same
                                                     player already passed in ends the game, a third repetition since the last pass is "no
                                                     result".  Implied by ko_rule != 0.  0 = two consecutive passes only. */
  double root_ending_bonus_points;                /* rootEndingBonusPoints (0.5 in the stock configs): at the root, moves into territory the net's
                                                     ownership head is sure about (|ownership| >= 0.95) lose up to this many points of score
                                                     utility unless they capture / touch the opponent / connect groups that are not pass-alive
                                                     (Search::getEndingWhiteScoreBonus, searchhelpers.cpp:351-420; area scoring).  Needs the
                                                     root to be evaluated by the net (root_num_symmetries_to_sample >= 2 or the cache off). */
} kgb_selfplay_config;

typedef struct kgb_selfplay_stats {
  uint64_t total_visits;     /* playouts completed (sum over games of root visits) */
  uint64_t total_moves;      /* root moves played */
  uint64_t games_finished;
  uint64_t black_wins;
  uint64_t nodes_allocated;
  uint64_t sum_leaf_depth;   /* sum over playouts of the leaf depth */
  uint64_t ladder_searches;  /* ladder searches run for feature planes 14-17 */
  uint64_t ladder_nodes;     /* moves played inside those searches */
  uint64_t stalled_waves;    /* game-waves that finished no playout: ladder searches still running (ladder_nodes_per_wave), every
                                playout of the wave ended on an existing edge, or one of the root's extra symmetric evaluations */
  uint64_t instant_playouts; /* graph search: playouts that ended on an edge catch-up or a cycle and needed no evaluation */
  uint64_t nn_cache_hits;    /* playouts whose leaf evaluation came from the evaluation cache */
  uint64_t nn_cache_stores;
} kgb_selfplay_stats;

/* ScoreValue::expectedWhiteScoreValue (neuralnet/nninputs.cpp:160-192) on the host, with the table the device loop uploads:
 * n independent lookups.  Test hook for the score-utility table. */
KGB_API int kgb_expected_white_score_value(int n, const double* mean, const double* stdev, const double* center, const double* scale,
                                           const double* sqrt_board_area, double* out);
/* The first n 32-bit outputs of the reference's Rand(seed_string) as the library's host generator produces them (core/rand.h:149-152,
 * core/rand.cpp:279-320).  Test hook (row a25): the reference's own self-test vector for "abc" (core/rand.cpp:386-415) pins it. */
KGB_API int kgb_rand_uint32_stream(const char* seed_string, int n, uint32_t* out);
/* The value-weighting CDF table the device loop uploads (Search's DistributionTable over tdistcdf(z, 3), search.cpp:131-137):
 * n must be 2000.  Test hook (row a20). */
KGB_API int kgb_value_weight_cdf_table(double* out, int n);
/* Search::getPlaySelectionValues of game g's root (searchresults.cpp:66-330), indexed by move position (-1 = no child):
 * the weights the root move is drawn from (reduced weights, LCB bonus, subtract / prune applied). */
KGB_API int kgb_selfplay_get_play_selection_values(kgb_selfplay* sp, int game, double* values);
/* TEST HOOK: `count` consecutive draws of Search::chooseIndexWithTemperature (searchhelpers.cpp:12-76) from the device Rand
 * seeded with seed_string. */
KGB_API int kgb_test_choose_index_with_temperature(const char* seed_string, const double* relative_probs, int n, double temperature,
                                                   double only_below_prob, int count, int32_t* chosen);
/* TEST HOOK (row a23): Board::simpleRepetitionBoundGt(move, bound) after every move of a stream (x, y, player 1 black / 2 white; x < 0 pass)
 * - the predicate that decides whether a graph-search node may be shared between histories (game/board.cpp:2853-2888). */
KGB_API int kgb_test_repetition_bound(int x_len, int y_len, int num_moves, int bound, const int8_t* moves_xyp, uint8_t* out);
/* TEST HOOK (row a3): replay num_games games (moves_xy [game][max_moves][2]: x,y; -1,-1 pass; -2 = end of that game; black first) through
 * the device ko rules (ko_rule 0 simple, 1 positional, 2 situational, 3 spight; area scoring).  Per move: flags (1 game over, 2 no result,
 * 4 a pass by the next player would end the phase), legality of every point for the next player (incl. ko and superko bans), and
 * the superko-banned points - BoardHistory::makeBoardMoveAssumeLegal / isLegal / passWouldEndPhase (game/boardhistory.cpp). */
KGB_API int kgb_test_history_replay(int x_len, int y_len, int ko_rule, int multi_stone_suicide_legal, int num_games, int max_moves, const int8_t* moves_xy,
                                    uint8_t* flags, uint8_t* legal_next, uint8_t* super_ko_banned);
/* TEST HOOK (rows a22/a25): the device loop's root-policy temperature + Dirichlet noise on a given policy (-1 = illegal), with the
 * device Rand initialised from seed_string like the reference's Rand(seed_string): the counterpart of
 * Search::maybeAddPolicyNoiseAndTemp / addDirichletNoise (searchhelpers.cpp:78-215). */
KGB_API int kgb_test_root_policy_noise(const char* seed_string, int x_len, int y_len, int policy_size, int turn_number, int noise_enabled,
                                       double concentration, double weight, double temperature, double temperature_early, double halflife,
                                       const float* policy_in, float* policy_out);
KGB_API int kgb_selfplay_create(kgb_handle* handle, const kgb_selfplay_config* config, kgb_selfplay** out);
KGB_API void kgb_selfplay_free(kgb_selfplay* sp);
/* Enqueue `steps` playout waves on the handle's stream (asynchronous; kgb_handle_sync() to wait). */
KGB_API int kgb_selfplay_run(kgb_selfplay* sp, int steps);
KGB_API int kgb_selfplay_get_stats(kgb_selfplay* sp, kgb_selfplay_stats* out);
/* Root position of game g: colors[Y*X] (0 empty, 1 black, 2 white); info[6] = move number, black-to-move, ko point
 * (y*32+x or -1), black stones captured, white stones captured, root visits. */
KGB_API int kgb_selfplay_get_game(kgb_selfplay* sp, int game, uint8_t* colors, int32_t* info);
/* Root children of game g, indexed by move position 0..X*Y (pass last): edge visit counts, NN policy (-1 illegal), and each
 * child's utilityAvg (white's perspective; 0 where there is no child). */
KGB_API int kgb_selfplay_get_root_children(kgb_selfplay* sp, int game, int32_t* visits, float* policy, double* utility_avg);
/* NodeStats moments (searchnode.h:17-41; white's perspective) of game g's root and of its children by move position:
 * winLossValueAvg, noResultValueAvg, scoreMeanAvg, scoreMeanSqAvg, leadAvg.  child_stats[(X*Y+1)*5] (0 where no child), root_stats[5]. */
KGB_API int kgb_selfplay_get_root_value_stats(kgb_selfplay* sp, int game, double* child_stats, double* root_stats);
/* Game recording (replaces the bookkeeping around the search in Play::runGame, program/play.cpp:1757-1936).
 * With kgb_selfplay_config.debug_hold_at_max_visits = 1 a game whose search reached max_visits idles until it is released; the
 * host reads the finished search (getters above and below), releases, and the next wave lets the device choose and play the move
 * as usual.  games_mask[num_games] (1 = release) or NULL = all.  A released game holds again at its next finished search. */
KGB_API int kgb_selfplay_release(kgb_selfplay* sp, const uint8_t* games_mask);
/* Komi per game.  The reference's GameInitializer draws a komi for every game (program/play.cpp:330-420: komiMean, komiStdev,
 * komiBigStdevProb, komiAuto ...); here the host draws and the device applies: komi[num_games] (multiples of 0.5) becomes the komi of
 * each slot's NEXT game, taken when the slot's game in progress ends; also_current_games != 0 replaces the komi of the games in
 * progress too (for games that have not begun to search).  kgb_selfplay_get_komi: komi of the game in progress and of the slot's
 * last finished game (either pointer may be NULL).  The evaluation cache keys on the mover's komi (NNInputs::getHash,
 * neuralnet/nninputs.cpp:869-943 via BoardHistory::getSituationRulesAndKoHash), so games of different komi never share an entry. */
KGB_API int kgb_selfplay_set_komi(kgb_selfplay* sp, const float* komi, int also_current_games);
KGB_API int kgb_selfplay_get_komi(kgb_selfplay* sp, float* current, float* last_finished);
/* Board size and ko / suicide rules per game.  The reference's GameInitializer draws them for every game (program/play.cpp:330-650:
 * bSizes / bSizeRelProbs, koRules, multiStoneSuicideLegals; BASELINE config 4 "mixed 9/13/19 board sizes"); here the host draws and the
 * device applies, like the komi.  setup[num_games][4] = board X, board Y (2..the context's nn_x_len / nn_y_len), ko rule (0-3 as in
 * kgb_selfplay_config.ko_rule), multi-stone suicide legal (0/1).  A game's board occupies the top-left corner of the evaluator's frame:
 * move positions stay y * nn_x_len + x (NNPos::locToPos, neuralnet/nninputs.cpp:27-33), input plane 0 marks the board and the conv
 * trunk masks everything outside it (nneval.cpp:874-883 does the same with nnXLen > board size).  The values become each slot's NEXT
 * game's; also_current_games != 0 applies them to the games in progress too, which must not have started (no move, no visit) -
 * KGB_ERR_INVALID otherwise, nothing changed for those.  A ko rule different from the configuration's needs full_history_rules = 1.
 * kgb_selfplay_get_game_setup: setup of the game in progress and of the slot's last finished game (either pointer may be NULL).
 * The evaluation cache keys on board size and rules as NNInputs::getHash does, so such games never share an entry. */
KGB_API int kgb_selfplay_set_game_setup(kgb_selfplay* sp, const int32_t* setup, int also_current_games);
KGB_API int kgb_selfplay_get_game_setup(kgb_selfplay* sp, int32_t* current, int32_t* last_finished);
/* Search limits per move.  Play::runGame gives every move its own limits (getSearchLimitsThisMove / runBotWithLimits, program/play.cpp:
 * 1093-1300): a "cheap search" (cheapSearchProb, cheapSearchVisits) or a visit count reduced in decided positions (reduceVisits), and for
 * cheap searches that are not recorded the root-only parameters off (no Dirichlet noise, root policy temperature 1, the tree's FPU
 * parameters at the root, no per-child visit floor, a single root symmetry).  Here the host draws (katago_b200/game_recorder.py) and the
 * device applies: visits[num_games][2] in [2, max_visits] and plain_root[num_games][2] (NULL = all 0) are taken by the root that follows
 * the slot's NEXT move - entry 0 if the game goes on, entry 1 if that move ends the game (first root of the slot's next game);
 * also_current_roots != 0 applies entry 0 to the current roots too, which must not have been searched yet (KGB_ERR_INVALID otherwise).
 * A game is held / moves when its root's visits reach ITS budget: kgb_selfplay_get_search_limits returns the current roots' budgets and
 * flags (either pointer may be NULL).  Deviation: the reference keeps the previous tree for unrecorded cheap searches, this loop always
 * starts a move on a cleared tree. */
KGB_API int kgb_selfplay_set_next_search_limits(kgb_selfplay* sp, const int32_t* visits, const uint8_t* plain_root, int also_current_roots);
KGB_API int kgb_selfplay_get_search_limits(kgb_selfplay* sp, int32_t* visits, uint8_t* plain_root);
/* Policy-initialised openings (PlaySettings::initGamesWithPolicy, policyInitAreaProp, policyInitAreaTemperature; PlayUtils::initializeGameUsingPolicy,
 * program/playutils.cpp:232-266, called from Play::runGame, play.cpp:1675-1700): the first num_moves[g] moves of slot g's NEXT game are drawn
 * from the net's own policy ^ (1 / temperature) of each position - one plain evaluation per move, no search, no noise; a slot in its opening is
 * never held for recording (the moves belong to the game's start history, not to its training turns).  The host draws the count (floor of an
 * exponential with mean board area * policyInitAreaProp).  also_current_games != 0: the games in progress take them too; they must not have
 * started (KGB_ERR_INVALID otherwise).  kgb_selfplay_get_policy_init: moves_left[num_games] (> 0 = still in its opening), count[num_games] and
 * moves[num_games][max_moves] = the opening of the game in progress (move positions, pass = nn_x_len * nn_y_len; max_moves <= 512); any
 * pointer may be NULL.  Not restated: the separately noised komi during the opening and the komi compensation after it
 * (compensateAfterPolicyInitProb needs searches before the game). */
KGB_API int kgb_selfplay_set_policy_init(kgb_selfplay* sp, const int32_t* num_moves, double temperature, int also_current_games);
KGB_API int kgb_selfplay_get_policy_init(kgb_selfplay* sp, int32_t* moves_left, int32_t* count, int16_t* moves, int max_moves);
/* FOR TESTING: the evaluation-cache key (the loop's NNInputs::getHash) of the leaf that slot `game` sent to the evaluator in the
 * last wave; only meaningful with nn_cache_size_power_of_two > 0. */
KGB_API int kgb_selfplay_get_leaf_cache_key(kgb_selfplay* sp, int game, uint64_t* key2);
/* Empties the loop's evaluation cache, ordered on the handle's stream: call it with kgb_handle_commit_weights - cached outputs
 * belong to the previous net (the reference gives every NNEvaluator its own NNCacheTable, nneval.cpp:129-130). */
KGB_API int kgb_selfplay_clear_nn_cache(kgb_selfplay* sp);
/* Root visits of every game (visits[num_games]): a game is held when its entry has reached max_visits. */
KGB_API int kgb_selfplay_get_root_visits(kgb_selfplay* sp, int32_t* visits);
/* What extractQValueTargets / computeNNRawStats (play.cpp:859-914) read besides the NodeStats: visits of the root's child NODES by
 * move position (0 = no child), and the root's own evaluation root_nn_stats[5] = winLoss, noResult, scoreMean, scoreMeanSq, lead (white). */
KGB_API int kgb_selfplay_get_root_extra(kgb_selfplay* sp, int game, int32_t* child_node_visits, double* root_nn_stats);
/* symmetries[num_games]: the symmetry (0-7) under which each game's row of the LAST wave was evaluated.  The wave is the loop's batcher (row a9:
 * one row per game, NNEvaluator::serve's batching replaced by the fixed wave); like the reference with nnRandomize = true every row draws its
 * own symmetry (nneval.cpp:698-707) - here from splitmix64 of (seed, game, game index, move, root visits) - except for a root's sampled
 * symmetries (root_num_symmetries_to_sample) and debug_fixed_symmetry. */
KGB_API int kgb_selfplay_get_nn_symmetries(kgb_selfplay* sp, int32_t* symmetries);
/* entropy[num_games]: the entropy of every current root's policy as the net gave it (averaged over the root's symmetric evaluations), BEFORE the
 * root policy temperature and the Dirichlet noise - NNRawStats::policyEntropy of computeNNRawStats (program/play.cpp:890-914; training global 59).
 * The reference takes it from a fresh single-symmetry evaluation of the root; here it is the root's own evaluation. */
KGB_API int kgb_selfplay_get_root_raw_policy_entropy(kgb_selfplay* sp, double* entropy);
/* The last root move of slot `game`: info[4] = move position (X*Y = pass), flags (1 = it ended the game | 2 = without result | 4 = by
 * the move limit), the move number it was played at, the slot's game index.  If it ended the game: final_score = white minus black
 * with komi, final_colors[Y*X] and final_area[Y*X] (0 none, 1 black, 2 white; Board::calculateArea with every flag on, which under
 * area scoring without tax is both the ownership and the full area of FinishedGameData). */
KGB_API int kgb_selfplay_get_last_move(kgb_selfplay* sp, int game, int32_t* info, float* final_score, uint8_t* final_colors, uint8_t* final_area);
/* The NN input row (NHWC [X*Y][22] + 19 globals) the last wave wrote for game g - what NNInputs::fillRowV7 would produce
 * for that leaf (planes listed in DESIGN.md §8; used by the feature parity tests). */
KGB_API int kgb_selfplay_get_nn_row(kgb_selfplay* sp, int game, float* spatial, float* global);
/* The NN input row (NNInputs::fillRowV7, NHWC, no symmetry) of the game's CURRENT ROOT, kept on the device from the wave that evaluated
 * the root until the next move: what TrainingWriteBuffers::addRow stores for the turn (dataio/trainingwrite.cpp:463-478).  Valid once the
 * root has been evaluated by the net (always, when root_num_symmetries_to_sample >= 2 or the evaluation cache is off). */
KGB_API int kgb_selfplay_get_root_row(kgb_selfplay* sp, int game, float* spatial, float* global);
/* The moves from the root to the leaf the last wave selected for game g (x,y pairs, -1,-1 = pass; at most max_len pairs are
 * written, *len_out is the full length) and whether that wave delivered a finished leaf (see ladder_nodes_per_wave). */
KGB_API int kgb_selfplay_get_leaf_path(kgb_selfplay* sp, int game, int32_t* moves_xy, int32_t max_len, int32_t* len_out, int32_t* valid_out);
/* Re-seed every game's search-thread generator like the reference's Rand(seed_string) (tests: reproduce a reference search whose
 * SearchThread seed string is known). */
KGB_API int kgb_selfplay_set_search_rand(kgb_selfplay* sp, const char* seed_string);
/* Desynchronise the games (bench / test support): every game plays its own random number (0..max_moves) of uniformly random
 * legal non-pass moves from its current root and clears its tree - positions "from random legal play-outs" (SURVEY.md §8d). */
KGB_API int kgb_selfplay_random_openings(kgb_selfplay* sp, int max_moves);
/* Play a fixed move list on EVERY game's root (x,y pairs, -1,-1 = pass; colours alternate) and clear the trees. */
KGB_API int kgb_selfplay_play_moves(kgb_selfplay* sp, const int8_t* moves_xy, int num_moves);
/* The same for one game only (games of different board sizes need different lists; match play mirrors the opponent's moves into the
 * other net's loop).  A move off the game's board is KGB_ERR_INVALID.  Unlike kgb_selfplay_play_moves, a move that ends the game - by
 * the rules or by max_moves - is treated like the loop's own: result readable with kgb_selfplay_get_last_move, the slot's next game
 * starts (next setup and komi), later moves of the list go to that game. */
KGB_API int kgb_selfplay_play_moves_game(kgb_selfplay* sp, int game, const int8_t* moves_xy, int num_moves);
/* Timing hook for bench.py's tree/board roofline entry: runs `iters` waves of ONLY the select(+board+featurize) and backup
 * kernels (evaluator outputs of the last wave are reused) and returns their CUDA-event averages per launch. */
KGB_API int kgb_selfplay_time_tree_kernels(kgb_selfplay* sp, int iters, float* ms_select, float* ms_backup);
/* Profiling aid: SM-clock spans of every game's block in the LAST select launch, cycles[game][8] = whole block, root move + tree reset
 * (only in waves where the game's visit budget was spent), warp 0 (descent + leaf features), ladder searches, then warp 0 split into
 * descent, liberties + legality, area (Benson), feature-row writes.  The launch lasts as
 * long as its slowest block; this shows which games those are and what they were doing.  clear != 0 zeroes the buffer afterwards. */
KGB_API int kgb_selfplay_debug_cycles(kgb_selfplay* sp, int64_t* cycles, int clear);
/* Kernel launches per playout wave (evaluator launches + 2). */
KGB_API int kgb_selfplay_launches_per_step(const kgb_selfplay* sp);

/* Zobrist data of the reference Board (game/board.cpp:151-216, Rand seeded "Board::initHash()"), regenerated by the backend's
 * own Rand restatement: board_hash[y][x][colour (0 black, 1 white)][2 x u64] and the empty-board hash size_hash[2]. */
KGB_API int kgb_zobrist_tables(int x_size, int y_size, uint64_t* board_hash, uint64_t* size_hash);

/* FOR TESTING (rows a1/a2): replay move streams on the device board.  moves[b][m] = {x, y (or -1,-1 = pass), pla (1 black,
 * 2 white)}; after every move returns stones, simple-ko point {x,y or -1,-1}, capture counters {black stones captured,
 * white stones captured}, liberty class of every stone (1,2,3, 0 = more / empty), Board::isLegal of every point for
 * the NEXT player, Board::pos_hash {hash0, hash1}, and Board::calculateArea (all flags on; 0 none, 1 black, 2 white). */
KGB_API int kgb_test_board_replay(int x_size, int y_size, int num_boards, int num_moves, int multi_stone_suicide_legal, const int8_t* moves,
                          uint8_t* colors, int8_t* ko, int16_t* caps, uint8_t* lib_class, uint8_t* legal_next, uint64_t* pos_hash, uint8_t* area);

/* Kernel-level timing hook for bench.py's roofline object: runs ONE convolution layer (random weights/inputs resident in
 * HBM, the production epilogue of a residual unit's first conv: BN + mish + mask -> fp16) `iters` times on a stream and returns the CUDA-event average per launch. */
KGB_API int kgb_bench_conv(int ky, int kx, int in_c, int out_c, int n, int nn_x_len, int nn_y_len, int use_fp16, int warmup, int iters,
                   float* ms_per_launch);
/* The same with the epilogue chosen: epilogue_kind 0 = fp32 raw output only (heads / gpool conv), 1 = BN + mish + mask -> fp16 (first conv
 * of a residual unit), 2 = + residual stream read and rewritten in place (second conv of a unit, a nested block's post conv), 3 = raw
 * stream + activated operand (a nested block's pre conv).  `rotate` >= 1 independent input / output buffer sets are cycled through so
 * that the timed loop's working set exceeds L2. */
KGB_API int kgb_bench_conv_ex(int ky, int kx, int in_c, int out_c, int n, int nn_x_len, int nn_y_len, int use_fp16, int epilogue_kind, int rotate,
                      int warmup, int iters, float* ms_per_launch);
/* Test hook for the fused epilogues (kinds 1-3 above) of one convolution layer: NHWC input [n][y][x][in_c], optional NHWC residual
 * [n][y][x][out_c], per-channel BN scale / bias (NULL = 1 / 0), activation (0 identity, 1 relu, 2 mish); returns the raw stream after
 * the launch (kinds 2, 3) and the activated fp16 operand, both NHWC [n][y][x][out_c] as floats.  Also checks that every pad row of the
 * activated operand was written as zero (reference semantics: the conv's zero padding, eigenbackend.cpp:448-701). */
KGB_API int kgb_test_conv_epilogue(int ky, int kx, int in_c, int out_c, const float* weights, int n, int nn_x_len, int nn_y_len, int use_fp16,
                           int epilogue_kind, const float* input, const float* residual_in, const float* bn_scale, const float* bn_bias,
                           int activation, float* raw_out, float* act_out);

#ifdef __cplusplus
}
#endif
#endif /* KGB200_H_ */
