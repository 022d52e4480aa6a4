/*
 * decagon_b200 -- C ABI of the B200-native Decagon hot path (libdecagon_b200.so).
 *
 * The reference (jrectorb/decagon) has no FFI: its "plugin API" for this path is a Python
 * object surface plus tf.Session.run(fetches, feed_dict).  Each entry point below names the
 * reference interface it replaces (file:line under the reference tree).  The Python shim in
 * decagon_b200/ (same class names as decagon/deep/*.py) binds these with ctypes; see
 * INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative dgn_status on failure;
 *     dgn_last_error() returns a thread-local, human-readable description.
 *   - all pointers are HOST pointers unless the name ends in _dev; arrays are caller-owned
 *     and copied during the call (feed values are copied in on every session.run in the
 *     reference too; here the graph is copied ONCE, not per step).
 *   - one graph handle = one GPU = one host thread at a time (the reference drives
 *     everything from one Python thread, DecagonTrainer.py:35-42).
 *   - node types t < n_types, groups g < n_groups are (row type i, col type j) pairs in the
 *     caller's dict order, flat relation id r enumerates groups in that order then k
 *     (minibatch.py:45-54).
 *   - dense matrices cross this boundary row-major float32, exactly as TF would fetch them.
 */
#ifndef DECAGON_B200_H
#define DECAGON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dgn_graph dgn_graph;

typedef enum {
    DGN_OK = 0,
    DGN_ERR_INVALID = -1,   /* bad argument / call order (ValueError / AssertionError upstream) */
    DGN_ERR_CUDA = -2,      /* CUDA runtime failure, message carries cudaGetErrorString */
    DGN_ERR_NO_DEVICE = -3, /* no CUDA device: there is NO CPU fallback */
    DGN_ERR_UNSUPPORTED = -4
} dgn_status;

/* decoder kinds, model.py:90-137 */
typedef enum { DGN_DEC_INNERPRODUCT = 0, DGN_DEC_DISTMULT = 1, DGN_DEC_BILINEAR = 2, DGN_DEC_DEDICOM = 3 } dgn_decoder;
/* optimizer.py:109 (hinge, active) / :110,122-127 (xent) */
typedef enum { DGN_LOSS_HINGE = 0, DGN_LOSS_XENT = 1 } dgn_loss;
/* parameter kinds; names of the reference variables in brackets */
typedef enum {
    DGN_PARAM_W1 = 0,      /* layers.py:80-83   weights_%d of GraphConvolutionSparseMulti, [F_j, hidden1] */
    DGN_PARAM_W2 = 1,      /* layers.py:104-107 weights_%d of GraphConvolutionMulti,       [hidden1, hidden2] */
    DGN_PARAM_DEC_GLOBAL = 2, /* layers.py:127-128 global_interaction (dedicom only), [hidden2, hidden2], k ignored */
    DGN_PARAM_DEC_LOCAL = 3   /* layers.py:129-133 local_variation_%d [hidden2] (dedicom);
                                 :156-160 relation_%d [hidden2] (distmult); :181-184 relation_%d [hidden2,hidden2] (bilinear) */
} dgn_param;
/* per-node-type / per-group tensors that session.run can fetch */
typedef enum {
    DGN_TENSOR_HIDDEN1 = 0,     /* model.py:74-75  hidden1[t]            [n_t, hidden1] */
    DGN_TENSOR_EMBEDDINGS = 1,  /* model.py:85-88  embeddings[t]         [n_t, hidden2] */
    DGN_TENSOR_LAYER1_GROUP = 2,/* layers.py:92-93 output of GraphConvolutionSparseMulti for group g [n_i, hidden1] */
    DGN_TENSOR_LAYER2_GROUP = 3,/* layers.py:116-117 embeddings_reltyp entry of group g             [n_i, hidden2] */
    DGN_TENSOR_GRAD_EMBEDDINGS = 4 /* dL/d embeddings[t] of the last train step (optimizer.grads_vars debugging) */
} dgn_tensor;

const char *dgn_last_error(void);
int dgn_version(void);
int dgn_device_count(int *count_out);

/* ---- host-only helpers (no device needed) ---------------------------------------------- */

/* Canonical CSR of one relation from the COO tuple the iterator feeds
 * (minibatch.py:259-267 feeds preprocess_graph's tuple, minibatch.py:80-93).  Entries are
 * ordered by (row, col); duplicates are kept as separate entries in input order.
 * rowptr_out has n_rows+1 entries, col_out / val_out nnz entries. */
int dgn_csr_from_coo(int32_t n_rows, int32_t n_cols, int64_t nnz, const int32_t *coo_rows, const int32_t *coo_cols,
                     const float *vals, int32_t *rowptr_out, int32_t *col_out, float *val_out);

/* uint32 CDF thresholds of degrees^0.75 for the negative sampler
 * (optimizer.py:40-47 fixed_unigram_candidate_sampler, distortion 0.75). */
int dgn_sampler_thresholds(const double *degrees, int32_t n, uint32_t *thresholds_out);

/* ---- graph lifetime --------------------------------------------------------------------- */

/* DecagonModel.__init__ (model.py:48-62): edge_types -> groups / K, num_feat -> feat_dim,
 * FLAGS.hidden1 / hidden2 (model.py:68,80).
 * Supported: hidden1 in 1 .. 128, hidden2 in 1 .. 32 (anything else: DGN_ERR_UNSUPPORTED).  The device works on 32, 64 or
 * 128 hidden and 32 embedding columns; other sizes are zero-padded INSIDE the library (columns of W1 / W2, rows of W2,
 * rows / columns of the decoder variables), which is the caller's model term by term: padded columns stay exactly zero
 * through forward, backward and Adam.  Every array that crosses this boundary has the caller's shape.  (The keep bit of
 * element (row, col) of the layer-2 input is bit row * stride + col of the relation's dropout stream, stride = the
 * padded hidden1.)
 * NOT supported: a per-relation activation inside the graph-convolution layers.  GraphConvolutionSparseMulti /
 * GraphConvolutionMulti default to act = tf.nn.relu applied to every relation's product BEFORE add_n
 * (layers.py:73,91,99,115); DecagonModel always constructs them with act = lambda x: x (model.py:71,82) and applies ONE
 * relu after the sum over the groups of a node type (model.py:74-75), which is what this library computes.  Summing
 * the relations in registers is only possible because nothing non-linear sits between a relation's product and the
 * sum; the Python layer classes raise NotImplementedError for any other act. */
int dgn_graph_create(dgn_graph **out, int device, int n_types, const int32_t *n_nodes, const int32_t *feat_dim,
                     int n_groups, const int32_t *group_ij, const int32_t *group_K, const int32_t *group_decoder,
                     int hidden1, int hidden2);
int dgn_graph_destroy(dgn_graph *g);

/* placeholders['adj_mats_%d,%d,%d'] (DecagonDataSet.py:112-115; fed at minibatch.py:262):
 * normalised adjacency of flat relation r as COO float32.  Set once (or again after the
 * iterator is rebuilt); NOT per step. */
int dgn_graph_set_relation(dgn_graph *g, int r, int32_t n_rows, int32_t n_cols, int64_t nnz, const int32_t *coo_rows,
                           const int32_t *coo_cols, const float *vals);
/* placeholders['feat_%d'] (DecagonDataSet.py:117-118; minibatch.py:264): sparse node features
 * of one type as COO float32.  Identity features (every BASELINE config) take the row-gather
 * fast path (layers.py:89 becomes a masked row gather of W1_k).  Any other matrix (the public data's multi-hot
 * drug features, DecagonPublicDataNodeFeaturesBuilder.py:34-51) is canonicalised to CSR sorted by (row, col);
 * layer 1 then runs P1_k = (X (.) m_k / q) W1_k followed by the SpMM on P1_k, dropout bit e of a relation =
 * non-zero e in that order (dropout_sparse, layers.py:23-31), and dW1_k = (X (.) m_k / q)^T (A_k^T dS1). */
int dgn_graph_set_features(dgn_graph *g, int type, int32_t n_rows, int32_t n_cols, int64_t nnz,
                           const int32_t *coo_rows, const int32_t *coo_cols, const float *vals);
/* DecagonOptimizer's per-relation unigram tables (optimizer.py:36-47): degrees[i][k] of the
 * row type of relation r. */
int dgn_sampler_set_degrees(dgn_graph *g, int r, const double *degrees, int32_t n);
/* Upload everything set so far and build the device-side layout; required before any
 * compute call and after every dgn_graph_set_* call. */
int dgn_graph_finalize(dgn_graph *g);

/* canonical device CSR of relation r copied back (parity tests: must equal scipy's) */
int dgn_graph_relation_nnz(dgn_graph *g, int r, int64_t *nnz_out);
int dgn_graph_get_csr(dgn_graph *g, int r, int32_t *rowptr_out, int32_t *col_out, float *val_out);

/* ---- one process per GPU (the reference is single-device) ----------------------
 * The relations of every group that takes the staged path (many small relations: the 964 drug-drug side
 * effects and their transposes) are partitioned over `world` ranks of ONE NVSwitch box: a rank keeps the
 * adjacency, layer-1 / layer-2 weights, dropout words and Adam state of its own relations only.  The other
 * groups and every decoder variable are replicated.  Per step three partial sums ([n_i, hidden1],
 * [n_i, hidden2], [n_j, hidden1] per partitioned group) are exchanged: every rank publishes its partial in
 * its exchange arena and reads the peers' arenas over NVLink (CUDA IPC mappings), summing in rank order, so
 * all ranks hold bit-identical embeddings.  Call order: dgn_graph_create, dgn_comm_init, dgn_graph_set_* for
 * ALL relations on every rank, dgn_graph_finalize, dgn_comm_handle -> all-gather of the 64-byte handles by
 * the host program -> dgn_comm_connect, then parameters and compute.  dgn_encoder_forward and dgn_train_step
 * are collective.  With world > 1 the parameter arena exists after dgn_graph_finalize. */
int dgn_comm_init(dgn_graph *g, int rank, int world);
int dgn_comm_handle(dgn_graph *g, void *handle_out_64_bytes);
int dgn_comm_connect(dgn_graph *g, const void *handles /* [world][64] */);
/* owner rank of flat relation r (-1: its group is replicated); dgn_params_get / dgn_grads_get return
 * zeros for the encoder weights of relations another rank owns, dgn_params_set skips them */
int dgn_relation_owner(dgn_graph *g, int r, int *owner_out);
/* host-only: the longest-processing-time assignment the library uses (weights = nnz + columns) */
int dgn_partition_relations(const int64_t *weights, int32_t K, int32_t world, int32_t *owner_out);

/* ---- parameters (tf.Variable) ------------------------------------------------------------ */

/* k >= 0: one relation's variable; k == -1: all K variables of the group, stacked.
 * n = number of floats at ptr (checked). */
int dgn_params_set(dgn_graph *g, int kind, int group, int k, const float *values, int64_t n);
int dgn_params_get(dgn_graph *g, int kind, int group, int k, float *values_out, int64_t n);
int dgn_params_count(dgn_graph *g, int64_t *n_out); /* total trainable floats */
/* tf.global_variables_initializer() for the optimizer slots: m = v = 0, beta powers reset. */
int dgn_optimizer_reset(dgn_graph *g, float beta1, float beta2, float epsilon);

/* ---- compute ----------------------------------------------------------------------------- */

/* Encoder only (model.py:64-88): what fetching model.embeddings / hidden1 runs.
 * dropout is the fed placeholders['dropout'] (rate, float32); masks come from the
 * Philox4x32-10 streams keyed by (seed, step) -- see DESIGN.md "Randomness". */
int dgn_encoder_forward(dgn_graph *g, float dropout, uint64_t seed, uint32_t step);

/* One session.run([opt_op, cost, batch_edge_type_idx]) (DecagonTrainer.py:94-100):
 * encoder forward, minibatch scores of relation r for batch[B,2] and B negatives
 * (optimizer.py:51-57), loss, full backward, TF1 Adam on every variable (optimizer.py:111-113).
 * negatives == NULL: sampled on device from relation r's unigram^0.75 table (optimizer.py:36-49);
 * otherwise int64[B] row-node indices are used as given (parity runs).
 * apply_update == 0 computes loss and gradients only (optimizer.grads_vars).
 * loss_out (host float, may be NULL) is written before the call returns.  The call returns as soon as the LOSS has
 * reached the host (right after the decode kernel); the backward pass and the optimizer it queued are still running
 * and complete before any later call on this handle can observe a variable, gradient or embedding (one stream order;
 * dgn_sync waits for everything).  DGN_SYNC_LOSS=1 makes the call wait for the whole step instead. */
int dgn_train_step(dgn_graph *g, int r, const int32_t *batch, int32_t batch_size, const int64_t *negatives,
                   int loss_kind, float margin, float neg_weight, float learning_rate, float dropout, uint64_t seed,
                   uint32_t step, int apply_update, float *loss_out);

/* optimizer.outputs / neg_outputs / neg_samples of the last dgn_train_step (optimizer.py:49-57) */
int dgn_last_batch_outputs(dgn_graph *g, float *pos_out, float *neg_out, int64_t *neg_samples_out, int32_t batch_size);
/* gradient of the last step for one variable (optimizer.grads_vars, optimizer.py:114).  Only a step run
 * with apply_update == 0 materialises every gradient: with apply_update != 0 the Adam update of the
 * layer-1 weights of the many-relation groups is fused into the kernel that produces their gradient. */
int dgn_grads_get(dgn_graph *g, int kind, int group, int k, float *values_out, int64_t n);
/* keep != 0: every later dgn_train_step materialises every gradient even with apply_update != 0 (the fused
 * update is replaced by the separate Adam kernel; same parameters bit for bit) -- what fetching
 * [opt.opt_op, opt.grads_vars] in one session.run needs (optimizer.py:111-114).  Without it dgn_grads_get
 * fails for the layer-1 weights whose gradient the fused step never stored. */
int dgn_keep_gradients(dgn_graph *g, int keep);

/* optimizer.predictions (optimizer.py:87-106): Z_i loc glb loc Z_j^T of relation r from the
 * CURRENT embeddings (call dgn_encoder_forward first), row-major [n_i, n_j]. */
int dgn_predict_all_pairs(dgn_graph *g, int r, float *out);
/* Batched form for evaluateAll (DecagonAccuracyEvaluator.py:57-91): relations r0 .. r0+count-1
 * of one group into out_dev[count, n_i, n_j] (DEVICE pointer) */
int dgn_predict_relations_dev(dgn_graph *g, int r0, int count, float *out_dev);
/* sigma(P_r)[u*n_j+v] at the given coordinates (DecagonAccuracyEvaluator.py:123-186) without
 * materialising P_r on the host */
int dgn_predict_edges(dgn_graph *g, int r, const int32_t *edges, int32_t n_edges, int apply_sigmoid, float *out);

/* DecagonAccuracyEvaluator.evaluateAll (DecagonAccuracyEvaluator.py:57-91) in one call, after ONE
 * dgn_encoder_forward: edge e = (rel_k[e], edges[2e], edges[2e+1]) of `group` is scored like dgn_predict_edges
 * (one launch for all relations); scores_out (host, [n_edges], may be NULL) receives the scores in input order.
 * With labels (1 = positive, 0 = negative) and auroc_out / auprc_out non-NULL the pooled scores are sorted on
 * the device and sklearn.metrics.roc_auc_score / average_precision_score (:69-75) are evaluated there
 * (ties share one threshold; NaN when a class is empty, where sklearn raises ValueError). */
int dgn_evaluate_edges(dgn_graph *g, int group, int64_t n_edges, const int32_t *rel_k, const int32_t *edges,
                       const uint8_t *labels, int apply_sigmoid, float *scores_out, double *auroc_out,
                       double *auprc_out);

/* GreedyActiveLearner._getRankedPossibilities (main/ActiveLearner/GreedyActiveLearner.py:84-92: sigmoid of the
 * predictions of one relation, np.take at the candidate coordinates, np.argsort(...)[::-1]) without the [n_i, n_j]
 * matrix: the candidates edges[2e], edges[2e+1] of relation r are scored and sorted on the device;
 * order_out[0 .. top) = indices of the best candidates in descending score order (ties in input order),
 * scores_out (may be NULL) their scores. */
int dgn_rank_edges(dgn_graph *g, int r, const int32_t *edges, int64_t n_edges, int apply_sigmoid, int64_t top,
                   int32_t *order_out, float *scores_out);

/* model.embeddings[t], model.hidden1[t], per-group layer outputs (DecagonLogger.py:239-242) */
int dgn_tensor_get(dgn_graph *g, int which, int index, float *out, int64_t n);

/* Load saved embeddings (the .npy dump of DecagonLogger.py:232-248) so that the predict calls can
 * score them without running the encoder -- the use case of main/Predictor/NpPredictor.py:293-313.
 * Only DGN_TENSOR_EMBEDDINGS can be set. */
int dgn_tensor_set(dgn_graph *g, int which, int index, const float *values, int64_t n);

/* latent_inters[r] / latent_varies[r] (model.py:116-137) as dense [hidden2, hidden2] */
int dgn_relation_matrices(dgn_graph *g, int r, float *glb_out, float *loc_out);

int dgn_sync(dgn_graph *g);

/* ---- measurement hooks (bench.py) -------------------------------------------------------- */
/* CUDA-event time in ms (events recorded on the library's own stream) of phase `name`
 * accumulated since the last dgn_timing_reset and the number of brackets summed; `name` is a
 * PREFIX.  Phases: "mask", "spmm_fwd1/g<group>", "spmm_fwd2/g<group>", "spmm_bwd2/g<group>",
 * "spmm_bwd1/g<group>", "project/g<group>", "dw2/g<group>", "dh/g<group>", "exchange/g<group>" (multi-GPU: publish + wait
 * for every peer), "decode", "adam", "epilogue", "predict".
 * dgn_launch_count: kernels launched by the library since the last dgn_timing_reset. */
int dgn_timing_enable(dgn_graph *g, int enable);
int dgn_timing_reset(dgn_graph *g);
int dgn_timing_get(dgn_graph *g, const char *name, double *ms_out, int64_t *count_out);
int dgn_launch_count(dgn_graph *g, int64_t *launches_out);
/* graph_replays: training steps executed as ONE CUDA-graph launch (from the second step of a configuration on; the
 * kernels inside still count in dgn_launch_count); groups_rebuilt: how many (i, j) groups dgn_graph_finalize has
 * built so far -- a finalize after dgn_graph_set_relation rebuilds only the groups whose relations changed (the
 * active-learning rounds of RandomMaskingActiveLearner.py:150-200 re-mask the drug-drug relations only).  Either
 * pointer may be NULL. */
int dgn_counters(dgn_graph *g, int64_t *graph_replays_out, int64_t *groups_rebuilt_out);
/* recorded phase `index` (0 .. until DGN_ERR_INVALID): name, stream lane, start / stop in ms after the first
 * recorded phase began -- a timeline of one step across the two lanes (tools/timeline.py) */
int dgn_timeline_get(dgn_graph *g, int index, char *name_out, int name_cap, int *lane_out, double *start_ms_out,
                     double *stop_ms_out);
/* CUDA events on the library's stream around a whole region (bench.py times K steps with it) */
int dgn_timer_start(dgn_graph *g);
int dgn_timer_stop(dgn_graph *g, double *ms_out);
int dgn_memory_bytes(dgn_graph *g, int64_t *free_out, int64_t *total_out);

#ifdef __cplusplus
}
#endif
#endif
