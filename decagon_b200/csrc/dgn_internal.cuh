// Internal declarations shared by the translation units of libdecagon_b200.so.
// Everything device-side is written for sm_100a (B200); there is no CPU fallback.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/decagon_b200.h"

namespace dgn {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
struct Failure {
    int code;
};
#define DGN_FAIL(code_, ...)          \
    do {                              \
        ::dgn::set_error(__VA_ARGS__); \
        throw ::dgn::Failure{code_};  \
    } while (0)
#define DGN_REQUIRE(cond, ...) \
    do {                       \
        if (!(cond)) DGN_FAIL(DGN_ERR_INVALID, __VA_ARGS__); \
    } while (0)
#define CUDA_CHECK(expr)                                                                            \
    do {                                                                                            \
        cudaError_t err__ = (expr);                                                                 \
        if (err__ != cudaSuccess)                                                                   \
            DGN_FAIL(DGN_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                     cudaGetErrorString(err__));                                                    \
    } while (0)

// cudaFuncSetAttribute is per device: launchers remember which devices they have configured, not "done once"
// (one process may own several graph handles on different GPUs)
struct PerDeviceOnce {
    bool done[64] = {};
    bool first() {  // true the first time it is called on the current device
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) return true;
        const bool f = !done[dev];
        done[dev] = true;
        return f;
    }
};

// ---------------------------------------------------------------- host-side sparse structures
struct HostCsr {
    int n_rows = 0, n_cols = 0;
    std::vector<int> rowptr, col;
    std::vector<float> val;
    int64_t nnz() const { return (int64_t)col.size(); }
};
void csr_from_coo(int n_rows, int n_cols, int64_t nnz, const int32_t *rows, const int32_t *cols, const float *vals,
                  HostCsr &out);
void csr_transpose(const HostCsr &a, HostCsr &out);

struct DevCsr {
    int n_rows = 0;
    int64_t nnz = 0;
    int *rowptr = nullptr, *col = nullptr;
    float *val = nullptr;
};

// Rows cut into segments of at most seg_len non-zeros; one warp per segment.
// trivial: every row fits one segment and no table is stored (segment id == row id).
struct SegTable {
    bool trivial = true;
    int seg_len = 0;
    int n_seg = 0;
    int n_multi = 0;              // rows that span more than one segment
    int *seg_row = nullptr;       // [n_seg]
    int *seg_begin = nullptr;     // [n_seg]
    int *row_seg_ptr = nullptr;   // [n_rows + 1]
    int *multi_rows = nullptr;    // [n_multi]
};

// ---------------------------------------------------------------- device-side argument blocks
constexpr int kMaxGroupsPerType = 8;
constexpr int kMaxWorld = 8;          // ranks of one NVSwitch box
constexpr float kL2Eps = 1e-12f;  // tf.nn.l2_normalize epsilon (layers.py:93,117)

struct SpmmArgs {
    // sparse operand
    const int *rowptr, *col;
    const float *val;
    const int *seg_row, *seg_begin, *row_seg_ptr;  // null when the table is trivial
    int seg_len, n_seg, n_rows;
    // dense operand, panel layout [P][op_rows][32]
    const float *op;
    int op_rows;
    // result: direct rows go to out[P][out_rows][32]; rows split over segments go to
    // partial[n_seg][P][32] (force_partial: every segment goes to partial)
    float *out;
    int out_rows;
    float *partial;
    int force_partial;
    // dropout of the layer-1 feature rows (identity features): bit index = column (col_mask)
    // or = output row (row_mask); kept entries are scaled by `scale`
    const uint32_t *mask;
    int col_mask, row_mask;
    float scale;
};

struct StagedArgs {
    // per-relation CSR of the group, concatenated: rowptr[K][n_i + 1] holds offsets into col/val
    const int *rowptr, *col;
    const float *val;
    int K, n_i, n_j;
    // operand tiles: tile (p, k) = op[((p * K + k) * n_j) * 32 ...], n_j * 32 floats
    const float *op;
    int P;
    // work lists: slot s handles relations slot_rel[slot_ptr[s] .. slot_ptr[s+1])
    const int *slot_ptr, *slot_rel;
    int n_slots;
    float *partial;  // [n_slots][P][n_i][32]
    const uint32_t *mask;  // layer-1 dropout bits, bit index k * n_j + c (null: none)
    float scale;
};

// Staged SpMM v3 ("warp-task streams").  The 4 quarter-warps of a warp work on 4 rows in lockstep;
// rows are sorted by non-zero count (summed over the relations, or per relation) and consecutive
// groups of 4 are dealt to the warps in snake order; slot s of warp w is its s-th group.  A CTA owns a
// list of relations (slot table) and every warp w of it reads ONE contiguous stream: for each relation
// of the list in order, for each slot s, cnt(k, w, s) pair-steps; a pair-step is 4 x int4 = one
// {offset0, value0, offset1, value1} per quarter-warp (offset = byte offset of the operand row in a
// [rows][32] tile; padding = (0, 0.0f)).
struct TaskArgs {
    const int *hdr;      // [K][n_warps][4]: 8 x uint16 pair-step counts of the slots
    const int4 *ent;     // the streams
    const int *wstart;   // [n_slots][n_warps]: first pair-step of the warp's stream for this slot table
    const int *orow;     // result row of (w, s, quarter) or -1: [n_warps * rpq * 4] if orow_stride == 0 else [K][...]
    int orow_stride;
    int K, n_warps, rpq;
    int n_out_rows;      // rows of one relation's result
    int n_op_rows;       // rows of the dense operand (tile rows forward, dS rows backward)
    const float *op;     // forward: tile (p, k) at op + ((p * K + k) * n_op_rows) * 32;  backward: [P][n_op_rows][32]
    int P;
    const int *slot_ptr, *slot_rel;
    int n_slots;
    float *out;          // forward: partial [n_slots][P][n_out_rows][32];  backward: [P][K * n_out_rows][32]
    const uint32_t *mask;  // forward: bit k * n_op_rows + operand row;  backward: bit k * n_out_rows + result row
    float scale;
    // backward only: TF1 Adam applied to the rows as they are produced (out is then not written);
    // p / m / v have the layout of out
    float *adam_p, *adam_m, *adam_v;
    const struct StepDyn *dyn;  // Adam coefficients of the step (device memory)
};

struct EpiGroup {
    const float *partial;
    const int *row_seg_ptr;  // segment mode; null => slot mode
    int n_slots;             // slot mode: partial[n_slots][P][n_rows][32]
    int n_peers;             // > 0: peer mode, the sum over ranks (in rank order) of peer[r][P][n_rows][32]
    const float *peer[kMaxWorld];
    float *Y;                // [P][n_rows][32] normalised rows
    float *nrm;              // [n_rows]  sqrt(max(|S|^2, eps))
};
struct EpiArgs {
    int n_rows, n_groups, relu;
    EpiGroup g[kMaxGroupsPerType];
    float *out;  // [P][n_rows][32]
};

struct L2BwdArgs {  // dS = l2norm backward of one group
    const float *Y, *nrm, *dY;
    float *dS;
    int n_rows;
};

struct ReluBwdGroup {
    const float *part;  // [n_chunks][P][n_rows][32]
    int n_chunks;
    int n_peers;        // > 0: the sum over ranks (in rank order) of peer[r][P][n_rows][32] instead
    const float *peer[kMaxWorld];
};
struct ReluBwdArgs {
    int n_rows, n_groups;
    ReluBwdGroup g[kMaxGroupsPerType];
    const float *H;  // [P][n_rows][32]
    float *dA;       // [P][n_rows][32]
};

struct DenseArgs {
    // one group, layer 2.  H: [P1][n_j][32]; W2: [K][D1][D2] row-major; P2/G2: [K*n_j][D2]
    const float *H, *W2;
    float *P2;        // forward output
    const float *G2;  // backward input
    float *dW2;       // [K][n_rb][D1*D2]: one partial per row block (n_rb == 1: the gradient itself)
    float *dHpart;    // [n_slots][P1][n_j][32]: one partial per slot of relations
    const uint32_t *mask;  // [K*n_j*P1] words or null
    float scale;
    int K, n_j;
    int n_rb;     // row blocks of dense_row_block(D1) rows
    int n_slots;  // persistent CTAs per row block; slot s owns relations [s K / n_slots, (s+1) K / n_slots)
};

// first-layer product with general (non-identity) sparse features, features.cu
struct FeatArgs {
    const int *rowptr, *col;  // CSR view: X by node row (forward) or X^T by feature (backward)
    const float *val;
    const int *eid;           // backward view: position of the entry in X's canonical order; forward: null (= e)
    int n_rows;               // rows of the view = rows per relation of the output
    int in_rows;              // rows per relation of the dense input
    const float *in;          // [P][K * in_rows][32]
    float *out;               // [P][K * n_rows][32]
    int K, P;
    long long nnz;            // non-zeros of X: relation k owns dropout bits [k nnz, (k + 1) nnz)
    const uint32_t *mask;     // packed keep bits or null
    float scale;
};
void launch_feature_product(const FeatArgs &a, cudaStream_t s);

struct DecodeArgs {
    const float *Zi, *Zj;  // [n][32]
    long long *dZi, *dZj;  // 2^-40 fixed point (order-independent scatter-add)
    int n_i, n_j;
    const int *batch;   // [B][2]
    const long long *neg_in;  // [B] or null
    long long *neg_out;       // [B]
    const uint32_t *thr;      // sampler CDF of the row type for this relation, n_thr entries
    int n_thr;
    int B, decoder, loss_kind;
    float margin, neg_weight;
    const float *glb, *loc;   // decoder parameters of this relation (see decode.cu)
    float *g_glb, *g_loc;     // their gradients
    float *pos_out, *neg_score_out, *loss_out;
    uint32_t seed_lo, seed_hi, step, relation;
    float *scratch;      // [kDecodeCtas][32 * 32 + 1] per-CTA dM and loss partials
    unsigned *ticket;    // zero between launches
};
constexpr int kDecodeCtas = 16;
constexpr int kMaxExchanges = 16;  // flag slots per source rank

// Everything that changes from one training step to the next, in DEVICE memory: the kernels read it from here, so
// the launches themselves are identical every step and the whole step can be replayed as one CUDA graph.  The host
// fills a pinned copy and moves it (together with the batch) with ONE host-to-device copy per step.
struct StepDyn {
    uint32_t step, seed_lo, seed_hi, threshold;  // dropout streams: Philox key / counter words, keep threshold
    float alpha, omb1, omb2, eps;                // TF1 Adam: alpha = lr sqrt(1 - b2^t) / (1 - b1^t)
    uint32_t seq;                                // serial number of the step: stored next to the loss, so that the host can tell
                                                 // this step's loss from the previous one in its pinned buffer
    uint32_t stamp[kMaxExchanges];               // multi-GPU: the stamp each exchange publishes / waits for this step
    DecodeArgs dec;                              // minibatch, relation, decoder variables, loss
};

struct PredictArgs {
    const float *Zi, *Zj;
    int n_i, n_j, decoder;
    const float *glb;         // group-level parameter (dedicom R) or null
    const float *loc;         // first relation's local parameter
    long long loc_stride;     // floats between consecutive relations' local parameters
    int count;
    float *out;               // [count][n_i][n_j]
};

// TF-1.8 ApplyAdam on one element (optimizer.py:111-113; training_ops.cc ApplyAdam: m += (g - m)(1 - b1);
// v += (g g - v)(1 - b2); p -= m alpha / (sqrt(v) + eps)).  Explicit round-to-nearest intrinsics, no FMA
// contraction: the separate adam_kernel and the update fused into spmm_tstaged_kernel round identically (and like
// the float32 numpy restatement in oracle/decagon_oracle.py AdamTF1).
__device__ __forceinline__ void adam_update(float &p, float &m, float &v, float g, float alpha, float omb1, float omb2,
                                            float eps) {
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), omb1));
    v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), omb2));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(m, alpha), __fadd_rn(__fsqrt_rn(v), eps)));
}

// ---------------------------------------------------------------- kernel launchers (host)
void launch_spmm(const SpmmArgs &a, int P, cudaStream_t s);
void launch_seg_reduce(const SpmmArgs &a, const int *multi_rows, int n_multi, int P, cudaStream_t s);
void launch_spmm_staged(const StagedArgs &a, cudaStream_t s);
size_t staged_smem_bytes(int n_j);
constexpr int kS3Warps = 31;   // forward v3: 31 consumer warps + 1 TMA producer warp
constexpr int kTsWarps = 32;   // backward: 32 consumer warps, operand resident in shared memory
void launch_spmm_staged3(const TaskArgs &a, cudaStream_t s);
void launch_spmm_tstaged(const TaskArgs &a, cudaStream_t s);
bool staged3_supported(int n_i, int n_j, int K);
bool tstaged_supported(int n_i, int n_j, int K, int P);
bool staged_supported(int n_i, int n_j, int K);
void launch_node_epilogue(const EpiArgs &a, int P, cudaStream_t s);
void launch_l2norm_bwd(const L2BwdArgs &a, int P, cudaStream_t s);
void launch_relu_bwd(const ReluBwdArgs &a, int P, cudaStream_t s);
constexpr int kMaxMaskBatch = 8;
struct MaskBatch {  // layer-1 keep bits of up to kMaxMaskBatch groups, one launch
    int n;
    uint32_t *words[kMaxMaskBatch];
    long long n_words[kMaxMaskBatch], bits_per_rel[kMaxMaskBatch];
    const int *rel_ids[kMaxMaskBatch];
};
void launch_gen_mask_multi(const MaskBatch &mb, uint32_t stream_id, const StepDyn *dyn, cudaStream_t s);
// step_offset: the masks of step dyn->step + step_offset (1: drawn ahead, while the current step's backward finishes)
void launch_gen_mask(uint32_t *words, long long n_words, long long bits_per_rel, int words_per_rel_or_0, const int *rel_ids,
                     uint32_t stream_id, const StepDyn *dyn, uint32_t step_offset, cudaStream_t s);
int dense_row_block(int D1, int which);
void launch_project(const DenseArgs &a, int D1, int D2, cudaStream_t s);
void launch_dw2(const DenseArgs &a, int D1, int D2, cudaStream_t s);
void launch_dw2_reduce(const float *part, float *out, int K, int n_chunks, int elems, cudaStream_t s);
void launch_dh(const DenseArgs &a, int D1, int D2, cudaStream_t s);
// tcgen05 versions (dense_tc.cu): a.n_rb = row tiles of 128 (project, dh) / chunks of row tiles per relation (dw2)
bool dense_tc_supported(int D1, int D2);
int dense_tc_tiles(int n_j);
void launch_project_tc(const DenseArgs &a, int D1, cudaStream_t s);
void launch_dw2_tc(const DenseArgs &a, int D1, cudaStream_t s);
void launch_dh_tc(const DenseArgs &a, int D1, cudaStream_t s);
void launch_decode(const StepDyn *dyn, cudaStream_t s);  // arguments: dyn->dec
constexpr int kMaxTypes = 8;
struct FixedBatch {
    int count;
    long long *q[kMaxTypes];
    float *out[kMaxTypes];
    size_t n[kMaxTypes];
};
void launch_fixed_to_float_clear(const FixedBatch &fb, cudaStream_t s);
void launch_fixed_to_float(const long long *q, float *out, size_t n, cudaStream_t s);
void launch_predict(const PredictArgs &a, cudaStream_t s);
void launch_predict_tc(const PredictArgs &a, int n_sm, cudaStream_t s);  // tcgen05, 3 x TF32 (predict_tc.cu)
void launch_predict_edges(const PredictArgs &a, const int *edges, int n_edges, int apply_sigmoid, float *out,
                          cudaStream_t s);
void launch_predict_edges_multi(const PredictArgs &a, const int *rel_k, const int *edges, long long n_edges,
                                int apply_sigmoid, float *out, cudaStream_t s);  // evaluate.cu
size_t auc_sort_bytes(long long n);
size_t rank_sort_bytes(long long n);
void launch_rank(const float *scores, long long n, int *idx_in, float *sorted_scores, int *order, void *tmp, size_t tmp_bytes,
                 cudaStream_t s);
void launch_auc(const float *scores, const unsigned char *labels, long long n, float *sorted_scores,
                unsigned char *sorted_labels, void *tmp, size_t tmp_bytes, double *out, cudaStream_t s);
void launch_adam(float *p, const float *g, float *m, float *v, long long n, const StepDyn *dyn, cudaStream_t s);
// multi-GPU exchange (node.cu): ordered sum of the local partials into this rank's exchange buffer; then
// "exchange x reached stamp" is stored into every peer's flag array and awaited from every peer
void launch_publish(const float *partial, int n_chunks, size_t floats, float *out, cudaStream_t s);
void launch_signal_wait(uint32_t *const *peer_flags_dev, uint32_t *my_flags, int rank, int world, int x, const StepDyn *dyn,
                        unsigned long long timeout_ns, int *error_flag, cudaStream_t s);
void launch_relation_matrices(int decoder, const float *glb, const float *loc, float *glb_out, float *loc_out,
                              cudaStream_t s);

}  // namespace dgn
