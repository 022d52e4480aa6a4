// C ABI of libdecagon_b200.so: graph object, device layout, and the orchestration of one
// encoder forward / one training step / all-pairs scoring.  See include/decagon_b200.h for the
// reference interface each entry point replaces.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <numeric>
#include <queue>
#include <tuple>

#include "dgn_internal.cuh"
#include "philox.cuh"

using namespace dgn;

namespace {

template <typename T>
T *dev_alloc(size_t n) {
    T *p = nullptr;
    if (n == 0) n = 1;
    CUDA_CHECK(cudaMalloc(&p, n * sizeof(T)));
    return p;
}
template <typename T>
T *dev_upload(const std::vector<T> &v) {
    T *p = dev_alloc<T>(v.size());
    if (!v.empty()) CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}
template <typename T>
void dev_free(T *&p) {
    if (p) cudaFree(p);
    p = nullptr;
}

void free_csr(DevCsr &c) {
    dev_free(c.rowptr);
    dev_free(c.col);
    dev_free(c.val);
}
void free_seg(SegTable &t) {
    dev_free(t.seg_row);
    dev_free(t.seg_begin);
    dev_free(t.row_seg_ptr);
    dev_free(t.multi_rows);
    t = SegTable();
}

DevCsr upload_csr(const HostCsr &h) {
    DevCsr d;
    d.n_rows = h.n_rows;
    d.nnz = h.nnz();
    d.rowptr = dev_upload(h.rowptr);
    d.col = dev_upload(h.col);
    d.val = dev_upload(h.val);
    return d;
}

// every_row: rows without non-zeros still get one (empty) segment so that their result is written
SegTable build_segments(const HostCsr &h, int seg_len, bool every_row) {
    SegTable t;
    t.seg_len = seg_len;
    int max_row = 0;
    for (int r = 0; r < h.n_rows; ++r) max_row = std::max(max_row, h.rowptr[r + 1] - h.rowptr[r]);
    if (every_row && max_row <= seg_len) {
        t.trivial = true;
        t.n_seg = h.n_rows;
        return t;
    }
    t.trivial = false;
    std::vector<int> seg_row, seg_begin, row_seg_ptr((size_t)h.n_rows + 1, 0), multi;
    for (int r = 0; r < h.n_rows; ++r) {
        const int b = h.rowptr[r], e = h.rowptr[r + 1];
        int n = 0;
        // (capping the segments per row at 16 -- longer segments for hub rows, shorter ordered sums in the epilogue --
        // was measured: the gather kernels get a tail, 80 -> 138 us for the PPI forward; not kept)
        for (int s = b; s < e; s += seg_len) {
            seg_row.push_back(r);
            seg_begin.push_back(s);
            ++n;
        }
        if (n == 0 && every_row) {
            seg_row.push_back(r);
            seg_begin.push_back(b);
            n = 1;
        }
        if (n > 1) multi.push_back(r);
        row_seg_ptr[(size_t)r + 1] = row_seg_ptr[r] + n;
    }
    t.n_seg = (int)seg_row.size();
    t.n_multi = (int)multi.size();
    t.seg_row = dev_upload(seg_row);
    t.seg_begin = dev_upload(seg_begin);
    t.row_seg_ptr = dev_upload(row_seg_ptr);
    t.multi_rows = dev_upload(multi);
    return t;
}

struct SlotTable {
    int n_slots = 0;
    int *ptr = nullptr, *rel = nullptr;
    std::vector<std::vector<int>> lists;  // host copy: relations of every slot, ascending
};

SlotTable upload_slots(const std::vector<std::vector<int>> &lists) {
    std::vector<int> ptr(1, 0), rel;
    for (auto &l : lists) {
        rel.insert(rel.end(), l.begin(), l.end());
        ptr.push_back((int)rel.size());
    }
    SlotTable t;
    t.n_slots = (int)lists.size();
    t.ptr = dev_upload(ptr);
    t.rel = dev_upload(rel);
    t.lists = lists;
    return t;
}

// longest-processing-time assignment of relations to persistent CTAs, balanced by weight;
// inside a slot relations stay in ascending order (fixed summation order)
SlotTable build_slots(const std::vector<long long> &weight, int n_slots) {
    const int K = (int)weight.size();
    n_slots = std::max(1, std::min(n_slots, K));
    std::vector<int> order(K);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return weight[x] > weight[y]; });
    typedef std::pair<long long, int> Load;
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (int s = 0; s < n_slots; ++s) heap.push(Load(0, s));
    std::vector<std::vector<int>> lists(n_slots);
    for (int k : order) {
        Load l = heap.top();
        heap.pop();
        lists[l.second].push_back(k);
        heap.push(Load(l.first + weight[k], l.second));
    }
    for (auto &l : lists) std::sort(l.begin(), l.end());
    return upload_slots(lists);
}

// every list of `coarse` cut into `parts` contiguous pieces of about equal weight (the pieces of slot s
// are slots s * parts .. s * parts + parts - 1): a warp stream laid out for `coarse` is contiguous for
// the pieces too
SlotTable split_slots(const SlotTable &coarse, int parts, const std::vector<long long> &weight) {
    std::vector<std::vector<int>> lists;
    for (auto &l : coarse.lists) {
        long long total = 0;
        for (int k : l) total += weight[k];
        size_t i = 0;
        long long done = 0;
        for (int part = 0; part < parts; ++part) {
            std::vector<int> piece;
            const long long target = total * (part + 1) / parts;
            while (i < l.size() && (part == parts - 1 || done + weight[l[i]] / 2 < target)) {
                done += weight[l[i]];
                piece.push_back(l[i++]);
            }
            lists.push_back(piece);
        }
    }
    return upload_slots(lists);
}

// Warp-task streams of the staged v3 kernels (TaskArgs in dgn_internal.cuh)
struct TaskCsr {
    int n_warps = 0, rpq = 0, orow_stride = 0;
    int *hdr = nullptr, *orow = nullptr;
    int4 *ent = nullptr;
    // host side, between plan and layout
    std::vector<int> h_hdr, h_orow;
    std::vector<long long> rel_steps;     // pair-steps of the longest warp stream per relation (balance weight)
    std::vector<long long> start;         // [K][n_warps] first pair-step of (k, w) in the laid-out streams
};
void free_task(TaskCsr &c) {
    dev_free(c.hdr);
    dev_free(c.orow);
    dev_free(c.ent);
    c = TaskCsr();
}

// Step 1: row order, slot counts and the balance weights.  rels: K matrices with n_rows rows each.
// Rows are sorted by length (summed over the relations, or per relation when per_rel); consecutive
// groups of 4 go to the n_warps warps in snake order; slot s of warp w is its s-th group.  round_rpq:
// slots per warp rounded up to an even count (the forward kernel is compiled for 2, 4, 6, 8).
// by_address: rows keep their address order instead (4 consecutive rows per warp and slot: contiguous 512-byte
// pieces of the output / optimizer state per access, but the lock-step padding of unsorted rows).
TaskCsr plan_task_csr(const std::vector<HostCsr> &rels, int n_rows, int n_warps, bool per_rel, bool round_rpq,
                      bool by_address = false) {
    const int K = (int)rels.size();
    const int n_groups = (n_rows + 3) / 4;
    int rpq = std::max(1, (n_groups + n_warps - 1) / n_warps);
    if (round_rpq) rpq = (rpq + 1) / 2 * 2;
    DGN_REQUIRE(rpq <= 8, "staged spmm: %d rows per quarter-warp (at most 8)", rpq);
    const int n_ws = n_warps * rpq;  // (warp, slot) pairs
    TaskCsr out;
    out.n_warps = n_warps, out.rpq = rpq, out.orow_stride = per_rel ? n_ws * 4 : 0;
    out.rel_steps.assign(K, 0);
    out.h_hdr.assign((size_t)K * n_warps * 4, 0);
    out.h_orow.assign((size_t)(per_rel ? K : 1) * n_ws * 4, -1);
    std::vector<long long> weight((size_t)n_rows, 0);
    std::vector<int> order((size_t)n_rows), map((size_t)n_ws * 4);
    auto make_map = [&]() {  // map[(w * rpq + s) * 4 + quarter] = row
        std::iota(order.begin(), order.end(), 0);
        if (!by_address) std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return weight[x] > weight[y]; });
        std::fill(map.begin(), map.end(), -1);
        for (int j = 0; j < n_groups; ++j) {
            const int s = j / n_warps, jj = j % n_warps;
            const int w = (s & 1) ? n_warps - 1 - jj : jj;
            for (int i = 0; i < 4 && 4 * j + i < n_rows; ++i) map[(size_t)(w * rpq + s) * 4 + i] = order[(size_t)4 * j + i];
        }
    };
    if (!per_rel) {
        for (auto &c : rels)
            for (int u = 0; u < n_rows; ++u) weight[u] += c.rowptr[u + 1] - c.rowptr[u];
        make_map();
        std::copy(map.begin(), map.end(), out.h_orow.begin());
    }
    for (int k = 0; k < K; ++k) {
        const HostCsr &c = rels[k];
        if (per_rel) {
            for (int u = 0; u < n_rows; ++u) weight[u] = c.rowptr[u + 1] - c.rowptr[u];
            make_map();
            std::copy(map.begin(), map.end(), out.h_orow.begin() + (size_t)k * n_ws * 4);
        }
        const int *m = per_rel ? map.data() : out.h_orow.data();
        for (int w = 0; w < n_warps; ++w) {
            int *h = out.h_hdr.data() + ((size_t)k * n_warps + w) * 4;
            long long steps = 0;
            for (int s = 0; s < rpq; ++s) {
                int cnt = 0;
                for (int q = 0; q < 4; ++q) {
                    const int u = m[(size_t)(w * rpq + s) * 4 + q];
                    if (u >= 0) cnt = std::max(cnt, c.rowptr[u + 1] - c.rowptr[u]);
                }
                const int n2 = (cnt + 1) / 2;
                DGN_REQUIRE(n2 <= 0xffff, "staged spmm: row with %d non-zeros", cnt);
                h[s >> 1] |= n2 << ((s & 1) * 16);
                steps += n2;
            }
            out.rel_steps[k] = std::max(out.rel_steps[k], steps);
        }
    }
    return out;
}

// Step 2: the streams, contiguous per (slot of `slots`, warp) over the slot's relations, and the upload.
void layout_task_csr(TaskCsr &t, const std::vector<HostCsr> &rels, const SlotTable &slots) {
    const int K = (int)rels.size(), n_warps = t.n_warps, rpq = t.rpq;
    const int n_ws = n_warps * rpq;
    long long total = 0;
    for (int k = 0; k < K; ++k)
        for (int w = 0; w < n_warps; ++w)
            for (int s = 0; s < rpq; ++s) total += (t.h_hdr[((size_t)k * n_warps + w) * 4 + (s >> 1)] >> ((s & 1) * 16)) & 0xffff;
    total += (long long)slots.n_slots * n_warps * 8 + 64;  // alignment of the stream starts + prefetch overrun
    DGN_REQUIRE(total < (long long)INT32_MAX / 4, "staged spmm: stream offsets overflow int32");
    std::vector<int4> ent((size_t)total * 4, make_int4(0, 0, 0, 0));
    t.start.assign((size_t)K * n_warps, 0);
    long long cursor = 0;  // pair-steps
    for (auto &list : slots.lists)
        for (int w = 0; w < n_warps; ++w) {
            cursor = (cursor + 7) / 8 * 8;
            for (int k : list) {
                const HostCsr &c = rels[k];
                const int *m = t.h_orow.data() + (t.orow_stride ? (size_t)k * n_ws * 4 : 0);
                t.start[(size_t)k * n_warps + w] = cursor;
                for (int s = 0; s < rpq; ++s) {
                    const int n2 = (t.h_hdr[((size_t)k * n_warps + w) * 4 + (s >> 1)] >> ((s & 1) * 16)) & 0xffff;
                    for (int q = 0; q < 4; ++q) {
                        const int u = m[(size_t)(w * rpq + s) * 4 + q];
                        if (u < 0) continue;
                        const int b = c.rowptr[u], n = c.rowptr[u + 1] - b;
                        for (int i = 0; i < n; ++i) {
                            int4 &x = ent[(size_t)(cursor + (i >> 1)) * 4 + q];
                            int bits;
                            memcpy(&bits, &c.val[b + i], sizeof(float));
                            if (i & 1) x.z = c.col[b + i] << 7, x.w = bits;
                            else x.x = c.col[b + i] << 7, x.y = bits;
                        }
                    }
                    cursor += n2;
                }
            }
        }
    t.hdr = dev_upload(t.h_hdr);
    t.orow = dev_upload(t.h_orow);
    t.ent = dev_upload(ent);
    t.h_hdr.clear(), t.h_hdr.shrink_to_fit();
    t.h_orow.clear(), t.h_orow.shrink_to_fit();
}

// first pair-step of every (slot, warp) stream for a slot table whose lists are contiguous pieces of the
// lists the streams were laid out for
int *upload_wstart(const TaskCsr &t, const SlotTable &slots) {
    std::vector<int> w((size_t)slots.n_slots * t.n_warps, 0);
    for (int s = 0; s < slots.n_slots; ++s)
        if (!slots.lists[s].empty())
            for (int wi = 0; wi < t.n_warps; ++wi) w[(size_t)s * t.n_warps + wi] = (int)t.start[(size_t)slots.lists[s][0] * t.n_warps + wi];
    return dev_upload(w);
}

struct NodeType {
    int n = 0, F = 0;
    bool feat_set = false, identity = false;
    HostCsr feat;                // general sparse features (canonical CSR), empty for the identity
    DevCsr X, Xt;                // device: by node row / by feature
    int *Xt_eid = nullptr;       // entry of Xt -> its position in X
    std::vector<int> row_groups, col_groups;
    float *H = nullptr, *Z = nullptr, *dZ = nullptr, *dA = nullptr;
    long long *dZq = nullptr;  // dZ accumulated by the decode kernel in fixed point
    int lane = 0;  // stream lane of the per-type kernels (epilogues, relu backward)
};

// last writer of a tensor: consumers on the other lane wait for the event
struct Dep {
    cudaEvent_t ev = nullptr;
    int lane = 0;
};

struct Group {
    int i = 0, j = 0, K = 0, decoder = 0, r0 = 0;
    // multi-GPU: the relations of a group of many small relations are partitioned over the ranks; Kl of
    // the K relations live here (loc: their group-wide indices, ascending; loc_index: inverse or -1).
    // Encoder state (adjacency, W1, W2, masks) is sized by Kl; decoder variables are replicated (K).
    bool partitioned = false;
    int Kl = 0;
    std::vector<int> loc, loc_index, owner;
    int *rel_ids = nullptr;  // device: flat relation id of every local relation (dropout streams are keyed by it)
    struct Exchange {        // partial sums that every rank needs from every rank: S1, S2, dH
        int id = -1;
        size_t floats = 0, off = 0;  // two parity buffers of `floats` at byte offset off of the comm arena
        uint32_t stamp = 0;
    } xch[3];
    int n_i = 0, n_j = 0, F_j = 0;
    std::vector<HostCsr> rel;
    std::vector<bool> rel_set;
    long long nnz = 0;
    // device sparse structures
    DevCsr relcsr, fwd, bwd;
    SegTable fwd_seg, bwd_seg;
    bool staged = false;
    bool dirty = true;  // a relation / the features / the partition changed since the device structures were built
    int lane = 0;
    SlotTable slots1, slots2;
    int staged_version = 3;      // 2: spmm_staged_kernel, 3: spmm_staged3_kernel (position order, mbarrier pipeline)
    bool tstaged = false;        // backward products through spmm_tstaged_kernel
    bool w1_grad_stale = false;  // the last step fused Adam into the gradient kernel: grads of W1 were never written
    TaskCsr task_fwd, task_bwd;
    SlotTable slots_bwd;
    int *wstart1 = nullptr, *wstart2 = nullptr, *wstart_bwd = nullptr;
    // second copy of the backward streams with the rows in ADDRESS order, used by the layer-1 backward (its output rows
    // are the 128-byte rows of dW1 / of the optimizer state: contiguous per warp instead of scattered); unused
    // (n_warps == 0) unless DGN_BWD_ROW_ORDER is "split" (the default)
    TaskCsr task_bwd1;
    SlotTable slots_bwd1;
    int *wstart_bwd1 = nullptr;
    // parameter arena offsets (floats)
    size_t w1_off = 0, w2_off = 0, glb_off = 0, loc_off = 0, loc_per_rel = 0;
    // work buffers
    float *part1 = nullptr, *part2 = nullptr, *Y1 = nullptr, *n1 = nullptr, *Y2 = nullptr, *n2 = nullptr;
    float *P2 = nullptr, *dS = nullptr, *G2 = nullptr, *bwd_partial = nullptr, *dW2part = nullptr, *dHpart = nullptr;
    float *rows1 = nullptr, *rows2 = nullptr;  // partitioned gather-path group: this rank's row sums [P][n_i][32] (what it publishes)
    uint32_t *mask1 = nullptr, *mask2 = nullptr;  // mask2: the buffer of mask2buf this step reads
    uint32_t *mask2buf[2] = {nullptr, nullptr};
    long long mask1_words = 0, mask2_words = 0;
    bool dense_tc = false;  // layer-2 contractions on tcgen05 (dense_tc.cu)
    bool gen_feat = false;  // column type has general sparse features: layer 1 runs on P1 = X W1_k / G1 = A_k^T dS1
    float *P1buf = nullptr, *G1buf = nullptr;  // [P1][K * n_j][32]
    long long feat_nnz = 0;
    int n_rb = 1, n_rb_pd = 1, slots_proj = 1, slots_dh = 1, slots_dw2 = 1;  // dense layer-2 kernels: row blocks, CTAs per block
    std::vector<uint32_t *> thr;  // per relation
    std::vector<int> thr_n;
};

struct Phase {
    std::string name;
    cudaEvent_t start, stop;
    int lane = 0;
};

}  // namespace

struct dgn_graph {
    int device = 0, n_sm = 148;
    int n_types = 0, n_groups = 0, R = 0, d1 = 0, d2 = 0, P1 = 0;
    // hidden1 / hidden2 as the caller passed them (model.py:68,80 take any FLAGS value); the device works on
    // d1 in {32, 64, 128} and d2 = 32 columns: smaller sizes are zero-padded at this boundary.  Zero columns of
    // W1 / W2, zero rows of W2 and zero rows / columns of the decoder variables give zero hidden / embedding columns
    // (relu(0) = 0, zeros add nothing to a norm), zero gradients for the padding and (m = v = 0) no Adam movement, so
    // the padded model IS the caller's model, term by term.
    int d1u = 0, d2u = 0;
    std::vector<NodeType> types;
    std::vector<Group> groups;
    std::vector<std::pair<int, int>> flat;  // r -> (group, k)
    bool finalized = false;
    bool allow_staged = true, allow_tstaged = true;
    int staged_version = 3;
    cudaStream_t stream = nullptr;   // lane 0: groups of many small relations (staged kernels), decode, Adam
    // lanes 1 .. kMaxLanes - 1: the other groups, ONE LANE PER GROUP (their short, latency-bound kernels overlap each
    // other); ordered against lane 0 and against each other by events per tensor
    static const int kMaxLanes = 5;
    cudaStream_t side[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // side[0] unused
    int n_lanes = 2;
    cudaStream_t stream3 = nullptr;  // layer-2 keep words of every group (integer ALU work beside lane 1's gathers)
    cudaEvent_t mask_go = nullptr, mask_done = nullptr, ahead_go = nullptr, ahead_done = nullptr;
    // Layer-2 keep words drawn AHEAD: the masks of step t + 1 (same seed and rate assumed) are generated into the other
    // buffer while step t's backward finishes on the side lanes; a step whose (seed, step, rate) match finds them ready
    bool mask_ahead = false;  // DGN_MASK_AHEAD=1 enables (measured: 1.812 vs 1.721 ms per step -- the generation lengthens the
                              // backward's tail by more than it shortens the next step's start)
    int mask_cur = 0;
    bool ahead_valid = false;
    uint64_t ahead_seed = 0;
    uint32_t ahead_step = 0, ahead_thr = 0;
    bool own_stream = false;
    int rank = 0, world = 1;
    bool arena_ready = false;
    // multi-GPU exchange arena: [flags: kMaxWorld x kMaxExchanges uint32][buffers]; peers map it through CUDA IPC
    unsigned char *comm = nullptr;
    size_t comm_bytes = 0;
    unsigned char *peer_comm[kMaxWorld] = {};
    uint32_t **peer_flags_dev = nullptr;
    bool connected = false;
    int n_exchanges = 0;
    unsigned long long exchange_timeout_ns = 600ull * 1000000000ull;  // DGN_EXCHANGE_TIMEOUT_S, wall clock
    int *exchange_error = nullptr;   // device: 1 + rank of the peer whose stamp never arrived
    int *exchange_error_host = nullptr;
    bool two_lanes = true;
    bool fuse_adam = true;  // Adam of the layer-1 weights inside the kernel that produces their gradient
    bool gate_lane0 = true;  // the staged kernels of lane 0 wait for the side groups' kernels of the same layer (see run_forward)
    bool gather_row_sums = false;  // gather path: segments reduced to row sums BEFORE the epilogue instead of inside it
                                   // (DGN_GATHER_ROWSUMS=1; measured 1.764 vs 1.720 ms per step on one GPU: two more launches per group)
    bool keep_grads = false;  // dgn_keep_gradients: every gradient is materialised (no fused Adam)
    std::vector<cudaEvent_t> dep_events;  // pool, reused every step
    size_t dep_next = 0;
    // parameters
    size_t n_params = 0, dec_off = 0;
    float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
    float beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, b1p = 0.9f, b2p = 0.999f;
    // per-step block [StepDyn | negatives | batch]: pinned host ring + ONE device mirror, one H2D copy per step
    static const int kRing = 4;
    unsigned char *step_host[kRing] = {nullptr, nullptr, nullptr, nullptr};
    unsigned char *step_dev = nullptr;
    size_t step_neg_off = 0, step_batch_off = 0;
    cudaEvent_t ring_ev[kRing];
    int ring_cap = 0, ring_pos = 0;
    StepDyn *dyn_dev = nullptr;
    int *batch_dev = nullptr;
    long long *neg_dev = nullptr, *neg_out = nullptr;
    // CUDA graphs of the training step, one per (dropout rate, update / gradients mode, exchange parities)
    struct StepGraph {
        int seen = 0;
        cudaGraphExec_t exec = nullptr;
        long long launches = 0;
    };
    std::map<std::tuple<uint32_t, int, int, int>, StepGraph> step_graphs;
    bool use_graphs = true;   // DGN_CUDA_GRAPH=0: issue every kernel from the host each step
    bool capturing = false;
    long long graph_replays = 0;
    long long groups_rebuilt = 0;  // build_group calls so far (tests: incremental finalize)
    float *pos_out = nullptr, *negs_out = nullptr, *loss_dev = nullptr, *loss_host = nullptr;
    uint32_t step_seq = 0;           // serial number of the last issued training step
    bool early_loss = true;          // DGN_SYNC_LOSS=1: dgn_train_step waits for the whole step before it returns the loss
    float *decode_scratch = nullptr;
    unsigned *decode_ticket = nullptr;
    int last_B = 0;
    bool dzq_clean = false;  // the fixed-point dZ accumulators are zero (cleared by the conversion kernel)
    // measurement
    bool timing = false;
    cudaEvent_t timer_start = nullptr, timer_stop = nullptr;
    std::vector<Phase> phases;
    long long launches = 0;
};

namespace {

struct PhaseScope {
    dgn_graph *g;
    cudaEvent_t stop = nullptr;
    cudaStream_t st;
    PhaseScope(dgn_graph *g_, const char *name, int group = -1, int lane = 0) : g(g_) {
        st = lane >= 1 && g->two_lanes && g->side[lane] ? g->side[lane] : g->stream;
        if (!g->timing) return;
        Phase p;
        p.name = name;
        if (group >= 0) p.name += "/g" + std::to_string(group);
        p.lane = lane;
        CUDA_CHECK(cudaEventCreate(&p.start));
        CUDA_CHECK(cudaEventCreate(&p.stop));
        CUDA_CHECK(cudaEventRecord(p.start, st));
        stop = p.stop;
        g->phases.push_back(p);
    }
    ~PhaseScope() {
        if (stop) cudaEventRecord(stop, st);
    }
};

cudaStream_t lane_stream(dgn_graph *g, int lane) { return lane >= 1 && g->two_lanes ? g->side[lane] : g->stream; }
// the tensor was just written on `lane`
void produced(dgn_graph *g, Dep &d, int lane) {
    if (!g->two_lanes) return;
    if (g->dep_next == g->dep_events.size()) {
        cudaEvent_t e;
        CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        g->dep_events.push_back(e);
    }
    d.ev = g->dep_events[g->dep_next++];
    d.lane = lane;
    CUDA_CHECK(cudaEventRecord(d.ev, lane_stream(g, lane)));
}
// the next kernel on `lane` reads the tensor
void consume(dgn_graph *g, const Dep &d, int lane) {
    if (!g->two_lanes || d.ev == nullptr || d.lane == lane) return;
    CUDA_CHECK(cudaStreamWaitEvent(lane_stream(g, lane), d.ev, 0));
}
// every side lane continues after what is queued on lane 0
void fork_lanes(dgn_graph *g) {
    if (!g->two_lanes) return;
    Dep d;
    produced(g, d, 0);
    for (int lane = 1; lane < g->n_lanes; ++lane) consume(g, d, lane);
}
// lane 0 continues after everything queued on the side lanes (and vice versa when both)
void join_lanes(dgn_graph *g, bool both) {
    if (!g->two_lanes) return;
    for (int lane = 1; lane < g->n_lanes; ++lane) {
        Dep d;
        produced(g, d, lane);
        consume(g, d, 0);
    }
    if (both) fork_lanes(g);
}

size_t panel_floats(int P, long long rows) { return (size_t)P * (size_t)rows * 32; }

// after a stream synchronisation: did an exchange give up on a peer?
void check_exchange(dgn_graph *g) {
    if (g->world == 1 || !g->exchange_error) return;
    CUDA_CHECK(cudaMemcpy(g->exchange_error_host, g->exchange_error, sizeof(int), cudaMemcpyDeviceToHost));
    if (*g->exchange_error_host != 0) {
        const int peer = *g->exchange_error_host - 1;
        CUDA_CHECK(cudaMemset(g->exchange_error, 0, sizeof(int)));
        DGN_FAIL(DGN_ERR_CUDA, "multi-GPU exchange timed out waiting for rank %d (every rank must call dgn_encoder_forward / "
                 "dgn_train_step in lock step; DGN_EXCHANGE_TIMEOUT_S sets the limit)", peer);
    }
}

void check_finalized(dgn_graph *g) { DGN_REQUIRE(g && g->finalized, "graph is not finalized (call dgn_graph_finalize)"); }

void free_group_device(Group &G) {
    free_csr(G.relcsr);
    free_csr(G.fwd);
    free_csr(G.bwd);
    free_seg(G.fwd_seg);
    free_seg(G.bwd_seg);
    dev_free(G.slots1.ptr);
    dev_free(G.slots1.rel);
    dev_free(G.slots2.ptr);
    dev_free(G.slots2.rel);
    dev_free(G.slots_bwd.ptr);
    dev_free(G.slots_bwd.rel);
    free_task(G.task_fwd);
    free_task(G.task_bwd);
    dev_free(G.slots_bwd1.ptr);
    dev_free(G.slots_bwd1.rel);
    free_task(G.task_bwd1);
    dev_free(G.wstart_bwd1);
    dev_free(G.wstart1);
    dev_free(G.wstart2);
    dev_free(G.wstart_bwd);
    G.slots1 = SlotTable(), G.slots2 = SlotTable(), G.slots_bwd = SlotTable(), G.slots_bwd1 = SlotTable();
    dev_free(G.rel_ids);
    float **bufs[] = {&G.rows1, &G.rows2, &G.part1, &G.part2, &G.Y1, &G.n1, &G.Y2, &G.n2, &G.P2, &G.dS, &G.G2, &G.bwd_partial, &G.dW2part, &G.dHpart, &G.P1buf, &G.G1buf};
    for (float **b : bufs) dev_free(*b);
    dev_free(G.mask1);
    dev_free(G.mask2buf[0]);
    dev_free(G.mask2buf[1]);
    G.mask2 = nullptr;
}

void build_group(dgn_graph *g, Group &G) {
    free_group_device(G);
    const int K = G.Kl, n_i = G.n_i, n_j = G.n_j, P1 = g->P1;  // K: the LOCAL relations from here on
    std::vector<HostCsr> local_copy;
    if (G.partitioned) {
        local_copy.reserve(K);
        for (int k : G.loc) local_copy.push_back(G.rel[k]);
    }
    const std::vector<HostCsr> &rel = G.partitioned ? local_copy : G.rel;
    {
        std::vector<int> ids((size_t)K + 33, 0);  // + 33: the last word of the packed layer-1 mask may walk past up to 32 relation ends
        for (int l = 0; l < K; ++l) ids[l] = G.r0 + G.loc[l];
        G.rel_ids = dev_upload(ids);
    }
    G.nnz = 0;
    for (auto &c : rel) G.nnz += c.nnz();
    DGN_REQUIRE(G.nnz < (long long)INT32_MAX, "group (%d,%d): %lld non-zeros do not fit int32 offsets", G.i, G.j, G.nnz);
    DGN_REQUIRE((long long)K * std::max(n_j, G.F_j) < (long long)INT32_MAX / 64, "group (%d,%d): K * n_j too large", G.i, G.j);

    // per-relation CSR, concatenated (row pointers hold offsets into the group arrays)
    HostCsr cat_rel;
    cat_rel.rowptr.reserve((size_t)K * (n_i + 1));
    cat_rel.col.reserve((size_t)G.nnz);
    cat_rel.val.reserve((size_t)G.nnz);
    for (int k = 0; k < K; ++k) {
        const int base = (int)cat_rel.col.size();
        for (int u = 0; u <= n_i; ++u) cat_rel.rowptr.push_back(base + rel[k].rowptr[u]);
        cat_rel.col.insert(cat_rel.col.end(), rel[k].col.begin(), rel[k].col.end());
        cat_rel.val.insert(cat_rel.val.end(), rel[k].val.begin(), rel[k].val.end());
    }
    cat_rel.n_rows = K * (n_i + 1) - 1;
    G.relcsr = upload_csr(cat_rel);

    // forward: [A_0 | A_1 | ...], n_i rows, column = k * n_j + c, entries ordered by (k, c)
    HostCsr fwd;
    fwd.n_rows = n_i;
    fwd.n_cols = K * n_j;
    fwd.rowptr.assign((size_t)n_i + 1, 0);
    for (int k = 0; k < K; ++k)
        for (int u = 0; u < n_i; ++u) fwd.rowptr[(size_t)u + 1] += rel[k].rowptr[u + 1] - rel[k].rowptr[u];
    for (int u = 0; u < n_i; ++u) fwd.rowptr[u + 1] += fwd.rowptr[u];
    fwd.col.resize((size_t)G.nnz);
    fwd.val.resize((size_t)G.nnz);
    {
        std::vector<int> cursor(fwd.rowptr.begin(), fwd.rowptr.end() - 1);
        for (int k = 0; k < K; ++k)
            for (int u = 0; u < n_i; ++u)
                for (int e = rel[k].rowptr[u]; e < rel[k].rowptr[u + 1]; ++e) {
                    const int dst = cursor[u]++;
                    fwd.col[dst] = k * n_j + rel[k].col[e];
                    fwd.val[dst] = rel[k].val[e];
                }
    }
    // backward: the transpose, K * n_j rows, column = u
    HostCsr bwd;
    csr_transpose(fwd, bwd);

    G.gen_feat = !g->types[G.j].identity;
    G.feat_nnz = G.gen_feat ? g->types[G.j].feat.nnz() : G.F_j;
    G.staged = g->allow_staged && staged_supported(n_i, n_j, G.K);  // by the group-wide K: every rank agrees
    const long long target_warps = (long long)g->n_sm * 64 * 2;
    // one quarter-warp per segment: aim at ~2 waves of quarter-warps, 32..2048 non-zeros each
    int seg_len = (int)std::min<long long>(2048, std::max<long long>(32, (G.nnz / (4 * target_warps) + 31) / 32 * 32));
    G.fwd = upload_csr(fwd);
    G.fwd_seg = build_segments(fwd, seg_len, false);
    G.bwd = upload_csr(bwd);
    G.bwd_seg = build_segments(bwd, std::min(256, std::max(32, seg_len)), true);
    if (!G.bwd_seg.trivial) G.bwd_partial = dev_alloc<float>(panel_floats(P1, G.bwd_seg.n_seg));

    G.staged_version = g->staged_version;
    if (G.staged && G.staged_version == 3 && !staged3_supported(n_i, n_j, G.K)) G.staged_version = 2;
    G.tstaged = G.staged && g->allow_tstaged && tstaged_supported(n_i, n_j, G.K, P1);
    std::vector<long long> w_fwd(K), w_bwd(K);
    for (int k = 0; k < K; ++k) w_fwd[k] = w_bwd[k] = rel[k].nnz() + rel[k].n_cols;
    if (G.staged && G.staged_version == 3) {
        G.task_fwd = plan_task_csr(rel, n_i, kS3Warps, false, true);
        for (int k = 0; k < K; ++k) w_fwd[k] = G.task_fwd.rel_steps[k] + 64;
        // layer 1 runs P1 panel CTAs per slot, layer 2 one: the layer-2 slots are pieces of the layer-1 slots
        G.slots1 = build_slots(w_fwd, std::max(1, g->n_sm / P1));
        G.slots2 = split_slots(G.slots1, P1, w_fwd);
        layout_task_csr(G.task_fwd, rel, G.slots1);
        G.wstart1 = upload_wstart(G.task_fwd, G.slots1);
        G.wstart2 = upload_wstart(G.task_fwd, G.slots2);
    } else if (G.staged) {
        G.slots1 = build_slots(w_fwd, std::max(1, g->n_sm / P1));
        G.slots2 = build_slots(w_fwd, g->n_sm);
    }
    if (G.tstaged) {
        std::vector<HostCsr> relt(K);
        for (int k = 0; k < K; ++k) csr_transpose(rel[k], relt[k]);
        // DGN_BWD_ROW_ORDER: "length" = rows of a relation sorted by length for both backward products (4 % lock-step
        // padding), "address" = address order for both (24 %), default "split" = layer 2 by length (gather-bound: the
        // padding costs 125 -> 141 us) and layer 1 by address (bound by the dW1 / optimizer-state traffic: contiguous
        // rows bring 444 -> 406 us); measured at the polypharmacy shape, DESIGN.md section 3
        const char *row_order = getenv("DGN_BWD_ROW_ORDER");
        const char mode = row_order ? row_order[0] : 's';
        G.task_bwd = plan_task_csr(relt, n_j, kTsWarps, true, false, mode == 'a');
        for (int k = 0; k < K; ++k) w_bwd[k] = G.task_bwd.rel_steps[k] + 16;
        G.slots_bwd = build_slots(w_bwd, g->n_sm);
        layout_task_csr(G.task_bwd, relt, G.slots_bwd);
        G.wstart_bwd = upload_wstart(G.task_bwd, G.slots_bwd);
        if (mode == 's') {
            G.task_bwd1 = plan_task_csr(relt, n_j, kTsWarps, true, false, true);
            for (int k = 0; k < K; ++k) w_bwd[k] = G.task_bwd1.rel_steps[k] + 16;
            G.slots_bwd1 = build_slots(w_bwd, g->n_sm);
            layout_task_csr(G.task_bwd1, relt, G.slots_bwd1);
            G.wstart_bwd1 = upload_wstart(G.task_bwd1, G.slots_bwd1);
        }
    }
    if (G.staged) {
        G.part1 = dev_alloc<float>((size_t)G.slots1.n_slots * panel_floats(P1, n_i));
        G.part2 = dev_alloc<float>((size_t)G.slots2.n_slots * panel_floats(1, n_i));
    } else {
        G.part1 = dev_alloc<float>(panel_floats(P1, G.fwd_seg.n_seg));
        G.part2 = dev_alloc<float>(panel_floats(1, G.fwd_seg.n_seg));
        G.rows1 = dev_alloc<float>(panel_floats(P1, n_i));
        G.rows2 = dev_alloc<float>(panel_floats(1, n_i));
    }
    G.Y1 = dev_alloc<float>(panel_floats(P1, n_i));
    G.n1 = dev_alloc<float>(n_i);
    G.Y2 = dev_alloc<float>(panel_floats(1, n_i));
    G.n2 = dev_alloc<float>(n_i);
    G.dS = dev_alloc<float>(panel_floats(P1, n_i));
    G.P2 = dev_alloc<float>(panel_floats(1, (long long)K * n_j));
    G.G2 = dev_alloc<float>(panel_floats(1, (long long)K * n_j));
    G.mask1_words = ((long long)K * G.feat_nnz + 31) / 32;  // identity features: one bit per row of W1_k
    if (G.gen_feat) {
        G.P1buf = dev_alloc<float>(panel_floats(P1, (long long)K * n_j));
        G.G1buf = dev_alloc<float>(panel_floats(P1, (long long)K * n_j));
    }
    G.mask2_words = (long long)K * n_j * P1;
    G.mask1 = dev_alloc<uint32_t>((size_t)G.mask1_words);
    G.mask2buf[0] = dev_alloc<uint32_t>((size_t)G.mask2_words);
    G.mask2buf[1] = dev_alloc<uint32_t>((size_t)G.mask2_words);
    G.mask2 = G.mask2buf[0];

    // dense layer-2 kernels: persistent CTAs, one per (row block, slot of relations)
    // tensor-core versions (dense_tc.cu) unless DGN_DENSE_FFMA=1 asks for the CUDA-core kernels of dense.cu
    const char *ffma = getenv("DGN_DENSE_FFMA");
    G.dense_tc = dense_tc_supported(g->d1, g->d2) && !(ffma && ffma[0] == '1');
    if (G.dense_tc) {
        const int n_rt = dense_tc_tiles(n_j);  // row tiles of 128 = threads of a CTA, two CTAs per SM
        G.n_rb_pd = n_rt;
        G.slots_proj = G.slots_dh = std::max(1, std::min(K, 2 * g->n_sm / n_rt));
        // dw2: CTA = (relation, chunk of row tiles); partials per chunk when the relations alone cannot fill the GPU
        G.n_rb = K >= 2 * g->n_sm ? 1 : std::max(1, std::min(n_rt, (2 * g->n_sm + K - 1) / std::max(K, 1)));
        G.slots_dw2 = 2 * g->n_sm;  // persistent CTAs
    } else {
        const int RB = dense_row_block(g->d1, 1), RBpd = dense_row_block(g->d1, 0);
        G.n_rb = (n_j + RB - 1) / RB;        // dw2
        G.n_rb_pd = (n_j + RBpd - 1) / RBpd;  // project, dh
        G.slots_dw2 = std::max(1, std::min(K, g->n_sm / G.n_rb));
        G.slots_proj = std::max(1, std::min(K, g->n_sm / G.n_rb_pd));
        G.slots_dh = std::max(1, std::min(K, g->n_sm / (P1 * G.n_rb_pd)));
    }
    if (G.n_rb > 1) G.dW2part = dev_alloc<float>((size_t)K * G.n_rb * g->d1 * g->d2);
    G.dHpart = dev_alloc<float>((size_t)G.slots_dh * panel_floats(P1, n_j));
}

uint32_t dropout_threshold(float rate) {
    const double t = ceil((double)rate * 4294967296.0);
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

const size_t kCommFlagBytes = (size_t)kMaxWorld * kMaxExchanges * sizeof(uint32_t);

void free_comm(dgn_graph *g) {
    for (int r = 0; r < kMaxWorld; ++r) {
        if (g->peer_comm[r] && r != g->rank) cudaIpcCloseMemHandle(g->peer_comm[r]);
        g->peer_comm[r] = nullptr;
    }
    if (g->peer_flags_dev) cudaFree(g->peer_flags_dev);
    g->peer_flags_dev = nullptr;
    if (g->comm) cudaFree(g->comm);
    g->comm = nullptr;
    g->comm_bytes = 0;
    g->connected = false;
}

// exchange buffers of the partitioned groups (called from finalize)
void build_comm(dgn_graph *g) {
    free_comm(g);
    if (g->world == 1) return;
    size_t off = kCommFlagBytes;
    int id = 0;
    for (auto &G : g->groups) {
        if (!G.partitioned) continue;
        const size_t fl[3] = {panel_floats(g->P1, G.n_i), panel_floats(1, G.n_i), panel_floats(g->P1, G.n_j)};
        for (int x = 0; x < 3; ++x) {
            DGN_REQUIRE(id < kMaxExchanges, "too many partitioned groups");
            G.xch[x].id = id++;
            G.xch[x].floats = fl[x];
            G.xch[x].off = off;
            G.xch[x].stamp = 0;
            off += 2 * fl[x] * sizeof(float);
        }
    }
    g->n_exchanges = id;
    g->comm_bytes = off;
    CUDA_CHECK(cudaMalloc(&g->comm, off));
    CUDA_CHECK(cudaMemset(g->comm, 0, off));
    g->peer_comm[g->rank] = g->comm;
}

// Sum the local partials (fixed order) into this rank's buffer of exchange x, publish it and wait for
// every peer; afterwards peer_ptr(r) of every rank may be read.  Collective: every rank calls it in the
// same order.
void exchange(dgn_graph *g, Group &G, int x, const float *partial, int n_chunks, cudaStream_t s) {
    DGN_REQUIRE(g->connected, "dgn_comm_connect was not called");
    Group::Exchange &X = G.xch[x];  // X.stamp: this step's stamp (begin_step)
    float *mine = reinterpret_cast<float *>(g->comm + X.off) + (X.stamp & 1) * X.floats;
    launch_publish(partial, n_chunks, X.floats, mine, s);
    launch_signal_wait(g->peer_flags_dev, reinterpret_cast<uint32_t *>(g->comm), g->rank, g->world, X.id, g->dyn_dev, g->exchange_timeout_ns,
                       g->exchange_error, s);
    g->launches += 2;
}
const float *peer_ptr(dgn_graph *g, const Group &G, int x, int r) {
    const Group::Exchange &X = G.xch[x];
    return reinterpret_cast<const float *>(g->peer_comm[r] + X.off) + (X.stamp & 1) * X.floats;
}

// Issue order of the groups: lane 0 first; on lane 1 the groups that lane 0 waits for first (row type on lane 0 in
// the forward pass and the layer-1 backward, column type on lane 0 in the layer-2 backward).
std::vector<int> lane_order(dgn_graph *g, bool by_col_type) {
    std::vector<int> order;
    for (int gi = 0; gi < g->n_groups; ++gi)
        if (g->groups[gi].lane == 0) order.push_back(gi);
    for (int pass = 0; pass < 2; ++pass)
        for (int gi = 0; gi < g->n_groups; ++gi) {
            const Group &G = g->groups[gi];
            if (G.lane == 0) continue;
            const int t = by_col_type ? G.j : G.i;
            if ((g->types[t].lane == 0) == (pass == 0)) order.push_back(gi);
        }
    return order;
}

// Per-step dependency state: one Dep per tensor that crosses lanes.
struct StepDeps {
    std::vector<Dep> S1, S2, dH;   // per group: layer-1 / layer-2 partial sums, dH partials
    std::vector<Dep> H, Z, dZ, dA; // per node type
};

void run_forward(dgn_graph *g, float rate, StepDeps &D, bool masks2_ready = false) {
    const int P1 = g->P1;
    const bool drop = rate > 0.f;
    const float keep = 1.f - rate;
    const float scale = drop ? 1.f / keep : 1.f;
    D.S1.assign(g->n_groups, Dep()), D.S2.assign(g->n_groups, Dep()), D.dH.assign(g->n_groups, Dep());
    D.H.assign(g->n_types, Dep()), D.Z.assign(g->n_types, Dep()), D.dZ.assign(g->n_types, Dep()), D.dA.assign(g->n_types, Dep());
    g->dep_next = 0;
    fork_lanes(g);  // the side lanes start after everything queued so far (previous step's Adam, the per-step block);
                    // every call ends with them joined into lane 0, so nothing is pending there
    if (drop) {
        if (g->two_lanes && !masks2_ready) {
            CUDA_CHECK(cudaEventRecord(g->mask_go, g->stream));  // after the join: last step's readers of mask2 are done
            CUDA_CHECK(cudaStreamWaitEvent(g->stream3, g->mask_go, 0));
            for (int side = 1; side >= 0; --side)  // the side groups' (small) masks first
                for (auto &G : g->groups) {
                    if ((G.lane != 0) != (side == 1)) continue;
                    launch_gen_mask(G.mask2, G.mask2_words, (long long)G.n_j * g->d1, G.n_j * P1, G.rel_ids, kStreamDropout2, g->dyn_dev,
                                    0, g->stream3);
                    g->launches++;
                }
            CUDA_CHECK(cudaEventRecord(g->mask_done, g->stream3));
        }
        // layer-1 keep bits of every group: one launch per batch of groups on lane 0 (lane 1 continues behind the join
        // below); layer-2 keep words: see further down
        {
            PhaseScope ph(g, "mask", -1, 0);
            MaskBatch mb = {};
            auto flush = [&]() {
                launch_gen_mask_multi(mb, kStreamDropout1, g->dyn_dev, g->stream);
                if (mb.n) g->launches++;
                mb.n = 0;
            };
            for (auto &G : g->groups) {
                mb.words[mb.n] = G.mask1, mb.n_words[mb.n] = G.mask1_words, mb.bits_per_rel[mb.n] = G.feat_nnz, mb.rel_ids[mb.n] = G.rel_ids;
                if (++mb.n == kMaxMaskBatch) flush();
            }
            flush();
        }
        if (g->two_lanes) {
            fork_lanes(g);  // the layer-1 keep bits of every group were drawn on lane 0
        } else {
            for (auto &G : g->groups) {
                launch_gen_mask(G.mask2, G.mask2_words, (long long)G.n_j * g->d1, G.n_j * P1, G.rel_ids, kStreamDropout2, g->dyn_dev,
                                0, g->stream);
                g->launches++;
            }
        }
    }
    bool mask_waited[dgn_graph::kMaxLanes];
    for (bool &w : mask_waited) w = !(drop && g->two_lanes) || masks2_ready;
    auto wait_mask2 = [&](int lane) {
        if (mask_waited[lane]) return;
        CUDA_CHECK(cudaStreamWaitEvent(lane_stream(g, lane), g->mask_done, 0));
        mask_waited[lane] = true;
    };
    auto spmm_fwd = [&](Group &G, const float *op, int P, long long op_rows, float *part, const SlotTable &slots,
                        const int *wstart, const uint32_t *mask) {
        cudaStream_t s = lane_stream(g, G.lane);
        if (G.staged && G.staged_version == 3) {
            TaskArgs a = {};
            a.hdr = G.task_fwd.hdr, a.ent = G.task_fwd.ent, a.orow = G.task_fwd.orow, a.orow_stride = 0;
            a.wstart = wstart;
            a.K = G.Kl, a.n_warps = G.task_fwd.n_warps, a.rpq = G.task_fwd.rpq;
            a.n_out_rows = G.n_i, a.n_op_rows = G.n_j;
            a.op = op, a.P = P;
            a.slot_ptr = slots.ptr, a.slot_rel = slots.rel, a.n_slots = slots.n_slots;
            a.out = part, a.mask = mask, a.scale = scale;
            launch_spmm_staged3(a, s);
        } else if (G.staged) {
            StagedArgs a = {};
            a.rowptr = G.relcsr.rowptr, a.col = G.relcsr.col, a.val = G.relcsr.val;
            a.K = G.Kl, a.n_i = G.n_i, a.n_j = G.n_j;
            a.op = op, a.P = P;
            a.slot_ptr = slots.ptr, a.slot_rel = slots.rel, a.n_slots = slots.n_slots;
            a.partial = part, a.mask = mask, a.scale = scale;
            launch_spmm_staged(a, s);
        } else {
            SpmmArgs a = {};
            a.rowptr = G.fwd.rowptr, a.col = G.fwd.col, a.val = G.fwd.val;
            a.seg_row = G.fwd_seg.seg_row, a.seg_begin = G.fwd_seg.seg_begin, a.row_seg_ptr = G.fwd_seg.row_seg_ptr;
            a.seg_len = G.fwd_seg.seg_len, a.n_seg = G.fwd_seg.n_seg, a.n_rows = G.n_i;
            a.op = op, a.op_rows = (int)op_rows;
            a.partial = part, a.force_partial = 1;
            a.mask = mask, a.col_mask = mask != nullptr, a.scale = scale;
            if (g->gather_row_sums || G.partitioned) {
                // row sums [P][n_i][32] (what a partitioned group publishes, and what the epilogue reads with one load
                // per row instead of a walk over the row's segments): rows of one segment are written directly, hub
                // rows are reduced in order by a warp each, rows without a non-zero stay zero
                float *rows = part == G.part1 ? G.rows1 : G.rows2;
                CUDA_CHECK(cudaMemsetAsync(rows, 0, panel_floats(P, G.n_i) * sizeof(float), s));
                a.out = rows, a.out_rows = G.n_i, a.force_partial = 0;
                launch_spmm(a, P, s);
                if (G.fwd_seg.n_multi > 0) {
                    a.mask = nullptr, a.col_mask = 0;
                    launch_seg_reduce(a, G.fwd_seg.multi_rows, G.fwd_seg.n_multi, P, s);
                    g->launches++;
                }
            } else {
                launch_spmm(a, P, s);
            }
        }
        g->launches++;
    };
    auto epilogue = [&](int t, int layer) {
        NodeType &T = g->types[t];
        EpiArgs e = {};
        e.n_rows = T.n;
        e.relu = layer == 1;
        e.out = layer == 1 ? T.H : T.Z;
        DGN_REQUIRE((int)T.row_groups.size() <= kMaxGroupsPerType, "more than %d groups share row type %d", kMaxGroupsPerType, t);
        for (int gi : T.row_groups) {
            Group &G = g->groups[gi];
            consume(g, layer == 1 ? D.S1[gi] : D.S2[gi], T.lane);
            EpiGroup &eg = e.g[e.n_groups++];
            eg.partial = layer == 1 ? G.part1 : G.part2;
            eg.row_seg_ptr = G.staged ? nullptr : G.fwd_seg.row_seg_ptr;
            eg.n_slots = layer == 1 ? G.slots1.n_slots : G.slots2.n_slots;
            if (!G.staged && g->gather_row_sums) {
                eg.partial = layer == 1 ? G.rows1 : G.rows2;
                eg.row_seg_ptr = nullptr;
                eg.n_slots = 1;
            }
            if (G.partitioned) {
                eg.n_peers = g->world;
                for (int r = 0; r < g->world; ++r) eg.peer[r] = peer_ptr(g, G, layer - 1, r);
            }
            eg.Y = layer == 1 ? G.Y1 : G.Y2;
            eg.nrm = layer == 1 ? G.n1 : G.n2;
        }
        PhaseScope ph(g, "epilogue", -1, T.lane);
        launch_node_epilogue(e, layer == 1 ? P1 : 1, lane_stream(g, T.lane));
        g->launches++;
        produced(g, layer == 1 ? D.H[t] : D.Z[t], T.lane);
    };
    // lane 0 groups first: their kernels are the long ones.  On lane 1 the groups whose ROW type is summed on lane 0
    // come first: lane 0 waits for exactly those partial sums before it can go on
    // The persistent staged kernels of lane 0 fill every SM (1024 threads x 64 registers): whatever lane 1 has queued
    // starves until they finish.  So lane 1's kernels of a layer are issued FIRST, where they overlap lane 0's mask
    // generation (integer ALU against L2-bound gathers) and its tensor-core projection (which leaves room), and the
    // staged kernel of the layer waits for them ("gate"): lane 1 then holds its results by the time lane 0 needs them.
    std::vector<int> order;
    {
        const std::vector<int> lo = lane_order(g, false);
        for (int gi : lo)
            if (g->groups[gi].lane != 0) order.push_back(gi);
        for (int gi : lo)
            if (g->groups[gi].lane == 0) order.push_back(gi);
    }
    auto gate = [&](Group &G, std::vector<Dep> &deps) {  // lane 0 waits for every side group's partial sums
        if (G.lane != 0 || !G.staged || !g->gate_lane0) return;
        for (int q = 0; q < g->n_groups; ++q)
            if (g->groups[q].lane != 0) consume(g, deps[q], 0);
        wait_mask2(0);  // or the staged kernel would starve the mask generation as well
    };
    for (int gi : order) {
        Group &G = g->groups[gi];
        gate(G, D.S1);
        {
            PhaseScope ph(g, "spmm_fwd1", gi, G.lane);
            if (G.gen_feat) {  // P1_k = (X_j (.) m_k / q) W1_k, then the SpMM on P1 like layer 2 on P2
                NodeType &Tj = g->types[G.j];
                FeatArgs f = {};
                f.rowptr = Tj.X.rowptr, f.col = Tj.X.col, f.val = Tj.X.val, f.eid = nullptr;
                f.n_rows = G.n_j, f.in_rows = G.F_j, f.in = g->params + G.w1_off, f.out = G.P1buf;
                f.K = G.Kl, f.P = P1, f.nnz = G.feat_nnz, f.mask = drop ? G.mask1 : nullptr, f.scale = scale;
                launch_feature_product(f, lane_stream(g, G.lane));
                g->launches++;
                spmm_fwd(G, G.P1buf, P1, (long long)G.Kl * G.n_j, G.part1, G.slots1, G.wstart1, nullptr);
            } else {
                spmm_fwd(G, g->params + G.w1_off, P1, (long long)G.Kl * G.F_j, G.part1, G.slots1, G.wstart1, drop ? G.mask1 : nullptr);
            }
        }
        if (G.partitioned) {  // its own phase: the wait for the slowest rank is not SpMM time
            PhaseScope ph(g, "exchange", gi, G.lane);
            if (G.staged) exchange(g, G, 0, G.part1, G.slots1.n_slots, lane_stream(g, G.lane));
            else exchange(g, G, 0, G.rows1, 1, lane_stream(g, G.lane));
        }
        produced(g, D.S1[gi], G.lane);
    }
    for (int lane = 0; lane < dgn_graph::kMaxLanes; ++lane)
        for (int t = 0; t < g->n_types; ++t)
            if (g->types[t].lane == lane) epilogue(t, 1);
    for (int gi : order) {
        Group &G = g->groups[gi];
        consume(g, D.H[G.j], G.lane);
        wait_mask2(G.lane);
        {
            PhaseScope ph(g, "project", gi, G.lane);
            DenseArgs a = {};
            a.H = g->types[G.j].H, a.W2 = g->params + G.w2_off, a.P2 = G.P2;
            a.mask = drop ? G.mask2 : nullptr, a.scale = scale, a.K = G.Kl, a.n_j = G.n_j;
            a.n_rb = G.n_rb_pd, a.n_slots = G.slots_proj;
            if (G.dense_tc) launch_project_tc(a, g->d1, lane_stream(g, G.lane));
            else launch_project(a, g->d1, g->d2, lane_stream(g, G.lane));
            g->launches++;
        }
        gate(G, D.S2);
        {
            PhaseScope ph(g, "spmm_fwd2", gi, G.lane);
            spmm_fwd(G, G.P2, 1, (long long)G.Kl * G.n_j, G.part2, G.slots2, G.wstart2, nullptr);
        }
        if (G.partitioned) {
            PhaseScope ph(g, "exchange", gi, G.lane);
            if (G.staged) exchange(g, G, 1, G.part2, G.slots2.n_slots, lane_stream(g, G.lane));
            else exchange(g, G, 1, G.rows2, 1, lane_stream(g, G.lane));
        }
        produced(g, D.S2[gi], G.lane);
    }
    for (int lane = 0; lane < dgn_graph::kMaxLanes; ++lane)
        for (int t = 0; t < g->n_types; ++t)
            if (g->types[t].lane == lane) epilogue(t, 2);
}

struct AdamStep {  // valid (alpha != 0) when the update may be fused into the kernels that produce the gradients
    float alpha = 0.f, omb1 = 0.f, omb2 = 0.f, eps = 0.f;
};

// Adam of this group's layer-1 weights runs inside the kernel that produces their gradient (which then never
// reaches HBM): staged backward path, identity features, and the caller did not ask for the gradients
bool adam_fused(const dgn_graph *g, const Group &G, const AdamStep &adam) {
    return adam.alpha != 0.f && G.tstaged && !G.gen_feat && g->fuse_adam && !g->keep_grads;
}

void run_backward(dgn_graph *g, float rate, StepDeps &D, const AdamStep &adam, bool draw_ahead = false) {
    const int P1 = g->P1;
    const bool drop = rate > 0.f;
    const float scale = drop ? 1.f / (1.f - rate) : 1.f;
    auto spmm_bwd = [&](Group &G, int P, float *out, long long out_rows, const uint32_t *row_mask, bool fuse_adam, bool layer1) {
        cudaStream_t s = lane_stream(g, G.lane);
        if (G.tstaged) {
            const bool by_address = layer1 && G.task_bwd1.n_warps > 0;  // the address-ordered copy of the streams
            const TaskCsr &T = by_address ? G.task_bwd1 : G.task_bwd;
            const SlotTable &S = by_address ? G.slots_bwd1 : G.slots_bwd;
            TaskArgs a = {};
            a.hdr = T.hdr, a.ent = T.ent, a.orow = T.orow, a.orow_stride = T.orow_stride;
            a.wstart = by_address ? G.wstart_bwd1 : G.wstart_bwd;
            a.K = G.Kl, a.n_warps = T.n_warps, a.rpq = T.rpq;
            a.n_out_rows = G.n_j, a.n_op_rows = G.n_i;
            a.op = G.dS, a.P = P;
            a.slot_ptr = S.ptr, a.slot_rel = S.rel, a.n_slots = S.n_slots;
            a.out = out, a.mask = row_mask, a.scale = scale;
            if (fuse_adam) {
                const size_t off = (size_t)(out - g->grads);
                a.adam_p = g->params + off, a.adam_m = g->adam_m + off, a.adam_v = g->adam_v + off;
                a.dyn = g->dyn_dev;
            }
            launch_spmm_tstaged(a, s);
            g->launches++;
            return;
        }
        SpmmArgs a = {};
        a.rowptr = G.bwd.rowptr, a.col = G.bwd.col, a.val = G.bwd.val;
        a.seg_row = G.bwd_seg.seg_row, a.seg_begin = G.bwd_seg.seg_begin, a.row_seg_ptr = G.bwd_seg.row_seg_ptr;
        a.seg_len = G.bwd_seg.seg_len, a.n_seg = G.bwd_seg.n_seg, a.n_rows = (int)out_rows;
        a.op = G.dS, a.op_rows = G.n_i;
        a.out = out, a.out_rows = (int)out_rows, a.partial = G.bwd_partial;
        a.mask = row_mask, a.row_mask = row_mask != nullptr, a.scale = scale;
        launch_spmm(a, P, s);
        g->launches++;
        if (G.bwd_seg.n_multi > 0) {
            launch_seg_reduce(a, G.bwd_seg.multi_rows, G.bwd_seg.n_multi, P, s);
            g->launches++;
        }
    };
    // backward of layer 2: on lane 1 the groups whose COLUMN type is summed on lane 0 come first (lane 0 waits for
    // their dH partials); backward of layer 1: any order
    const std::vector<int> order = lane_order(g, false), order2 = lane_order(g, true);
    std::vector<std::pair<int, DenseArgs>> deferred_dw2;
    auto run_dw2 = [&](int gi, DenseArgs a) {
        Group &G = g->groups[gi];
        cudaStream_t s = lane_stream(g, G.lane);
        PhaseScope ph(g, "dw2", gi, G.lane);
        a.n_rb = G.n_rb;
        a.n_slots = G.slots_dw2;
        if (G.dense_tc) launch_dw2_tc(a, g->d1, s);
        else launch_dw2(a, g->d1, g->d2, s);
        g->launches++;
        if (G.n_rb > 1) {
            launch_dw2_reduce(G.dW2part, g->grads + G.w2_off, G.Kl, G.n_rb, g->d1 * g->d2, s);
            g->launches++;
        }
    };
    // ---- layer 2
    for (int gi : order2) {
        Group &G = g->groups[gi];
        cudaStream_t s = lane_stream(g, G.lane);
        consume(g, D.dZ[G.i], G.lane);
        consume(g, D.H[G.j], G.lane);
        {
            PhaseScope ph(g, "epilogue", -1, G.lane);
            L2BwdArgs l = {G.Y2, G.n2, g->types[G.i].dZ, G.dS, G.n_i};
            launch_l2norm_bwd(l, 1, s);
            g->launches++;
        }
        {
            PhaseScope ph(g, "spmm_bwd2", gi, G.lane);
            spmm_bwd(G, 1, G.G2, (long long)G.Kl * G.n_j, nullptr, false, false);
        }
        DenseArgs a = {};
        a.H = g->types[G.j].H, a.W2 = g->params + G.w2_off, a.G2 = G.G2;
        a.mask = drop ? G.mask2 : nullptr, a.scale = scale, a.K = G.Kl, a.n_j = G.n_j;
        a.n_rb = G.n_rb;
        a.dW2 = G.n_rb > 1 ? G.dW2part : g->grads + G.w2_off;
        a.dHpart = G.dHpart;
        // dW2 of a lane-0 group feeds nothing but Adam: it is queued behind the group's layer-1 backward, where its
        // tensor-core kernel (which leaves room on the SMs) runs beside lane 1's remaining kernels
        if (g->two_lanes && G.lane == 0 && G.dense_tc) deferred_dw2.push_back({gi, a});
        else run_dw2(gi, a);
        {
            PhaseScope ph(g, "dh", gi, G.lane);
            a.n_slots = G.slots_dh, a.n_rb = G.n_rb_pd;
            if (G.dense_tc) launch_dh_tc(a, g->d1, s);
            else launch_dh(a, g->d1, g->d2, s);
            g->launches++;
        }
        if (G.partitioned) {
            PhaseScope ph(g, "exchange", gi, G.lane);
            exchange(g, G, 2, G.dHpart, G.slots_dh, s);
        }
        produced(g, D.dH[gi], G.lane);
    }
    for (int lane = 0; lane < dgn_graph::kMaxLanes; ++lane)
        for (int t = 0; t < g->n_types; ++t) {
            NodeType &T = g->types[t];
            if (T.lane != lane) continue;
            ReluBwdArgs r = {};
            r.n_rows = T.n;
            r.H = T.H, r.dA = T.dA;
            consume(g, D.H[t], T.lane);
            for (int gi : T.col_groups) {
                consume(g, D.dH[gi], T.lane);
                r.g[r.n_groups].part = g->groups[gi].dHpart;
                r.g[r.n_groups].n_chunks = g->groups[gi].slots_dh;
                if (g->groups[gi].partitioned) {
                    r.g[r.n_groups].n_peers = g->world;
                    for (int q = 0; q < g->world; ++q) r.g[r.n_groups].peer[q] = peer_ptr(g, g->groups[gi], 2, q);
                }
                r.n_groups++;
            }
            PhaseScope ph(g, "epilogue", -1, T.lane);
            launch_relu_bwd(r, P1, lane_stream(g, T.lane));
            g->launches++;
            produced(g, D.dA[t], T.lane);
        }
    // ---- layer 1
    for (int gi : order) {
        Group &G = g->groups[gi];
        cudaStream_t s = lane_stream(g, G.lane);
        consume(g, D.dA[G.i], G.lane);
        {
            PhaseScope ph(g, "epilogue", -1, G.lane);
            L2BwdArgs l = {G.Y1, G.n1, g->types[G.i].dA, G.dS, G.n_i};
            launch_l2norm_bwd(l, P1, s);
            g->launches++;
        }
        PhaseScope ph(g, "spmm_bwd1", gi, G.lane);
        if (G.gen_feat) {  // G1_k = A_k^T dS1, then dW1_k = (X_j (.) m_k / q)^T G1_k
            spmm_bwd(G, P1, G.G1buf, (long long)G.Kl * G.n_j, nullptr, false, true);
            NodeType &Tj = g->types[G.j];
            FeatArgs f = {};
            f.rowptr = Tj.Xt.rowptr, f.col = Tj.Xt.col, f.val = Tj.Xt.val, f.eid = Tj.Xt_eid;
            f.n_rows = G.F_j, f.in_rows = G.n_j, f.in = G.G1buf, f.out = g->grads + G.w1_off;
            f.K = G.Kl, f.P = P1, f.nnz = G.feat_nnz, f.mask = drop ? G.mask1 : nullptr, f.scale = scale;
            launch_feature_product(f, s);
            g->launches++;
            continue;
        }
        spmm_bwd(G, P1, g->grads + G.w1_off, (long long)G.Kl * G.F_j, drop ? G.mask1 : nullptr, adam_fused(g, G, adam), true);
    }
    if (draw_ahead) {
        // next step's layer-2 keep words into the OTHER buffer, on the mask stream, from the moment lane 0 has issued
        // its last persistent kernel: the integer-ALU work runs beside the tensor-core dW2 kernel (few warps per SM)
        // and the side lanes' tail instead of in front of the next step's first staged kernel
        CUDA_CHECK(cudaEventRecord(g->ahead_go, g->stream));
        CUDA_CHECK(cudaStreamWaitEvent(g->stream3, g->ahead_go, 0));
        for (int side = 1; side >= 0; --side)
            for (auto &G : g->groups) {
                if ((G.lane != 0) != (side == 1)) continue;
                launch_gen_mask(G.mask2buf[1 - g->mask_cur], G.mask2_words, (long long)G.n_j * g->d1, G.n_j * P1, G.rel_ids,
                                kStreamDropout2, g->dyn_dev, 1, g->stream3);
                g->launches++;
            }
        CUDA_CHECK(cudaEventRecord(g->ahead_done, g->stream3));
    }
    for (auto &d : deferred_dw2) run_dw2(d.first, d.second);
    if (draw_ahead) CUDA_CHECK(cudaStreamWaitEvent(g->stream, g->ahead_done, 0));
    join_lanes(g, false);
}

// logical [K][rows][32 P] row-major <-> device [P][K * rows][32]
// row-major [rows, d] (d <= 32 * P columns; the panels' remaining columns are padding) <-> panel layout [P][rows][32]
void pack_panels(const float *src, float *dst, long long stacked_rows, int P, int d) {
    if (d < 32 * P) std::fill(dst, dst + (size_t)P * stacked_rows * 32, 0.f);
    for (long long r = 0; r < stacked_rows; ++r)
        for (int c = 0; c < d; ++c) dst[((size_t)(c >> 5) * stacked_rows + r) * 32 + (c & 31)] = src[(size_t)r * d + c];
}
void unpack_panels(const float *src, float *dst, long long stacked_rows, int P, int d) {
    for (long long r = 0; r < stacked_rows; ++r)
        for (int c = 0; c < d; ++c) dst[(size_t)r * d + c] = src[((size_t)(c >> 5) * stacked_rows + r) * 32 + (c & 31)];
}

struct ParamSpan {
    size_t off;         // floats into the arena (row-major kinds) or arena offset of the group's W1 block
    long long count;    // floats covered
    bool panels;        // W1: panel layout, needs (un)packing
    long long row0, rows, stacked_rows;
};

// k is the LOCAL index for the encoder weights (W1 / W2) and the group-wide index for decoder variables
ParamSpan locate_param(dgn_graph *g, int kind, int group, int k) {
    Group &G = g->groups[group];
    ParamSpan s = {};
    const bool encoder = kind == DGN_PARAM_W1 || kind == DGN_PARAM_W2;
    const long long Kk = encoder ? G.Kl : G.K;
    const long long nk = k < 0 ? Kk : 1, k0 = k < 0 ? 0 : k;
    switch (kind) {
        case DGN_PARAM_W1:
            s.panels = true;
            s.off = G.w1_off;
            s.row0 = k0 * G.F_j, s.rows = nk * G.F_j, s.stacked_rows = (long long)G.Kl * G.F_j;
            s.count = s.rows * g->d1;
            break;
        case DGN_PARAM_W2:
            s.off = G.w2_off + (size_t)k0 * g->d1 * g->d2;
            s.count = nk * g->d1 * g->d2;
            break;
        case DGN_PARAM_DEC_GLOBAL:
            DGN_REQUIRE(G.decoder == DGN_DEC_DEDICOM, "group %d has no global_interaction (decoder kind %d)", group, G.decoder);
            s.off = G.glb_off;
            s.count = (long long)g->d2 * g->d2;
            break;
        case DGN_PARAM_DEC_LOCAL:
            DGN_REQUIRE(G.loc_per_rel > 0, "group %d (innerproduct) has no per-relation decoder variable", group);
            s.off = G.loc_off + (size_t)k0 * G.loc_per_rel;
            s.count = nk * (long long)G.loc_per_rel;
            break;
        default: DGN_FAIL(DGN_ERR_INVALID, "unknown parameter kind %d", kind);
    }
    return s;
}

// floats of one relation's variable of this kind as they cross the C ABI
long long param_floats_per_relation(dgn_graph *g, int kind, int group) {
    Group &G = g->groups[group];
    switch (kind) {
        case DGN_PARAM_W1: return (long long)G.F_j * g->d1u;
        case DGN_PARAM_W2: return (long long)g->d1u * g->d2u;
        case DGN_PARAM_DEC_GLOBAL: return (long long)g->d2u * g->d2u;
        case DGN_PARAM_DEC_LOCAL:
            return G.decoder == DGN_DEC_BILINEAR ? (long long)g->d2u * g->d2u : G.loc_per_rel ? (long long)g->d2u : 0;
        default: DGN_FAIL(DGN_ERR_INVALID, "unknown parameter kind %d", kind);
    }
    return 0;
}

void arena_write(dgn_graph *g, float *arena, const ParamSpan &s, const float *values) {
    if (!s.panels) {
        CUDA_CHECK(cudaMemcpy(arena + s.off, values, (size_t)s.count * sizeof(float), cudaMemcpyHostToDevice));
        return;
    }
    // panel p of rows [row0, row0 + rows) is contiguous on the device
    std::vector<float> tmp((size_t)s.count);
    pack_panels(values, tmp.data(), s.rows, g->P1, g->d1u);
    for (int p = 0; p < g->P1; ++p)
        CUDA_CHECK(cudaMemcpy(arena + s.off + ((size_t)p * s.stacked_rows + s.row0) * 32, tmp.data() + (size_t)p * s.rows * 32,
                              (size_t)s.rows * 32 * sizeof(float), cudaMemcpyHostToDevice));
}

void arena_read(dgn_graph *g, const float *arena, const ParamSpan &s, float *values) {
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    if (!s.panels) {
        CUDA_CHECK(cudaMemcpy(values, arena + s.off, (size_t)s.count * sizeof(float), cudaMemcpyDeviceToHost));
        return;
    }
    std::vector<float> tmp((size_t)s.count);
    for (int p = 0; p < g->P1; ++p)
        CUDA_CHECK(cudaMemcpy(tmp.data() + (size_t)p * s.rows * 32, arena + s.off + ((size_t)p * s.stacked_rows + s.row0) * 32,
                              (size_t)s.rows * 32 * sizeof(float), cudaMemcpyDeviceToHost));
    unpack_panels(tmp.data(), values, s.rows, g->P1, g->d1u);
}

void drop_step_graphs(dgn_graph *g) {
    for (auto &kv : g->step_graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    g->step_graphs.clear();
}

void ensure_batch_capacity(dgn_graph *g, int B) {
    if (B <= g->ring_cap) return;
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    drop_step_graphs(g);  // they hold the old addresses
    const int cap = std::max(B, 1024);
    g->step_neg_off = (sizeof(StepDyn) + 15) / 16 * 16;
    g->step_batch_off = g->step_neg_off + (size_t)cap * sizeof(long long);
    const size_t total = g->step_batch_off + (size_t)cap * 2 * sizeof(int);
    for (int i = 0; i < dgn_graph::kRing; ++i) {
        if (g->step_host[i]) cudaFreeHost(g->step_host[i]);
        CUDA_CHECK(cudaMallocHost(&g->step_host[i], total));
        memset(g->step_host[i], 0, total);
    }
    dev_free(g->step_dev);
    dev_free(g->neg_out);
    dev_free(g->pos_out);
    dev_free(g->negs_out);
    g->step_dev = dev_alloc<unsigned char>(total);
    CUDA_CHECK(cudaMemset(g->step_dev, 0, total));
    g->dyn_dev = reinterpret_cast<StepDyn *>(g->step_dev);
    g->neg_dev = reinterpret_cast<long long *>(g->step_dev + g->step_neg_off);
    g->batch_dev = reinterpret_cast<int *>(g->step_dev + g->step_batch_off);
    g->neg_out = dev_alloc<long long>(cap);
    g->pos_out = dev_alloc<float>(cap);
    g->negs_out = dev_alloc<float>(cap);
    g->ring_cap = cap;
}

// Next pinned slot of the ring with the per-step header filled in: dropout stream words, Adam coefficients (when
// given) and the stamps of the exchanges this call will run (n_exchanges_per_group: 2 = forward only, 3 = a whole
// training step).  The caller adds the decode arguments / batch and calls commit_step.
int begin_step(dgn_graph *g, float rate, uint64_t seed, uint32_t step, int n_exchanges_per_group) {
    ensure_batch_capacity(g, 1);
    const int slot = g->ring_pos;
    g->ring_pos = (g->ring_pos + 1) % dgn_graph::kRing;
    CUDA_CHECK(cudaEventSynchronize(g->ring_ev[slot]));
    StepDyn *h = reinterpret_cast<StepDyn *>(g->step_host[slot]);
    memset(h, 0, sizeof(StepDyn));
    h->step = step, h->seed_lo = (uint32_t)(seed & 0xffffffffu), h->seed_hi = (uint32_t)(seed >> 32);
    h->threshold = dropout_threshold(rate);
    for (auto &G : g->groups)
        if (G.partitioned)
            for (int x = 0; x < n_exchanges_per_group; ++x) h->stamp[G.xch[x].id] = ++G.xch[x].stamp;
    return slot;
}
void commit_step(dgn_graph *g, int slot, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(g->step_dev, g->step_host[slot], bytes, cudaMemcpyHostToDevice, g->stream));
    CUDA_CHECK(cudaEventRecord(g->ring_ev[slot], g->stream));
}

// all-pairs scores: tensor cores (tcgen05, 3 x TF32) unless DGN_PREDICT_FFMA=1 asks for the CUDA-core kernel
void run_predict(dgn_graph *g, const PredictArgs &a) {
    const char *e = getenv("DGN_PREDICT_FFMA");
    if (e && e[0] == '1') launch_predict(a, g->stream);
    else launch_predict_tc(a, g->n_sm, g->stream);
    g->launches++;
}

PredictArgs predict_args(dgn_graph *g, int r, int count) {
    DGN_REQUIRE(r >= 0 && r + count <= g->R && count >= 1, "relation range [%d, %d) out of range", r, r + count);
    const int gi = g->flat[r].first, k = g->flat[r].second;
    DGN_REQUIRE(g->flat[r + count - 1].first == gi, "relations %d..%d span more than one group", r, r + count - 1);
    Group &G = g->groups[gi];
    PredictArgs a = {};
    a.Zi = g->types[G.i].Z, a.Zj = g->types[G.j].Z;
    a.n_i = G.n_i, a.n_j = G.n_j, a.decoder = G.decoder;
    a.glb = G.decoder == DGN_DEC_DEDICOM ? g->params + G.glb_off : nullptr;
    a.loc = G.loc_per_rel ? g->params + G.loc_off + (size_t)k * G.loc_per_rel : nullptr;
    a.loc_stride = (long long)G.loc_per_rel;
    a.count = count;
    return a;
}

void free_arena(dgn_graph *g) {
    dev_free(g->params);
    dev_free(g->grads);
    dev_free(g->adam_m);
    dev_free(g->adam_v);
    g->arena_ready = false;
}

// parameter arena: [W1 of every group | W2 of every group | decoder variables], encoder parts sized by
// the LOCAL relation count of each group
void layout_arena(dgn_graph *g) {
    size_t off = 0;
    for (auto &G : g->groups) {
        G.w1_off = off;
        off += (size_t)G.Kl * G.F_j * g->d1;
    }
    for (auto &G : g->groups) {
        G.w2_off = off;
        off += (size_t)G.Kl * g->d1 * g->d2;
    }
    g->dec_off = off;
    for (auto &G : g->groups) {
        if (G.decoder == DGN_DEC_DEDICOM) {
            G.glb_off = off;
            off += (size_t)g->d2 * g->d2;
        }
        G.loc_per_rel = G.decoder == DGN_DEC_BILINEAR ? (size_t)g->d2 * g->d2 : G.decoder == DGN_DEC_INNERPRODUCT ? 0 : (size_t)g->d2;
        G.loc_off = off;
        off += (size_t)G.K * G.loc_per_rel;
    }
    if (g->arena_ready && off == g->n_params) return;
    free_arena(g);
    g->n_params = off;
    g->params = dev_alloc<float>(off);
    g->grads = dev_alloc<float>(off);
    g->adam_m = dev_alloc<float>(off);
    g->adam_v = dev_alloc<float>(off);
    CUDA_CHECK(cudaMemset(g->params, 0, off * sizeof(float)));
    CUDA_CHECK(cudaMemset(g->grads, 0, off * sizeof(float)));
    CUDA_CHECK(cudaMemset(g->adam_m, 0, off * sizeof(float)));
    CUDA_CHECK(cudaMemset(g->adam_v, 0, off * sizeof(float)));
    g->arena_ready = true;
}

void set_local_relations(Group &G, const std::vector<int> &loc) {
    G.loc = loc;
    G.Kl = (int)loc.size();
    G.loc_index.assign(G.K, -1);
    for (int l = 0; l < G.Kl; ++l) G.loc_index[loc[l]] = l;
}

void check_arena(dgn_graph *g) {
    DGN_REQUIRE(g->arena_ready, "with more than one rank the parameters exist after dgn_graph_finalize (the partition needs the relations)");
}

}  // namespace

extern "C" int dgn_partition_relations(const int64_t *weights, int32_t K, int32_t world, int32_t *owner_out) {
    // longest-processing-time assignment, ties by index: every rank computes the same owners
    if (!weights || !owner_out || K < 0 || world < 1) {
        set_error("dgn_partition_relations: bad argument");
        return DGN_ERR_INVALID;
    }
    std::vector<int> order(K);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return weights[x] > weights[y]; });
    typedef std::pair<long long, int> Load;
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (int p = 0; p < world; ++p) heap.push(Load(0, p));
    for (int k : order) {
        Load l = heap.top();
        heap.pop();
        owner_out[k] = l.second;
        heap.push(Load(l.first + weights[k], l.second));
    }
    return DGN_OK;
}

#define DGN_API_BEGIN try {
#define DGN_API_END                         \
    return DGN_OK;                          \
    }                                       \
    catch (const Failure &f) { return f.code; } \
    catch (const std::bad_alloc &) {        \
        set_error("host allocation failed"); \
        return DGN_ERR_INVALID;             \
    }

extern "C" int dgn_device_count(int *count_out) {
    DGN_API_BEGIN
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count_out = n;
    DGN_API_END
}

extern "C" int dgn_graph_create(dgn_graph **out, int device, int n_types, const int32_t *n_nodes, const int32_t *feat_dim,
                                int n_groups, const int32_t *group_ij, const int32_t *group_K,
                                const int32_t *group_decoder, int hidden1, int hidden2) {
    DGN_API_BEGIN
    DGN_REQUIRE(out && n_nodes && feat_dim && group_ij && group_K && group_decoder, "null argument");
    DGN_REQUIRE(n_types > 0 && n_groups > 0, "need at least one node type and one group");
    if (hidden2 < 1 || hidden2 > 32) DGN_FAIL(DGN_ERR_UNSUPPORTED, "hidden2 = %d is not supported (1 .. 32)", hidden2);
    if (hidden1 < 1 || hidden1 > 128) DGN_FAIL(DGN_ERR_UNSUPPORTED, "hidden1 = %d is not supported (1 .. 128)", hidden1);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        DGN_FAIL(DGN_ERR_NO_DEVICE, "no CUDA device is visible; decagon_b200 has no CPU fallback");
    }
    DGN_REQUIRE(device >= 0 && device < n_dev, "device %d out of range (%d visible)", device, n_dev);
    CUDA_CHECK(cudaSetDevice(device));
    std::unique_ptr<dgn_graph> g(new dgn_graph());
    g->device = device;
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    g->n_sm = prop.multiProcessorCount;
    g->n_types = n_types, g->n_groups = n_groups, g->d1u = hidden1, g->d2u = hidden2;
    g->d1 = hidden1 <= 32 ? 32 : hidden1 <= 64 ? 64 : 128, g->d2 = 32, g->P1 = g->d1 / 32;  // the kernels' panel counts: 1, 2, 4
    g->types.resize(n_types);
    for (int t = 0; t < n_types; ++t) {
        DGN_REQUIRE(n_nodes[t] > 0 && feat_dim[t] > 0, "node type %d: empty", t);
        g->types[t].n = n_nodes[t];
        g->types[t].F = feat_dim[t];
    }
    g->groups.resize(n_groups);
    int r = 0;
    for (int gi = 0; gi < n_groups; ++gi) {
        Group &G = g->groups[gi];
        G.i = group_ij[2 * gi], G.j = group_ij[2 * gi + 1], G.K = group_K[gi], G.decoder = group_decoder[gi];
        DGN_REQUIRE(G.i >= 0 && G.i < n_types && G.j >= 0 && G.j < n_types, "group %d: node types (%d, %d) out of range", gi, G.i, G.j);
        DGN_REQUIRE(G.K > 0, "group %d: no relations", gi);
        DGN_REQUIRE(G.decoder >= DGN_DEC_INNERPRODUCT && G.decoder <= DGN_DEC_DEDICOM, "Unknown decoder type %d", G.decoder);
        G.n_i = n_nodes[G.i], G.n_j = n_nodes[G.j], G.F_j = feat_dim[G.j];
        G.r0 = r;
        G.rel.resize(G.K);
        G.rel_set.assign(G.K, false);
        G.thr.assign(G.K, nullptr);
        G.thr_n.assign(G.K, 0);
        for (int k = 0; k < G.K; ++k) g->flat.push_back(std::make_pair(gi, k));
        r += G.K;
        g->types[G.i].row_groups.push_back(gi);
        g->types[G.j].col_groups.push_back(gi);
        std::vector<int> all(G.K);
        std::iota(all.begin(), all.end(), 0);
        set_local_relations(G, all);
    }
    g->R = r;
    layout_arena(g.get());  // one rank: every relation is local (dgn_comm_init re-partitions at finalize)
    {
        int lo = 0, hi = 0;
        CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_CHECK(cudaStreamCreateWithPriority(&g->stream, cudaStreamNonBlocking, hi));
        for (int l = 1; l < dgn_graph::kMaxLanes; ++l) CUDA_CHECK(cudaStreamCreateWithPriority(&g->side[l], cudaStreamNonBlocking, lo));
        CUDA_CHECK(cudaStreamCreateWithPriority(&g->stream3, cudaStreamNonBlocking, lo));
        CUDA_CHECK(cudaEventCreateWithFlags(&g->mask_go, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&g->mask_done, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&g->ahead_go, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&g->ahead_done, cudaEventDisableTiming));
    }
    g->own_stream = true;
    for (int i = 0; i < dgn_graph::kRing; ++i) CUDA_CHECK(cudaEventCreateWithFlags(&g->ring_ev[i], cudaEventDisableTiming));
    g->loss_dev = dev_alloc<float>(2);  // [loss, serial number of the step]
    g->exchange_error = dev_alloc<int>(1);
    CUDA_CHECK(cudaMemset(g->exchange_error, 0, sizeof(int)));
    CUDA_CHECK(cudaMallocHost(&g->exchange_error_host, sizeof(int)));
    *g->exchange_error_host = 0;
    if (const char *t = getenv("DGN_EXCHANGE_TIMEOUT_S")) g->exchange_timeout_ns = (unsigned long long)(atof(t) * 1e9);
    g->decode_scratch = dev_alloc<float>((size_t)kDecodeCtas * (32 * 32 + 1));
    g->decode_ticket = dev_alloc<unsigned>(1);
    CUDA_CHECK(cudaMemset(g->decode_ticket, 0, sizeof(unsigned)));
    CUDA_CHECK(cudaMallocHost(&g->loss_host, 2 * sizeof(float)));
    g->loss_host[0] = g->loss_host[1] = 0.f;
    if (const char *e = getenv("DGN_SYNC_LOSS")) g->early_loss = e[0] != '1';
    for (int t = 0; t < n_types; ++t) {
        NodeType &T = g->types[t];
        T.H = dev_alloc<float>(panel_floats(g->P1, T.n));
        T.dA = dev_alloc<float>(panel_floats(g->P1, T.n));
        T.Z = dev_alloc<float>(panel_floats(1, T.n));
        T.dZ = dev_alloc<float>(panel_floats(1, T.n));
        T.dZq = dev_alloc<long long>(panel_floats(1, T.n));
        CUDA_CHECK(cudaMemset(T.dZ, 0, panel_floats(1, T.n) * sizeof(float)));
    }
    const char *env = getenv("DGN_DISABLE_STAGED");
    g->allow_staged = !(env && env[0] == '1');
    env = getenv("DGN_STAGED_VERSION");
    if (env && env[0] == '2') g->staged_version = 2;
    env = getenv("DGN_FUSE_ADAM");
    g->fuse_adam = !(env && env[0] == '0');
    env = getenv("DGN_SINGLE_STREAM");
    g->two_lanes = !(env && env[0] == '1');
    env = getenv("DGN_MASK_AHEAD");
    g->mask_ahead = env && env[0] == '1';
    env = getenv("DGN_GATHER_ROWSUMS");
    g->gather_row_sums = env && env[0] == '1';
    env = getenv("DGN_CUDA_GRAPH");
    g->use_graphs = !(env && env[0] == '0');
    env = getenv("DGN_DISABLE_TSTAGED");
    g->allow_tstaged = !(env && env[0] == '1');
    *out = g.release();
    DGN_API_END
}

extern "C" int dgn_graph_destroy(dgn_graph *g) {
    DGN_API_BEGIN
    if (!g) return DGN_OK;
    cudaSetDevice(g->device);
    cudaDeviceSynchronize();
    for (auto &G : g->groups) {
        free_group_device(G);
        for (auto &t : G.thr) dev_free(t);
    }
    for (auto &T : g->types) {
        dev_free(T.H);
        dev_free(T.Z);
        dev_free(T.dZ);
        dev_free(T.dZq);
        dev_free(T.dA);
        free_csr(T.X);
        free_csr(T.Xt);
        dev_free(T.Xt_eid);
    }
    free_arena(g);
    free_comm(g);
    drop_step_graphs(g);
    dev_free(g->step_dev);
    dev_free(g->neg_out);
    dev_free(g->pos_out);
    dev_free(g->negs_out);
    dev_free(g->loss_dev);
    dev_free(g->exchange_error);
    if (g->exchange_error_host) cudaFreeHost(g->exchange_error_host);
    dev_free(g->decode_scratch);
    dev_free(g->decode_ticket);
    if (g->loss_host) cudaFreeHost(g->loss_host);
    for (int i = 0; i < dgn_graph::kRing; ++i) {
        if (g->step_host[i]) cudaFreeHost(g->step_host[i]);
        cudaEventDestroy(g->ring_ev[i]);
    }
    for (auto &p : g->phases) {
        cudaEventDestroy(p.start);
        cudaEventDestroy(p.stop);
    }
    for (auto &e : g->dep_events) cudaEventDestroy(e);
    if (g->own_stream && g->stream) cudaStreamDestroy(g->stream);
    for (int l = 1; l < dgn_graph::kMaxLanes; ++l)
        if (g->side[l]) cudaStreamDestroy(g->side[l]);
    if (g->stream3) cudaStreamDestroy(g->stream3);
    if (g->mask_go) cudaEventDestroy(g->mask_go);
    if (g->mask_done) cudaEventDestroy(g->mask_done);
    if (g->ahead_go) cudaEventDestroy(g->ahead_go);
    if (g->ahead_done) cudaEventDestroy(g->ahead_done);
    delete g;
    DGN_API_END
}

extern "C" int dgn_graph_set_relation(dgn_graph *g, int r, int32_t n_rows, int32_t n_cols, int64_t nnz,
                                      const int32_t *coo_rows, const int32_t *coo_cols, const float *vals) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    DGN_REQUIRE(r >= 0 && r < g->R, "relation %d out of range (R = %d)", r, g->R);
    Group &G = g->groups[g->flat[r].first];
    DGN_REQUIRE(n_rows == G.n_i && n_cols == G.n_j, "relation %d: shape %d x %d does not match group (%d,%d): %d x %d", r, n_rows,
                n_cols, G.i, G.j, G.n_i, G.n_j);
    csr_from_coo(n_rows, n_cols, nnz, coo_rows, coo_cols, vals, G.rel[g->flat[r].second]);
    G.rel_set[g->flat[r].second] = true;
    G.dirty = true;  // only this group's device structures are rebuilt by the next dgn_graph_finalize
    g->finalized = false;
    DGN_API_END
}

extern "C" int dgn_graph_set_features(dgn_graph *g, int type, int32_t n_rows, int32_t n_cols, int64_t nnz,
                                      const int32_t *coo_rows, const int32_t *coo_cols, const float *vals) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    DGN_REQUIRE(type >= 0 && type < g->n_types, "node type %d out of range", type);
    NodeType &T = g->types[type];
    DGN_REQUIRE(n_rows == T.n && n_cols == T.F, "features of type %d: shape %d x %d, expected %d x %d", type, n_rows, n_cols, T.n, T.F);
    bool identity = n_rows == n_cols && nnz == n_rows;
    std::vector<char> seen((size_t)n_rows, 0);
    for (int64_t e = 0; identity && e < nnz; ++e) {
        identity = coo_rows[e] == coo_cols[e] && coo_rows[e] >= 0 && coo_rows[e] < n_rows && vals[e] == 1.f && !seen[coo_rows[e]];
        if (identity) seen[coo_rows[e]] = 1;
    }
    T.identity = identity;
    T.feat = HostCsr();
    if (!identity) csr_from_coo(n_rows, n_cols, nnz, coo_rows, coo_cols, vals, T.feat);  // canonical (row, col) order
    T.feat_set = true;
    for (auto &G : g->groups)
        if (G.j == type) G.dirty = true;
    g->finalized = false;
    DGN_API_END
}

extern "C" int dgn_sampler_set_degrees(dgn_graph *g, int r, const double *degrees, int32_t n) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    DGN_REQUIRE(r >= 0 && r < g->R, "relation %d out of range", r);
    Group &G = g->groups[g->flat[r].first];
    const int k = g->flat[r].second;
    DGN_REQUIRE(n == G.n_i, "relation %d: %d degrees, row type has %d nodes (range_max = len(degrees[i][k]), optimizer.py:45)", r, n, G.n_i);
    std::vector<uint32_t> thr((size_t)n);
    int rc = dgn_sampler_thresholds(degrees, n, thr.data());
    if (rc != DGN_OK) return rc;
    CUDA_CHECK(cudaSetDevice(g->device));
    dev_free(G.thr[k]);
    G.thr[k] = dev_upload(thr);
    G.thr_n[k] = n;
    DGN_API_END
}

extern "C" int dgn_graph_finalize(dgn_graph *g) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    CUDA_CHECK(cudaSetDevice(g->device));
    for (auto &G : g->groups) {
        for (int k = 0; k < G.K; ++k) DGN_REQUIRE(G.rel_set[k], "relation (%d,%d,%d) was never set", G.i, G.j, k);
        DGN_REQUIRE(g->types[G.j].feat_set, "features of node type %d were never set", G.j);
        DGN_REQUIRE(!g->types[G.j].identity || G.F_j == G.n_j, "identity features need feat_dim == n_nodes for type %d", G.j);
    }
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    drop_step_graphs(g);  // they hold the addresses of the old device layout
    g->ahead_valid = false;  // rebuilt groups get fresh mask buffers
    for (auto &T : g->types) {
        free_csr(T.X);
        free_csr(T.Xt);
        dev_free(T.Xt_eid);
        if (!T.feat_set || T.identity) continue;
        // X^T by feature, every entry remembering its position in X (= its dropout bit)
        HostCsr Xt;
        csr_transpose(T.feat, Xt);
        std::vector<int> eid(T.feat.col.size());
        {
            std::vector<int> cursor(Xt.rowptr.begin(), Xt.rowptr.end() - 1);
            for (int r = 0; r < T.feat.n_rows; ++r)
                for (int e = T.feat.rowptr[r]; e < T.feat.rowptr[r + 1]; ++e) eid[cursor[T.feat.col[e]]++] = e;
        }
        T.X = upload_csr(T.feat);
        T.Xt = upload_csr(Xt);
        T.Xt_eid = dev_upload(eid);
    }
    for (auto &G : g->groups) {
        // multi-GPU: groups that take the staged path (many small relations) are partitioned by relation
        // (also those whose operand tiles exceed shared memory and take the gather path: config #5)
        G.partitioned = g->world > 1 && G.K >= 8 * g->world;
        std::vector<int> loc;
        if (G.partitioned) {
            std::vector<int64_t> w(G.K);
            for (int k = 0; k < G.K; ++k) w[k] = G.rel[k].nnz() + G.rel[k].n_cols;
            G.owner.assign(G.K, 0);
            dgn_partition_relations(w.data(), G.K, g->world, G.owner.data());
            for (int k = 0; k < G.K; ++k)
                if (G.owner[k] == g->rank) loc.push_back(k);
        } else {
            loc.resize(G.K);
            std::iota(loc.begin(), loc.end(), 0);
        }
        if (loc != G.loc) G.dirty = true;
        set_local_relations(G, loc);
    }
    layout_arena(g);
    // incremental: only the groups whose relations (or features, or partition) changed are rebuilt -- an
    // active-learning round that re-masks the drug-drug relations (RandomMaskingActiveLearner._applyMask,
    // RandomMaskingActiveLearner.py:184-200) leaves the protein-protein structures on the device untouched
    for (auto &G : g->groups)
        if (G.dirty) {
            build_group(g, G);
            G.dirty = false;
            g->groups_rebuilt++;
        }
    build_comm(g);
    // no device allocation may happen while a peer waits for this rank inside an exchange: take the batch
    // buffers now (a larger batch later re-allocates; keep the ranks in step around such a change)
    ensure_batch_capacity(g, g->world > 1 ? 4096 : 1024);
    // stream lanes: the groups of many small relations (persistent one-CTA-per-SM kernels) on lane 0, the
    // rest on lane 1 so that their short kernels fill the gaps; per-type kernels follow their row groups
    bool any_staged = false;
    for (auto &G : g->groups) any_staged = any_staged || G.staged;
    // The gate (lane 0's persistent kernels wait for the side groups of the layer so that they do not starve them) pays
    // on one GPU, where those kernels run for 160-280 us; with several ranks they are short and the gate only
    // serialises the two lanes.  DGN_GATE overrides.
    g->gate_lane0 = g->world == 1;
    if (const char *e = getenv("DGN_GATE")) g->gate_lane0 = e[0] != '0';
    // With several ranks the side groups get one lane each (round robin when there are more groups than lanes): lane 0's
    // kernels shrink with the rank count and the side groups become the critical path, so their short kernels should
    // overlap each other.  On one GPU a single side lane measured 1 % faster (the persistent lane-0 kernels fill the
    // SMs either way).  DGN_SIDE_LANES overrides.
    int side_lanes = g->world > 1 ? dgn_graph::kMaxLanes - 1 : 1;
    if (const char *e = getenv("DGN_SIDE_LANES")) side_lanes = std::max(1, std::min(side_lanes, atoi(e)));
    int next = 0;
    g->n_lanes = 1;
    for (auto &G : g->groups) {
        G.lane = 0;
        if (g->two_lanes && any_staged && !G.staged) {
            G.lane = 1 + next++ % side_lanes;
            g->n_lanes = std::max(g->n_lanes, G.lane + 1);
        }
    }
    for (auto &T : g->types) {
        T.lane = -1;
        for (int gi : T.row_groups)
            if (g->groups[gi].lane == 0) T.lane = 0;
        if (T.lane < 0 && !T.row_groups.empty()) T.lane = g->groups[T.row_groups[0]].lane;
        if (T.lane < 0 || !g->two_lanes || !any_staged) T.lane = 0;
    }
    g->finalized = true;
    DGN_API_END
}

extern "C" int dgn_graph_relation_nnz(dgn_graph *g, int r, int64_t *nnz_out) {
    DGN_API_BEGIN
    DGN_REQUIRE(g && r >= 0 && r < g->R, "relation %d out of range", r);
    *nnz_out = g->groups[g->flat[r].first].rel[g->flat[r].second].nnz();
    DGN_API_END
}

extern "C" int dgn_graph_get_csr(dgn_graph *g, int r, int32_t *rowptr_out, int32_t *col_out, float *val_out) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(r >= 0 && r < g->R, "relation %d out of range", r);
    CUDA_CHECK(cudaSetDevice(g->device));
    Group &G = g->groups[g->flat[r].first];
    const int k = g->flat[r].second;
    std::vector<int> rp((size_t)G.n_i + 1);
    CUDA_CHECK(cudaMemcpy(rp.data(), G.relcsr.rowptr + (size_t)k * (G.n_i + 1), rp.size() * sizeof(int), cudaMemcpyDeviceToHost));
    const int base = rp[0], nnz = rp[G.n_i] - base;
    for (int u = 0; u <= G.n_i; ++u) rowptr_out[u] = rp[u] - base;
    CUDA_CHECK(cudaMemcpy(col_out, G.relcsr.col + base, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(val_out, G.relcsr.val + base, (size_t)nnz * sizeof(float), cudaMemcpyDeviceToHost));
    DGN_API_END
}

namespace {
// k: group-wide relation index or -1 (all K stacked).  Encoder weights of relations another rank owns are
// skipped on write and read back as zeros (dgn_relation_owner tells who has them).
void param_io(dgn_graph *g, float *arena, int kind, int group, int k, float *values, int64_t n, bool write) {
    DGN_REQUIRE(g && values, "null argument");
    CUDA_CHECK(cudaSetDevice(g->device));
    check_arena(g);
    DGN_REQUIRE(group >= 0 && group < g->n_groups, "group %d out of range", group);
    Group &G = g->groups[group];
    DGN_REQUIRE(k >= -1 && k < G.K, "relation index %d out of range for group %d (K = %d)", k, group, G.K);
    const long long per = param_floats_per_relation(g, kind, group);
    const bool encoder = kind == DGN_PARAM_W1 || kind == DGN_PARAM_W2;
    const long long nk = kind == DGN_PARAM_DEC_GLOBAL ? 1 : (k < 0 ? G.K : 1);
    DGN_REQUIRE(n == nk * per, "parameter kind %d group %d k %d: got %lld floats, expected %lld", kind, group, k, (long long)n, nk * per);
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    // zero-padded on the device (W1 pads inside arena_write / arena_read: its panels are packed there anyway)
    const bool padded = kind != DGN_PARAM_W1 && (g->d2u != g->d2 || (kind == DGN_PARAM_W2 && g->d1u != g->d1));
    if (!padded && (!encoder || !G.partitioned)) {
        ParamSpan s = locate_param(g, kind, group, k);
        if (write) arena_write(g, arena, s, values);
        else arena_read(g, arena, s, values);
        return;
    }
    // logical [rows_u, d2u] <-> device [rows_p, 32] of ONE relation's variable
    long long rows_u = 1, rows_p = 1;
    if (kind == DGN_PARAM_W2) rows_u = g->d1u, rows_p = g->d1;
    else if (kind == DGN_PARAM_DEC_GLOBAL || G.decoder == DGN_DEC_BILINEAR) rows_u = g->d2u, rows_p = g->d2;
    std::vector<float> wide(padded ? (size_t)rows_p * g->d2 : 0);
    const int k0 = k < 0 ? 0 : k;
    for (int kk = k0; kk < k0 + nk; ++kk) {
        float *v = values + (size_t)(kk - k0) * per;
        const int l = encoder && G.partitioned ? G.loc_index[kk] : kk;
        if (l < 0) {
            if (!write) memset(v, 0, (size_t)per * sizeof(float));
            continue;
        }
        ParamSpan s = locate_param(g, kind, group, l);
        if (!padded) {
            if (write) arena_write(g, arena, s, v);
            else arena_read(g, arena, s, v);
            continue;
        }
        DGN_REQUIRE(s.count == rows_p * g->d2 && per == rows_u * g->d2u, "internal: padded span of kind %d", kind);
        if (write) {
            std::fill(wide.begin(), wide.end(), 0.f);
            for (long long r = 0; r < rows_u; ++r)
                std::copy(v + r * g->d2u, v + (r + 1) * g->d2u, wide.begin() + r * g->d2);
            arena_write(g, arena, s, wide.data());
        } else {
            arena_read(g, arena, s, wide.data());
            for (long long r = 0; r < rows_u; ++r)
                std::copy(wide.begin() + r * g->d2, wide.begin() + r * g->d2 + g->d2u, v + r * g->d2u);
        }
    }
}
}  // namespace

extern "C" int dgn_params_set(dgn_graph *g, int kind, int group, int k, const float *values, int64_t n) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    param_io(g, g->params, kind, group, k, const_cast<float *>(values), n, true);
    DGN_API_END
}

extern "C" int dgn_params_get(dgn_graph *g, int kind, int group, int k, float *values_out, int64_t n) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    param_io(g, g->params, kind, group, k, values_out, n, false);
    DGN_API_END
}

extern "C" int dgn_grads_get(dgn_graph *g, int kind, int group, int k, float *values_out, int64_t n) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    if (kind == DGN_PARAM_W1 && group >= 0 && group < g->n_groups && g->groups[group].w1_grad_stale)
        DGN_FAIL(DGN_ERR_INVALID,
                 "group %d: the last step fused the Adam update of the layer-1 weights into the kernel that produces their "
                 "gradient, which was never stored; call dgn_keep_gradients(g, 1) before the step (or run it with apply_update = 0)",
                 group);
    param_io(g, g->grads, kind, group, k, values_out, n, false);
    DGN_API_END
}

extern "C" int dgn_keep_gradients(dgn_graph *g, int keep) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    g->keep_grads = keep != 0;
    DGN_API_END
}

extern "C" int dgn_relation_owner(dgn_graph *g, int r, int *owner_out) {
    DGN_API_BEGIN
    DGN_REQUIRE(g && owner_out, "null argument");
    DGN_REQUIRE(r >= 0 && r < g->R, "relation %d out of range", r);
    Group &G = g->groups[g->flat[r].first];
    DGN_REQUIRE(g->world == 1 || g->finalized, "the partition exists after dgn_graph_finalize");
    *owner_out = !G.partitioned ? -1 : G.owner[g->flat[r].second];
    DGN_API_END
}

extern "C" int dgn_comm_handle(dgn_graph *g, void *handle_out) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(handle_out && g->comm, "no exchange arena (one rank, or no partitioned group)");
    CUDA_CHECK(cudaSetDevice(g->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI passes IPC handles as 64 bytes");
    cudaIpcMemHandle_t h;
    CUDA_CHECK(cudaIpcGetMemHandle(&h, g->comm));
    memcpy(handle_out, &h, sizeof(h));
    DGN_API_END
}

extern "C" int dgn_comm_connect(dgn_graph *g, const void *handles) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(handles && g->comm, "no exchange arena (one rank, or no partitioned group)");
    CUDA_CHECK(cudaSetDevice(g->device));
    std::vector<uint32_t *> flags(g->world);
    for (int r = 0; r < g->world; ++r) {
        if (r != g->rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, static_cast<const unsigned char *>(handles) + (size_t)r * sizeof(h), sizeof(h));
            void *p = nullptr;
            CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            g->peer_comm[r] = static_cast<unsigned char *>(p);
        }
        flags[r] = reinterpret_cast<uint32_t *>(g->peer_comm[r]);
    }
    if (g->peer_flags_dev) cudaFree(g->peer_flags_dev);
    CUDA_CHECK(cudaMalloc(&g->peer_flags_dev, flags.size() * sizeof(uint32_t *)));
    CUDA_CHECK(cudaMemcpy(g->peer_flags_dev, flags.data(), flags.size() * sizeof(uint32_t *), cudaMemcpyHostToDevice));
    g->connected = true;
    DGN_API_END
}

extern "C" int dgn_comm_init(dgn_graph *g, int rank, int world) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    DGN_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "rank %d of %d (at most %d ranks)", rank, world, kMaxWorld);
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    g->rank = rank, g->world = world;
    for (auto &G : g->groups) G.dirty = true;
    g->finalized = false;
    if (world > 1) free_arena(g);  // the layout depends on the partition, which needs the relations: see finalize
    DGN_API_END
}

extern "C" int dgn_params_count(dgn_graph *g, int64_t *n_out) {
    DGN_API_BEGIN
    DGN_REQUIRE(g && n_out, "null argument");
    int64_t n = 0;  // floats as the caller counts them (hidden2 columns, not the padded 32)
    for (int gi = 0; gi < g->n_groups; ++gi) {
        Group &G = g->groups[gi];
        n += (int64_t)G.Kl * (param_floats_per_relation(g, DGN_PARAM_W1, gi) + param_floats_per_relation(g, DGN_PARAM_W2, gi));
        n += (int64_t)G.K * param_floats_per_relation(g, DGN_PARAM_DEC_LOCAL, gi);
        if (G.decoder == DGN_DEC_DEDICOM) n += param_floats_per_relation(g, DGN_PARAM_DEC_GLOBAL, gi);
    }
    *n_out = n;
    DGN_API_END
}

extern "C" int dgn_optimizer_reset(dgn_graph *g, float beta1, float beta2, float epsilon) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    check_arena(g);
    CUDA_CHECK(cudaMemset(g->adam_m, 0, g->n_params * sizeof(float)));
    CUDA_CHECK(cudaMemset(g->adam_v, 0, g->n_params * sizeof(float)));
    g->beta1 = beta1, g->beta2 = beta2, g->eps = epsilon;
    g->b1p = beta1, g->b2p = beta2;
    DGN_API_END
}

extern "C" int dgn_encoder_forward(dgn_graph *g, float dropout, uint64_t seed, uint32_t step) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(dropout >= 0.f && dropout < 1.f, "dropout rate %g outside [0, 1)", dropout);
    CUDA_CHECK(cudaSetDevice(g->device));
    const int slot = begin_step(g, dropout, seed, step, 2);
    commit_step(g, slot, sizeof(StepDyn));
    StepDeps deps;
    run_forward(g, dropout, deps);
    join_lanes(g, false);
    DGN_API_END
}

namespace {

// forward + decode + backward + optimizer of one training step, issued on the library's streams (directly, or
// while lane 0 is being captured into a CUDA graph).  Everything that differs between two steps is read by the
// kernels from the per-step block in device memory (StepDyn), so the sequence of launches depends only on
// (dropout on / off and its rate, update or not, gradients kept or not).
void issue_train_step(dgn_graph *g, float dropout, bool apply_update, bool masks2_ready, bool draw_ahead, bool early_loss) {
    cudaStream_t s = g->stream;
    StepDeps deps;
    run_forward(g, dropout, deps, masks2_ready);
    {
        for (int t = 0; t < g->n_types; ++t) consume(g, deps.Z[t], 0);
        PhaseScope ph(g, "decode");
        if (g->n_params > g->dec_off)
            CUDA_CHECK(cudaMemsetAsync(g->grads + g->dec_off, 0, (g->n_params - g->dec_off) * sizeof(float), s));
        launch_decode(g->dyn_dev, s);
        g->launches++;
        if (early_loss) {
            // the loss is final once decode has run: it goes to pinned host memory now, together with the step's serial
            // number; the host polls for that number and returns from dgn_train_step while the backward pass and the
            // optimizer are still running (every later call is ordered behind them on the stream)
            CUDA_CHECK(cudaMemcpyAsync(g->loss_host, g->loss_dev, 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
        }
        if (g->n_types <= kMaxTypes) {
            FixedBatch fb = {};
            for (auto &T : g->types) fb.q[fb.count] = T.dZq, fb.out[fb.count] = T.dZ, fb.n[fb.count] = (size_t)panel_floats(1, T.n), fb.count++;
            launch_fixed_to_float_clear(fb, s);
            g->launches++;
        } else {
            for (auto &T : g->types) {
                launch_fixed_to_float(T.dZq, T.dZ, panel_floats(1, T.n), s);
                g->launches++;
            }
        }
        for (int t = 0; t < g->n_types; ++t) produced(g, deps.dZ[t], 0);
    }
    AdamStep adam;
    adam.alpha = apply_update ? 1.f : 0.f;  // only "is there an update": the coefficients are in the per-step block
    run_backward(g, dropout, deps, adam, draw_ahead);
    if (apply_update) {
        PhaseScope ph(g, "adam");
        // every variable whose update was not fused into the kernel that produced its gradient
        size_t begin = 0;
        auto flush = [&](size_t end) {
            if (end > begin) {
                launch_adam(g->params + begin, g->grads + begin, g->adam_m + begin, g->adam_v + begin, (long long)(end - begin),
                            g->dyn_dev, s);
                g->launches++;
            }
        };
        for (auto &Gq : g->groups)
            if (adam_fused(g, Gq, adam)) {
                flush(Gq.w1_off);
                begin = Gq.w1_off + (size_t)Gq.Kl * Gq.F_j * g->d1;
            }
        flush(g->n_params);
    }
}

}  // namespace

extern "C" int dgn_train_step(dgn_graph *g, int r, const int32_t *batch, int32_t batch_size, const int64_t *negatives,
                              int loss_kind, float margin, float neg_weight, float learning_rate, float dropout,
                              uint64_t seed, uint32_t step, int apply_update, float *loss_out) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(batch && batch_size > 0, "empty batch");
    DGN_REQUIRE(r >= 0 && r < g->R, "batch_edge_type_idx %d out of range (R = %d)", r, g->R);
    DGN_REQUIRE(loss_kind == DGN_LOSS_HINGE || loss_kind == DGN_LOSS_XENT, "unknown loss kind %d", loss_kind);
    DGN_REQUIRE(dropout >= 0.f && dropout < 1.f, "dropout rate %g outside [0, 1)", dropout);
    CUDA_CHECK(cudaSetDevice(g->device));
    Group &G = g->groups[g->flat[r].first];
    const int k = g->flat[r].second;
    DGN_REQUIRE(negatives != nullptr || G.thr[k] != nullptr,
                "relation %d: no negatives given and no degree table set (dgn_sampler_set_degrees)", r);
    for (int b = 0; b < batch_size; ++b) {
        DGN_REQUIRE(batch[2 * b] >= 0 && batch[2 * b] < G.n_i && batch[2 * b + 1] >= 0 && batch[2 * b + 1] < G.n_j,
                    "batch edge %d = (%d, %d) outside %d x %d", b, batch[2 * b], batch[2 * b + 1], G.n_i, G.n_j);
        if (negatives) DGN_REQUIRE(negatives[b] >= 0 && negatives[b] < G.n_i, "negative sample %d = %lld outside [0, %d)", b, (long long)negatives[b], G.n_i);
    }
    cudaStream_t s = g->stream;
    ensure_batch_capacity(g, batch_size);

    // ---- the per-step block: one pinned slot, one host-to-device copy
    const int slot = begin_step(g, dropout, seed, step, 3);
    unsigned char *host = g->step_host[slot];
    StepDyn *h = reinterpret_cast<StepDyn *>(host);
    h->seq = ++g->step_seq;
    if (apply_update) {
        // TF 1.8 ApplyAdam: alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t), float32
        h->alpha = learning_rate * sqrtf(1.f - g->b2p) / (1.f - g->b1p);
        h->omb1 = 1.f - g->beta1, h->omb2 = 1.f - g->beta2, h->eps = g->eps;
    }
    {
        DecodeArgs &a = h->dec;
        a.Zi = g->types[G.i].Z, a.Zj = g->types[G.j].Z, a.dZi = g->types[G.i].dZq, a.dZj = g->types[G.j].dZq;
        a.n_i = G.n_i, a.n_j = G.n_j;
        a.batch = g->batch_dev, a.neg_in = negatives ? g->neg_dev : nullptr, a.neg_out = g->neg_out;
        a.thr = G.thr[k], a.n_thr = G.thr_n[k];
        a.B = batch_size, a.decoder = G.decoder, a.loss_kind = loss_kind, a.margin = margin, a.neg_weight = neg_weight;
        a.glb = G.decoder == DGN_DEC_DEDICOM ? g->params + G.glb_off : nullptr;
        a.loc = G.loc_per_rel ? g->params + G.loc_off + (size_t)k * G.loc_per_rel : nullptr;
        a.g_glb = G.decoder == DGN_DEC_DEDICOM ? g->grads + G.glb_off : nullptr;
        a.g_loc = G.loc_per_rel ? g->grads + G.loc_off + (size_t)k * G.loc_per_rel : nullptr;
        a.pos_out = g->pos_out, a.neg_score_out = g->negs_out, a.loss_out = g->loss_dev;
        a.seed_lo = (uint32_t)(seed & 0xffffffffu), a.seed_hi = (uint32_t)(seed >> 32), a.step = step, a.relation = (uint32_t)r;
        a.scratch = g->decode_scratch, a.ticket = g->decode_ticket;
    }
    if (negatives) memcpy(host + g->step_neg_off, negatives, (size_t)batch_size * sizeof(long long));
    memcpy(host + g->step_batch_off, batch, (size_t)batch_size * 2 * sizeof(int));
    commit_step(g, slot, g->step_batch_off + (size_t)batch_size * 2 * sizeof(int));

    if (!g->dzq_clean)  // normally left clean by the previous step's conversion kernel
        for (auto &T : g->types) CUDA_CHECK(cudaMemsetAsync(T.dZq, 0, panel_floats(1, T.n) * sizeof(long long), s));
    AdamStep mode;
    mode.alpha = apply_update ? 1.f : 0.f;
    for (auto &Gq : g->groups) Gq.w1_grad_stale = !Gq.gen_feat && adam_fused(g, Gq, mode);
    // layer-2 keep words drawn ahead by the previous step?
    const bool ahead_on = g->mask_ahead && g->two_lanes && dropout > 0.f;
    const uint32_t thr = dropout_threshold(dropout);
    const bool masks2_ready = ahead_on && g->ahead_valid && g->ahead_seed == seed && g->ahead_step == step && g->ahead_thr == thr;
    if (masks2_ready) g->mask_cur ^= 1;
    for (auto &Gq : g->groups) Gq.mask2 = Gq.mask2buf[g->mask_cur];

    // ---- the step itself: a CUDA graph replay when this configuration has been seen before
    const bool early = loss_out != nullptr && g->early_loss && !g->timing;
    bool done = false;
    if (g->use_graphs && !g->timing) {
        int parity = 0;
        for (auto &Gq : g->groups)
            if (Gq.partitioned)
                for (int x = 0; x < 3; ++x) parity |= (int)(Gq.xch[x].stamp & 1u) << Gq.xch[x].id;
        uint32_t rate_bits;
        memcpy(&rate_bits, &dropout, sizeof(rate_bits));
        const int mode_bits = (apply_update ? 1 : 0) | (masks2_ready ? 2 : 0) | (ahead_on ? 4 : 0) | (g->mask_cur << 3) | (early ? 16 : 0);
        dgn_graph::StepGraph &sg = g->step_graphs[std::make_tuple(rate_bits, mode_bits, g->keep_grads ? 1 : 0, parity)];
        if (sg.exec == nullptr && sg.seen++ >= 1) {
            // second occurrence: capture (the first ran directly: lazy module loading and function attributes are done)
            const long long before = g->launches;
            cudaGraph_t graph = nullptr;
            CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            g->capturing = true;
            try {
                issue_train_step(g, dropout, apply_update != 0, masks2_ready, ahead_on, early);
            } catch (...) {
                g->capturing = false;
                cudaStreamEndCapture(s, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            g->capturing = false;
            CUDA_CHECK(cudaStreamEndCapture(s, &graph));
            sg.launches = g->launches - before;
            g->launches = before;
            cudaError_t e = cudaGraphInstantiate(&sg.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) {
                sg.exec = nullptr;
                DGN_FAIL(DGN_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
            }
        }
        if (sg.exec != nullptr) {
            CUDA_CHECK(cudaGraphLaunch(sg.exec, s));
            g->launches += sg.launches;
            g->graph_replays++;
            done = true;
        }
    }
    if (!done) issue_train_step(g, dropout, apply_update != 0, masks2_ready, ahead_on, early);
    g->ahead_valid = ahead_on;
    g->ahead_seed = seed, g->ahead_step = step + 1, g->ahead_thr = thr;
    g->dzq_clean = g->n_types <= kMaxTypes;
    g->last_B = batch_size;

    if (apply_update) {
        g->b1p *= g->beta1;
        g->b2p *= g->beta2;
    }
    if (loss_out && early) {
        // session.run([opt_op, cost]) returns when the COST is known; the update it queued is complete before any
        // later call can observe a variable (same stream)
        const volatile uint32_t *seq_host = reinterpret_cast<const volatile uint32_t *>(g->loss_host + 1);
        for (long long spin = 0; *seq_host != g->step_seq; ++spin) {
            if ((spin & 0xfff) == 0xfff) {  // every 4096 polls: did the stream stop (error, or done without our number)?
                cudaError_t e = cudaStreamQuery(s);
                if (e == cudaSuccess && *seq_host != g->step_seq) DGN_FAIL(DGN_ERR_CUDA, "the step finished without delivering its loss");
                if (e != cudaSuccess && e != cudaErrorNotReady) DGN_FAIL(DGN_ERR_CUDA, "training step failed: %s", cudaGetErrorString(e));
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        *loss_out = *const_cast<const volatile float *>(g->loss_host);
    } else if (loss_out) {
        CUDA_CHECK(cudaMemcpyAsync(g->loss_host, g->loss_dev, sizeof(float), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        check_exchange(g);
        *loss_out = *g->loss_host;
    }
    DGN_API_END
}

extern "C" int dgn_last_batch_outputs(dgn_graph *g, float *pos_out, float *neg_out, int64_t *neg_samples_out,
                                      int32_t batch_size) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(batch_size == g->last_B && batch_size > 0, "batch size %d does not match the last step (%d)", batch_size, g->last_B);
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    if (pos_out) CUDA_CHECK(cudaMemcpy(pos_out, g->pos_out, (size_t)batch_size * sizeof(float), cudaMemcpyDeviceToHost));
    if (neg_out) CUDA_CHECK(cudaMemcpy(neg_out, g->negs_out, (size_t)batch_size * sizeof(float), cudaMemcpyDeviceToHost));
    if (neg_samples_out)
        CUDA_CHECK(cudaMemcpy(neg_samples_out, g->neg_out, (size_t)batch_size * sizeof(long long), cudaMemcpyDeviceToHost));
    DGN_API_END
}

extern "C" int dgn_predict_all_pairs(dgn_graph *g, int r, float *out) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(out, "null output");
    CUDA_CHECK(cudaSetDevice(g->device));
    PredictArgs a = predict_args(g, r, 1);
    const size_t n = (size_t)a.n_i * a.n_j;
    float *tmp = dev_alloc<float>(n);
    a.out = tmp;
    try {
        run_predict(g, a);
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
        CUDA_CHECK(cudaMemcpy(out, tmp, n * sizeof(float), cudaMemcpyDeviceToHost));
    } catch (...) {
        cudaFree(tmp);
        throw;
    }
    cudaFree(tmp);
    DGN_API_END
}

extern "C" int dgn_predict_relations_dev(dgn_graph *g, int r0, int count, float *out_dev) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(out_dev, "null output");
    CUDA_CHECK(cudaSetDevice(g->device));
    PredictArgs a = predict_args(g, r0, count);
    a.out = out_dev;
    PhaseScope ph(g, "predict");
    run_predict(g, a);
    DGN_API_END
}

extern "C" int dgn_predict_edges(dgn_graph *g, int r, const int32_t *edges, int32_t n_edges, int apply_sigmoid, float *out) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(n_edges >= 0 && (n_edges == 0 || (edges && out)), "null argument");
    if (n_edges == 0) return DGN_OK;
    CUDA_CHECK(cudaSetDevice(g->device));
    PredictArgs a = predict_args(g, r, 1);
    for (int e = 0; e < n_edges; ++e)
        DGN_REQUIRE(edges[2 * e] >= 0 && edges[2 * e] < a.n_i && edges[2 * e + 1] >= 0 && edges[2 * e + 1] < a.n_j,
                    "edge %d = (%d, %d) outside %d x %d", e, edges[2 * e], edges[2 * e + 1], a.n_i, a.n_j);
    int *e_dev = dev_alloc<int>((size_t)n_edges * 2);
    float *o_dev = dev_alloc<float>(n_edges);
    try {
        CUDA_CHECK(cudaMemcpyAsync(e_dev, edges, (size_t)n_edges * 2 * sizeof(int), cudaMemcpyHostToDevice, g->stream));
        launch_predict_edges(a, e_dev, n_edges, apply_sigmoid, o_dev, g->stream);
        g->launches++;
        CUDA_CHECK(cudaMemcpyAsync(out, o_dev, (size_t)n_edges * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
    } catch (...) {
        cudaFree(e_dev);
        cudaFree(o_dev);
        throw;
    }
    cudaFree(e_dev);
    cudaFree(o_dev);
    DGN_API_END
}

// evaluateAll in one call: edges of many relations of one group, optional pooled AUROC / AUPRC on the device
extern "C" int dgn_evaluate_edges(dgn_graph *g, int group, int64_t n_edges, const int32_t *rel_k, const int32_t *edges,
                                  const uint8_t *labels, int apply_sigmoid, float *scores_out, double *auroc_out,
                                  double *auprc_out) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(group >= 0 && group < (int)g->groups.size(), "group %d out of range", group);
    DGN_REQUIRE(n_edges >= 0 && (n_edges == 0 || (rel_k && edges)), "null argument");
    const bool want_auc = auroc_out || auprc_out;
    DGN_REQUIRE(!want_auc || labels || n_edges == 0, "labels are needed for AUROC / AUPRC");
    if (auroc_out) *auroc_out = nan("");
    if (auprc_out) *auprc_out = nan("");
    if (n_edges == 0) return DGN_OK;
    CUDA_CHECK(cudaSetDevice(g->device));
    Group &G = g->groups[group];
    int r0 = -1;
    for (int r = 0; r < g->R && r0 < 0; ++r)
        if (g->flat[r].first == group) r0 = r;
    PredictArgs a = predict_args(g, r0, 1);
    for (int64_t e = 0; e < n_edges; ++e) {
        DGN_REQUIRE(rel_k[e] >= 0 && rel_k[e] < G.K, "edge %lld: relation %d outside group of %d", (long long)e, rel_k[e], G.K);
        DGN_REQUIRE(edges[2 * e] >= 0 && edges[2 * e] < a.n_i && edges[2 * e + 1] >= 0 && edges[2 * e + 1] < a.n_j,
                    "edge %lld = (%d, %d) outside %d x %d", (long long)e, edges[2 * e], edges[2 * e + 1], a.n_i, a.n_j);
    }
    int *k_dev = nullptr, *e_dev = nullptr;
    float *s_dev = nullptr, *ss_dev = nullptr;
    unsigned char *l_dev = nullptr, *sl_dev = nullptr;
    void *tmp = nullptr;
    double *res_dev = nullptr;
    auto release = [&]() {
        cudaFree(k_dev), cudaFree(e_dev), cudaFree(s_dev), cudaFree(ss_dev), cudaFree(l_dev), cudaFree(sl_dev);
        cudaFree(tmp), cudaFree(res_dev);
    };
    try {
        PhaseScope ph(g, "evaluate");
        k_dev = dev_alloc<int>((size_t)n_edges), e_dev = dev_alloc<int>((size_t)n_edges * 2);
        s_dev = dev_alloc<float>((size_t)n_edges);
        CUDA_CHECK(cudaMemcpyAsync(k_dev, rel_k, (size_t)n_edges * sizeof(int), cudaMemcpyHostToDevice, g->stream));
        CUDA_CHECK(cudaMemcpyAsync(e_dev, edges, (size_t)n_edges * 2 * sizeof(int), cudaMemcpyHostToDevice, g->stream));
        launch_predict_edges_multi(a, k_dev, e_dev, n_edges, apply_sigmoid, s_dev, g->stream);
        g->launches++;
        if (scores_out)
            CUDA_CHECK(cudaMemcpyAsync(scores_out, s_dev, (size_t)n_edges * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
        double res[4] = {0, 0, 0, 0};
        if (want_auc) {
            const size_t tmp_bytes = auc_sort_bytes(n_edges);
            l_dev = dev_alloc<unsigned char>((size_t)n_edges), sl_dev = dev_alloc<unsigned char>((size_t)n_edges);
            ss_dev = dev_alloc<float>((size_t)n_edges), res_dev = dev_alloc<double>(4);
            tmp = dev_alloc<unsigned char>(tmp_bytes ? tmp_bytes : 1);
            CUDA_CHECK(cudaMemcpyAsync(l_dev, labels, (size_t)n_edges, cudaMemcpyHostToDevice, g->stream));
            launch_auc(s_dev, l_dev, n_edges, ss_dev, sl_dev, tmp, tmp_bytes, res_dev, g->stream);
            g->launches += 2;
            CUDA_CHECK(cudaMemcpyAsync(res, res_dev, sizeof(res), cudaMemcpyDeviceToHost, g->stream));
        }
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
        if (auroc_out) *auroc_out = res[0];
        if (auprc_out) *auprc_out = res[1];
    } catch (...) {
        release();
        throw;
    }
    release();
    DGN_API_END
}

// GreedyActiveLearner._getRankedPossibilities: candidates of ONE relation scored on the device and ranked there
extern "C" int dgn_rank_edges(dgn_graph *g, int r, const int32_t *edges, int64_t n_edges, int apply_sigmoid, int64_t top,
                              int32_t *order_out, float *scores_out) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(n_edges >= 0 && n_edges < (int64_t)INT32_MAX && top >= 0 && top <= n_edges, "bad sizes: %lld candidates, top %lld",
                (long long)n_edges, (long long)top);
    DGN_REQUIRE(n_edges == 0 || (edges && (top == 0 || order_out)), "null argument");
    if (n_edges == 0 || top == 0) return DGN_OK;
    CUDA_CHECK(cudaSetDevice(g->device));
    PredictArgs a = predict_args(g, r, 1);
    for (int64_t e = 0; e < n_edges; ++e)
        DGN_REQUIRE(edges[2 * e] >= 0 && edges[2 * e] < a.n_i && edges[2 * e + 1] >= 0 && edges[2 * e + 1] < a.n_j,
                    "candidate %lld = (%d, %d) outside %d x %d", (long long)e, edges[2 * e], edges[2 * e + 1], a.n_i, a.n_j);
    int *e_dev = nullptr, *idx = nullptr, *order = nullptr;
    float *s_dev = nullptr, *ss_dev = nullptr;
    void *tmp = nullptr;
    auto release = [&]() { cudaFree(e_dev), cudaFree(idx), cudaFree(order), cudaFree(s_dev), cudaFree(ss_dev), cudaFree(tmp); };
    try {
        PhaseScope ph(g, "rank");
        e_dev = dev_alloc<int>((size_t)n_edges * 2);
        idx = dev_alloc<int>((size_t)n_edges), order = dev_alloc<int>((size_t)n_edges);
        s_dev = dev_alloc<float>((size_t)n_edges), ss_dev = dev_alloc<float>((size_t)n_edges);
        const size_t tmp_bytes = rank_sort_bytes(n_edges);
        tmp = dev_alloc<unsigned char>(tmp_bytes ? tmp_bytes : 1);
        CUDA_CHECK(cudaMemcpyAsync(e_dev, edges, (size_t)n_edges * 2 * sizeof(int), cudaMemcpyHostToDevice, g->stream));
        launch_predict_edges(a, e_dev, (int)n_edges, apply_sigmoid, s_dev, g->stream);
        launch_rank(s_dev, n_edges, idx, ss_dev, order, tmp, tmp_bytes, g->stream);
        g->launches += 3;
        CUDA_CHECK(cudaMemcpyAsync(order_out, order, (size_t)top * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
        if (scores_out) CUDA_CHECK(cudaMemcpyAsync(scores_out, ss_dev, (size_t)top * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
    } catch (...) {
        release();
        throw;
    }
    release();
    DGN_API_END
}

extern "C" int dgn_tensor_get(dgn_graph *g, int which, int index, float *out, int64_t n) {
    DGN_API_BEGIN
    check_finalized(g);
    DGN_REQUIRE(out, "null output");
    CUDA_CHECK(cudaSetDevice(g->device));
    const float *src = nullptr;
    long long rows = 0;
    int P = 1;
    switch (which) {
        case DGN_TENSOR_HIDDEN1:
        case DGN_TENSOR_EMBEDDINGS:
        case DGN_TENSOR_GRAD_EMBEDDINGS: {
            DGN_REQUIRE(index >= 0 && index < g->n_types, "node type %d out of range", index);
            NodeType &T = g->types[index];
            rows = T.n;
            if (which == DGN_TENSOR_HIDDEN1) src = T.H, P = g->P1;
            else src = which == DGN_TENSOR_EMBEDDINGS ? T.Z : T.dZ;
        } break;
        case DGN_TENSOR_LAYER1_GROUP:
        case DGN_TENSOR_LAYER2_GROUP: {
            DGN_REQUIRE(index >= 0 && index < g->n_groups, "group %d out of range", index);
            Group &G = g->groups[index];
            rows = G.n_i;
            if (which == DGN_TENSOR_LAYER1_GROUP) src = G.Y1, P = g->P1;
            else src = G.Y2;
        } break;
        default: DGN_FAIL(DGN_ERR_INVALID, "unknown tensor id %d", which);
    }
    const bool layer1 = which == DGN_TENSOR_HIDDEN1 || which == DGN_TENSOR_LAYER1_GROUP;
    const long long cols = layer1 ? g->d1u : g->d2u;  // what the caller sees; the device holds 32 * P columns
    DGN_REQUIRE(n == rows * cols, "tensor %d[%d]: got room for %lld floats, need %lld", which, index, (long long)n, rows * cols);
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    std::vector<float> tmp((size_t)rows * 32 * P);
    CUDA_CHECK(cudaMemcpy(tmp.data(), src, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
    unpack_panels(tmp.data(), out, rows, P, (int)cols);  // the padding columns (zeros) stay behind
    DGN_API_END
}

extern "C" int dgn_tensor_set(dgn_graph *g, int which, int index, const float *values, int64_t n) {
    DGN_API_BEGIN
    DGN_REQUIRE(g && values, "null argument");
    DGN_REQUIRE(which == DGN_TENSOR_EMBEDDINGS, "only the embeddings can be set (tensor id %d)", which);
    DGN_REQUIRE(index >= 0 && index < g->n_types, "node type %d out of range", index);
    NodeType &T = g->types[index];
    DGN_REQUIRE(n == (int64_t)T.n * g->d2u, "embeddings[%d]: got %lld floats, need %lld", index, (long long)n, (long long)T.n * g->d2u);
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    if (g->d2u == g->d2) {  // one panel = row-major
        CUDA_CHECK(cudaMemcpy(T.Z, values, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    } else {
        std::vector<float> wide((size_t)T.n * g->d2, 0.f);
        for (long long r = 0; r < T.n; ++r) std::copy(values + r * g->d2u, values + (r + 1) * g->d2u, wide.begin() + r * g->d2);
        CUDA_CHECK(cudaMemcpy(T.Z, wide.data(), wide.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    DGN_API_END
}

extern "C" int dgn_relation_matrices(dgn_graph *g, int r, float *glb_out, float *loc_out) {
    DGN_API_BEGIN
    DGN_REQUIRE(g && glb_out && loc_out, "null argument");
    DGN_REQUIRE(r >= 0 && r < g->R, "relation %d out of range", r);
    CUDA_CHECK(cudaSetDevice(g->device));
    Group &G = g->groups[g->flat[r].first];
    const int k = g->flat[r].second;
    const size_t n = (size_t)g->d2 * g->d2;
    float *tmp = dev_alloc<float>(2 * n);
    try {
        launch_relation_matrices(G.decoder, G.decoder == DGN_DEC_DEDICOM ? g->params + G.glb_off : nullptr,
                                 G.loc_per_rel ? g->params + G.loc_off + (size_t)k * G.loc_per_rel : nullptr, tmp, tmp + n, g->stream);
        g->launches++;
        CUDA_CHECK(cudaStreamSynchronize(g->stream));
        std::vector<float> both(2 * n);
        CUDA_CHECK(cudaMemcpy(both.data(), tmp, 2 * n * sizeof(float), cudaMemcpyDeviceToHost));
        for (int m = 0; m < 2; ++m)  // the leading [hidden2, hidden2] block of the device's [32, 32]
            for (int r = 0; r < g->d2u; ++r)
                std::copy(both.begin() + m * n + (size_t)r * g->d2, both.begin() + m * n + (size_t)r * g->d2 + g->d2u,
                          (m ? loc_out : glb_out) + (size_t)r * g->d2u);
    } catch (...) {
        cudaFree(tmp);
        throw;
    }
    cudaFree(tmp);
    DGN_API_END
}

extern "C" int dgn_sync(dgn_graph *g) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    check_exchange(g);
    DGN_API_END
}

extern "C" int dgn_timing_enable(dgn_graph *g, int enable) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    g->timing = enable != 0;
    DGN_API_END
}

extern "C" int dgn_timing_reset(dgn_graph *g) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    for (auto &p : g->phases) {
        cudaEventDestroy(p.start);
        cudaEventDestroy(p.stop);
    }
    g->phases.clear();
    g->launches = 0;
    DGN_API_END
}

extern "C" int dgn_timing_get(dgn_graph *g, const char *name, double *ms_out, int64_t *count_out) {
    DGN_API_BEGIN
    DGN_REQUIRE(g && name && ms_out, "null argument");
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    double total = 0.0;
    int64_t count = 0;
    const size_t len = strlen(name);
    for (auto &p : g->phases)
        if (p.name.compare(0, len, name) == 0) {
            float ms = 0.f;
            CUDA_CHECK(cudaEventElapsedTime(&ms, p.start, p.stop));
            total += ms;
            ++count;
        }
    *ms_out = total;
    if (count_out) *count_out = count;
    DGN_API_END
}

// phase `index` of the recorded list: name, stream lane and start / stop in ms after the first recorded phase began
extern "C" int dgn_timeline_get(dgn_graph *g, int index, char *name_out, int name_cap, int *lane_out, double *start_ms_out,
                                double *stop_ms_out) {
    DGN_API_BEGIN
    DGN_REQUIRE(g && name_out && name_cap > 0 && lane_out && start_ms_out && stop_ms_out, "null argument");
    if (index < 0 || index >= (int)g->phases.size()) return DGN_ERR_INVALID;
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaDeviceSynchronize());
    const Phase &p = g->phases[index];
    float a = 0.f, b = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&a, g->phases[0].start, p.start));
    CUDA_CHECK(cudaEventElapsedTime(&b, g->phases[0].start, p.stop));
    snprintf(name_out, (size_t)name_cap, "%s", p.name.c_str());
    *lane_out = p.lane, *start_ms_out = a, *stop_ms_out = b;
    DGN_API_END
}

extern "C" int dgn_counters(dgn_graph *g, int64_t *graph_replays_out, int64_t *groups_rebuilt_out) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    if (graph_replays_out) *graph_replays_out = g->graph_replays;
    if (groups_rebuilt_out) *groups_rebuilt_out = g->groups_rebuilt;
    DGN_API_END
}

extern "C" int dgn_launch_count(dgn_graph *g, int64_t *launches_out) {
    DGN_API_BEGIN
    DGN_REQUIRE(g && launches_out, "null argument");
    *launches_out = g->launches;
    DGN_API_END
}

// whole-region timer on the library's own stream (bench.py: CUDA events around K steps)
extern "C" int dgn_timer_start(dgn_graph *g) {
    DGN_API_BEGIN
    DGN_REQUIRE(g, "null graph");
    CUDA_CHECK(cudaSetDevice(g->device));
    if (!g->timer_start) {
        CUDA_CHECK(cudaEventCreate(&g->timer_start));
        CUDA_CHECK(cudaEventCreate(&g->timer_stop));
    }
    CUDA_CHECK(cudaEventRecord(g->timer_start, g->stream));
    DGN_API_END
}

extern "C" int dgn_timer_stop(dgn_graph *g, double *ms_out) {
    DGN_API_BEGIN
    DGN_REQUIRE(g && ms_out && g->timer_start, "timer was not started");
    CUDA_CHECK(cudaSetDevice(g->device));
    CUDA_CHECK(cudaEventRecord(g->timer_stop, g->stream));
    CUDA_CHECK(cudaEventSynchronize(g->timer_stop));
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, g->timer_start, g->timer_stop));
    *ms_out = ms;
    DGN_API_END
}

extern "C" int dgn_memory_bytes(dgn_graph *g, int64_t *free_out, int64_t *total_out) {
    DGN_API_BEGIN
    DGN_REQUIRE(g && free_out && total_out, "null argument");
    CUDA_CHECK(cudaSetDevice(g->device));
    size_t f = 0, t = 0;
    CUDA_CHECK(cudaMemGetInfo(&f, &t));
    *free_out = (int64_t)f, *total_out = (int64_t)t;
    DGN_API_END
}
