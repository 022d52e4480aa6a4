#!/usr/bin/env python
"""Headline benchmark: full-batch training steps on the polypharmacy-shape graph (BASELINE.json
config #3) -- 19 085 proteins, 645 drugs, 964 side-effect types (1932 relation matrices with
transposed twins), hidden 64/32, batch 512, dropout 0.1, hinge loss, TF1 Adam.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one reference ``session.run([opt_op, cost, batch_edge_type_idx])``
(``main/Trainer/DecagonTrainer.py:94-100``): encoder forward over the whole graph, 512-edge
decode of one relation, full backward, Adam on all ~87 M parameters.  Prints ONE JSON line
(contract in the task statement).  ``oracle/`` is touched only for the CPU baseline legs.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 3  # dropout / negative-sampler streams of the throughput runs
HYPER = dict(hidden1=64, hidden2=32, batch_size=512, dropout=0.1, lr=1e-3, margin=0.1, val_test_size=0.05)
PLACEHOLDERS = {k: k for k in ['batch', 'batch_edge_type_idx', 'batch_row_edge_type', 'batch_col_edge_type', 'dropout']}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------- workload
def build_workload(config, scale, decoder=None):
    from decagon_b200 import datasets
    from decagon_b200.deep.minibatch import EdgeMinibatchIterator
    t0 = time.time()
    if config == 'toy':
        # config #1 (main.py's decoders) or config #2: every group on the same decoder kind
        inputs = datasets.toy_graph({g: decoder for g in datasets.DEFAULT_DECODERS} if decoder else None)
    else:
        inputs = datasets.polypharmacy_graph(scale=scale)
    np.random.seed(0)
    it = EdgeMinibatchIterator(inputs.adj_mats, inputs.feat, inputs.edge_types, {}, batch_size=HYPER['batch_size'],
                               val_test_size=HYPER['val_test_size'])
    log('workload built in %.1fs' % (time.time() - t0))
    return inputs, it


def steps_per_epoch(it):
    """Length of one reference epoch (``while not iterator.end()``, DecagonTrainer.py:54):
    4 * (sum of full batches of the side-effect relations and of the PPI twin) + O(1), SURVEY.md a3."""
    B = it.batch_size
    free = sum(len(it.train_edges[1, 1][k]) // B for k in range(it.edge_types[1, 1]))
    if it.edge_types[0, 0] > 1:
        free += len(it.train_edges[0, 0][1]) // B
    return 4 * free + 4


def glorot_params(inputs, hidden1, hidden2, seed=1):
    rng = np.random.RandomState(seed)

    def glorot(shape, fan_in, fan_out):
        a = np.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-a, a, size=shape).astype(np.float32)

    p = {'W1': {}, 'W2': {}, 'R': {}, 'D': {}}
    for g, K in inputs.edge_types.items():
        F = inputs.num_feat[g[1]]
        p['W1'][g] = glorot((K, F, hidden1), F, hidden1)
        p['W2'][g] = glorot((K, hidden1, hidden2), hidden1, hidden2)
        kind = inputs.edge_type2decoder[g]
        if kind == 'dedicom':
            p['R'][g] = glorot((hidden2, hidden2), hidden2, hidden2)
        if kind in ('dedicom', 'distmult'):
            p['D'][g] = glorot((K, hidden2), hidden2, 1)
        elif kind == 'bilinear':
            p['D'][g] = glorot((K, hidden2, hidden2), hidden2, hidden2)
    return p


def next_batch(it):
    fd = it.next_minibatch_feed_dict(PLACEHOLDERS)
    return int(fd['batch_edge_type_idx']), np.ascontiguousarray(fd['batch'], dtype=np.int32)


def algorithmic_bytes(inputs, it, hidden1, hidden2, is_local=None):
    """fp32 values, int32 indices, CSR, dense operand counted once per relation,
    no cache credit.  Returns {phase: bytes per launch}.  is_local(g, k): the relations THIS rank's launch
    processes (multi-GPU: the bytes of one rank's kernel, not of the whole graph)."""
    out = {}
    for gi, (g, K_all) in enumerate(inputs.edge_types.items()):
        n_i, n_j = inputs.n_nodes[g[0]], inputs.n_nodes[g[1]]
        mine = [k for k in range(K_all) if is_local is None or is_local(g, k)]
        K = len(mine)
        nnz = sum(len(it.adj_train[g][k][1]) for k in mine)
        csr = nnz * 8 + K * (n_i + 1) * 4
        csr_t = nnz * 8 + K * (n_j + 1) * 4
        out['spmm_fwd1/g%d' % gi] = csr + K * n_j * hidden1 * 4 + n_i * hidden1 * 4
        out['spmm_fwd2/g%d' % gi] = csr + K * hidden1 * hidden2 * 4 + n_j * hidden1 * 4 + n_i * hidden2 * 4
        out['spmm_bwd2/g%d' % gi] = csr_t + n_i * hidden2 * 4 + K * hidden1 * hidden2 * 4 + n_j * hidden1 * 4
        out['spmm_bwd1/g%d' % gi] = csr_t + n_i * hidden1 * 4 + K * n_j * hidden1 * 4
    return out


def load_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the newest committed
    ncu --set full capture (profiles/rNN_traffic.json; one GPU, polypharmacy shape)."""
    for name in ('r02_traffic.json', 'r01_traffic.json'):
        try:
            return json.load(open(os.path.join(ROOT, 'profiles', name)))
        except Exception:
            continue
    return {}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        good = [s for s in self.samples if len(s) == 6 and s[0].isdigit()]
        if not good:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] == 'Active' for s in good)]
        return {'sm_mhz': float(np.median([int(s[0]) for s in good])), 'sm_max_mhz': float(good[0][1]), 'reasons': reasons,
                'samples': len(good)}


# ----------------------------------------------------------------------------- CPU legs
def cpu_port_steps(inputs, it, params, n_steps, n_warmup, budget_s=150.0):
    """The reference-style torch-CPU port (oracle/torch_ref.py): FULL train steps (all relation
    matrices, forward + backward + Adam), float32, all host threads.  A full step takes ~10 s at
    the polypharmacy shape, so the number of timed steps is bounded by a time budget:
    n = min(n_steps, budget / warm-up step time).  Returns (seconds per step, n timed, text)."""
    import torch
    from oracle import decagon_oracle as O
    from oracle.torch_ref import TorchDecagon
    torch.set_num_threads(os.cpu_count())
    graph = O.Graph.from_iterator(it, inputs.edge_type2decoder, HYPER['hidden1'], HYPER['hidden2'], dtype=np.float32)
    model = TorchDecagon(graph, params, torch.float32, lr=HYPER['lr'])
    rng = np.random.RandomState(2)
    masks = None

    def one(step):
        g = [(0, 0), (0, 1), (1, 0), (1, 1)][step % 4]
        edges = it.train_edges[g][0]
        batch = edges[rng.randint(0, len(edges), HYPER['batch_size'])]
        negs = rng.randint(0, graph.n_nodes[g[0]], HYPER['batch_size'])
        t0 = time.perf_counter()
        model.train_step(g, 0, batch, negs, rate=0.0, masks=masks, margin=HYPER['margin'])
        return time.perf_counter() - t0

    warm = [one(s) for s in range(n_warmup)]
    n = int(max(1, min(n_steps, budget_s / max(warm[-1], 1e-9))))
    times = [one(n_warmup + s) for s in range(n)]
    text = ('%d full train steps (all %d relation matrices; %d requested, bounded by a %.0f s budget) after %d warm-up, '
            'torch-CPU float32 port of the TF graph (oracle/torch_ref.py), dropout masks not drawn'
            % (n, sum(inputs.edge_types.values()), n_steps, budget_s, n_warmup))
    return float(np.mean(times)), n, text


# ----------------------------------------------------------------------------- main legs
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    inputs, it = build_workload(args.config, args.scale, args.decoder)
    params = glorot_params(inputs, HYPER['hidden1'], HYPER['hidden2'])
    n_warm = max(args.warmup, 1)
    sec, n_timed, sample = cpu_port_steps(inputs, it, params, args.steps, n_warm, args.reference_budget)
    spe = steps_per_epoch(it)
    value = 1.0 / sec / spe
    line = {
        'impl': 'reference', 'metric': 'train_epochs_per_s', 'value': value, 'unit': 'epochs/s', 'n_gpus': args.gpus,
        'steps': n_timed, 'warmup': n_warm, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'steps_per_s': 1.0 / sec, 'steps_per_epoch': spe,
        'config': workload_config(args, inputs),
        'cpu_baseline': {'value': value, 'unit': 'epochs/s', 'cores': os.cpu_count(), 'kind': 'port',
                         'sample': sample, 'steps_per_s': 1.0 / sec},
        'e2e': {'value': value, 'unit': 'epochs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, inputs):
    toy = 'toy (config #1)' if not args.decoder else 'toy, every group %s (config #2)' % args.decoder
    return {'workload': ('polypharmacy-shape synthetic (BASELINE config #%d)' % (3 if args.scale == 1 else 5)) if args.config == 'poly' else toy,
            'n_proteins': inputs.n_nodes[0], 'n_drugs': inputs.n_nodes[1],
            'relation_matrices': sum(inputs.edge_types.values()), 'hidden': [HYPER['hidden1'], HYPER['hidden2']],
            'batch': HYPER['batch_size'], 'dropout': HYPER['dropout'], 'loss': 'hinge', 'optimizer': 'adam(tf1)',
            'scale': args.scale,
            'l2': ('per-step working set (>4 GB) exceeds the 126 MB L2; no explicit flush' if args.config == 'poly' else
                   'working set ~1 MB, L2-resident by nature: latency-bound, steps/s and parity only, no roofline claim')}


def construct_placeholders(edge_types):
    """DecagonDataSet._getPlaceholdersDict (DecagonDataSet.py:84-120) with tf -> decagon_b200.tf_compat."""
    from decagon_b200 import tf_compat as tf
    ph = {
        'batch': tf.placeholder(tf.int32, name='batch'),
        'batch_edge_type_idx': tf.placeholder(tf.int32, shape=(), name='batch_edge_type_idx'),
        'batch_row_edge_type': tf.placeholder(tf.int32, shape=(), name='batch_row_edge_type'),
        'batch_col_edge_type': tf.placeholder(tf.int32, shape=(), name='batch_col_edge_type'),
        'degrees': tf.placeholder(tf.int32),
        'dropout': tf.placeholder_with_default(0., shape=()),
    }
    ph.update({'adj_mats_%d,%d,%d' % (i, j, k): tf.sparse_placeholder(tf.float32)
               for i, j in edge_types for k in range(edge_types[i, j])})
    ph.update({'feat_%d' % i: tf.sparse_placeholder(tf.float32) for i, _ in edge_types})
    return ph


def build_trainable(inputs, it):
    """DecagonTrainableBuilder.build (DecagonTrainableBuilder.py:56-118) against the drop-in classes."""
    from decagon_b200 import tf_compat as tf
    from decagon_b200.deep import inits
    from decagon_b200.deep.model import DecagonModel
    from decagon_b200.deep.optimizer import DecagonOptimizer
    tf.FLAGS.hidden1, tf.FLAGS.hidden2, tf.FLAGS.learning_rate = HYPER['hidden1'], HYPER['hidden2'], HYPER['lr']
    placeholders = construct_placeholders(inputs.edge_types)
    inits.set_seed(1)
    model = DecagonModel(placeholders, inputs.num_feat, inputs.nonzero_feat, inputs.edge_types, inputs.edge_type2decoder)
    with tf.name_scope('optimizer'):
        opt = DecagonOptimizer(model.embeddings, model.latent_inters, model.latent_varies, inputs.degrees,
                               inputs.edge_types, inputs.edge_type2dim, placeholders, margin=HYPER['margin'],
                               neg_sample_weights=1., batch_size=HYPER['batch_size'])
    return placeholders, model, opt


def model_params(model):
    """The model's initial variables as the engine's parameter dict."""
    p = {'W1': {}, 'W2': {}, 'R': {}, 'D': {}}
    for g, K in model.edge_types.items():
        p['W1'][g] = np.stack([model.layer1[g].vars['weights_%d' % k].initial for k in range(K)])
        p['W2'][g] = np.stack([model.layer2[g].vars['weights_%d' % k].initial for k in range(K)])
        dec = model.edge_type2decoder[g]
        if dec.kind == 'dedicom':
            p['R'][g] = dec.vars['global_interaction'].initial
            p['D'][g] = np.stack([dec.vars['local_variation_%d' % k].initial for k in range(K)])
        elif dec.kind in ('distmult', 'bilinear'):
            p['D'][g] = np.stack([dec.vars['relation_%d' % k].initial for k in range(K)])
    return p


def equivalence_steps(eng, batches, n_types):
    """One forward from the initial parameters, three training steps, one more forward:
    (losses, embeddings after the steps, embeddings before them)."""
    eng.forward(HYPER['dropout'], SEED, 99)
    Z0 = [eng.embeddings(t) for t in range(n_types)]
    losses = [float(eng.train_step(r, b, loss='hinge', margin=HYPER['margin'], lr=HYPER['lr'], dropout=HYPER['dropout'],
                                   seed=SEED, step=i)) for i, (r, b) in enumerate(batches)]
    eng.forward(0.0, SEED, 0)
    return losses, [eng.embeddings(t) for t in range(n_types)], Z0


def run_ours(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    from decagon_b200 import tf_compat as tf
    from decagon_b200.engine import Engine

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # control plane only (handle exchange, barriers, max over ranks of the timings): the per-step data
        # path is the library's own peer-memory exchange over NVLink, not a torch.distributed collective
        dist.init_process_group('gloo')

    inputs, it = build_workload(args.config, args.scale, args.decoder)
    t0 = time.time()
    # the reference's own construction sequence: placeholders, model, optimizer, session
    placeholders, model, opt = build_trainable(inputs, it)
    sess = tf.Session(seed=SEED, device=local_rank)
    np.random.seed(2)
    it.shuffle()
    n_types = len(inputs.n_nodes)

    def feed():  # what DecagonTrainer.train builds per step (DecagonTrainer.py:86-93)
        return it.update_feed_dict(it.next_minibatch_feed_dict(placeholders), HYPER['dropout'], placeholders)

    def as_batch(fd):
        return int(fd[placeholders['batch_edge_type_idx']]), np.ascontiguousarray(fd[placeholders['batch']], dtype=np.int32)

    eq_batches = [as_batch(feed()) for _ in range(3)]
    equivalence = None
    ref_eq = None
    if world > 1 and not args.no_equivalence:
        # multi-GPU equivalence: rank 0 alone first runs three steps on an UN-partitioned engine
        # (no device allocation may happen next to a live partitioned engine that peers wait on)
        if rank == 0:
            ref = Engine(inputs.n_nodes, inputs.num_feat, inputs.edge_types, inputs.edge_type2decoder, HYPER['hidden1'],
                         HYPER['hidden2'], device=local_rank)
            ref.load_iterator(it, inputs.degrees)
            ref.set_params(model_params(model))
            ref.reset_optimizer()
            ref_eq = equivalence_steps(ref, eq_batches, n_types)
            ref.close()
            del ref
        dist.barrier()

    sess.run(tf.global_variables_initializer())
    sess.run(model.embeddings[1], feed_dict=feed())  # builds, partitions (world > 1), uploads and initialises the engine
    eng = model.engine
    log('engine loaded in %.1fs, %d parameters' % (time.time() - t0, eng.n_params()))

    if world > 1 and not args.no_equivalence:
        import hashlib
        losses_p, Z_p, Z0_p = equivalence_steps(eng, eq_batches, n_types)
        digests = [None] * world
        dist.all_gather_object(digests, hashlib.sha256(b''.join(z.tobytes() for z in Z_p)).hexdigest())
        if rank == 0:
            rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / np.abs(b).max())
            equivalence = {
                'what': '3 training steps + 1 forward from identical parameters: %d-rank partitioned engine vs the '
                        'un-partitioned engine on rank 0' % world,
                'embeddings_first_forward_rel_err': max(rel(a, b) for a, b in zip(Z0_p, ref_eq[2])),
                'loss_first_step_rel_err': abs(losses_p[0] - ref_eq[0][0]) / abs(ref_eq[0][0]),
                'loss_rel_err_max': max(abs(a - b) / abs(b) for a, b in zip(losses_p, ref_eq[0])),
                'embeddings_after_3_adam_steps_rel_err': max(rel(a, b) for a, b in zip(Z_p, ref_eq[1])),
                'note': 'same parameters (first forward, first loss): the float32 contract 1e-5; after Adam steps the '
                        'trajectories separate because the first updates move every weight by ~lr * sign(g), also where '
                        'g is zero up to re-association',
                'ranks_bit_identical': all(d == digests[0] for d in digests),
                'tolerance': 1e-5, 'losses': losses_p, 'losses_one_gpu': ref_eq[0]}
        sess.run(tf.global_variables_initializer())  # back to the initial parameters and zeroed Adam slots

    fetches = [opt.opt_op, opt.cost, opt.batch_edge_type_idx]
    kw = dict(loss='hinge', margin=HYPER['margin'], lr=HYPER['lr'], dropout=HYPER['dropout'], seed=SEED)
    for _ in range(max(args.warmup, 3)):
        sess.run(fetches, feed_dict=feed())
    step = 1000000  # the engine-level legs below use their own step numbers (dropout / sampler stream keys)

    def barrier():
        eng.sync()
        if dist is not None:
            dist.barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # (1) device-resident throughput: graph, parameters and optimizer state live in HBM; the only
    # per-step input is the 4 KB batch index block; CUDA events on the library's stream
    batches = [as_batch(feed()) for _ in range(args.steps)]
    for r, batch in [as_batch(feed()) for _ in range(6)]:
        # warm the launch configuration of this leg too: a step that does not fetch the loss replays a different CUDA
        # graph than the Session steps above (no early loss copy); its first two occurrences issue / capture it, and
        # with several ranks there are two such graphs (the exchange buffers alternate)
        eng.train_step(r, batch, step=step, want_loss=False, **kw)
        step += 1
    barrier()
    eng.timing_reset()
    eng.timer_start()
    for r, batch in batches:
        eng.train_step(r, batch, step=step, want_loss=False, **kw)
        step += 1
    dev_ms = eng.timer_stop()
    launches = eng.launch_count()
    barrier()

    # (2) end to end through the drop-in boundary, exactly the reference trainer's loop body
    # (DecagonTrainer.py:86-100): iterator -> feed dict with all 1932 adjacency entries -> Session.run with host
    # buffers (H2D of the batch inside), the step, D2H of the loss, every step
    barrier()
    t0 = time.perf_counter()
    losses = []
    for _ in range(args.steps):
        outs = sess.run(fetches, feed_dict=feed())
        losses.append(outs[1])
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.summary()

    # (3) per-kernel times (CUDA events per phase) for the roofline.  On one GPU they come from a second engine over the
    # same HBM-resident workload that issues every kernel on ONE stream: with the stream lanes of the timed engine a
    # phase bracket also contains whatever the kernel waits for (a persistent kernel of lane 0 cannot start on an SM
    # while a tensor-core CTA of a side lane still holds it -- seen as 590 instead of 406 us for the layer-1
    # backward), which is scheduling, not kernel time.  With several ranks the partitioned engine itself is used.
    timed = eng
    if world == 1:
        os.environ['DGN_SINGLE_STREAM'] = '1'
        try:
            eng = Engine(inputs.n_nodes, inputs.num_feat, inputs.edge_types, inputs.edge_type2decoder, HYPER['hidden1'],
                         HYPER['hidden2'], device=local_rank)
        finally:
            del os.environ['DGN_SINGLE_STREAM']
        eng.load_iterator(it, inputs.degrees)
        eng.set_params(model_params(model))
        eng.reset_optimizer()
        for r, batch in batches[:3]:
            eng.train_step(r, batch, step=step, want_loss=False, **kw)
            step += 1
        eng.sync()
    eng.timing(True)
    eng.timing_reset()
    n_prof = min(args.steps, 10)
    for r, batch in batches[:n_prof]:
        eng.train_step(r, batch, step=step, want_loss=False, **kw)
        step += 1
    eng.sync()
    flat_index = {gk: r for r, gk in enumerate(eng.flat)}
    alg = algorithmic_bytes(inputs, it, HYPER['hidden1'], HYPER['hidden2'],
                            (lambda g, k: eng.relation_owner(flat_index[(g, k)]) in (-1, rank)) if world > 1 else None)
    phases = {}
    names = list(alg) + ['project', 'dw2', 'dh', 'mask', 'epilogue', 'decode', 'adam', 'exchange']
    names += ['%s/g%d' % (n, gi) for n in ('project', 'dw2', 'dh') for gi in range(len(inputs.edge_types))]
    for name in names:
        ms, n = eng.timing_get(name)
        if n:
            phases[name] = {'ms_per_step': ms / n_prof}
            if name in alg:
                phases[name]['bytes'] = alg[name]
                phases[name]['gbs'] = alg[name] / (ms / n_prof * 1e-3) / 1e9
    eng.timing(False)
    if eng is not timed:
        eng.close()
        eng = timed

    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('hbm_gbs', 6650.0)
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s (B200_PROFILING.md)'
    spmm = {k: v for k, v in phases.items() if k.startswith('spmm_fwd')}
    top = max(spmm, key=lambda k: spmm[k]['ms_per_step'])
    traffic = None  # DRAM bytes per launch of that kernel from the committed ncu --set full capture (1 GPU, poly shape)
    try:
        if world == 1 and args.config == 'poly' and args.scale == 1:
            traffic = load_traffic().get(top)
    except Exception:
        pass
    spe = steps_per_epoch(it)
    sps = args.steps / (dev_ms * 1e-3)
    e2e_sps = args.steps / e2e_s
    line = {
        'metric': 'train_epochs_per_s', 'value': sps / spe, 'unit': 'epochs/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
        'parallelism': ('relation-partitioned many-relation groups over %d ranks, 3 peer-memory exchanges per step, other '
                        'groups replicated' % world) if world > 1 else 'single GPU',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'steps_per_s': sps, 'steps_per_epoch': spe,
        'config': workload_config(args, inputs),
        'clocks': clocks,
        'e2e': {'value': e2e_sps / spe, 'unit': 'epochs/s', 'steps_per_s': e2e_sps, 'ms_per_step': 1e3 / e2e_sps,
                'through': 'tf_compat.Session.run([opt.opt_op, opt.cost, opt.batch_edge_type_idx], '
                           'iterator.update_feed_dict(iterator.next_minibatch_feed_dict(...)))',
                'h2d_bytes_per_step': HYPER['batch_size'] * 2 * 4, 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'kernel': top, 'achieved': spmm[top]['gbs'], 'peak': peak, 'unit': 'GB/s',
                     'frac': spmm[top]['gbs'] / peak, 'traffic': traffic, 'peak_source': peak_src,
                     'algorithmic_bytes': spmm[top]['bytes'], 'ms': spmm[top]['ms_per_step']},
        'kernels': phases,
        'kernels_note': ('CUDA events around every kernel of a training step, 10 steps, warm; one GPU: issued on one stream by a '
                         'second engine over the same workload, so a bracket holds the kernel and nothing it waits for'
                         if world == 1 else 'CUDA events around every phase on the stream lanes of the partitioned engine'),
        'loss_first_last': [float(losses[0]), float(losses[-1])],
    }
    if world > 1:
        line['roofline']['per_rank'] = 'algorithmic bytes and time of ONE rank\'s launch (its own relations of the partitioned group)'
    if equivalence is not None:
        line['equivalence'] = equivalence
    # The longest kernel of the step is the layer-1 backward SpMM of the many-relation group with the TF1 Adam of its
    # weights fused in (the gradient never reaches HBM).  SURVEY.md 8(d): transposed CSR + dS + Adam state, "2.1 GB if
    # fused into the dW epilogue" = read p, m, v and write p, m, v: 6 x 4 B per parameter instead of the dW1 write.
    try:
        bw = max((k for k in phases if k.startswith('spmm_bwd1')), key=lambda k: phases[k]['ms_per_step'])
        gi = int(bw.rsplit('g', 1)[1])
        g = list(inputs.edge_types)[gi]
        K, n_j = inputs.edge_types[g], inputs.n_nodes[g[1]]
        if world == 1 and K >= 8 and HYPER['dropout'] >= 0:
            w1 = K * n_j * HYPER['hidden1'] * 4
            fused = phases[bw]['bytes'] - w1 + 6 * w1
            line['roofline_bwd1_fused_adam'] = {
                'bound': 'hbm', 'kernel': bw + ' (spmm_tstaged_kernel<2>, Adam fused)', 'algorithmic_bytes': fused,
                'ms': phases[bw]['ms_per_step'], 'achieved': fused / phases[bw]['ms_per_step'] / 1e6, 'peak': peak,
                'unit': 'GB/s', 'frac': fused / phases[bw]['ms_per_step'] / 1e6 / peak,
                'traffic': load_traffic().get(bw) if args.config == 'poly' and args.scale == 1 else None}
    except Exception:
        pass
    if world == 1 and args.config == 'poly' and args.scale == 1:
        # BASELINE config #4: all drug pairs x all drug-drug relation matrices from the current embeddings
        # (tcgen05 3 x TF32 kernel, output written to HBM; bound by the 3.2 GB it writes)
        try:
            import torch
            g11 = (1, 1)
            K11, n1 = eng.K[g11], inputs.n_nodes[1]
            buf = torch.empty((K11, n1, n1), dtype=torch.float32, device='cuda:%d' % local_rank)
            eng.forward(0.0, SEED, step)
            for _ in range(2):
                eng.predict_relations_dev(eng.flat_index[(g11, 0)], K11, buf.data_ptr())
            eng.sync()
            eng.timer_start()
            for _ in range(5):
                eng.predict_relations_dev(eng.flat_index[(g11, 0)], K11, buf.data_ptr())
            ap_ms = eng.timer_stop() / 5
            line['all_pairs'] = {'workload': 'config #4: %d relation matrices x %d x %d scores' % (K11, n1, n1), 'ms': ap_ms,
                                 'kernel': 'predict_tc_kernel (tcgen05.mma kind::tf32, 3 x TF32 split)',
                                 'output_bytes': K11 * n1 * n1 * 4, 'write_gbs': K11 * n1 * n1 * 4 / ap_ms / 1e6,
                                 'frac_of_hbm_peak': K11 * n1 * n1 * 4 / ap_ms / 1e6 / peak}
            del buf
        except Exception as exc:  # the headline line must not depend on this extra
            line['all_pairs'] = {'error': str(exc)[:200]}
    if world == 1 and not args.no_cpu_baseline:
        sec, _, sample = cpu_port_steps(inputs, it, glorot_params(inputs, HYPER['hidden1'], HYPER['hidden2']), 2, 1, 25.0)
        line['cpu_baseline'] = {'value': 1.0 / sec / spe, 'unit': 'epochs/s', 'steps_per_s': 1.0 / sec,
                                'cores': os.cpu_count(), 'kind': 'port',
                                'sample': sample}
    print(json.dumps(line), flush=True)
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='poly', choices=['poly', 'toy'])
    ap.add_argument('--scale', type=int, default=1)
    ap.add_argument('--decoder', default=None, choices=['innerproduct', 'distmult', 'bilinear', 'dedicom'],
                    help='toy config only: put every group on this decoder (BASELINE config #2)')
    ap.add_argument('--reference-budget', type=float, default=150.0, help='seconds of timed CPU steps')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-equivalence', action='store_true', help='skip the N-GPU vs 1-GPU equivalence steps (N > 1)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
