"""Timeline of one training step at the polypharmacy shape: every recorded phase with its stream lane and its
start / stop in microseconds after the step's first phase (CUDA events on the library's two streams).
    python tools/timeline.py > gpurun_out/timeline.txt"""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np

import bench
from decagon_b200.engine import Engine

H = bench.HYPER
inputs, it = bench.build_workload('poly', 1)
eng = Engine(inputs.n_nodes, inputs.num_feat, inputs.edge_types, inputs.edge_type2decoder, H['hidden1'], H['hidden2'])
eng.load_iterator(it, inputs.degrees)
eng.set_params(bench.glorot_params(inputs, H['hidden1'], H['hidden2']))
eng.reset_optimizer()
np.random.seed(2)
it.shuffle()
kw = dict(loss='hinge', margin=H['margin'], lr=H['lr'], dropout=H['dropout'], seed=bench.SEED)
for step in range(4):
    r, batch = bench.next_batch(it)
    eng.train_step(r, batch, step=step, **kw)
eng.sync()
eng.timing(True)
eng.timing_reset()
eng.timer_start()
for step in range(4, 7):
    r, batch = bench.next_batch(it)
    eng.train_step(r, batch, step=step, want_loss=False, **kw)
total = eng.timer_stop()
tl = eng.timeline()
print('3 steps with phase events: %.1f us per step' % (total * 1000 / 3))
n = len(tl) // 3
base = tl[2 * n][2]
for name, lane, a, b in tl[2 * n:]:
    print('%-16s lane %d  %8.1f -> %8.1f  (%6.1f us)' % (name, lane, (a - base) * 1000, (b - base) * 1000, (b - a) * 1000))
