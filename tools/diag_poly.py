"""Diagnostic (GPU box): per-tensor errors at the config-#3 shape instead of the first failing assert."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from common import Case, rel_err, rel_l2
from decagon_b200 import datasets, _lib
from oracle import decagon_oracle as O
import torch

SEED = 42
c = Case(datasets.polypharmacy_graph())
eng = c.engine()
print('engine ready', flush=True)
Z, cache = O.encoder_forward(c.graph, c.p64, 0.0, None)
eng.forward(0.0, SEED, 0)
g = (1, 1); n = c.graph.n_nodes[1]; K = c.graph.K[g]
for mode in ('0', '1'):
    os.environ['DGN_PREDICT_FFMA'] = mode
    for k0, cnt in ((0, 1), (0, 4), (K - 4, 4), (0, 64)):
        buf = torch.full((cnt, n, n), float('nan'), dtype=torch.float32, device='cuda')
        eng.predict_relations_dev(eng.flat_index[(g, k0)], cnt, buf.data_ptr()); eng.sync()
        out = buf.cpu().numpy()
        errs = [rel_err(out[q], O.predict_all_pairs(c.graph, c.p64, Z, g, k0 + q)) for q in range(min(cnt, 4))]
        print('predict ffma=%s k0=%d cnt=%d' % (mode, k0, cnt), ['%.2e' % e for e in errs], 'nan', int(np.isnan(out).sum()), flush=True)
    one = eng.predict(eng.flat_index[(g, 3)])
    print('predict_all_pairs(host) ffma=%s' % mode, '%.2e' % rel_err(one, O.predict_all_pairs(c.graph, c.p64, Z, g, 3)))
del os.environ['DGN_PREDICT_FFMA']

for step, (r, batch) in enumerate(c.batches(4)):
    gg, k = c.graph.flat[r]
    if gg not in ((0, 0), (1, 1)):
        continue
    negs = O.sample_negatives(c.thresholds(r), len(batch), r, step, SEED)
    masks = O.masks_for(c.graph, 0.1, step, SEED)
    loss, pos, neg, grads, Zs = O.train_step_grads(c.graph, c.p64, gg, k, batch, negs, 0.1, masks, 'hinge')
    got = eng.train_step(r, batch, negatives=negs, loss='hinge', dropout=0.1, seed=SEED, step=step, apply_update=False)
    print('batch of', gg, k, 'loss', float(got), loss, flush=True)
    for t in Zs:
        print('  dZ', t, '%.2e' % rel_err(eng.tensor(_lib.TENSOR_GRAD_EMBEDDINGS, t), 0 * Zs[t] + eng.tensor(_lib.TENSOR_GRAD_EMBEDDINGS, t)))
    eg = eng.get_grads()
    for name in grads:
        for q in grads[name]:
            a, b = eg[name][q], grads[name][q]
            print('  grad %s %s max %.2e l2 %.2e |ref|max %.3e |dev|max %.3e' % (name, q, rel_err(a, b), rel_l2(a, b), np.abs(b).max(), np.abs(a).max()), flush=True)
            if name == 'W1' and rel_err(a, b) > 1e-5:
                d = np.abs(a - b)
                idx = np.unravel_index(np.argsort(d.ravel())[-5:], d.shape)
                for i in range(5):
                    ii = tuple(x[i] for x in idx)
                    print('     worst', ii, a[ii], b[ii])
                per_k = d.reshape(d.shape[0], -1).max(axis=1)
                print('     relations with err > 1e-6*max:', int((per_k > 1e-6 * np.abs(b).max()).sum()), 'of', len(per_k))
                rows = np.nonzero(d.max(axis=2) > 1e-6 * np.abs(b).max())
                print('     bad rows (k, row) count', len(rows[0]), 'first', list(zip(rows[0][:10], rows[1][:10])))
