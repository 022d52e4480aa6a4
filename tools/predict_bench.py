"""BASELINE config #4: all drug pairs x all drug-drug relations (Z M_r Z^T) at the polypharmacy shape.
Times dgn_predict_relations_dev (tcgen05 kernel, or the CUDA-core kernel with DGN_PREDICT_FFMA=1) and checks a
few relations against a float64 evaluation of the same formula.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from decagon_b200.engine import Engine
    small = '--small' in sys.argv
    args = type('A', (), {'config': 'toy' if small else 'poly', 'scale': 1})()
    inputs, it = bench.build_workload(args.config, args.scale)
    eng = Engine(inputs.n_nodes, inputs.num_feat, inputs.edge_types, inputs.edge_type2decoder, 64, 32)
    eng.load_iterator(it, inputs.degrees)
    eng.set_params(bench.glorot_params(inputs, 64, 32))
    rng = np.random.RandomState(7)
    Z = {t: (rng.standard_normal((inputs.n_nodes[t], 32)) * 0.3).astype(np.float32) for t in inputs.n_nodes}
    for t in Z:
        eng.set_embeddings(t, Z[t])
    g = (1, 1)
    r0 = eng.flat_index[(g, 0)]
    count = eng.K[g]
    n = inputs.n_nodes[1]
    out = torch.empty((count, n, n), dtype=torch.float32, device='cuda')
    torch.cuda.synchronize()
    for _ in range(2):
        eng.predict_relations_dev(r0, count, out.data_ptr())
    eng.sync()
    reps = 5
    eng.timer_start()
    for _ in range(reps):
        eng.predict_relations_dev(r0, count, out.data_ptr())
    ms = eng.timer_stop() / reps
    worst = 0.0
    for k in (0, 1, count // 2, count - 1):
        glb, loc = eng.relation_matrices(r0 + k)
        M = loc.astype(np.float64) @ glb.astype(np.float64) @ loc.astype(np.float64)
        ref = Z[1].astype(np.float64) @ M @ Z[1].astype(np.float64).T
        got = out[k].cpu().numpy().astype(np.float64)
        worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
    out_bytes = count * n * n * 4
    print(json.dumps({'kernel': 'predict_ffma' if os.environ.get('DGN_PREDICT_FFMA') == '1' else 'predict_tc (tcgen05, 3xTF32)',
                      'relations': count, 'n': n, 'ms': ms, 'output_GB': out_bytes / 1e9, 'write_GBs': out_bytes / ms / 1e6,
                      'max_rel_err_vs_f64': worst}), flush=True)
    assert worst <= 1e-5, worst


if __name__ == '__main__':
    main()
