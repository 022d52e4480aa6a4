"""Multi-GPU equivalence check: run under torchrun with one rank per GPU,

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/multigpu_check.py

Every rank trains the mini polypharmacy-shape graph for a few steps with the relations of the drug-drug
group partitioned over the ranks, rank 0 also runs the same steps on an un-partitioned engine; losses,
embeddings and the parameters each rank owns must agree (1e-5 on the same parameters, 1e-4 after six Adam
steps), and all ranks must hold bit-identical
embeddings.  (Needs GPUs: not collected by pytest; the host-side partition logic is covered on CPU by
tests/test_partition.py.)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import torch.distributed as dist
    from common import Case, mini_poly, rel_err
    from decagon_b200 import _lib
    from decagon_b200.engine import Engine

    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('gloo')  # control plane only: the data path is the library's own peer-memory exchange
    t0 = time.time()

    def say(msg):
        print('[rank %d %6.1fs] %s' % (rank, time.time() - t0, msg), flush=True)

    case = Case(mini_poly(n_types=24), batch_size=128)
    say('case built')
    inputs = case.inputs

    def make(partitioned):
        eng = Engine(inputs.n_nodes, inputs.num_feat, inputs.edge_types, inputs.edge_type2decoder, hidden1=case.hidden1, device=local)
        if partitioned:
            eng.comm_init(rank, world)
        eng.load_iterator(case.it, inputs.degrees)
        if partitioned:
            eng.connect(dist)
        eng.set_params(case.p32)
        eng.reset_optimizer()
        return eng

    batches = case.batches(6)
    # Rank 0 first runs the steps on an un-partitioned engine, alone.  (No device allocation may happen on a
    # GPU while a peer waits for it inside an exchange: a second engine is never built next to a live
    # partitioned one.)
    ref_loss, ref_Z, ref_Z0, ref_p = [], {}, {}, None
    if rank == 0:
        ref = make(False)
        ref.forward()
        ref_Z0 = {t: ref.embeddings(t) for t in inputs.n_nodes}
        for step, (r, batch) in enumerate(batches):
            ref_loss.append(float(ref.train_step(r, batch, dropout=0.1, seed=11, step=step)))
        ref.forward()
        ref_Z = {t: ref.embeddings(t) for t in inputs.n_nodes}
        ref_p = ref.get_params()
        ref.close()
        say('reference run done')
    dist.barrier()
    part = make(True)
    say('partitioned engine connected')
    part.forward()
    Z0 = {t: part.embeddings(t) for t in inputs.n_nodes}
    losses = [float(part.train_step(r, batch, dropout=0.1, seed=11, step=step)) for step, (r, batch) in enumerate(batches)]
    say('steps done: %s' % losses[:2])
    part.forward()
    Z = {t: part.embeddings(t) for t in inputs.n_nodes}
    say('forward done')
    # bit-identical embeddings on every rank
    for t in Z:
        everyone = [None] * world
        dist.all_gather_object(everyone, Z[t].tobytes())
        assert all(e == everyone[0] for e in everyone), 'embeddings of type %d differ between ranks' % t
    if rank == 0:
        # same parameters: the partitioned forward and the first loss agree to the fp32 contract (1e-5)
        first = max([rel_err(Z0[t], ref_Z0[t]) for t in Z0] + [abs(losses[0] - ref_loss[0]) / abs(ref_loss[0])])
        print('multigpu_check: first forward / first loss rel-err %.2e' % first)
        assert first <= 1e-5, first
        # six Adam steps later: Adam's m / sqrt(v) amplifies re-association differences of tiny gradients, the
        # trajectories agree to 1e-4
        worst = max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(losses, ref_loss))
        for t in Z:
            worst = max(worst, rel_err(Z[t], ref_Z[t]))
        # parameters: replicated ones everywhere, partitioned ones on their owner
        pp = part.get_params()
        flat = {gk: r for r, gk in enumerate(part.flat)}
        for name in pp:
            for g in pp[name]:
                a, b = pp[name][g], ref_p[name][g]
                if name in ('W1', 'W2'):
                    own = np.array([part.relation_owner(flat[(g, k)]) in (-1, rank) for k in range(part.K[g])])
                    a, b = a[own], b[own]
                if a.size:
                    worst = max(worst, rel_err(a, b))
        print('multigpu_check: world %d, losses %s, worst rel-err vs one GPU %.2e' % (world, losses[:3], worst))
        assert worst <= 1e-4, worst
    dist.barrier()
    if rank == 0:
        print('multigpu_check OK')
    part.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
