import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
from common import Case, mini_poly
from decagon_b200.engine import Engine
nt = int(sys.argv[1])
case = Case(mini_poly(n_types=nt))
inputs = case.inputs
print('case', flush=True)
eng = Engine(inputs.n_nodes, inputs.num_feat, inputs.edge_types, inputs.edge_type2decoder, hidden1=case.hidden1)
eng.load_iterator(case.it, inputs.degrees)
eng.set_params(case.p32); eng.reset_optimizer()
print('engine', flush=True)
for step, (r, batch) in enumerate(case.batches(3)):
    print('step', step, r, batch.shape, flush=True)
    l = eng.train_step(r, batch, dropout=0.1, seed=11, step=step)
    print('  loss', l, flush=True)
eng.forward(); print('fwd ok', eng.embeddings(1)[0, :3], flush=True)
