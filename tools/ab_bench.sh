#!/bin/bash
# A/B on the GPU box: the in-tree library and every variants/lib_*.so named on the command line, one short bench each
for name in main "$@"; do
    if [ "$name" = main ]; then unset DGN_LIB_PATH; else export DGN_LIB_PATH=$PWD/variants/lib_$name.so; fi
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || tail -5 gpurun_out/ab_$name.err
    python - "$name" <<'PY'
import json, sys
d = json.load(open('gpurun_out/ab_%s.json' % sys.argv[1]))
k = d['kernels']
print(sys.argv[1], 'step %.4f ms  e2e %.4f ms' % (d['ms_per_step'], d['e2e']['ms_per_step']),
      {n: round(k[n]['ms_per_step'] * 1000, 1) for n in ('spmm_fwd1/g2', 'spmm_fwd2/g2', 'spmm_bwd2/g2', 'spmm_bwd1/g2', 'project/g2', 'dw2/g2', 'dh/g2') if n in k})
PY
done
