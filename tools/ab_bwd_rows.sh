#!/bin/bash
# A/B of the backward row order (GPU box): length-sorted rows per relation (default) against address order
# (DGN_BWD_ROW_ORDER=address).  Parity of the switched path first, then two bench lines.
DGN_BWD_ROW_ORDER=address python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider \
    -k "fused_adam_updates_w1 or grads_mini_staged" > gpurun_out/ab_rows_pytest.log 2>&1
tail -2 gpurun_out/ab_rows_pytest.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_rows_default.json 2> gpurun_out/ab_rows_default.err
DGN_BWD_ROW_ORDER=address python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_rows_address.json 2> gpurun_out/ab_rows_address.err
python - <<'PY'
import json
for name in ('default', 'address'):
    try:
        d = json.load(open('gpurun_out/ab_rows_%s.json' % name))
        ph = d.get('kernels', {})
        print(name, 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'],
              {k: round(v['ms_per_step'] * 1e3, 1) for k, v in ph.items() if k in ('spmm_bwd1/g2', 'spmm_bwd2/g2', 'spmm_fwd1/g2', 'spmm_fwd2/g2')})
    except Exception as e:
        print(name, 'unreadable', e)
PY
