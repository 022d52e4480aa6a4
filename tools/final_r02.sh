#!/bin/bash
# Final single-GPU collection of round 2 (GPU box): tests, headline bench, toy configs, reference arm, profiles.
python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_final.log 2>&1; tail -3 gpurun_out/r02_gputest_final.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || tail -5 gpurun_out/r02_bench_n1.err
python bench.py --config toy --steps 100 --warmup 10 > gpurun_out/r02_bench_toy.json 2> gpurun_out/r02_bench_toy.err
for d in innerproduct distmult bilinear dedicom; do
    python bench.py --config toy --decoder $d --steps 100 --warmup 10 > gpurun_out/r02_bench_toy_$d.json 2> gpurun_out/r02_bench_toy_$d.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
tools/profile_r02.sh > gpurun_out/r02_profile.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    try:
        d = json.load(open(f))
        print(f, d.get('impl', 'ours'), 'ms/step %.4f' % d['ms_per_step'], 'e2e', d.get('e2e', {}).get('ms_per_step'), 'steps/s', d.get('steps_per_s'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
