#!/bin/bash
# Round-2 profile captures on the GPU box (B200_PROFILING.md recipe): the plain command first, then the launch list,
# then one --set full capture of the staged SpMM kernels and one of the tensor-core kernels.  Outputs: gpurun_out/r02_*
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err || { tail -5 gpurun_out/r02_prof_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spmm_staged3|spmm_tstaged" -s 8 -c 4 -f -o gpurun_out/r02_spmm $CMD > gpurun_out/r02_ncu_spmm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"project_ts|dw2_tc|dh_tc|spmm_seg" -s 24 -c 12 -f -o gpurun_out/r02_dense $CMD > gpurun_out/r02_ncu_dense.log 2>&1
ncu -i gpurun_out/r02_spmm.ncu-rep --page raw --csv > gpurun_out/r02_spmm_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_dense.ncu-rep --page raw --csv > gpurun_out/r02_dense_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
python tools/timeline.py > gpurun_out/r02_timeline.txt 2>&1 || true
