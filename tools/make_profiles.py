#!/usr/bin/env python
"""gpurun_out/r02_* (written on the GPU box by tools/profile_r02.sh) -> the tracked summaries under profiles/:
launch list, per-kernel share of a step, ncu --set full tables of the staged SpMM and of the side / tensor-core kernels,
DRAM traffic per launch (what bench.py reports as roofline.traffic)."""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
WANT = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('smsp__inst_executed.sum', 'warp instructions'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'), ('launch__registers_per_thread', 'registers'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'shared-memory wavefronts'),
        ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'), ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
        ('sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
        ('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor (hmma) active %')]


def short(name):
    return name.split('(')[0].replace('void ', '').replace('unnamed>::', '').replace('dgn::', '').strip()


def raw_table(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rec = collections.OrderedDict(kernel=short(d['Kernel Name']), grid=d.get('launch__grid_size', ''))
        for key, label in WANT:
            if key in d and d[key] != '':
                rec[label] = '%s %s' % (d[key], u.get(key, ''))
        stalls = []
        for h in hdr:
            if 'issue_stalled' in h and 'per_issue_active' in h:
                try:
                    v = float(d[h].replace(',', ''))
                except ValueError:
                    continue
                if v >= 0.5:
                    stalls.append((v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
        rec['top stalls (warps per issue)'] = ', '.join('%s %.1f' % (n, v) for v, n in sorted(stalls, reverse=True)[:4])
        out.append((rec, d, u))
    return out


def to_bytes(value, unit):
    v = float(value.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def md(records, title, note):
    lines = ['# %s' % title, '', note, '']
    for rec, _, _ in records:
        lines.append('## `%s`  (grid %s)' % (rec['kernel'], rec['grid']))
        lines.append('')
        lines.append('| metric | value |')
        lines.append('|---|---|')
        for k, v in rec.items():
            if k not in ('kernel', 'grid'):
                lines.append('| %s | %s |' % (k, v))
        lines.append('')
    return '\n'.join(lines)


def main():
    os.makedirs(PROF, exist_ok=True)
    # launch list + shares of one step
    src = os.path.join(OUT, 'r02_launches.csv')
    shutil.copy(src, os.path.join(PROF, 'r02_launches.csv'))
    rows = list(csv.reader(open(src)))
    hdr, launches = None, []
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
        elif hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            launches.append((short(d['Kernel Name']), float(d['Metric Value'])))
    starts = [i for i, (n, _) in enumerate(launches) if 'gen_mask_multi' in n]
    step = launches[starts[-2]:starts[-1]]
    total = sum(t for _, t in step)
    agg = collections.OrderedDict()
    for n, t in step:
        agg.setdefault(n, [0.0, 0])
        agg[n][0] += t
        agg[n][1] += 1
    with open(os.path.join(PROF, 'r02_step_shares.md'), 'w') as f:
        f.write('# One training step at the polypharmacy shape, kernel by kernel (ncu launch list, round 2)\n\n'
                '`ncu --metrics gpu__time_duration.sum --clock-control none` of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` '
                '(the plain command exited 0 first).  ncu serialises the kernels and replays them with cold caches: the SHARES below are '
                'what is comparable with the CUDA-event phases of the bench line, not the absolute times.  %d launches per step '
                '(replayed as one CUDA graph in the timed runs), %.1f us serialised.\n\n| kernel | launches | us | share |\n|---|---|---|---|\n'
                % (len(step), total / 1e3))
        for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write('| `%s` | %d | %.1f | %.1f %% |\n' % (n, c, t / 1e3, 100 * t / total))
    # ncu --set full tables
    spmm = raw_table(os.path.join(OUT, 'r02_spmm_raw.csv'))
    note = ('`ncu --set full --clock-control none --import-source on -k regex:"spmm_staged3|spmm_tstaged"` of `python bench.py --steps 2 '
            '--warmup 3 --no-cpu-baseline` (polypharmacy shape, one B200; read with `ncu -i ... --page raw --csv`).  Times under ncu are '
            'cold-cache and serialised; the CUDA-event times of the same kernels inside a step are in `r02_bench_n1.json`.')
    open(os.path.join(PROF, 'r02_spmm_ncu.md'), 'w').write(md(spmm, 'Staged SpMM kernels, round 2 (ncu --set full)', note))
    side = raw_table(os.path.join(OUT, 'r02_dense_raw.csv'))
    open(os.path.join(PROF, 'r02_side_ncu.md'), 'w').write(md(side, 'Gather-path and tensor-core layer-2 kernels, round 2 (ncu --set full)',
                                                              note.replace('spmm_staged3|spmm_tstaged', 'project_ts|dw2_tc|dh_tc|spmm_seg')))
    # DRAM traffic per launch of the staged kernels, keyed by the bench's phase names
    traffic = {'note': 'dram__bytes_read.sum + dram__bytes_write.sum per launch from profiles/r02_spmm_ncu.md (ncu --set full, polypharmacy '
                       'shape, 1 GPU)'}
    seen = collections.Counter()
    for rec, d, u in spmm:
        k = rec['kernel']
        b = to_bytes(d['dram__bytes_read.sum'], u['dram__bytes_read.sum']) + to_bytes(d['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
        if 'staged3' in k:
            traffic['spmm_fwd1/g2' if seen[k] == 0 else 'spmm_fwd2/g2'] = b
        elif 'tstaged_kernel<1>' in k:
            traffic['spmm_bwd2/g2'] = b
        elif 'tstaged_kernel<2>' in k:
            traffic['spmm_bwd1/g2'] = b
        seen[k] += 1
    json.dump(traffic, open(os.path.join(PROF, 'r02_traffic.json'), 'w'), indent=1)
    for name in ('r02_timeline.txt',):
        if os.path.exists(os.path.join(OUT, name)):
            shutil.copy(os.path.join(OUT, name), os.path.join(PROF, name))
    print(open(os.path.join(PROF, 'r02_step_shares.md')).read())
    print(json.dumps(traffic, indent=1))


if __name__ == '__main__':
    main()
