#!/usr/bin/env python
"""cuobjdump -sass excerpts of the hot kernels for profiles/ (B200_PROFILING.md: the mnemonics that prove tcgen05 / TMA).

    python tools/sass_excerpt.py > profiles/r02_sass_excerpts.txt

Per kernel: instruction count, opcode histogram, counts of the Blackwell-specific mnemonics (UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async) and
the hottest loop: the backward branch whose body holds the most LDS.128 / UTCHMMA instructions."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'decagon_b200', 'libdecagon_b200.so')
KERNELS = ['spmm_staged3_kernelILi6', 'spmm_tstaged_kernelILi2', 'spmm_tstaged_kernelILi1', 'project_ts_kernel', 'dw2_tc_kernelILi64',
           'dh_tc_kernelILi64', 'predict_tc_kernel', 'spmm_seg_kernelILi2', 'decode_kernel']
SPECIAL = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTCATOMSWS', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'LDGSTS', 'LDS.128',
           'LDS.64', 'STS.128', 'LDG.E.128', 'STG.E.128', 'FFMA', 'CCTL', 'UBLKPF']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout.split('\n')
    funcs, cur = {}, None
    for line in sass:
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    print('cuobjdump -sass %s (sm_100a), %d functions' % (os.path.relpath(LIB, ROOT), len(funcs)))
    for want in KERNELS:
        for name, lines in funcs.items():
            if want not in name:
                continue
            ins = []
            for l in lines:
                m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
                if m:
                    ins.append((int(m.group(1), 16), m.group(2).strip()))
            ops = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0] for _, t in ins)
            print('\n' + '=' * 110)
            print('%s\n%d instructions' % (name, len(ins)))
            print('opcodes:', ', '.join('%s %d' % kv for kv in ops.most_common(22)))
            print('marker mnemonics:', ', '.join('%s %d' % (s, sum(1 for _, t in ins if s in t)) for s in SPECIAL
                                                 if any(s in t for _, t in ins)))
            # hottest loop: backward branch with the most LDS.128 / UTCHMMA in its body
            addr = {a: i for i, (a, _) in enumerate(ins)}
            best = None
            for i, (a, t) in enumerate(ins):
                m = re.search(r'\bBRA\S*\s+(?:.*\s)?0x([0-9a-f]+)', t)
                if not m:
                    continue
                tgt = int(m.group(1), 16)
                if tgt in addr and addr[tgt] < i:
                    body = ins[addr[tgt]:i + 1]
                    score = sum(1 for _, x in body if 'LDS.128' in x or 'UTCHMMA' in x or 'LDTM' in x)
                    if score and len(body) < 400 and (best is None or score / len(body) > best[0]):
                        best = (score / len(body), body)
            if best:
                print('hottest loop (%d instructions):' % len(best[1]))
                for a, t in best[1][:140]:
                    print('    /*%04x*/  %s' % (a, t))
                if len(best[1]) > 140:
                    print('    ... (%d more)' % (len(best[1]) - 140))


if __name__ == '__main__':
    main()
