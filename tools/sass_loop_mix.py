#!/usr/bin/env python
"""Instruction mix of the innermost hot loop of one kernel in a cuobjdump -sass listing.

    cuobjdump -sass mcmc_dynamics_b200/_lib/mcd_kernels_fast.o > /tmp/fast.sass
    python tools/sass_loop_mix.py /tmp/fast.sass 'lnlike_kernelILi1ELi0ELi1ELi0ELb0ELb0E'

Finds every backward branch, takes the innermost loop body with the most FP64 instructions, and prints counts per
opcode plus the FP64-pipe issue cycles per warp under the measured costs of profiles/r01_microbench.md
(DFMA with three distinct register operands 3, other FP64 2, MUFU.*64H 2).
"""
import collections
import re
import sys


def main():
    path, key = sys.argv[1], sys.argv[2]
    lines = open(path).read().split('\n')
    start = next(i for i, l in enumerate(lines) if 'Function :' in l and key in l)
    end = next((i for i in range(start + 1, len(lines)) if 'Function :' in lines[i]), len(lines))
    ins = []
    for l in lines[start:end]:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, text) in enumerate(ins):
        m = re.search(r'\bBRA\b.*?(0x[0-9a-f]+)', text)
        if m:
            target = int(m.group(1), 16)
            if target < a and target in addr_index:
                loops.append((addr_index[target], i))
    best = None
    for lo, hi in loops:
        if any((l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi for l2, h2 in loops):
            continue                     # not innermost
        body = ins[lo:hi + 1]
        fp64 = sum(1 for _, t in body if re.search(r'\b(DFMA|DMUL|DADD|DSETP|MUFU\.\w*64H)\b', t))
        if best is None or fp64 > best[0]:
            best = (fp64, body)
    fp64, body = best
    counts = collections.Counter()
    cycles = 0
    for _, t in body:
        t = re.sub(r'^@!?U?P\d+\s+', '', t)
        op = t.split()[0]
        base = op.split('.')[0]
        if base == 'MUFU':
            base = op
        counts[base] += 1
        if base == 'DFMA':
            regs = re.findall(r'\bR\d+\b', t)[1:]
            distinct = len(set(regs))
            cycles += 3 if (distinct >= 3 and 'UR' not in t and 'c[' not in t and not re.search(r'[ -]\d+\.?\d*e?[+-]?\d*\b(?!\])', t.split(',', 1)[1] if ',' in t else '')) else 2
        elif base in ('DMUL', 'DADD', 'DSETP') or '64H' in base:
            cycles += 2
    print('loop of %d instructions at 0x%x..0x%x' % (len(body), body[0][0], body[-1][0]))
    for op, n in counts.most_common():
        print('  %-14s %d' % (op, n))
    print('FP64-pipe instructions %d, estimated pipe cycles per warp per iteration %d' % (fp64, cycles))


if __name__ == '__main__':
    main()
