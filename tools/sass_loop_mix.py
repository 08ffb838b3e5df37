#!/usr/bin/env python
"""Instruction mix of the innermost hot loop (the star loop) of the likelihood kernels, from SASS.

    python tools/sass_loop_mix.py /tmp/fast.sass 'lnlike_kernelILi1ELi0ELi1ELi0ELb0ELb0E'     # one kernel, printed
    python tools/sass_loop_mix.py --profiles r02      # the shipped build -> profiles/r02_sass_counts.json
                                                      # + profiles/r02_sass_loop_<kernel>.txt (the loop listings)

For a kernel it finds every backward branch, takes the innermost loop body with the most FP64 instructions
and counts the opcodes on its hot path (rare-branch code inside the loop is skipped, see `hot_path`), plus
the FP64-pipe issue cycles per warp under the measured costs of
profiles/r01_microbench.md (DFMA with three distinct register operands 3, other FP64 2, MUFU.*64H 2).
`bench.py` reads `fp64_pipe_instr_per_term` from the JSON instead of carrying literals.
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FP64_RE = re.compile(r'\b(DFMA|DMUL|DADD|DSETP|DMNMX)\b')
MUFU64_RE = re.compile(r'\bMUFU\.\w*64H\b')

#: kernels reported in profiles/: name -> (mangled template arguments <ROT, FREE, BG, MATH, SEG, FUSE>, stars per loop iteration)
#: stars per iteration = 2 * MCD_PAIRS (no background) or 2 * MCD_BG_PAIRS (mixtures), csrc/mcd_kernels.cu
KERNELS = {
    'lnlike<RADIAL,FIXED,BG_NONE,FAST>': ('ILi1ELi0ELi0ELi0ELb0ELb0E', 4),
    'lnlike<RADIAL,FREE,BG_NONE,FAST>': ('ILi1ELi1ELi0ELi0ELb0ELb0E', 4),
    'lnlike<CONSTANT,FIXED,BG_NONE,FAST>': ('ILi0ELi0ELi0ELi0ELb0ELb0E', 4),
    'lnlike<RADIAL,FIXED,BG_FIXED_PMEMBER,FAST>': ('ILi1ELi0ELi1ELi0ELb0ELb0E', 4),
    'lnlike<RADIAL,FIXED,BG_FIXED_DENSITY,FAST>': ('ILi1ELi0ELi2ELi0ELb0ELb0E', 4),
    'lnlike<RADIAL,FIXED,BG_GAUSSIAN,FAST>': ('ILi1ELi0ELi3ELi0ELb0ELb0E', 4),
    'lnlike<CONSTANT,FIXED,BG_FIXED_PMEMBER,FAST>': ('ILi0ELi0ELi1ELi0ELb0ELb0E', 4),
}


def function_instructions(lines, key):
    start = next(i for i, l in enumerate(lines) if 'Function :' in l and key in l)
    end = next((i for i in range(start + 1, len(lines)) if 'Function :' in lines[i]), len(lines))
    ins = []
    for l in lines[start:end]:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def hot_path(body):
    """The instructions the loop executes when no rare branch is taken = the cheapest path from the loop head to
    the backward branch, where every instruction costs 1 and an out-of-line CALL (the mixture kernels'
    extended-range evaluation) costs 1000.  Predicated forward branches may or may not be taken, unconditional
    ones are; predicated non-branch instructions occupy an issue slot either way and are counted."""
    n = len(body)
    index = {a: i for i, (a, _) in enumerate(body)}
    cost = [float('inf')] * (n + 1)
    prev = [None] * (n + 1)
    cost[0] = 0.0
    for i in range(n):                      # forward edges only: a topological order is the address order
        if cost[i] == float('inf'):
            continue
        addr, text = body[i]
        step = 1000.0 if re.search(r'\bCALL\b', text) else 1.0
        m = re.search(r'\bBRA\b.*?(0x[0-9a-f]+)', text)
        targets = []
        if m and int(m.group(1), 16) > addr and int(m.group(1), 16) in index:
            targets.append(index[int(m.group(1), 16)])
            if re.match(r'^@', text):
                targets.append(i + 1)
        else:
            targets.append(i + 1)
        for t in targets:
            if cost[i] + step < cost[t]:
                cost[t] = cost[i] + step
                prev[t] = i
    path, i = [], n
    while prev[i] is not None:
        i = prev[i]
        path.append(body[i])
    return path[::-1]


def loop_mix(lines, key):
    ins = function_instructions(lines, key)
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
        fp64 = sum(1 for _, t in body if FP64_RE.search(t) or MUFU64_RE.search(t))
        if best is None or fp64 > best[0]:
            best = (fp64, body)
    _, body = best
    body = hot_path(body)
    counts = collections.Counter()
    cycles = 0
    three_operand = 0
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
            rest = t.split(',', 1)[1] if ',' in t else ''
            immediate = re.search(r'[ -]\d+\.?\d*e?[+-]?\d*\b(?!\])', rest)
            if distinct >= 3 and 'UR' not in t and 'c[' not in t and not immediate:
                cycles += 3
                three_operand += 1
            else:
                cycles += 2
        elif base in ('DMUL', 'DADD', 'DSETP', 'DMNMX') or '64H' in base:
            cycles += 2
    fp64 = sum(n for op, n in counts.items() if op in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX'))
    mufu = sum(n for op, n in counts.items() if '64H' in op)
    return {'body': body, 'counts': counts, 'instructions': len(body), 'fp64': fp64, 'mufu64': mufu,
            'dfma_three_operand': three_operand, 'pipe_cycles': cycles}


def profiles(tag):
    obj = os.path.join(ROOT, 'mcmc_dynamics_b200', '_lib', 'mcd_kernels_fast.o')
    sass = subprocess.run(['cuobjdump', '-sass', obj], check=True, capture_output=True, text=True).stdout
    lines = sass.split('\n')
    out = {'source': 'cuobjdump -sass mcmc_dynamics_b200/_lib/mcd_kernels_fast.o (the shipped build), innermost star loop; '
                     'tools/sass_loop_mix.py --profiles ' + tag,
           'cost_model': 'FP64-pipe issue cycles per warp: DFMA with three distinct register operands 3, other FP64 2, '
                         'MUFU.*64H 2 (profiles/r01_microbench.md)',
           'kernels': {}}
    for name, (key, stars) in KERNELS.items():
        mix = loop_mix(lines, 'lnlike_kernel' + key)
        other = mix['instructions'] - mix['fp64'] - mix['mufu64']
        out['kernels'][name] = {
            'mangled': '_ZN3mcd13lnlike_kernel' + key, 'stars_per_iteration': stars,
            'loop_instructions': mix['instructions'],
            'fp64_pipe_instr_per_term': mix['fp64'] / stars, 'mufu64_per_term': mix['mufu64'] / stars,
            'other_instr_per_term': other / stars, 'dfma_three_operand_per_term': mix['dfma_three_operand'] / stars,
            'est_pipe_cycles_per_warp_term': mix['pipe_cycles'] / stars,
            'opcodes': dict(mix['counts'].most_common()),
        }
        short = re.sub(r'[^A-Za-z0-9]+', '_', name).strip('_')
        with open(os.path.join(ROOT, 'profiles', '%s_sass_loop_%s.txt' % (tag, short)), 'w') as f:
            f.write('// %s  (%s)\n// innermost star loop, %d stars per iteration\n' % (name, out['kernels'][name]['mangled'], stars))
            for addr, text in mix['body']:
                f.write('/*%04x*/  %s ;\n' % (addr, text))
    with open(os.path.join(ROOT, 'profiles', tag + '_sass_counts.json'), 'w') as f:
        json.dump(out, f, indent=1)
    for name, k in out['kernels'].items():
        print('%-48s FP64 %.1f + MUFU %.1f + other %.1f per term, est. %.0f pipe cycles per warp-term' % (
            name, k['fp64_pipe_instr_per_term'], k['mufu64_per_term'], k['other_instr_per_term'],
            k['est_pipe_cycles_per_warp_term']))


def main():
    if sys.argv[1] == '--profiles':
        return profiles(sys.argv[2] if len(sys.argv) > 2 else 'r02')
    path, key = sys.argv[1], sys.argv[2]
    mix = loop_mix(open(path).read().split('\n'), key)
    body = mix['body']
    print('loop of %d instructions at 0x%x..0x%x' % (len(body), body[0][0], body[-1][0]))
    for op, n in mix['counts'].most_common():
        print('  %-14s %d' % (op, n))
    print('FP64-pipe instructions %d (+ %d MUFU.64H), estimated pipe cycles per warp per iteration %d' % (
        mix['fp64'], mix['mufu64'], mix['pipe_cycles']))


if __name__ == '__main__':
    main()
