#!/usr/bin/env python
"""Summarise an ncu report (read here, without a GPU) into the small CSV kept under profiles/:

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r02_lnlike_kernel_ncu_metrics.csv \
        [--kernel lnlike_kernel] [--traffic 'lnlike<RADIAL,FIXED,BG_NONE,FAST>' n_stars walkers_per_call]

One row per metric, one column per profiled launch of the selected kernel.  With --traffic the per-launch
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches) are also written to
profiles/r02_ncu_traffic.json under the given kernel key, which is where bench.py takes `roofline.traffic` from.
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    'gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.sum',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
    'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_dynamic',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
    'lts__t_bytes.sum', 'nvlrx__bytes.sum', 'nvltx__bytes.sum',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
]

TO_BYTES = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def raw_page(path):
    """Rows of `ncu --page raw --csv` (one per launch) or of a `--csv --log-file` metric list."""
    if path.endswith('.ncu-rep'):
        text = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], check=True, capture_output=True, text=True).stdout
    else:
        text = open(path).read()
    lines = [l for l in text.split('\n') if l.startswith('"')]
    rows = list(csv.reader(io.StringIO('\n'.join(lines))))
    header = rows[0]
    if 'Metric Name' in header:                       # long format: one row per (launch, metric)
        k_id, k_name, k_metric, k_unit, k_value = (header.index(c) for c in ('ID', 'Kernel Name', 'Metric Name',
                                                                            'Metric Unit', 'Metric Value'))
        launches = {}
        for r in rows[1:]:
            d = launches.setdefault(r[k_id], {'Kernel Name': (r[k_name], '')})
            d[r[k_metric]] = (r[k_value], r[k_unit])
        return list(launches.values())
    units = rows[1]
    out = []
    for r in rows[2:]:
        out.append({h: (v, u) for h, v, u in zip(header, r, units)})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('out')
    ap.add_argument('--kernel', default='lnlike_kernel')
    ap.add_argument('--traffic', nargs=3, metavar=('KEY', 'N_STARS', 'WALKERS_PER_CALL'))
    args = ap.parse_args()
    launches = [l for l in raw_page(args.report) if args.kernel in l['Kernel Name'][0]]
    if not launches:
        raise SystemExit('no launch of %s in %s' % (args.kernel, args.report))
    with open(args.out, 'w') as f:
        f.write('metric,unit,' + ','.join('launch_%d' % (i + 1) for i in range(len(launches))) + '\n')
        f.write('kernel,,' + ','.join('"%s"' % l['Kernel Name'][0] for l in launches) + '\n')
        for m in METRICS:
            if m in launches[0]:
                f.write('%s,%s,%s\n' % (m, launches[0][m][1], ','.join(l[m][0].replace(',', '') for l in launches)))
    print('wrote', args.out, '(%d launches)' % len(launches))
    if args.traffic:
        total = []
        for l in launches:
            b = 0.0
            for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                value, unit = l[m]
                b += float(value.replace(',', '')) * TO_BYTES.get(unit, 1.0)
            total.append(b)
        path = os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')
        try:
            doc = json.load(open(path))
        except Exception:
            doc = {}
        doc[args.traffic[0]] = {'n_stars': int(args.traffic[1]), 'walkers_per_call': int(args.traffic[2]),
                                'dram_bytes_per_launch': sum(total) / len(total), 'launches': len(total),
                                'source': os.path.basename(args.out)}
        json.dump(doc, open(path, 'w'), indent=1)
        print('traffic %.4g bytes per launch ->' % (sum(total) / len(total)), path)


if __name__ == '__main__':
    main()
