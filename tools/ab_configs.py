#!/usr/bin/env python
"""Time the lnprob launch of BASELINE configurations through the C ABI only (ctypes), so that
MCD_B200_LIB=<another build> A/B runs measure exactly the library named:

    [MCD_B200_LIB=scratch_ab/x/libmcd_b200.so] [MCD_GEOMETRY=tile,tiles_per_chunk] \
        python tools/ab_configs.py c3 c3b c4 mix mixgb [--calls 200]

Per configuration: device time per half-ensemble call (CUDA events around `mcd_lnprob_device` launches on
one stream, theta resident), host-buffer time per call (`mcd_lnprob`), terms/s of both, launch geometry.
`mix` / `mixgb` are the 2e6-star x 512-walker mixture workloads of tools/probe/mixture_large.py.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mcmc_dynamics_b200 import configs, synthetic  # noqa: E402


def mixture_large(gb):
    from mcmc_dynamics_b200.analysis import ModelFit, ModelFitGB
    from mcmc_dynamics_b200.background import Gaussian
    n = 2_000_000
    cols, truth = synthetic.mock_cluster(n, seed=2, as_reader=False)
    cols, _ = synthetic.add_background(cols, truth, seed=102)
    truth = dict(truth, v_back=5.0, sigma_back=55.0, f_back=0.3)
    data = synthetic.reader_from_columns(cols)
    m = ModelFitGB(data) if gb else ModelFit(data, background=Gaussian(5.0, 55.0))
    configs.fix_centre(m, truth)
    return ('ModelFitGB 2e6 stars' if gb else 'ModelFit + fixed background 2e6 stars'), m, truth, 1024


BUILDERS = {'c1': configs.config_c1, 'c2': configs.config_c2, 'c3': configs.config_c3, 'c3b': configs.config_c3b,
            'c4': configs.config_c4, 'c5': configs.config_c5, 'c5free': lambda: configs.config_c5(free=True),
            # one of eight star shards of C5: what each rank runs at N = 8
            'c5s': lambda: configs.config_c5(n_stars=1_250_000),
            'c5q': lambda: configs.config_c5(n_stars=2_500_000), 'c5h': lambda: configs.config_c5(n_stars=5_000_000),
            'mix': lambda: mixture_large(False), 'mixgb': lambda: mixture_large(True)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('configs', nargs='+', choices=sorted(BUILDERS))
    ap.add_argument('--calls', type=int, default=200)
    args = ap.parse_args()
    for key in args.configs:
        name, model, truth, n_walkers = BUILDERS[key]()
        n = model.n_data
        half = n_walkers // 2
        theta = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=5, scale=0.05)[:half]
        packed = model.pack()
        lib = packed._lib
        calls = max(5, min(args.calls, int(2e11 / (half * n))))
        th = torch.as_tensor(theta, device='cuda:0')
        out = torch.empty(half, dtype=torch.float64, device='cuda:0')
        stream = torch.cuda.current_stream().cuda_stream

        def launch():
            rc = lib.mcd_lnprob_device(packed.handle, th.data_ptr(), half, out.data_ptr(), stream)
            assert rc == 0, lib.mcd_last_error()
        for _ in range(5):
            launch()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(calls):
                launch()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / calls * 1e-3)
        host = model.lnprob(theta)
        assert np.array_equal(host, out.cpu().numpy())
        from oracle import harness                   # checker: three walkers against the NumPy oracle
        err = harness.relative_error(host[:3], harness.oracle_for(model).lnprob_many(theta[:3])) if n <= 3_000_000 else float('nan')
        host_s = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            for _ in range(calls):
                model.lnprob(theta)
            host_s = min(host_s, (time.perf_counter() - t0) / calls)
        info = packed.info()
        print('%-6s %-62s | device %8.1f us/call %.3e terms/s | host buffers %8.1f us/call %.3e terms/s | grid %d x %d, '
              'wl %d | lnprob[0] %.9e | rel err vs oracle %.1e' % (key, name, 1e6 * best, half * n / best, 1e6 * host_s, half * n / host_s,
                                         info['last_grid_x'], info['last_grid_y'], info['last_walker_tile'], host[0], err),
              flush=True)
        packed.close()


if __name__ == '__main__':
    main()
