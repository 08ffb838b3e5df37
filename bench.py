#!/usr/bin/env python
"""Headline benchmark: walker*star log-likelihood terms per second of the ensemble lnprob path.

Workload (BASELINE.json configs[4], the one the metric's 1/2/4/8-GPU sweep is quoted on; it fits one
GPU): synthetic 10^7-star cluster x 1024 walkers, ModelFit (Lynden-Bell rotation + Plummer
dispersion), fixed centre, float64.  One *step* = one emcee iteration of the red/blue stretch move =
two half-ensemble lnprob calls of 512 walkers each over every star = 1.024e10 terms.  With N > 1
GPUs the SAME catalogue is star-sharded over the ranks (strong scaling) and the per-walker partial
sums are all-reduced over NCCL after every call.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0.  `value` is device-timed (CUDA events, inputs resident in HBM);
`e2e` goes through the public host-buffer API (H2D of theta and D2H of lnprob inside the timed
region); `roofline` and `cpu_baseline` are described in DESIGN.md.  Before anything is timed the run
checks itself (`checks`): the fused cross-GPU reduction against kernel + NCCL all_reduce on the full
workload, and the sharded path against the NumPy oracle on a 2e5-star prefix of the catalogue;
`steps_per_s` holds MEASURED emcee iterations per second of the device-resident and the host stretch-move
samplers on the workload; `configs` (N = 1) runs BASELINE.json configurations C1-C4 in-process with parity,
terms/s, steps/s of both samplers and the CPU oracle beside each.
"""
import argparse
import json
import multiprocessing
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'walker*star lnlike terms/s'
UNIT = 'terms/s'
SEED = 4


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=60)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--stars', type=int, default=10_000_000)
    ap.add_argument('--walkers', type=int, default=1024)
    ap.add_argument('--math', default='fast', choices=['fast', 'plain'])
    ap.add_argument('--free-centre', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample-stars', type=int, default=250_000)
    ap.add_argument('--no-configs', action='store_true', help='skip the C1-C4 block (N = 1 only)')
    ap.add_argument('--no-samplers', action='store_true', help='skip the measured steps/s of the samplers')
    ap.add_argument('--check-stars', type=int, default=200_000, help='catalogue prefix of the oracle check')
    return ap.parse_args()


def bytes_per_star(args):
    """Packed float64 columns the likelihood kernel streams per star."""
    return 40


def collective_mode():
    return 'nccl' if os.environ.get('MCD_COLLECTIVE', 'fused') == 'nccl' else 'fused'


def workload_config(args, world):
    """The workload both arms (`--impl b200` and `--impl reference`) describe: a pure function of the
    command line and the number of ranks, so that the two JSON lines carry the same `config`."""
    half = args.walkers // 2
    base, extra = divmod(args.stars, world)
    stars_per_gpu = base + (1 if extra else 0)
    resident = stars_per_gpu * bytes_per_star(args)
    cores = usable_cores()
    return {
        'workload': 'C5: synthetic {0:.0e}-star cluster x {1} walkers, ModelFit {2} centre, star-sharded'.format(
            args.stars, args.walkers, 'free' if args.free_centre else 'fixed'),
        'n_stars': args.stars, 'n_walkers': args.walkers, 'walkers_per_call': half,
        'calls_per_step': 2, 'model': 'ModelFit', 'free_centre': bool(args.free_centre), 'math': args.math,
        'seed': SEED, 'stars_per_gpu': stars_per_gpu,
        'collective': 'none' if world == 1 else (
            'in-kernel one-shot all-reduce of %d f64 per call over NVLink peer memory (symmetric memory), fused into '
            'the likelihood kernel' % half if collective_mode() == 'fused' else
            'nccl all_reduce(sum) of %d f64 per call' % half),
        'l2': 'flushed between steps (512 MiB write)' if resident < (256 << 20) else
              'inputs larger than L2 (%.0f MB per GPU)' % (resident / 1e6),
        # what the CPU arm (cpu_baseline / --impl reference) evaluates per step: a bounded sample of the workload
        'cpu_arm_sample': {'n_stars_sampled': min(args.stars, args.cpu_sample_stars), 'walkers_sampled': 4 * cores,
                           'calls': 'one lnprob call per walker, process pool over %d cores' % cores},
    }


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle (literal NumPy restatement of the reference), one lnprob call per walker,
# fanned out over a process pool exactly like analysis/runner.py:398-403
# ----------------------------------------------------------------------------------------------
_CPU_ORACLE = None


def _cpu_init(columns, parameters_table, fixed):
    global _CPU_ORACLE
    from oracle import reference_np as ref
    params = ref.default_params(parameters_table)
    for p in params:
        if p.name in fixed:
            p.fixed = True
            p.value = fixed[p.name]
    _CPU_ORACLE = ref.OracleModelFit(columns, parameters=params)


def _cpu_one(theta_row):
    return _CPU_ORACLE.lnprob(theta_row)


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(args, steps, warmup):
    """Times the CPU reference path on a bounded sample of the workload.  Returns a dict with
    terms/s for the pool (all cores) and for a single process."""
    from mcmc_dynamics_b200 import synthetic
    n_sample = min(args.stars, args.cpu_sample_stars)
    columns, truth = synthetic.mock_cluster(n_sample, seed=SEED, as_reader=False)
    fixed = {} if args.free_centre else {'ra_center': truth['ra_center'], 'dec_center': truth['dec_center']}
    names = [n for n in ('v_sys', 'sigma_max', 'a', 'v_maxx', 'ra_center', 'dec_center', 'v_maxy', 'r_peak')
             if n not in fixed]
    cores = usable_cores()
    per_step = 4 * cores
    theta = synthetic.initial_ball(truth, names, per_step, seed=5)
    _cpu_init(columns, 'model', fixed)
    # single process, n_threads = 1 (the reference default)
    t0 = time.perf_counter()
    for row in theta[:2]:
        _cpu_one(row)
    single = 2 * n_sample / (time.perf_counter() - t0)
    ctx = multiprocessing.get_context('fork')
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(_cpu_one, list(theta[:cores]))
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_one, list(theta))
        elapsed = time.perf_counter() - t0
    value = steps * per_step * n_sample / elapsed
    # the same sample through the compiled C restatement (oracle/oracle_c.c, OpenMP over walkers):
    # what the reference's arithmetic costs without NumPy temporaries and the Python interpreter
    c_port = None
    try:
        from oracle import c_port as cport
        if cport.available():
            co = cport.COracle(_CPU_ORACLE)
            c_threads = cport.set_threads(cores)           # torchrun exports OMP_NUM_THREADS=1
            co.lnprob_many(theta[:cores])
            t0 = time.perf_counter()
            co.lnprob_many(theta)
            c_all = per_step * n_sample / (time.perf_counter() - t0)
            t0 = time.perf_counter()
            co.lnprob_many(theta[:1])
            c_port = {'value': c_all, 'unit': UNIT, 'threads': c_threads,
                      'one_walker_call_value': n_sample / (time.perf_counter() - t0),
                      'what': 'oracle/oracle_c.c (gcc -O2 -fopenmp), literal C restatement incl. per-call geometry'}
    except Exception as exc:                                # the C port is optional
        c_port = {'error': str(exc)}
    return {
        'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'n_stars_sampled': n_sample,
        'walkers_sampled': per_step,
        'steps_per_s_extrapolated': value / (float(args.walkers) * float(args.stars)),
        'sample': '{0} lnprob calls (one walker each, process pool of {1}) over a {2}-star catalogue from the workload generator, '
                  'x{3} steps; NumPy oracle = literal restatement of the reference'.format(per_step, cores, n_sample,
                                                                                           steps),
        'single_process_value': single, 'ms_per_step': 1e3 * elapsed / steps, 'c_port': c_port,
    }


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._thread = None
        self._proc = None

    def __enter__(self):
        # one long-lived nvidia-smi in loop mode (a fresh process per sample takes ~1 s on these boxes)
        try:
            self._proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self._proc = None
            return self
        self._thread = threading.Thread(target=self._read, daemon=True)
        self._thread.start()
        return self

    def _read(self):
        for line in self._proc.stdout:
            if self._stop.is_set():
                break
            cells = [c.strip() for c in line.strip().split(',')]
            if len(cells) >= 7:
                self.rows.append((time.perf_counter(), cells))

    def __exit__(self, *exc):
        self._stop.set()
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except Exception:
                self._proc.kill()
            self._thread.join(timeout=5)

    def summary(self, t0=None, t1=None):
        """Median SM clock and the throttle reasons seen between perf_counter times t0 and t1 (the
        timed region); falls back to every sample taken under load if the window caught none."""
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [cells for (t, cells) in self.rows if t0 is None or (t0 <= t <= t1)]
        if not rows:
            rows = [cells for (_, cells) in self.rows]
        for row in rows:
            try:
                sm.append(float(row[0]))
                smax.append(float(row[1]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, row[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(np.max(smax)), 'reasons': sorted(reasons),
                'samples': len(sm)}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def half_guess(args):
    return args.walkers // 2


def relative_error(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    if not np.array_equal(np.isinf(got), np.isinf(want)) or np.any(np.isnan(got)):
        return float('inf')
    fin = np.isfinite(want)
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(got[fin] - want[fin]) / np.maximum(1.0, np.abs(want[fin]))))


def load_profile_json(name):
    try:
        with open(os.path.join(ROOT, 'profiles', name)) as f:
            return json.load(f)
    except Exception:
        return None


def self_checks(args, rank, world, local_rank, device, like, packed, halves_dev, truth, free_names):
    """Run before anything is timed; every rank takes part, rank 0 reports.

    * fused_vs_nccl: the full workload's first half-ensemble through the fused in-kernel exchange and
      through shard kernel + NCCL all_reduce (the reference semantics: ONE sum over all stars,
      analysis/runner.py:264-271); 1e-12.
    * vs_oracle: a 2e5-star prefix of the same catalogue, sharded over the ranks exactly like the
      workload, three walkers, against the NumPy oracle of the whole prefix on rank 0; 1e-9 (north_star).
    * lnprob_digest: the first four values of the workload's lnprob, so that lines taken at different
      GPU counts can be compared by eye."""
    import torch
    import torch.distributed as dist
    from mcmc_dynamics_b200 import sharded, synthetic
    from mcmc_dynamics_b200.analysis import ModelFit
    checks = {}
    full = like.lnprob_tensor(halves_dev[0])
    torch.cuda.synchronize(device)
    checks['lnprob_digest'] = [float(x) for x in full[:4].cpu().numpy()]
    if world > 1:
        partial = packed.lnprob_partial_tensor(halves_dev[0])
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
        checks['fused_vs_nccl' if like.fused else 'nccl_vs_nccl'] = relative_error(full.cpu().numpy(),
                                                                                  partial.cpu().numpy())
        gathered = [torch.empty_like(full) for _ in range(world)]
        dist.all_gather(gathered, full)
        checks['bit_identical_across_ranks'] = bool(all(torch.equal(g, gathered[0]) for g in gathered))
    # the oracle leg on a prefix of the catalogue (same generator, same seed: a prefix of an n-star mock
    # cluster is not a smaller mock cluster, so it is cut from a catalogue generated at its own size)
    n_small = min(args.check_stars, args.stars)
    columns, small_truth = synthetic.mock_cluster(n_small, seed=SEED, as_reader=False)
    shard = sharded.shard_columns(columns, rank, world)
    small = ModelFit(synthetic.reader_from_columns(shard), device=local_rank, math_mode=args.math)
    for name in ('ra_center', 'dec_center'):
        small.parameters[name].set(value=small_truth[name], fixed=not args.free_centre)
    small_like = sharded.ShardedLikelihood(small, fused=like.fused, max_walkers=64)
    theta = synthetic.initial_ball(small_truth, free_names, 4, seed=6)
    theta[3, free_names.index('sigma_max')] = -1.0               # one prior-rejected walker: exactly -inf
    got_tensor = small_like.lnprob_tensor(torch.as_tensor(theta, device=device)).cpu().numpy()
    got_host = small_like.lnprob(theta) if world > 1 else small.lnprob(theta)
    if rank == 0:
        from oracle import reference_np as ref                   # checker only
        params = ref.default_params('model')
        for par in params:
            if par.name in ('ra_center', 'dec_center') and not args.free_centre:
                par.fixed, par.value = True, small_truth[par.name]
        want = ref.OracleModelFit(columns, parameters=params).lnprob_many(theta)
        checks['vs_oracle'] = relative_error(got_tensor, want)
        checks['vs_oracle_host_buffers'] = relative_error(got_host, want)
        checks['oracle_check'] = '%d stars (sharded %d-way like the workload) x 4 walkers incl. one prior-rejected' % (
            n_small, world)
        assert checks['vs_oracle'] < 1e-9 and checks['vs_oracle_host_buffers'] < 1e-9, checks
        for key in ('fused_vs_nccl', 'nccl_vs_nccl'):
            assert checks.get(key, 0.0) < 1e-12, checks
        assert checks.get('bit_identical_across_ranks', True), checks
    small.pack().close()
    return checks


def measured_samplers(args, world, device, model, like, theta_host, n_steps):
    """emcee iterations per second, measured: the device-resident stretch move and the host stretch move
    (`Runner.__call__` -> `run_mcmc`, analysis/runner.py:332-443) on the workload.  A first short run of
    each engine warms it up (graph capture, scratch sizing), the second is timed by wall clock."""
    import torch
    import torch.distributed as dist
    from mcmc_dynamics_b200 import sampler as samplers
    out = {}

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize(device)
        elapsed = time.perf_counter() - t0
        t = torch.tensor([elapsed], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if world == 1:
        model(n_walkers=args.walkers, n_steps=3, pos=theta_host, sampler='device', seed=1, prefix=None)
        out['device'] = n_steps / timed(lambda: model(n_walkers=args.walkers, n_steps=n_steps, pos=theta_host,
                                                      sampler='device', seed=1, prefix=None))
        # the same engine without the per-run set-up (state upload, initial lnprob of all walkers, graph capture)
        dev = samplers.DeviceEnsembleSampler(args.walkers, theta_host.shape[1], model.pack(), seed=1)
        dev.run_mcmc(theta_host, 3, store=False)
        out['device_steady_state'] = n_steps / timed(lambda: dev.run_mcmc(None, n_steps, store=False))
        out['device_engine'] = dev.engine[0]
        dev.close()
        model(n_walkers=args.walkers, n_steps=2, pos=theta_host, sampler='host', seed=1, prefix=None)
        out['host'] = n_steps / timed(lambda: model(n_walkers=args.walkers, n_steps=n_steps, pos=theta_host,
                                                    sampler='host', seed=1, prefix=None))
        out['api'] = "Runner.__call__(n_walkers, n_steps, sampler='device' | 'host')"
    else:
        if like.fused:
            dev = like.device_sampler(args.walkers, seed=1)
            dev.run_mcmc(theta_host, 3, store=False)
            # the engine alone (graph captured, nothing stored), then a run as a user makes it: chain stored,
            # which allocates the chain and re-captures the graph once
            out['device_steady_state'] = n_steps / timed(lambda: dev.run_mcmc(None, n_steps, store=False))
            out['device'] = n_steps / timed(lambda: dev.run_mcmc(None, n_steps))
            dev.close()
        host = samplers.HostEnsembleSampler(args.walkers, theta_host.shape[1], like.lnprob, seed=1)
        pos, lnp, _ = host.run_mcmc(theta_host, 2, store=False)
        out['host'] = n_steps / timed(lambda: host.run_mcmc(pos, n_steps, log_prob0=lnp))
        out['api'] = ('ShardedLikelihood.device_sampler(...).run_mcmc (replicated ensemble, fused exchange per half-step) | '
                      'HostEnsembleSampler over ShardedLikelihood.lnprob on every rank')
    out['n_steps'] = n_steps
    return out


def configs_block(device_index):
    """BASELINE.json configurations C1-C4 in-process (milliseconds each): parity against the oracle, terms/s
    of the half-ensemble lnprob call with theta resident and through the host-buffer C ABI, measured steps/s
    of both samplers, and the single-process CPU oracle beside them."""
    import torch
    from mcmc_dynamics_b200 import configs, synthetic
    from mcmc_dynamics_b200 import sampler as samplers
    from oracle import harness                                    # checker / CPU baseline only
    block = {}
    for key in ('C1', 'C2', 'C3', 'C3b', 'C4'):
        t_build = time.perf_counter()
        name, model, truth, n_walkers = configs.BUILDERS[key](device=device_index)
        packed = model.pack()
        t_build = time.perf_counter() - t_build
        n = model.n_data
        half = n_walkers // 2
        theta = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=5, scale=0.05)
        got = model.lnprob(theta[:half])
        n_check = min(half, 3 if n >= 100_000 else 8)
        oracle = harness.oracle_for(model)
        t0 = time.perf_counter()
        want = oracle.lnprob_many(theta[:n_check])
        cpu_s = (time.perf_counter() - t0) / n_check
        # device-timed through the C ABI itself (mcd_lnprob_device on torch's current stream): the torch
        # operator wrapper adds ~8 us of host work per call, which on C1/C2 would be timed instead of the GPU
        th_dev = torch.as_tensor(theta[:half], device='cuda:%d' % device_index)
        out_dev = torch.empty(half, dtype=torch.float64, device=th_dev.device)
        stream = torch.cuda.current_stream(th_dev.device).cuda_stream
        lib = packed._lib

        def launch():
            rc = lib.mcd_lnprob_device(packed.handle, th_dev.data_ptr(), half, out_dev.data_ptr(), stream)
            assert rc == 0, lib.mcd_last_error()
        reps = 200
        for _ in range(5):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            launch()
        e1.record()
        torch.cuda.synchronize()
        dev_s = e0.elapsed_time(e1) / reps * 1e-3
        assert np.array_equal(out_dev.cpu().numpy(), got)
        for _ in range(5):
            model.lnprob(theta[:half])
        t0 = time.perf_counter()
        for _ in range(reps):
            model.lnprob(theta[:half])
        host_s = (time.perf_counter() - t0) / reps
        steps = 300
        s = samplers.DeviceEnsembleSampler(n_walkers, model.n_fitted_parameters, packed, seed=1)
        s.run_mcmc(theta, 5, store=False)
        t0 = time.perf_counter()
        s.run_mcmc(None, steps, store=False)
        dev_steps = steps / (time.perf_counter() - t0)
        engine = s.engine
        s.close()
        h = samplers.HostEnsembleSampler(n_walkers, model.n_fitted_parameters, model.lnprob, seed=1)
        pos, lnp, _ = h.run_mcmc(theta, 3, store=False)
        t0 = time.perf_counter()
        h.run_mcmc(pos, 100, log_prob0=lnp, store=False)
        host_steps = 100 / (time.perf_counter() - t0)
        info = packed.info()
        block[key] = {
            'workload': name, 'n_stars': n, 'n_walkers': n_walkers, 'walkers_per_call': half,
            'max_rel_err_vs_oracle': harness.relative_error(got[:n_check], want), 'oracle_walkers': n_check,
            'terms_per_s': half * n / dev_s, 'us_per_call': 1e6 * dev_s,
            'e2e_terms_per_s': half * n / host_s, 'e2e_us_per_call': 1e6 * host_s,
            'steps_per_s': {'device': dev_steps, 'host': host_steps, 'cpu': 1.0 / (cpu_s * n_walkers),
                            'device_engine': '%s (%d CTAs per segment)' % engine if engine[0] == 'resident' else engine[0]},
            'cpu_terms_per_s': n / cpu_s, 'nominal_flops_per_term': info['flops_per_term'],
            'bytes_per_star': info['bytes_per_star'], 'construction_s': t_build,
        }
        assert block[key]['max_rel_err_vs_oracle'] < 1e-9, (key, block[key])
        packed.close()
        del model
    return block


def gpu_run(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised: the pool forks
        cpu_baseline = cpu_reference_run(args, steps=3, warmup=1)

    import torch
    import torch.distributed as dist
    from mcmc_dynamics_b200 import _native, synthetic, sharded
    from mcmc_dynamics_b200.analysis import ModelFit

    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the B200 path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    # ---- the workload: same catalogue on every rank, each keeps its contiguous shard ----------
    columns, truth = synthetic.mock_cluster(args.stars, seed=SEED, as_reader=False)
    shard = sharded.shard_columns(columns, rank, world)
    n_shard = len(shard['v'])
    del columns
    model = ModelFit(synthetic.reader_from_columns(shard), device=local_rank, math_mode=args.math)
    if not args.free_centre:
        model.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
        model.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    else:
        model.parameters['ra_center'].set(value=truth['ra_center'])
        model.parameters['dec_center'].set(value=truth['dec_center'])
    packed = model.pack()
    # MCD_COLLECTIVE=nccl forces kernel + NCCL all_reduce; default: the reduction fused into the kernel
    like = sharded.ShardedLikelihood(model, fused=collective_mode() == 'fused', max_walkers=max(1024, args.walkers))
    info = packed.info()
    assert info['bytes_per_star'] == bytes_per_star(args)

    half = args.walkers // 2
    theta_host = synthetic.initial_ball(truth, model.fitted_parameters, args.walkers, seed=5)
    halves_host = [np.ascontiguousarray(theta_host[:half]), np.ascontiguousarray(theta_host[half:])]
    halves_dev = [torch.as_tensor(h, device=device) for h in halves_host]

    checks = self_checks(args, rank, world, local_rank, device, like, packed, halves_dev, truth,
                         list(model.fitted_parameters))

    # inputs smaller than ~2x L2 are evicted between timed steps by writing a 512 MiB buffer
    bytes_resident = n_shard * info['bytes_per_star']
    flush = bytes_resident < (256 << 20)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=device) if flush else None

    def one_step():
        out = None
        for th in halves_dev:
            out = like.lnprob_tensor(th)
        return out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    lib = _native.load_library()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        for _ in range(max(args.warmup, 3)):
            one_step()
        sync_all()
        launches_before = packed.info()['launches']
        sync_all()
        wall0 = time.perf_counter()
        for k in range(args.steps):
            if flush:
                flush_buf.fill_(k & 0xff)
            starts[k].record()
            result = one_step()
            stops[k].record()
        sync_all()
        wall1 = time.perf_counter()
        wall = wall1 - wall0
        if wall < 0.5:
            time.sleep(0.3)       # let the sampler deliver what it measured during a short region
    clock_summary = clocks.summary(wall0, wall1)
    launches = packed.info()['launches'] - launches_before
    device_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    t = torch.tensor([device_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    device_ms = float(t.item())
    ms_per_step = device_ms / args.steps
    terms_per_step = float(args.walkers) * float(args.stars)
    value = terms_per_step / (ms_per_step * 1e-3)
    assert torch.isfinite(result).all(), 'benchmark theta must not be prior-rejected'

    # ---- dominant kernel alone: per-launch duration on the launching stream --------------------
    k_starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k_stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sync_all()
    for k in range(args.steps):
        if flush:
            flush_buf.fill_(k & 0xff)
        k_starts[k].record()
        like.lnprob_tensor(halves_dev[k & 1]) if like.fused else packed.lnprob_partial_tensor(halves_dev[k & 1])
        k_stops[k].record()
    torch.cuda.synchronize(device)
    kernel_ms = float(np.mean([s.elapsed_time(e) for s, e in zip(k_starts, k_stops)]))

    # ---- end to end through the host-buffer API -------------------------------------------------
    def e2e_step():
        out = None
        for th in halves_host:
            out = model.lnprob(th) if world == 1 else like.lnprob(th)
        return out

    for _ in range(3):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_out = e2e_step()
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = terms_per_step * args.steps / e2e_s
    # the two paths must agree bit for bit at N = 1 (same kernel), closely otherwise
    assert np.allclose(e2e_out, result.cpu().numpy(), rtol=1e-12, atol=0)

    # ---- measured emcee iterations per second ------------------------------------------------------
    steps_per_s = {'lnprob_calls_only': 1e3 / ms_per_step}
    if not args.no_samplers:
        steps_per_s.update(measured_samplers(args, world, device, model, like, theta_host, max(20, args.steps)))
    if cpu_baseline is not None:
        steps_per_s['cpu'] = cpu_baseline['steps_per_s_extrapolated']

    # ---- roofline denominators, measured in this process ----------------------------------------
    fp64 = np.zeros(2)
    rc = lib.mcd_measure_fp64_peak(local_rank, _native.as_double_ptr(fp64[0:1]), _native.as_double_ptr(fp64[1:2]))
    fp64_peak = float(fp64[0]) if rc == 0 else None
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    hbm_source = 'MEASURED_PEAKS.json' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'

    info = packed.info()
    terms_per_launch = float(half) * float(n_shard)
    flops_per_launch = terms_per_launch * info['flops_per_term']
    achieved_tflops = flops_per_launch / (kernel_ms * 1e-3) / 1e12
    bytes_per_launch = float(n_shard) * info['bytes_per_star'] + half * (info['n_theta'] + 1) * 8.0
    achieved_gbs = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    # FP64-pipe instructions issued per term by the kernel: counted in the SASS of the shipped build by
    # tools/sass_loop_mix.py (profiles/r02_sass_counts.json, listing excerpts beside it); utilisation of the
    # pipe by instruction count = instructions/s over the measured DFMA issue rate
    kernel_key = 'lnlike<RADIAL,%s,BG_NONE,%s>' % ('FREE' if args.free_centre else 'FIXED', args.math.upper())
    sass = (load_profile_json('r02_sass_counts.json') or {}).get('kernels', {}).get(kernel_key)
    pipe_instr = sass['fp64_pipe_instr_per_term'] if sass else None
    pipe_frac = None
    if fp64_peak and pipe_instr:
        pipe_frac = (terms_per_launch * pipe_instr / (kernel_ms * 1e-3)) / (fp64_peak * 1e12 / 2.0)
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu capture of this workload
    # (profiles/r02_ncu_traffic.json, written by tools/ncu_summary.py from the .ncu-rep); null otherwise
    traffic = None
    ncu = load_profile_json('r02_ncu_traffic.json') or {}
    entry = ncu.get(kernel_key)
    if entry and entry.get('n_stars') == n_shard and entry.get('walkers_per_call') == half:
        traffic = entry.get('dram_bytes_per_launch')

    config = workload_config(args, world)
    assert config['stars_per_gpu'] >= n_shard
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config,
        'checks': checks,
        'steps_per_s': steps_per_s,
        'wall_ms_per_step': 1e3 * wall / args.steps,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(2 * half * info['n_theta'] * 8),
                'd2h_bytes_per_step': int(2 * half * 8), 'ms_per_step': 1e3 * e2e_s / args.steps,
                'api': 'ModelFit.lnprob(theta ndarray) -> C ABI mcd_lnprob (host buffers)' if world == 1 else
                       ('ShardedLikelihood.lnprob(theta ndarray) -> C ABI mcd_lnprob_allreduce (host buffers; copy-in -> '
                        'shard kernel + in-kernel exchange as one CUDA graph, results polled in pinned memory)' if like.fused else
                        'ShardedLikelihood.lnprob(theta ndarray): pinned H2D, shard kernel, NCCL all_reduce, D2H')},
        'gpu_launches': int(launches),
        'clocks': clock_summary,
        'roofline': {
            'bound': 'fp64', 'achieved': achieved_tflops, 'peak': fp64_peak, 'unit': 'TFLOP/s',
            'frac': (achieved_tflops / fp64_peak) if fp64_peak else None,
            'traffic': traffic,
            'traffic_unit': 'bytes per launch (algorithmic: %.3g)' % bytes_per_launch,
            'kernel': 'mcd::lnlike_kernel<RADIAL,%s,BG_NONE,%s>' % ('FREE' if args.free_centre else 'FIXED',
                                                                    args.math.upper()),
            'kernel_ms': kernel_ms, 'terms_per_launch': terms_per_launch,
            'nominal_flops_per_term': info['flops_per_term'],
            'peak_source': 'mcd_measure_fp64_peak: DFMA chains measured in this run (MEASURED_PEAKS.json has no FP64 '
                           'figure)',
            'fp64_pipe_instr_per_term': pipe_instr, 'fp64_pipe_frac': pipe_frac,
            'fp64_pipe_instr_source': 'profiles/r02_sass_counts.json (tools/sass_loop_mix.py on the shipped build)'
                                      if sass else None,
            'hbm': {'achieved': achieved_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved_gbs / hbm_peak,
                    'bytes_per_star': info['bytes_per_star'], 'peak_source': hbm_source},
            'grid': [info['last_grid_x'], info['last_grid_y']], 'block': info['last_block'],
            'walkers_per_cta': info['last_walker_tile'],
        },
    }
    if cpu_baseline is not None:
        line['cpu_baseline'] = cpu_baseline
    if world == 1 and not args.no_configs:
        packed.close()
        line['configs'] = configs_block(local_rank)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def reference_run(args):
    """`--impl reference`: the reference's CPU implementation of the path (the NumPy oracle port: the
    reference itself needs astropy/emcee/asteval/lmfit, none of which exist in this image) on all
    host cores.  Rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', str(args.gpus)))
    res = cpu_reference_run(args, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': res['value'], 'unit': UNIT,
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args, world),
        'steps_per_s': {'cpu': res['steps_per_s_extrapolated']},
        'cpu_baseline': {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample', 'n_stars_sampled',
                                             'walkers_sampled', 'single_process_value', 'c_port')},
        'e2e': {'value': res['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == 'reference':
        reference_run(args)
    else:
        gpu_run(args)


if __name__ == '__main__':
    main()
