"""Star-sharded reduction logic on CPU: world size 2, gloo backend.  Each rank evaluates the ORACLE
on its contiguous star shard (standing in for the per-GPU partial kernel) and ShardedLikelihood
all-reduces the per-walker partial sums; the result must equal the whole-catalogue value on every
rank, including -inf for prior-rejected walkers."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


class _OracleShard(object):
    """Duck-typed model: `.device`, `.pack().lnprob_partial_tensor(theta)`."""
    device = 0

    def __init__(self, oracle):
        self.oracle = oracle

    def pack(self):
        return self

    def lnprob_partial_tensor(self, theta):
        return torch.from_numpy(self.oracle.lnprob_many(theta.numpy()))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from mcmc_dynamics_b200 import sharded, synthetic
        from oracle import reference_np as ref
        columns, truth = synthetic.mock_cluster(1001, seed=3, as_reader=False)

        def make(cols):
            params = ref.default_params('model')
            for p in params:
                if p.name in ('ra_center', 'dec_center'):
                    p.fixed, p.value = True, truth[p.name]
            return ref.OracleModelFit(cols, parameters=params)
        names = ['v_sys', 'sigma_max', 'a', 'v_maxx', 'v_maxy', 'r_peak']
        theta = synthetic.initial_ball(truth, names, 12, seed=1)
        theta[4, 1] = -3.0                                   # sigma_max < 0: rejected on every rank
        shard = sharded.shard_columns(columns, rank, world)
        like = sharded.ShardedLikelihood(_OracleShard(make(shard)), device='cpu')
        got = like.lnprob(theta)
        want = make(columns).lnprob_many(theta)
        ok = bool(got[4] == -np.inf and np.all(np.isfinite(np.delete(got, 4)))
                  and np.allclose(np.delete(got, 4), np.delete(want, 4), rtol=1e-12, atol=0))
        # the host stretch move over the sharded likelihood, as bench.py runs it at N > 1: every rank draws the
        # same random numbers and sees the same all-reduced sums, so the replicas must stay in lock step
        from mcmc_dynamics_b200 import sampler as samplers
        start = synthetic.initial_ball(truth, names, 16, seed=2)
        replica = samplers.HostEnsembleSampler(len(start), len(names), like.lnprob, seed=21)
        replica.run_mcmc(start, 25)
        whole = samplers.HostEnsembleSampler(len(start), len(names), make(columns).lnprob_many, seed=21)
        whole.run_mcmc(start, 25)
        ok = ok and bool(np.allclose(replica.lnprobability[:, 0], whole.lnprobability[:, 0], rtol=1e-12, atol=0))
        digest = float(np.sum(replica.chain)) + float(np.sum(replica.lnprobability)) + float(replica.naccepted.sum())
        lo, hi = sharded.shard_range(1001, rank, world)
        out.put((rank, ok, hi - lo, digest))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_partition_the_catalogue():
    from mcmc_dynamics_b200 import sharded
    for n in (0, 1, 7, 1000, 1001, 10_000_019):
        for world in (1, 2, 3, 8):
            edges = [sharded.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_two_rank_gloo_allreduce_matches_whole_catalogue():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] for r in results)
    assert sorted(r[2] for r in results) == [500, 501]
    assert results[0][3] == results[1][3]          # the two replicas of the host sampler took identical decisions


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (fused cross-GPU reduction)')
@pytest.mark.timeout(300)
def test_fused_in_kernel_allreduce_two_gpus():
    """tools/check_fused_allreduce.py under torchrun on two GPUs: fused == NCCL == whole catalogue,
    bit-identical across ranks."""
    import subprocess
    for attempt in range(2):           # a rendezvous port just released by another job can fail the first try
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
               '127.0.0.1', '--master-port', str(_free_port()), os.path.join(ROOT, 'tools', 'check_fused_allreduce.py')]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=140, env=dict(os.environ, N_STARS='100000'))
        if res.returncode == 0 or 'RESULT FAIL' in res.stdout:
            break
    assert res.returncode == 0 and 'RESULT PASS' in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
