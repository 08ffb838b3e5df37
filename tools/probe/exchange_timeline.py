"""Timeline of the fused cross-GPU reduction on the headline workload (run under torchrun, one rank per GPU, with
MCD_B200_LIB pointing at a -DMCD_KERNEL_PROFILE build): every rank's finishing CTA prints when its shard sums
were complete, when they were published to the peers and how long it then waited for the slowest rank.

    MCD_B200_LIB=scratch_ab/profile/libmcd_b200.so python -m torch.distributed.run --nproc-per-node 8 \
        --master-addr 127.0.0.1 tools/probe/exchange_timeline.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mcmc_dynamics_b200 import sharded, synthetic  # noqa: E402
from mcmc_dynamics_b200.analysis import ModelFit  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n_stars = int(os.environ.get('N_STARS', 10_000_000))
columns, truth = synthetic.mock_cluster(n_stars, seed=4, as_reader=False)
model = ModelFit(synthetic.reader_from_columns(sharded.shard_columns(columns, rank, world)), device=local)
del columns
model.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
model.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
like = sharded.ShardedLikelihood(model, fused=True, max_walkers=1024)
assert like.fused
theta = synthetic.initial_ball(truth, model.fitted_parameters, 512, seed=5)
for _ in range(8):
    out = like.lnprob(theta)          # host-buffer C ABI path: honours MCD_B200_LIB
dist.barrier()
if rank == 0:
    print('lnprob[0:2] =', out[:2], flush=True)
dist.destroy_process_group()
