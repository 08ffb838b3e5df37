"""Device time of the per-star kernel in model-curve mode (mcd_model_per_star_device) and end-to-end time of the
_calculate_lnlike hook (mcd_calculate_lnlike, host buffers) on a catalogue larger than L2; parity of their
composition against the fused likelihood kernel at that size."""
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np
import torch

from mcmc_dynamics_b200 import _native, synthetic
from mcmc_dynamics_b200.analysis import ModelFit

N = int(os.environ.get('N_STARS', 4000000))
data, truth = synthetic.mock_cluster(N, seed=4)
model = ModelFit(data)
model.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
model.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
packed = model.pack()
lib = _native.load_library()
theta = synthetic.initial_ball(truth, model.fitted_parameters, 1, seed=5)[0]
dev = torch.device('cuda:0')
th = torch.from_numpy(np.ascontiguousarray(theta)).to(dev)
v_los = torch.empty(N, dtype=torch.float64, device=dev)
sigma_los = torch.empty(N, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream


def curves():
    rc = lib.mcd_model_per_star_device(packed.handle, th.data_ptr(), v_los.data_ptr(), sigma_los.data_ptr(), stream)
    assert rc == 0, lib.mcd_last_error()


for _ in range(3):
    curves()
torch.cuda.synchronize()
reps = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    curves()
e1.record()
torch.cuda.synchronize()
us = 1e3 * e0.elapsed_time(e1) / reps
info = packed.info()
moved = N * (info['bytes_per_star'] + 16)
print('per_star_kernel, model-curve mode: %d stars, %.1f us per launch, %d B/star read + 16 written = %.0f GB/s'
      % (N, us, info['bytes_per_star'], moved / us / 1e3))

v_host, s_host = v_los.cpu().numpy(), sigma_los.cpu().numpy()
got = model._calculate_lnlike(v_host, s_host)
t0 = time.perf_counter()
for _ in range(3):
    got = model._calculate_lnlike(v_host, s_host)
hook_ms = 1e3 * (time.perf_counter() - t0) / 3
want = model.lnlike(theta)
print('_calculate_lnlike(rotation_model, dispersion_model) = %.12g, fused kernel lnlike = %.12g, rel. diff %.2e'
      % (got, want, abs(got - want) / abs(want)))
print('mcd_calculate_lnlike, host buffers (2 x %d doubles in, pageable): %.2f ms per call' % (N, hook_ms))
