"""Star-sharded likelihood across the GPUs of one box.

Per-star terms are independent and ``lnlike_w = sum_i term(w, i)``, so rank ``g`` of ``G`` packs the
contiguous star range ``[g N / G, (g + 1) N / G)`` of the catalogue, evaluates ALL walkers of a call
over its shard (``mcd_lnprob_partial_device``) and the per-walker partial sums -- ``n_walkers``
float64 values, 4 KiB for 512 walkers -- are combined with one ``all_reduce(SUM)`` over NCCL
(NVLink 5 / NVSwitch).  A walker outside its box prior is ``-inf`` on every rank, so the sum is
``-inf`` too.  This replaces the walker-parallel process pool of the reference
(``analysis/runner.py:398-403``) at multi-GPU scale.

``torch.distributed`` is plumbing: process group, NCCL communicator, device tensors.

With ``fused=True`` (default on CUDA) the collective is not a separate launch at all: the last CTA of
the likelihood kernel publishes the shard's sums into every rank's exchange buffer through NVLink
peer mappings (torch symmetric memory supplies the mappings), waits for the other shards and adds
them in rank order (``mcd_lnprob_allreduce_device``).  Every GPU then holds bit-identical values.
"""
import ctypes
import logging

import numpy as np

from . import _native

logger = logging.getLogger(__name__)


def shard_range(n_stars, rank, world_size):
    """Contiguous star range of `rank`: sizes differ by at most one star."""
    base, extra = divmod(int(n_stars), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_columns(columns, rank, world_size):
    n = len(next(iter(columns.values())))
    lo, hi = shard_range(n, rank, world_size)
    return {name: np.ascontiguousarray(values[lo:hi]) for name, values in columns.items()}


class ShardedLikelihood(object):
    """``lnprob`` of a model whose catalogue is split over the ranks of a process group.

    Parameters
    ----------
    model : a model object built on THIS rank's shard of the catalogue (device = this rank's GPU)
    group : torch.distributed process group or None (default group); with world size 1 no
        collective is issued.
    """

    def __init__(self, model, group=None, device=None, fused=True, max_walkers=4096):
        import torch
        import torch.distributed as dist
        self._torch = torch
        self._dist = dist
        self.model = model
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        # `device` is overridden only by the CPU (gloo) tests of the reduction logic
        self.device = torch.device('cuda', model.device) if device is None else torch.device(device)
        self._pinned_in = None
        self._pinned_out = None
        self.max_walkers = int(max_walkers)
        self.fused = False
        self._exchange = None
        if fused and self.world_size > 1 and self.device.type == 'cuda':
            self.fused = self._attach_exchange()

    def _attach_exchange(self):
        """Allocate the exchange buffer in symmetric memory, map the peers and hand the addresses to
        the handle.  Any failure (no peer access, old torch) falls back to the NCCL all-reduce."""
        torch, dist = self._torch, self._dist
        ok = True
        try:
            import torch.distributed._symmetric_memory as symm_mem
            lib = _native.load_library()
            nbytes = ctypes.c_int64()
            _native.check(lib.mcd_exchange_bytes(self.world_size, self.max_walkers, ctypes.byref(nbytes)))
            buf = symm_mem.empty((int(nbytes.value) // 8,), dtype=torch.float64, device=self.device)
            hdl = symm_mem.rendezvous(buf, group=self.group if self.group is not None else dist.group.WORLD)
            buf.zero_()
            torch.cuda.synchronize(self.device)
            ptrs = (ctypes.c_uint64 * self.world_size)(*[int(p) for p in hdl.buffer_ptrs])
            _native.check(lib.mcd_exchange_attach(self.model.pack().handle, int(hdl.rank), self.world_size, ptrs,
                                                  self.max_walkers))
            self._exchange = (buf, hdl)
        except Exception as exc:                      # noqa: BLE001 -- any failure means "use NCCL"
            logger.warning('fused cross-GPU reduction unavailable (%s); using the NCCL all-reduce', exc)
            ok = False
        # all ranks must agree, and nobody may publish before every buffer is zeroed
        flag = torch.tensor([1 if ok else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(flag.item())

    def lnprob_tensor(self, theta):
        """theta: [n_walkers, n_free] float64 CUDA tensor, identical on every rank.  Returns the
        full-catalogue lnprob on every rank (asynchronous on the current stream)."""
        packed = self.model.pack()
        if self.fused and theta.shape[0] <= self.max_walkers:
            return _native.load_torch_ops().lnprob_allreduce(packed.handle.value, theta)
        partial = packed.lnprob_partial_tensor(theta)
        if self.world_size > 1:
            self._dist.all_reduce(partial, op=self._dist.ReduceOp.SUM, group=self.group)
        return partial

    def device_sampler(self, n_walkers, seed, a=2.0):
        """Device-resident stretch-move sampler over the shards (fused mode only): every rank holds a
        replica of the ensemble, replays the same CUDA graph with the SAME `seed`, and the fused
        half-step kernels exchange the shard sums before accepting in place, so all replicas stay
        bit-identical without any further communication.  A collective: call ``run_mcmc`` with the
        same arguments on every rank."""
        if self.world_size > 1 and not self.fused:
            raise RuntimeError('the sharded device sampler needs the fused cross-GPU reduction')
        if n_walkers > self.max_walkers:
            raise ValueError('n_walkers exceeds max_walkers of the exchange buffer')
        from .sampler import DeviceEnsembleSampler
        return DeviceEnsembleSampler(n_walkers, self.model.n_fitted_parameters, self.model.pack(), a=a, seed=seed)

    def lnprob(self, theta):
        """Host array in, host array out (what a host sampler calls on every rank).  Fused mode: ONE C-ABI
        call, ``mcd_lnprob_allreduce`` -- pinned copy-in and the shard kernel with the cross-GPU exchange in its
        tail replayed as one CUDA graph, results written by the kernel into pinned memory.  NCCL mode: pinned staging, H2D, shard kernel,
        ``all_reduce``, D2H through torch."""
        torch = self._torch
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        n, p = theta.shape
        if self.fused and n <= self.max_walkers:
            packed = self.model.pack()
            if p != packed.n_theta:
                raise ValueError('theta must have shape (n_walkers, {0}), got {1}'.format(packed.n_theta, theta.shape))
            out = np.empty(n, dtype=np.float64)
            rc = packed._lib.mcd_lnprob_allreduce(packed.handle, _native.address(theta), n, _native.address(out))
            if rc != 0:
                _native.check(rc)
            return out
        on_gpu = self.device.type == 'cuda'
        if self._pinned_in is None or self._pinned_in.shape[0] < n or self._pinned_in.shape[1] != p:
            self._pinned_in = torch.empty((max(n, 64), p), dtype=torch.float64)
            self._pinned_out = torch.empty((max(n, 64),), dtype=torch.float64)
            if on_gpu:
                self._pinned_in = self._pinned_in.pin_memory()
                self._pinned_out = self._pinned_out.pin_memory()
        self._pinned_in[:n].copy_(torch.from_numpy(theta))
        dev = self._pinned_in[:n].to(self.device, non_blocking=True)
        out = self.lnprob_tensor(dev)
        self._pinned_out[:n].copy_(out, non_blocking=True)
        if on_gpu:
            torch.cuda.current_stream(self.device).synchronize()
        return self._pinned_out[:n].numpy().copy()
