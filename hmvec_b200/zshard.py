"""Redshift-axis sharding over the GPUs of one box (one process per GPU, torch.distributed).

Every stage of the path is independent per redshift (sigma^2 per z, gradient along M only, profiles per (z,M), HOD
per z, mass integrals reduce over M), so each rank owns a contiguous z-slab and runs the single-GPU path on it with
no data-path collective.  The two places where redshifts couple:
  * the mthresh<->ngal bisection stops when EVERY z has converged (utils.py:26) -> one 8-byte AND all-reduce of the
    per-iteration pass masks (`ZComm.all_reduce_and`);
  * the Limber integral runs along z (cosmology.py:903) -> one all-gather of each P(k,z) slab (`ZComm.all_gather_z`),
    NCCL over NVLink on the GPUs (gloo on CPU tensors in the host-logic tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def slab_bounds(nz, world):
    """Contiguous near-equal slabs: rank r owns [b[r], b[r+1]); the first nz % world ranks get one extra z."""
    base, extra = divmod(int(nz), int(world))
    sizes = [base + (1 if r < extra else 0) for r in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def slab(nz, rank, world):
    b = slab_bounds(nz, world)
    return slice(int(b[rank]), int(b[rank + 1]))


class ZComm(object):
    """Communicator for a z-sharded run.  `group=None` uses the default process group."""

    def __init__(self, nz_total, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nz_total = int(nz_total)
        self.bounds = slab_bounds(self.nz_total, self.world)
        self.slab = slice(int(self.bounds[self.rank]), int(self.bounds[self.rank + 1]))
        self.nz_local = self.slab.stop - self.slab.start
        self._shifts = {}

    def all_reduce_and(self, mask):
        """In-place bitwise AND of a 1-element int64 tensor over the ranks (the global bisection stop condition).
        NCCL has no bitwise reductions, so the 64 bits travel as 64 int32 flags reduced with MIN."""
        sh = self._shifts.get(mask.device)
        if sh is None:
            sh = self._shifts.setdefault(mask.device, torch.arange(64, dtype=torch.int64, device=mask.device))
        bits = ((mask.reshape(1) >> sh) & 1).to(torch.int32)
        dist.all_reduce(bits, op=dist.ReduceOp.MIN, group=self.group)
        mask.copy_((bits.to(torch.int64) << sh).sum().reshape(mask.shape))
        return mask

    def all_gather_rows(self, local, out):
        """local: contiguous [nz_local, n] slab -> out: preallocated contiguous [nz_total, n], in place, no staging
        copies when every rank owns the same number of redshifts (the usual case); otherwise via all_gather_z."""
        if self.nz_total == self.world * self.nz_local and local.is_contiguous() and out.is_contiguous():
            dist.all_gather_into_tensor(out, local, group=self.group)
        else:
            out.copy_(self.all_gather_z(local))
        return out

    def all_gather_host(self, local, to_host=True):
        """numpy [nz_local, n] slab -> [nz_total, n] on every rank (what a user of the drop-in API does with the per-rank
        get_power results before the Limber integral): upload, one all-gather over NVLink, download.  to_host=False
        skips the download and returns the CUDA tensor, which C_kk / C_kg / C_yy accept in place of a numpy table."""
        from . import _capi as capi
        h = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        t = h.to(dev, non_blocking=h.is_pinned()) if dev.type == "cuda" else h
        out = torch.empty((self.nz_total, t.shape[1]), dtype=torch.float64, device=dev)
        self.all_gather_rows(t, out)
        if dev.type != "cuda":
            return out.numpy()
        capi.count_h2d(h.numel() * 8)
        if not to_host:
            return out
        capi.count_d2h(out.numel() * 8)
        res = torch.empty(out.shape, dtype=torch.float64, pin_memory=True)
        res.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return res.numpy()

    def all_gather_z(self, local):
        """local: [..., nz_local, n] slab (z is dim -2) -> [..., nz_total, n] on every rank."""
        if local.shape[-2] != self.nz_local:
            raise ValueError("slab has %d redshifts, this rank owns %d" % (local.shape[-2], self.nz_local))
        lead, n = tuple(local.shape[:-2]), local.shape[-1]
        nmax = int(np.max(np.diff(self.bounds)))
        # z-major staging so that each rank's contribution is one contiguous block
        send = torch.zeros((nmax,) + lead + (n,), dtype=local.dtype, device=local.device)
        send[:self.nz_local] = local.movedim(-2, 0)
        recv = torch.empty((self.world * nmax,) + lead + (n,), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        if self.nz_total == self.world * nmax:
            full = recv
        else:
            parts = [recv[r * nmax: r * nmax + int(self.bounds[r + 1] - self.bounds[r])] for r in range(self.world)]
            full = torch.cat(parts, dim=0)
        return full.movedim(0, -2).contiguous()
