"""Redshift-axis sharding over the GPUs of one box (one process per GPU, torch.distributed).

Every stage of the path is independent per redshift (sigma^2 per z, gradient along M only, profiles per (z,M), HOD
per z, mass integrals reduce over M), so each rank owns a contiguous z-slab and runs the single-GPU path on it with
no data-path collective.  The two places where redshifts couple:
  * the mthresh<->ngal bisection stops when EVERY z has converged (utils.py:26) -> one 8-byte AND all-reduce of the
    per-iteration pass masks (`ZComm.all_reduce_and`);
  * the Limber integral runs along z (cosmology.py:903) -> one all-gather of the P(k,z) slabs.  On the GPUs of one
    box this is `PeerGather`: the kernel that forms the tables stores them straight into every rank's gathered table
    over NVLink (peer memory through CUDA IPC) and raises a step flag -- no collective launch.  NCCL
    (`ZComm.all_gather_rows` / `all_gather_z`) is the fallback (HMV_PEER_GATHER=0, or no peer access); gloo on CPU
    tensors in the host-logic tests.
"""
import ctypes as C
import os

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
        self._peer = {}

    def peer_gather(self, ncols):
        """The `PeerGather` for gathered tables of `ncols` doubles per redshift (created collectively on first use --
        every rank must ask for it at the same point), or None when peer access is unavailable or switched off."""
        if ncols not in self._peer:
            self._peer[ncols] = PeerGather.create(self, ncols)
        return self._peer[ncols]

    def close(self):
        """Collective: unmap and free the peer tables (after a barrier, so that no rank is still storing into them)."""
        if any(pg is not None for pg in self._peer.values()):
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
        for pg in self._peer.values():
            if pg is not None:
                pg.close()
        self._peer = {}

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
        pg = self.peer_gather(int(local.shape[1])) if local.is_cuda and local.is_contiguous() else None
        if pg is not None:
            out.copy_(pg.gather([local], None).view(self.nz_total, -1))
        elif self.nz_total == self.world * self.nz_local and local.is_contiguous() and out.is_contiguous():
            dist.all_gather_into_tensor(out, local, group=self.group)
        else:
            out.copy_(self.all_gather_z(local))
        return out

    def all_gather_host(self, local, to_host=True):
        """[nz_local, n] slab -> [nz_total, n] on every rank (what a user of the drop-in API does with the per-rank
        get_power results before the Limber integral).  `local` is a numpy array (upload, gather over NVLink, download)
        or the CUDA tensor HaloModel.get_power_device returns (no upload).  to_host=False skips the download and
        returns the CUDA tensor, which C_kk / C_kg / C_yy accept in place of a numpy table."""
        from . import _capi as capi
        if isinstance(local, torch.Tensor) and local.is_cuda:
            # already on the device (HaloModel.get_power_device): nothing to upload
            t, dev, h = local.to(torch.float64).contiguous(), local.device, None
        else:
            h = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
            t = h.to(dev, non_blocking=h.is_pinned()) if dev.type == "cuda" else h
        out = torch.empty((self.nz_total, t.shape[1]), dtype=torch.float64, device=dev)
        self.all_gather_rows(t, out)
        if dev.type != "cuda":
            return out.numpy()
        if h is not None:
            capi.count_h2d(h.numel() * 8)
        if not to_host:
            return out
        capi.count_d2h(out.numel() * 8)
        res = torch.empty(out.shape, dtype=torch.float64, pin_memory=True)
        res.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return res.numpy()

    def all_gather_tables(self, tables):
        """Several CUDA [nz_local, n] slabs (HaloModel.get_power_device) -> their [nz_total, n] tables on every rank in
        ONE exchange (one hmv_peer_scatter + hmv_peer_wait for all of them); one gather per table without peer access."""
        tables = [t.to(torch.float64).contiguous() for t in tables]
        n = int(tables[0].shape[1])
        pg = self.peer_gather(len(tables) * n) if (len(tables) <= 4 and all(t.is_cuda and t.shape == tables[0].shape
                                                                             for t in tables)) else None
        if pg is None:
            return [self.all_gather_host(t, to_host=False) for t in tables]
        full = pg.gather(tables, None)                       # [nz_total, nsp, n], valid until the gather after next
        return [full[:, s, :].contiguous() for s in range(len(tables))]

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


class _DevView(object):
    """A raw device allocation seen by torch (zero-copy, through the CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class PeerGather(object):
    """All-gather of z-slab tables by peer stores (hmv_peer_scatter / hmv_peer_wait, csrc/k_peer.cu).

    Each rank owns two [nz_total][ncols] tables (used alternately) and a flag array in one cudaMalloc'ed block that
    every other rank of the box maps through CUDA IPC.  `gather(a, b)` queues, on the current stream, one kernel that
    writes a (+ b) for this rank's rows into all tables and one that waits until every rank's rows of this step have
    arrived; it returns the local table of this step as a CUDA tensor, valid until the gather after next."""

    TIMEOUT_S = 20.0

    def __init__(self):
        self.base = None

    @classmethod
    def create(cls, zc, ncols):
        from . import _capi as capi
        if os.environ.get("HMV_PEER_GATHER", "1") == "0" or zc.world > 16 or not torch.cuda.is_available():
            return None
        if dist.get_backend(zc.group) != "nccl":
            return None
        self = cls()
        self.zc, self.ncols = zc, int(ncols)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.table_doubles = zc.nz_total * self.ncols
        self.flag_off = ((2 * self.table_doubles * 8 + 255) // 256) * 256
        nbytes = self.flag_off + 256
        base, handle = C.c_void_p(), C.create_string_buffer(64)
        ok = capi.lib.hmv_peer_alloc(nbytes, C.byref(base), handle) == 0
        handles = [None] * zc.world
        dist.all_gather_object(handles, handle.raw if ok else None, group=zc.group)
        ptrs = [None] * zc.world
        if ok and all(h is not None for h in handles):
            for r, h in enumerate(handles):
                if r == zc.rank:
                    ptrs[r] = base.value
                    continue
                q = C.c_void_p()
                if capi.lib.hmv_peer_open(h, C.byref(q)) != 0:
                    ok = False
                    break
                ptrs[r] = q.value
        else:
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=zc.group)        # all or nothing (also the start barrier)
        self.base, self.ptrs = base.value, ptrs
        if int(flag.item()) == 0:
            self.close()
            return None
        self.step = 0
        self.done = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._flags = (C.c_void_p * zc.world)(*[p + self.flag_off for p in ptrs])
        self._bufs = [(C.c_void_p * zc.world)(*[p + 8 * half * self.table_doubles for p in ptrs]) for half in (0, 1)]
        self._views = [torch.as_tensor(_DevView(self.base + 8 * half * self.table_doubles, self.table_doubles),
                                       device=self.device) for half in (0, 1)]
        return self

    def gather(self, a, b=None):
        """a, b: lists of nsp contiguous float64 CUDA tensors [nz_local, ncols // nsp] (b or its entries may be None).
        Returns the gathered [nz_total, nsp, ncols // nsp] table (row r of spectrum s = a[s][r] + b[s][r])."""
        from . import _capi as capi
        nsp, zc = len(a), self.zc
        n = int(a[0].shape[1])
        if nsp * n != self.ncols or any(t.shape != (zc.nz_local, n) for t in a):
            raise ValueError("PeerGather: tables must be %d x [%d, %d]" % (nsp, zc.nz_local, self.ncols // max(nsp, 1)))
        self.step += 1
        half = self.step & 1
        pa = (C.c_void_p * nsp)(*[capi.ptr(t).value for t in a])
        pb = (C.c_void_p * nsp)(*[capi.ptr(t).value if t is not None else None for t in b]) if b is not None else None
        st = capi.stream()
        capi.check(capi.lib.hmv_peer_scatter(zc.nz_local, n, nsp, pa, pb, zc.world, zc.rank, self._bufs[half],
                                             self._flags, int(zc.bounds[zc.rank]), self.step,
                                             C.c_void_p(self.done.data_ptr()), st), "hmv_peer_scatter")
        capi.check(capi.lib.hmv_peer_wait(C.c_void_p(self.base + self.flag_off), zc.world, self.step, self.TIMEOUT_S,
                                          C.c_void_p(self.status.data_ptr()), st), "hmv_peer_wait")
        return self._views[half].view(zc.nz_total, nsp, n)

    def check(self):
        """Synchronising: raise if a wait gave up (a rank never delivered its rows)."""
        s = int(self.status.item())
        if s:
            raise RuntimeError("PeerGather: rank %d's rows did not arrive within %.0f s" % (s - 1, self.TIMEOUT_S))

    def close(self):
        from . import _capi as capi
        if self.base is None:
            return
        try:
            torch.cuda.synchronize()
            for r, p in enumerate(self.ptrs):
                if p is not None and r != self.zc.rank:
                    capi.lib.hmv_peer_close(C.c_void_p(p))
            capi.lib.hmv_peer_free(C.c_void_p(self.base))
        except Exception:
            pass
        self.base = None
