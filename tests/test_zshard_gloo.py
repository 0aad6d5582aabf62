"""z-sharding host logic with two ranks over gloo (CPU): slab ownership, the padded all-gather of P(k,z) slabs and the
AND all-reduce of the bisection pass masks."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, nz, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hmvec_b200 import zshard
        zc = zshard.ZComm(nz)
        full = torch.arange(4 * nz * 6, dtype=torch.float64).reshape(4, nz, 6)
        got = zc.all_gather_z(full[:, zc.slab].contiguous())
        ok_gather = bool(torch.equal(got, full))
        got2 = zc.all_gather_z(full[0, zc.slab].contiguous())
        ok_gather2 = bool(torch.equal(got2, full[0]))
        out = torch.empty_like(full[1])
        zc.all_gather_rows(full[1, zc.slab].contiguous(), out)
        ok_gather2 = ok_gather2 and bool(torch.equal(out, full[1]))
        if nz % world == 0:
            # the Limber hand-over of the pipeline: the spectra packed z-major [nz_local][nq][n] -> ONE all-gather ->
            # [nz_total][nq][n], every spectrum a table with row stride nq n
            nzl, n = zc.nz_local, full.shape[-1]
            pack = torch.stack([full[i, zc.slab] for i in range(4)], dim=1).contiguous()
            allp = torch.empty((nz, 4, n), dtype=torch.float64)
            zc.all_gather_rows(pack.view(nzl, 4 * n), allp.view(nz, 4 * n))
            ok_gather2 = ok_gather2 and all(bool(torch.equal(allp[:, i], full[i])) for i in range(4))
        # the drop-in API's hand-over: per-rank numpy slabs -> the full numpy table on every rank (even and ragged slabs)
        host = zc.all_gather_host(full[2, zc.slab].numpy())
        ok_gather2 = ok_gather2 and isinstance(host, np.ndarray) and bool(np.array_equal(host, full[2].numpy()))
        # several tables in one call (the API's Limber hand-over); without peer access this is one gather per table,
        # and ZComm.close() is a no-op
        tabs = zc.all_gather_tables([full[i, zc.slab].contiguous() for i in (0, 3)])
        ok_gather2 = ok_gather2 and all(bool(np.array_equal(np.asarray(t), full[i].numpy())) for t, i in zip(tabs, (0, 3)))
        ok_gather2 = ok_gather2 and zc.peer_gather(6) is None
        zc.close()
        # rank 0 passes from iteration 3 on, rank 1 from iteration 5 on -> global first pass = iteration 5
        mask = torch.tensor([(~0) << (3 if rank == 0 else 5)], dtype=torch.int64)
        zc.all_reduce_and(mask)
        first = (int(mask.item()) & -int(mask.item())).bit_length() - 1
        q.put((rank, zc.slab.start, zc.slab.stop, ok_gather, ok_gather2, first))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nz", [7, 8])
def test_two_rank_gloo(nz):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + nz
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nz, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] == 0 and out[0][2] == out[1][1] and out[1][2] == nz
    assert all(o[3] and o[4] for o in out)
    assert [o[5] for o in out] == [5, 5]
