"""CPU, gloo, world_size 2: the region-parallel exchange (plan + all-gather + replicated blend inputs)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rtti_b200.region_parallel import RegionParallelPlan
    passes = [dict(kind=k) for k in "ABCDEEEE"]
    plan = RegionParallelPlan(passes, True)
    full = torch.arange(len(passes), dtype=torch.float32)[:, None, None, None] * torch.ones(1, 4, 8, 8)
    ok = True
    for feat in (True, False):
        local = plan.local_passes(feat)
        if feat and any(passes[p]["kind"] == "E" for p in local):
            ok &= 3 in local
            src = plan.injection_sources(local)
            ok &= all(src[k] == local.index(3) for k, p in enumerate(local) if passes[p]["kind"] == "E")
        eps_local = full[local] + 0.0
        got = plan.gather(eps_local, local, feat)
        ok &= bool(torch.equal(got, full))
    q.put((rank, ok, plan.local_passes(True)))
    dist.destroy_process_group()


def test_region_parallel_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    locals_ = dict((r, l) for r, _, l in res)
    assert sorted(set(locals_[0]) | set(locals_[1])) == list(range(8))
