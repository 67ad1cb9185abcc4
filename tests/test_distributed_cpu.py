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


class _GlooDist:
    """torch.distributed with the one collective gloo lacks (reduce_scatter) emulated by all_reduce + slice."""

    def __getattr__(self, name):
        return getattr(dist, name)

    @staticmethod
    def reduce_scatter_tensor(out, inp, group=None):
        t = inp.clone()
        dist.all_reduce(t, group=group)
        out.copy_(t.view(dist.get_world_size(group), -1)[dist.get_rank(group)])


def _stripe_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rtti_b200 import ops, stripe_parallel
    from tests import stripe_emu

    class GlooArena(stripe_emu.FakeArenaBase):
        """Halo rows travel by point-to-point messages (the peer stores of rtti_halo_exchange on the GPU)."""

        def exchange(self, pad, seq):
            assert seq == self.halo_seq
            rows = pad.shape[0] - 2
            reqs = []
            if self.rank > 0:
                reqs += [dist.isend(pad[1].clone(), self.rank - 1), dist.irecv(pad[0], self.rank - 1)]
            else:
                pad[0].zero_()
            if self.rank + 1 < self.world:
                reqs += [dist.isend(pad[rows].clone(), self.rank + 1), dist.irecv(pad[rows + 1], self.rank + 1)]
            else:
                pad[rows + 1].zero_()
            for r in reqs:
                r.wait()

    def reduce_over_ranks(key, val):
        t = val.clone()
        dist.all_reduce(t)
        return t

    for name, fn in stripe_emu.fake_ops(reduce_over_ranks).items():
        setattr(ops, name, fn)
    vae = stripe_emu.make_vae()
    h = wd = 8
    g = torch.Generator().manual_seed(5)
    zs = [torch.randn(1, 4, h, wd, generator=g) for _ in range(2)]
    wgt = torch.randn(1, 3, 8 * h, 8 * wd, generator=g)
    grad_fn = lambda img: torch.tanh(img) * wgt
    want = stripe_emu.autograd_reference(vae, zs, grad_fn)
    pad_bytes = stripe_parallel.stripe_pad_elems(vae.decoder, h // world, wd) * 4
    eng = stripe_parallel.StripedDecoderFwdBwd(vae, h, wd, "cpu", arena=GlooArena(world, rank, pad_bytes, dist.group.WORLD),
                                               dist=_GlooDist())
    ok, err = True, ""
    try:
        res = []
        for z in zs:
            img = eng.forward(z)
            res.append((img.clone(), eng.backward(grad_fn(img))))
        stripe_emu.assert_matches(res, want)
        gathered = [torch.empty_like(res[-1][1]) for _ in range(world)]
        dist.all_gather(gathered, res[-1][1])
        ok = all(torch.equal(gathered[0], x) for x in gathered)
    except Exception as e:   # noqa: BLE001 - reported to the parent
        ok, err = False, repr(e)
    q.put((rank, ok, err))
    dist.destroy_process_group()


def test_stripe_parallel_guidance_gloo_world2():
    """rtti_b200.stripe_parallel.StripedDecoderFwdBwd over real torch.distributed (gloo, 2 processes): collectives,
    group handling and the final broadcast, with the kernels emulated (tests/stripe_emu.py), against autograd."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stripe_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
