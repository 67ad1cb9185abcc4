"""Region-parallel sharding of the UNet passes of one denoising step across GPUs.

New functionality relative to the reference (which is single-GPU: models/region_diffusion_sdxl.py:779-821
runs the passes one after another). Within a step the passes A (uncond), B (base prompt + font sizes),
C/D (reference latent, uncond / base) and E_1..E_{N-1} (regions) are independent UNet evaluations, except
that on feature-injection steps every E pass consumes the self-attention Q/K and one resnet feature of
pass D (:1018-1061). So:

  * one process per GPU (torch.distributed, NCCL over NVLink); every rank holds the full weights;
  * each rank runs a subset of the passes as one batched UNet call; D is REPLICATED on every rank that
    owns an E pass on injection steps (0.4 GB of Q/K per step would have to cross NVLink otherwise);
  * one all-gather of the per-pass noise predictions ([4,h,w] fp16 = 128 KB each at 1024^2) per step, then
    the blend + CFG + scheduler update is replicated on every rank — it is deterministic, so the
    latents stay bit-identical across ranks without a broadcast.
"""
from typing import List

import torch


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def assign_passes(kinds: List[str], world: int, feat_inject: bool):
    """kinds: per pass 'A','B','C','D','E'. Returns (per-rank list of pass indices, owner rank per pass).
    Minimises the maximum number of passes per rank, counting the replicated D on E-owning ranks."""
    n = len(kinds)
    loads = [0] * world
    has_d = [False] * world
    assign = [[] for _ in range(world)]
    owner = [-1] * n
    d_idx = kinds.index("D") if "D" in kinds else -1
    need_d = feat_inject and d_idx >= 0
    order = [i for i, k in enumerate(kinds) if k == "E"] + [i for i, k in enumerate(kinds) if k not in ("E", "D")]
    for i in order:
        best, best_cost = 0, None
        for r in range(world):
            extra = 1 if (need_d and kinds[i] == "E" and not has_d[r]) else 0
            cost = loads[r] + 1 + extra
            if best_cost is None or cost < best_cost:
                best, best_cost = r, cost
        if need_d and kinds[i] == "E" and not has_d[best]:
            has_d[best] = True
            assign[best].append(d_idx)
            loads[best] += 1
            if owner[d_idx] < 0:
                owner[d_idx] = best
        assign[best].append(i)
        owner[i] = best
        loads[best] += 1
    if d_idx >= 0 and owner[d_idx] < 0:
        r = min(range(world), key=lambda q: loads[q])
        assign[r].append(d_idx)
        owner[d_idx] = r
        loads[r] += 1
    return [sorted(a) for a in assign], owner


class RegionParallelPlan:
    def __init__(self, passes, inject, group=None):
        self.passes = passes
        self.kinds = [p["kind"] for p in passes]
        self.inject = inject
        d = _dist()
        self.dist = d
        self.group = group
        self.world = d.get_world_size(group) if d else 1
        self.rank = d.get_rank(group) if d else 0
        self._cache = {}

    def _plan(self, feat_inject):
        key = bool(feat_inject and self.inject)
        if key not in self._cache:
            self._cache[key] = assign_passes(self.kinds, self.world, key)
        return self._cache[key]

    def local_passes(self, feat_inject):
        return self._plan(feat_inject)[0][self.rank]

    def injection_sources(self, local):
        """For each local batch entry the local index whose self-attention / resnet feature it uses:
        E entries point at the local copy of D, everything else at itself."""
        d_pos = [k for k, p in enumerate(local) if self.kinds[p] == "D"]
        out = []
        for k, p in enumerate(local):
            if self.kinds[p] == "E":
                assert d_pos, "an E pass on a feature-injection step needs pass D in the same batch"
                out.append(d_pos[0])
            else:
                out.append(k)
        return out

    def gather(self, eps_local, local, feat_inject):
        """eps_local [len(local), ...] -> eps of ALL passes in pass order, on every rank."""
        n = len(self.passes)
        if self.world == 1:
            if list(local) == list(range(n)):
                return eps_local
            out = eps_local.new_empty((n,) + tuple(eps_local.shape[1:]))
            out[torch.as_tensor(local, device=eps_local.device)] = eps_local
            return out
        assign, owner = self._plan(feat_inject)
        m = max(len(a) for a in assign)
        buf = eps_local.new_zeros((m,) + tuple(eps_local.shape[1:]))
        buf[: eps_local.shape[0]] = eps_local
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        self.dist.all_gather(parts, buf.contiguous(), group=self.group)
        out = eps_local.new_empty((n,) + tuple(eps_local.shape[1:]))
        for p in range(n):
            r = owner[p]
            out[p] = parts[r][assign[r].index(p)]
        return out


class PeerExchange:
    """Symmetric (peer-mapped) slot buffers for the fused gather+blend kernel (csrc/gather_blend.cu).

    Layout per rank: [256 B header: uint32 step flag][fp16 slots [2 parities][n_slots][n]].  The buffers are
    allocated with torch's symmetric memory (CUDA IPC / multicast-capable allocation, NVLink peer access);
    PyTorch is used for the rendezvous only — the exchange itself is the kernel's peer loads.
    Slot order: 0 = A (uncond), 1..N-1 = E_1..E_{N-1}, N = B (base), N+1 = C, N+2 = D."""
    HEADER = 256

    def __init__(self, passes, n, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.dist = dist
        self.group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n = n
        n_regions = sum(1 for p in passes if p["kind"] == "E") + 1
        self.n_regions = n_regions
        has_ref = any(p["kind"] == "C" for p in passes)
        self.n_slots = n_regions + 1 + (2 if has_ref else 0)
        self.slot_of_pass = []
        for p in passes:
            k = p["kind"]
            self.slot_of_pass.append({"A": 0, "B": n_regions, "C": n_regions + 1, "D": n_regions + 2}.get(k, 1 + p.get("region", 0)))
        nbytes = self.HEADER + 2 * self.n_slots * n * 2
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.handle = symm.rendezvous(self.buf, self.group.group_name)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(self.group)
        base = [int(p) for p in self.handle.buffer_ptrs]
        self.flag_ptrs = base
        self.slot_ptrs = [b + self.HEADER for b in base]
        self.slots = self.buf[self.HEADER:].view(torch.float16).view(2, self.n_slots, n)
        self.step_id = 0

    def publish(self, eps_local, local, owner):
        """Copy the noise predictions of the passes this rank OWNS into its slots for the next step id."""
        self.step_id += 1
        par = self.step_id & 1
        for k, p in enumerate(local):
            if owner[p] == self.rank:
                self.slots[par, self.slot_of_pass[p]].copy_(eps_local[k].reshape(-1))
        return self.step_id

    def check(self):
        """Raise if a fused exchange timed out waiting for a peer (error word set by the kernel)."""
        err = int(self.buf[4:8].view(torch.int32).item())
        if err != 0:
            raise RuntimeError("rtti_gather_blend_step: timed out waiting for a peer rank's noise predictions")

    def slot_owner(self, owner):
        out = [0] * self.n_slots
        for p, s in enumerate(self.slot_of_pass):
            out[s] = owner[p]
        return out
