"""CPU model of the RemoteQK hand-off protocol (region_parallel.RemoteQK + csrc/peer_push.cu), threads as ranks.

What is modelled — the ordering rules the CUDA side relies on, not its arithmetic:
  * the producer (the rank that runs pass D) writes event e of a pass into every consumer's SINGLE receive region e and
    then publishes `base + e + 1` to the consumer's flag word; a consumer waits until its flag reaches ITS OWN
    `base + e + 1` and then reads region e;
  * every rank of the group advances its sequence base by the number of events at the end of every injection-step
    pass, whatever its role (a rank without a role in one sampling call may be a consumer in the next);
  * the only synchronisation between steps is the gather+blend exchange: every owner publishes its step flag after its
    pass, every rank waits for all owners before it starts the next step.
The test drives several "sampling calls" with different pass assignments (roles change between calls) under random
delays and checks that every consumer read returns exactly the producer's data of the same step and event — i.e. no
stale data (flag satisfied early), no overwritten data (single buffer re-used too soon) and no dead-lock.
"""
import random
import threading
import time

import pytest

from rtti_b200.region_parallel import RegionParallelPlan

N_EVENTS = 7          # events per pass (71 for SDXL; the protocol does not depend on the number)
TIMEOUT = 20.0


class World:
    def __init__(self, n):
        self.n = n
        self.flag = [0] * n                 # arrival flag of rank r (written by the producer)
        self.base = [0] * n                 # sequence base of rank r (device-side word; advanced by its own stream)
        self.region = [[None] * N_EVENTS for _ in range(n)]   # single-buffered receive regions of rank r
        self.step_flag = [0] * n            # gather+blend: step published by rank r
        self.errors = []
        self.lock = threading.Lock()


def wait_until(pred, what, world):
    t0 = time.time()
    while not pred():
        if time.time() - t0 > TIMEOUT:
            with world.lock:
                world.errors.append(f"timeout: {what}")
            return False
        time.sleep(0.0002)
    return True


def rank_main(world, rank, calls, seed):
    rng = random.Random(seed * 1000 + rank)
    jitter = lambda: time.sleep(rng.random() * 0.002) if rng.random() < 0.5 else None
    step_id = 0
    for call_id, kinds in enumerate(calls):
        plan = RegionParallelPlan([dict(kind=k) for k in kinds], True, remote_qk=True)
        plan.world, plan.rank = world.n, rank
        local = plan.local_passes(True)
        role = plan.remote_role(local) if local else None
        assign, owner = plan._plan(True)
        owners = sorted(set(owner))
        for step in range(3):                                # three injection steps per sampling call
            step_id += 1
            tag = (call_id, step)
            jitter()
            if role is not None and role[0] == "src":
                for e in range(N_EVENTS):
                    jitter()
                    for d in role[2]:
                        world.region[d][e] = (tag, e)        # peer stores of the slab ...
                    for d in role[2]:
                        world.flag[d] = world.base[rank] + e + 1   # ... then the release store of the event number
            elif role is not None and role[0] == "dst":
                for e in range(N_EVENTS):
                    want = world.base[rank] + e + 1
                    if not wait_until(lambda: world.flag[rank] >= want, f"rank {rank} event {e} of {tag}", world):
                        return
                    jitter()                                 # the attention kernel reads the region some time later
                    got = world.region[rank][e]
                    if got != (tag, e):
                        with world.lock:
                            world.errors.append(f"rank {rank} read {got} for event {e} of {tag}")
            world.base[rank] += N_EVENTS                     # end_pass(): every rank, whatever its role
            # gather+blend exchange: owners publish, everyone waits for all owners (the only inter-step synchronisation)
            jitter()
            if rank in owners:
                world.step_flag[rank] = step_id
            sid = step_id
            if not wait_until(lambda: all(world.step_flag[o] >= sid for o in owners), f"rank {rank} exchange {tag}", world):
                return


@pytest.mark.parametrize("world_size,seed", [(2, 1), (4, 2), (8, 3), (8, 4)])
def test_remote_qk_protocol_under_random_skew(world_size, seed):
    # sampling calls with different region counts: the producer and the consumer set change between calls, and ranks
    # without any role in one call (world 8, 5 passes + ...) become consumers in the next
    calls = [list("ABCD") + ["E"] * 4, list("ABCD") + ["E"] * 9, list("ABCD") + ["E"] * 2, list("ABCD") + ["E"] * 7]
    world = World(world_size)
    threads = [threading.Thread(target=rank_main, args=(world, r, calls, seed)) for r in range(world_size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(TIMEOUT * 4)
    assert not any(t.is_alive() for t in threads), "dead-lock"
    assert not world.errors, world.errors[:5]


def test_bases_must_move_on_every_rank():
    """The rule the emulation above depends on, shown by breaking it: if ranks without a role skip the advance, a rank
    that becomes a consumer in a later call compares against a stale base and accepts an event before its data arrives."""
    world_size = 8
    calls = [list("ABCD") + ["E"] * 2, list("ABCD") + ["E"] * 7]     # ranks 2.. have no E pass in call 0, some do in call 1
    # static analysis of the numbers on the wire instead of a race: after call 0 (3 steps) the producer's base is
    # 3 * N_EVENTS; a rank that never advanced expects event 1 of call 1 as base 0 + 1 and would be satisfied by ANY flag
    # value >= 1 — e.g. the previous event's — before the matching data has been written.
    roles = []
    for c, kinds in enumerate(calls):
        row = []
        for r in range(world_size):
            plan = RegionParallelPlan([dict(kind=k) for k in kinds], True, remote_qk=True)
            plan.world, plan.rank = world_size, r
            local = plan.local_passes(True)
            row.append(plan.remote_role(local) if local else None)
        roles.append(row)
    late_consumers = [r for r in range(world_size) if roles[0][r] is None and roles[1][r] is not None and roles[1][r][0] == "dst"]
    assert late_consumers, "the scenario needs a rank that is idle in the first call and a consumer in the second"
    producer_base_after_call0 = 3 * N_EVENTS
    stale_expectation = 0 + 2                      # what such a rank would wait for at its event index 1 without advancing
    first_publish_of_call1 = producer_base_after_call0 + 1
    assert first_publish_of_call1 >= stale_expectation, "event 0's flag already satisfies the wait for event 1: data race"
