"""Host-side logic of the time sharding (grb200/sharding.py) under torch.distributed with the gloo backend,
world_size 2 and 3, on CPU tensors: halo exchange + loop-state ring.  The data path is replaced by a tiny
CPU stand-in with the same structure as the chain -- a finite-memory front (moving sum over `halo` rows) and
an infinite-memory tail (a first-order recurrence) -- so that a wrong halo, a wrong state order or a missing
hand-off changes the result.  The sharded run must equal the sequential one exactly (integer arithmetic)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))

ROWS, COLS, HALO, STEPS = 64, 5, 7, 3


def stream(world):
    g = torch.Generator().manual_seed(7)
    return torch.randint(-50, 50, (world * STEPS * ROWS, COLS), generator=g, dtype=torch.int64)


def front(rows_with_halo):
    """moving sum over HALO+1 rows: rows_with_halo = [HALO + n][COLS] -> [n][COLS]"""
    c = torch.cumsum(torch.cat([torch.zeros(1, COLS, dtype=torch.int64), rows_with_halo]), 0)
    return c[HALO + 1:] - c[:-(HALO + 1)]


def tail(f, state):
    """y[i] = 3*y[i-1] + f[i] (mod 2^31-1), state = y[-1]"""
    out = torch.empty_like(f)
    s = state.clone()
    for i in range(f.shape[0]):
        s = (3 * s + f[i]) % 2147483647
        out[i] = s
    return out, s


def sequential(world):
    x = stream(world)
    xh = torch.cat([torch.zeros(HALO, COLS, dtype=torch.int64), x])
    y, _ = tail(front(xh), torch.zeros(COLS, dtype=torch.int64))
    return y


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")   # the container hostname may not resolve
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from grb200 import sharding
    plan = sharding.TimeShardPlan(world, rank, ROWS, HALO)
    ring = sharding.RingExchanger(plan)
    x = stream(world)
    state = torch.zeros(COLS, dtype=torch.int64)
    outs = []
    prev_tail = torch.zeros(HALO, COLS, dtype=torch.int64)   # what this rank sends first: unused by step 0 / rank 0
    last = STEPS - 1
    for s in range(STEPS):
        b = plan.block_index(s)
        assert plan.abs_start(s) == b * ROWS
        mine = x[b * ROWS:(b + 1) * ROWS]
        halo_in = torch.zeros(HALO, COLS, dtype=torch.int64)
        works = ring.exchange_halo(mine[-HALO:].contiguous(), halo_in, s)
        ring.wait_all(works)
        # rank 0 receives the halo of the LAST rank's block of the same step, which is not its left neighbour in
        # time: block b-1 belongs to rank world-1 at step s-1 -> it must keep that one from the previous step
        if rank == 0:
            halo_use = prev_tail if s > 0 else torch.zeros(HALO, COLS, dtype=torch.int64)
            prev_tail = halo_in.clone()
        else:
            halo_use = halo_in
        f = front(torch.cat([halo_use, mine]))
        if ring.recv_state(state, s):
            pass
        y, state = tail(f, state if plan.has_left_state(s) else torch.zeros(COLS, dtype=torch.int64))
        ring.send_state(state, s, last)
        outs.append((b, y))
    ring.finish()
    counts = sharding.gather_counts(int(sum(int(y.sum()) for _, y in outs) % 1000), world, torch.device("cpu"))
    q.put((rank, [(b, y.numpy()) for b, y in outs], counts))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_time_shards_equal_sequential(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = sequential(world).numpy()
    blocks = {}
    for rank, outs, counts in got:
        for b, y in outs:
            blocks[b] = y
        if rank == 0:
            assert counts is not None and len(counts) == world
        else:
            assert counts is None
    assert sorted(blocks) == list(range(world * STEPS))
    y = np.concatenate([blocks[b] for b in sorted(blocks)])
    assert np.array_equal(y, want)


def test_plan_indexing():
    sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))
    from grb200 import sharding
    p = sharding.TimeShardPlan(8, 3, 12500, 126)
    assert p.block_index(0) == 3 and p.block_index(2) == 19 and p.abs_start(2) == 19 * 12500
    assert p.left == 2 and p.right == 4 and p.has_left_state(0)
    p0 = sharding.TimeShardPlan(8, 0, 12500, 126)
    assert p0.left == 7 and not p0.has_left_state(0) and p0.has_left_state(1)


def test_stream_policy_by_world_size():
    """One rank overlaps its front with its own previous tail; from two ranks on the front is ordered after it
    (stream policy)."""
    from grb200 import sharding
    assert not sharding.TimeShardPlan(1, 0, 100, 0).front_after_own_tail
    for world in (2, 3, 4, 8):
        for rank in range(world):
            p = sharding.TimeShardPlan(world, rank, 100, 10)
            assert p.front_after_own_tail
            assert p.block_index(3) == 3 * world + rank and p.abs_start(3) == (3 * world + rank) * 100
            assert p.left == (rank - 1) % world and p.right == (rank + 1) % world
    assert not sharding.TimeShardPlan(4, 0, 100, 10).has_left_state(0)
    assert sharding.TimeShardPlan(4, 1, 100, 10).has_left_state(0)
