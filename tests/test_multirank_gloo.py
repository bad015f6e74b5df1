"""N>1 path on CPU: two ranks over gloo, each owning a disjoint shard of the groups
(group_offset = rank * G), leader-announce all_gather -- must equal one process
owning all groups.  The engine here is the device code compiled for the host
(tests/emu); on the GPU box bench.py --gpus N runs the same logic over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from josefine_b200 import abi
from tests import parity

G_PER_RANK, R, STEPS, SEED = 6, 3, 40, 21


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.emu.emu import EmuEngine
    eng = EmuEngine.create(G_PER_RANK, R, seed=SEED, group_offset=rank * G_PER_RANK, flags=abi.F_STREAM_DIGEST)
    now = 100
    gathered_hist = []
    for chunk in range(STEPS // 10):
        eng.run(now, 100, 10, 1)
        now += 1000
        table = torch.tensor(eng.leader_table(), dtype=torch.int64)          # [G, 3] = term, leader, commit
        allt = [torch.zeros_like(table) for _ in range(world)]
        dist.all_gather(allt, table)                                          # the leader announce
        gathered_hist.append(torch.cat(allt).tolist())
    digest = torch.tensor([eng.state_digest() & ((1 << 62) - 1), eng.state_digest() >> 62], dtype=torch.int64)
    sums = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(sums, digest)
    if rank == 0:
        out_q.put((gathered_hist, [s.tolist() for s in sums]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shards_equal_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    hist, digests = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single process, all groups, the oracle
    from oracle.restated import RestatedCluster
    one = RestatedCluster.create(world * G_PER_RANK, R, seed=SEED, flags=abi.F_STREAM_DIGEST)
    now = 100
    for chunk in range(STEPS // 10):
        one.run(now, 100, 10, 1)
        now += 1000
        assert [list(t) for t in one.leader_table()] == hist[chunk]
    total = sum(lo + (hi << 62) for lo, hi in digests) & ((1 << 64) - 1)
    assert total == one.state_digest()            # digests are wrapping sums over replicas: shards add up
    assert any(l for (_, l, _) in one.leader_table())
