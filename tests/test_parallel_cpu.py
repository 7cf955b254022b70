"""CPU: the multi-GPU plumbing (one process per rank, batch shards, one weight broadcast, max-over-ranks
timing) exercised with world_size 2 over gloo."""
import os
import socket
from argparse import Namespace

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import RAFT_CFG


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from dkt_stereo_b200 import parallel
    from dkt_stereo_b200.raft_stereo import RAFTStereo
    r, l, w = parallel.init_from_env(backend="gloo")
    assert (r, l, w) == (rank, rank, world)
    torch.manual_seed(100 + rank)                           # ranks start with DIFFERENT weights
    model = RAFTStereo(Namespace(mixed_precision=False, **RAFT_CFG))
    before = torch.cat([p.detach().reshape(-1) for p in model.update_block.parameters()]).clone()
    nbytes = parallel.broadcast_weights(model, src=0)
    after = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.zeros_like(after) for _ in range(world)]
    dist.all_gather(gathered, after)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    changed = not torch.equal(before, torch.cat([p.detach().reshape(-1) for p in model.update_block.parameters()]))
    mx = parallel.all_reduce_max(10.0 + rank, torch.device("cpu"))
    lo, hi = parallel.shard_range(13, rank, world)
    assert parallel.all_gather_int(7 * rank + 1, torch.device("cpu")) == [1, 8]      # checksum exchange of bench.py
    parallel.barrier()
    q.put((rank, nbytes, same, changed, mx, lo, hi))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, n0, same0, ch0, mx0, lo0, hi0), (r1, n1, same1, ch1, mx1, lo1, hi1) = res
    assert n0 == n1 and n0 > 40_000_000                    # ~11.1 M fp32 parameters + buffers in one flat broadcast
    assert same0 and same1                                 # every rank ends with rank 0's weights
    assert not ch0 and ch1                                 # rank 0 unchanged, rank 1 overwritten
    assert mx0 == mx1 == 11.0                              # max over ranks (timing reduction)
    assert (lo0, hi0, lo1, hi1) == (0, 7, 7, 13)           # contiguous shards, remainder to the first ranks


def test_shard_range_partitions():
    from dkt_stereo_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
