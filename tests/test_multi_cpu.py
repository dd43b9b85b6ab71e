"""CPU suite, part 4: the N > 1 path (stream dealing, shard ranges, variable-size gather to rank 0 with
preallocated slots) on two gloo processes.  Payloads are stand-ins produced by the oracle; on the GPU box
bench.py drives the same brotli_b200.shard.ShardGather over NCCL with device tensors."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from brotli_b200.shard import ShardGather, gather_shards, shard_range, streams_of_rank
    from brotli_libs import Oracle
    from corpus import synth_text
    ora = Oracle()
    mine = streams_of_rank(5, rank, world)
    blobs = [ora.compress(synth_text(20000 + 1000 * j, seed=j), 5, 22) for j in mine]
    payload = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8)
    got = gather_shards(payload, dst=0)
    # the reusable form bench.py uses: capacity-sized buffer, several steps through the same slots
    g = ShardGather(payload.numel() + 100 * (rank + 1), "cpu")
    buf = torch.zeros(g.cap, dtype=torch.uint8)
    steps = []
    for step in range(3):
        nb = payload.numel() - step * 7 * (rank + 1)
        buf[:nb] = payload[:nb]
        out = g.gather(buf, nb)
        if rank == 0:
            assert g.sizes == [int(t.numel()) for t in out]
            steps.append([bytes(t.numpy().tobytes()) for t in out])
        else:
            assert out is None
    lo, hi = shard_range(1000003, rank, world)
    if rank == 0:
        q.put(([bytes(t.numpy().tobytes()) for t in got], steps, (lo, hi)))
    else:
        assert got is None and (lo, hi) == (500001, 1000003)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_shards_two_ranks():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from brotli_b200.shard import streams_of_rank
    from brotli_libs import Oracle
    from corpus import synth_text
    assert streams_of_rank(5, 0, 2) == [0, 2, 4] and streams_of_rank(5, 1, 2) == [1, 3]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, steps, rng0 = q.get(timeout=120)
    assert rng0 == (0, 500001)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ora = Oracle()
    for r in range(2):
        want = b"".join(ora.compress(synth_text(20000 + 1000 * j, seed=j), 5, 22) for j in streams_of_rank(5, r, 2))
        assert res[r] == want
        for step in range(3):
            assert steps[step][r] == want[:len(want) - step * 7 * (r + 1)]
