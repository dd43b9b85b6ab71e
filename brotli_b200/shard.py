"""Multi-GPU plumbing for independent streams: one process per GPU,
no collective on the data path; at the end of a step the variable-size compressed shards are
gathered to rank 0 (sizes by all_gather, payloads by grouped send/recv into preallocated
slots).  Works with any torch.distributed backend: NCCL on the GPU box (bench.py), gloo in the
CPU tests (tests/test_multi_cpu.py) -- the same class runs in both.

Partitioning rules of the reference's configs:
  * config 3: shard i of N = bytes [i * total / N, (i + 1) * total / N) as its own stream  -> shard_range()
  * config 5: stream j lives on GPU j mod N                                            -> streams_of_rank()
"""
import torch
import torch.distributed as dist


def streams_of_rank(num_streams, rank, world):
    """Stream j lives on GPU j mod world (BASELINE.json config 5 / SURVEY.md 8e)."""
    return list(range(rank, num_streams, world))


def shard_range(total, rank, world):
    """Byte range of shard `rank` of `world` (BASELINE.json config 3 / SURVEY.md 8e)."""
    return (rank * total) // world, ((rank + 1) * total) // world


class ShardGather(object):
    """Gathers one variable-size uint8 payload per rank to `dst`.  Everything is allocated once:
    a [world, cap] receive area on dst, the size vectors, and (CUDA) a pinned host copy of the
    sizes, so a step costs one all_gather of 8 bytes per rank, one event wait and one grouped
    send/recv -- no allocation, no per-rank serial wait."""

    def __init__(self, cap, device, dst=0, group=None):
        self.group, self.dst = group, dst
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device)
        capt = torch.tensor([int(cap)], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(capt, op=dist.ReduceOp.MAX, group=group)
        self.cap = int(capt.item())
        self.mine = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.sizes_dev = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        self.sizes_host = torch.zeros(self.world, dtype=torch.int64)
        if self.device.type == "cuda":
            self.sizes_host = self.sizes_host.pin_memory()
        self.recv = None
        if self.rank == dst:
            self.recv = torch.empty((self.world, self.cap), dtype=torch.uint8, device=self.device)
        self.sizes = [0] * self.world

    def gather(self, payload, nbytes):
        """payload: 1-D uint8 tensor on self.device holding this rank's bytes in [0, nbytes).
        On dst returns the list of per-rank views (into the receive area; dst's own is `payload`);
        None elsewhere."""
        if self.world == 1:
            self.sizes = [int(nbytes)]
            return [payload[:nbytes]]
        self.mine.fill_(int(nbytes))
        dist.all_gather_into_tensor(self.sizes_dev, self.mine, group=self.group)
        if self.rank != self.dst:
            if nbytes:
                dist.send(payload[:nbytes], dst=self.dst, group=self.group)
            return None
        self.sizes_host.copy_(self.sizes_dev, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        self.sizes = [int(x) for x in self.sizes_host.tolist()]
        ops, out = [], [None] * self.world
        for r in range(self.world):
            if r == self.dst:
                out[r] = payload[:nbytes]
                continue
            out[r] = self.recv[r, :self.sizes[r]]
            if self.sizes[r]:
                ops.append(dist.P2POp(dist.irecv, out[r], r, group=self.group))
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        return out


def gather_shards(payload, dst=0, group=None):
    """One-off form of ShardGather (allocates): list of per-rank payload tensors on dst, None elsewhere."""
    g = ShardGather(payload.numel(), payload.device, dst, group)
    return g.gather(payload, payload.numel())
