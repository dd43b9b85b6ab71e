"""Multi-GPU plumbing for independent streams: one process per GPU,
streams dealt round-robin, no collective on the data path; at the end the variable-size
compressed shards are gathered to rank 0 (sizes by all_gather, payloads by send/recv).
Works with any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def streams_of_rank(num_streams, rank, world):
    """Stream j lives on GPU j mod world (BASELINE.json config 5 / SURVEY.md 8e)."""
    return list(range(rank, num_streams, world))


def gather_shards(payload, dst=0, group=None):
    """payload: 1-D uint8 tensor (this rank's compressed bytes, on the backend's device).
    Returns the list of per-rank payload tensors on `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = payload.device
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([payload.numel()], dtype=torch.int64, device=dev), group=group)
    sizes = [int(s.item()) for s in sizes]
    if rank == dst:
        out = [None] * world
        out[dst] = payload
        reqs = []
        for r in range(world):
            if r == dst:
                continue
            out[r] = torch.empty(sizes[r], dtype=torch.uint8, device=dev)
            if sizes[r]:
                reqs.append(dist.irecv(out[r], src=r, group=group))
        for q in reqs:
            q.wait()
        return out
    if payload.numel():
        dist.send(payload, dst=dst, group=group)
    return None
