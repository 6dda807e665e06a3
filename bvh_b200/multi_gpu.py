"""Ray-batch sharding across GPUs.

The traversal has no exchange step: rays are independent units, the BVH is replicated (every rank runs
the same deterministic build), each rank traces a contiguous range of the global batch.  The only
collective is the gather of 16-byte hit records the north star asks for, issued per chunk on a side
stream so that it overlaps the traversal of the next chunk.  ``torch.distributed`` is plumbing only
(NCCL on GPUs, gloo in the CPU tests); the tracer is injected so the host logic is testable without
a device.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [begin, end) of ``total`` units for ``rank`` — contiguity keeps primary-ray
    coherence per GPU.  The first ``total % world`` ranks get one extra unit."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def chunk_bounds(count: int, chunks: int) -> list[tuple[int, int]]:
    chunks = max(1, min(chunks, count)) if count else 1
    return [shard_range(count, c, chunks) for c in range(chunks)]


class ShardedTracer:
    """Traces this rank's shard chunk by chunk and all-gathers the hit records.

    ``trace(begin, end, out)`` must enqueue the traversal of local rays [begin, end) into ``out``
    (a ``(end-begin, words)`` tensor) on the CURRENT stream.  All ranks must hold equally sized shards
    (pad the batch) because ``all_gather_into_tensor`` needs equal contributions.
    """

    def __init__(self, local_count: int, hit_words: int, dtype: torch.dtype, device: torch.device,
                 trace: Callable[[int, int, torch.Tensor], None], chunks: int = 4, group=None):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        self.trace = trace
        self.local_count = local_count
        self.bounds = chunk_bounds(local_count, chunks)
        self.local = torch.empty((local_count, hit_words), dtype=dtype, device=device)
        # gathered[c] is the concatenation over ranks of chunk c: (world * chunk_len, words), rank-major
        self.gathered = [torch.empty((self.world * (e - b), hit_words), dtype=dtype, device=device) for b, e in self.bounds]
        self.is_cuda = device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=device) if self.is_cuda else None

    def step(self) -> None:
        """One pass: trace every chunk, gather each as soon as it is done.  Returns with all work
        enqueued; the caller synchronises."""
        handles = []
        for c, (b, e) in enumerate(self.bounds):
            self.trace(b, e, self.local[b:e])
            if self.world == 1:
                continue
            if self.is_cuda:
                done = torch.cuda.Event()
                done.record()
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(done)
                    dist.all_gather_into_tensor(self.gathered[c], self.local[b:e], group=self.group)
            else:
                handles.append(dist.all_gather_into_tensor(self.gathered[c], self.local[b:e].contiguous(),
                                                           group=self.group, async_op=True))
        for h in handles:
            h.wait()
        if self.is_cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def global_hits(self) -> torch.Tensor:
        """The whole batch's hit records in global ray order: (world * local_count, words)."""
        if self.world == 1:
            return self.local
        views = [g.view(self.world, e - b, -1) for g, (b, e) in zip(self.gathered, self.bounds)]
        per_rank = [torch.cat([v[r] for v in views], dim=0) for r in range(self.world)]
        return torch.cat(per_rank, dim=0)


class FusedGatherTracer:
    """Traversal fused with the all-gather of its hit records over NVLink peer memory.

    The gathered hit array lives in symmetric memory (``torch.distributed._symmetric_memory``): every rank
    holds the same (world * local_count, words) buffer and knows the address of each peer's copy.  One
    launch of the traversal kernel (``bvhNN_intersect_rays_gather``) traces this rank's shard and stores
    each finished ray's 16-byte record straight into the shard's slot of EVERY rank's buffer — one multimem
    store through the NVSwitch multicast address when the fabric offers it, otherwise one peer store per
    rank — so the transfer rides along with the traversal instead of following it as an NCCL call.  A
    symmetric-memory barrier at the end of the step orders the remote stores before anyone reads.
    """

    def __init__(self, bvh, rays: torch.Tensor, hit_words: int, group=None, flags: int = 0, mode: str = "auto"):
        import torch.distributed._symmetric_memory as symm_mem
        self.bvh, self.rays, self.flags = bvh, rays, flags
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.local_count = rays.shape[0]
        self.gathered = symm_mem.empty((self.world * self.local_count, hit_words), dtype=torch.int32, device=rays.device)
        self.handle = symm_mem.rendezvous(self.gathered, self.group)
        self.peers = [int(p) for p in self.handle.buffer_ptrs]
        multicast = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        # Measured on B200 / NVSwitch (profiles/r01_gather_modes.txt): with two ranks the multicast store is the
        # faster form (5227 vs 5155 Mrays/s), with four it is the slower one (8846 vs 10049) — "auto" follows that.
        if mode == "peer" or (mode == "auto" and self.world > 2):
            multicast = 0
        if mode == "multicast" and not multicast:
            raise RuntimeError("no multicast address for the symmetric buffer on this fabric")
        self.multicast = multicast
        self.mode = "multicast" if multicast else "peer"
        self.bounds = [(0, self.local_count)]
        self.local = self.gathered[self.rank * self.local_count:(self.rank + 1) * self.local_count]

    def step(self) -> None:
        self.bvh.intersect_rays_gather(self.rays.data_ptr(), self.local_count, self.peers,
                                       self.rank * self.local_count, multicast_ptr=self.multicast, flags=self.flags)
        self.handle.barrier(channel=0)

    def global_hits(self) -> torch.Tensor:
        return self.gathered
