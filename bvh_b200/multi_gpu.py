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
    launch of the traversal kernel (``bvhNN_intersect_rays_gather``) traces this rank's shard and delivers
    every hit record into the shard's slot of EVERY rank's buffer while it runs: a warp stages the records of
    32 consecutive rays in shared memory and one lane sends the 512-byte block to each rank with a bulk
    asynchronous copy (``cp.async.bulk`` shared -> peer global), or — ``mode="multicast"`` — every record
    goes out as one ``multimem.st`` through the NVSwitch multicast address.  A symmetric-memory barrier at the
    end of the step orders the remote stores before anyone reads.

    Ordering (what a caller may rely on):
      * the kernel runs on the BVH handle's stream; ``step()`` makes it wait for the work already enqueued on
        torch's current stream and makes the current stream wait for the kernel before the barrier, so the
        barrier is ordered after the kernel's peer stores whichever stream the handle owns;
      * the gathered array is double-buffered: step k writes buffer k % 2, so a peer still reading the records
        of step k - 1 is never overwritten — the barrier that ends step k is only passed by ranks that are done
        with the buffer step k + 1 will write;
      * ``check()`` waits for the handle's stream and raises if a kernel watchdog fired (stale records).
    """

    def __init__(self, bvh, rays: torch.Tensor, hit_words: int, group=None, flags: int = 0, mode: str = "auto",
                 buffers: int = 2):
        import torch.distributed._symmetric_memory as symm_mem
        self.bvh, self.rays, self.flags = bvh, rays, flags
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.local_count = rays.shape[0]
        self.buffers = []
        for _ in range(max(1, buffers)):
            gathered = symm_mem.empty((self.world * self.local_count, hit_words), dtype=torch.int32, device=rays.device)
            handle = symm_mem.rendezvous(gathered, self.group)
            self.buffers.append((gathered, handle, [int(p) for p in handle.buffer_ptrs],
                                 int(getattr(handle, "multicast_ptr", 0) or 0)))
        has_multicast = all(b[3] for b in self.buffers)
        # "auto": staged bulk copies to every peer.  Round 1 measured per-record stores only (multicast ahead at two
        # ranks, peer stores ahead at four: profiles/r01_gather_modes.txt); the staged form replaces world_size
        # store instructions per ray by one bulk copy per rank and 32 rays.
        if mode == "multicast" and not has_multicast:
            raise RuntimeError("no multicast address for the symmetric buffer on this fabric")
        self.mode = "multicast" if mode == "multicast" else "peer"
        self.bounds = [(0, self.local_count)]
        self.steps = 0
        lib_stream = int(bvh.get_property("stream"))
        self.lib_stream = None
        if lib_stream != int(torch.cuda.current_stream(rays.device).cuda_stream):
            self.lib_stream = torch.cuda.ExternalStream(lib_stream, device=rays.device)
        self._select(0)

    def _select(self, k: int) -> None:
        self.gathered, self.handle, self.peers, multicast = self.buffers[k % len(self.buffers)]
        self.multicast = multicast if self.mode == "multicast" else 0
        self.local = self.gathered[self.rank * self.local_count:(self.rank + 1) * self.local_count]

    def step(self) -> None:
        self._select(self.steps)
        self.steps += 1
        current = torch.cuda.current_stream(self.rays.device)
        if self.lib_stream is not None:
            self.lib_stream.wait_stream(current)            # e.g. the previous step's barrier, the caller's ray upload
        self.bvh.intersect_rays_gather(self.rays.data_ptr(), self.local_count, self.peers,
                                       self.rank * self.local_count, multicast_ptr=self.multicast, flags=self.flags)
        if self.lib_stream is not None:
            current.wait_stream(self.lib_stream)            # the barrier below must follow the kernel's peer stores
        self.handle.barrier(channel=0)

    def check(self) -> None:
        """Waits for the handle's stream; raises ``BvhError`` if the traversal kernel's watchdog fired."""
        self.bvh.sync()

    def global_hits(self) -> torch.Tensor:
        """The gathered records of the most recent step."""
        return self.gathered
