"""Optional final gather of the sharded height maps WITHOUT kernels: a ring of slots in the root GPU's memory that
the other ranks fill with plain device-to-device copies over NVLink (copy engines), one `cudaMemcpyAsync` per chunk.

Why not NCCL send/recv for the streamed gather: NCCL moves data with copy *kernels*, and those compete for SMs with
the persistent FCD grids (profiles/stream_video_2gpu_r01.json: 28.6k -> 9.0k frames/s).  Here the data plane is
peer memory mapped through torch symmetric memory (cuMem / VMM handles -- the transport NCCL reports as P2P/CUMEM)
filled by the copy engines, and the control plane (slot filled / slot free) is a gloo process group on CPU tensors,
so the SMs of every GPU keep running the FCD kernels.  Measured on this pool (scripts/p2p_probe*.py, 2 GiB):
symmetric-memory peer copy 755 GB/s, NCCL send/recv 690 GB/s, a legacy cudaIpc mapping 36 GB/s (PCIe),
NCCL_P2P_USE_CUDA_MEMCPY=1 69 GB/s -- hence symmetric memory.

Frames stay independent (pydata/analyze.py:220-252 is a plain loop); this is the "optional final NCCL gather" of
SURVEY.md 8(e) for BASELINE.json configs[3], where the gathered stream (20k maps = 335 GB) cannot live in one GPU's
memory and the root therefore drains ring slots as they arrive.
"""
from __future__ import annotations

import threading
from typing import Callable, Optional

import torch
import torch.distributed as dist


class PeerRing:
    """root: owns `slots` buffers of `chunk` maps per source rank (symmetric allocation: (world * slots) chunks on
    every rank, only the root's are written) and a consumer thread that calls
    `consume(src, chunk_index, slot_tensor, n_frames, tag0, tag1)` for every filled slot, then frees it.
    other ranks: `push(local_chunk_tensor, n_frames)` copies a finished chunk into the next slot on a side stream.

    All ranks must construct it collectively (it exchanges IPC handles) and call `close()` collectively."""

    def __init__(self, chunk_shape, chunks_per_rank, root: int = 0, slots: int = 2, device=None,
                 consume: Optional[Callable] = None, dtype=torch.float32):
        import torch.distributed._symmetric_memory as symm_mem

        self.rank, self.world, self.root, self.slots = dist.get_rank(), dist.get_world_size(), root, slots
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.chunks_per_rank = list(chunks_per_rank)            # number of chunks every rank will push
        self.ctrl = dist.new_group(backend="gloo")              # control plane: CPU tensors, no GPU kernels
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._k = 0
        self._copied = [None] * slots                           # events: copy out of the local buffer finished
        self._thread = None
        self.consumed = 0
        # symmetric allocation: every rank allocates the ring, only the root's copy is used
        shape = (self.world, slots) + tuple(chunk_shape)
        self._symm = symm_mem.empty(shape, dtype=dtype, device=self.device)
        self._hdl = symm_mem.rendezvous(self._symm, dist.group.WORLD)
        if self.rank != root:
            peer = self._hdl.get_buffer(root, shape, dtype)     # the root's ring, mapped into this process
            self.views = [peer[self.rank, slot] for slot in range(slots)]
        else:
            self.ring = {src: [self._symm[src, slot] for slot in range(slots)]
                         for src in range(self.world) if src != root}
            self._consume = consume
            self._thread = threading.Thread(target=self._serve, daemon=True)
            self._thread.start()
        dist.barrier()

    # ------------------------------------------------------------------ senders
    def push(self, chunk: torch.Tensor, n_frames: int, ready: Optional[torch.cuda.Event] = None, tag0: int = 0,
             tag1: int = 0) -> torch.cuda.Event:
        """Queue the copy of chunk[:n_frames] into the next ring slot (after `ready`, default: everything queued so
        far on the current stream).  Returns the event that marks the end of the copy: the source buffer may be
        overwritten after it.  tag0 / tag1 travel with the "slot filled" message (e.g. checksums)."""
        assert self.rank != self.root
        k, slot = self._k, self._k % self.slots
        if k >= self.slots:                                     # wait until the root has drained this slot
            free = torch.zeros(1, dtype=torch.int64)
            dist.recv(free, src=self.root, group=self.ctrl, tag=slot)
        if ready is None:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            self.views[slot][:n_frames].copy_(chunk[:n_frames], non_blocking=True)     # peer copy, copy engine
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self._copied[slot] = (done, k, n_frames, int(tag0), int(tag1))
        self._k += 1
        # announce the previous chunk once its copy has landed (its wait overlaps the kernels queued meanwhile)
        self._announce(upto=k - 1)
        return done

    def _announce(self, upto: int) -> None:
        for slot in range(self.slots):
            item = self._copied[slot]
            if item is not None and item[1] <= upto:
                done, k, n_frames, t0, t1 = item
                done.synchronize()
                dist.send(torch.tensor([k, n_frames, t0, t1], dtype=torch.int64), dst=self.root, group=self.ctrl,
                          tag=self.slots + slot)
                self._copied[slot] = None

    def flush(self) -> None:
        if self.rank != self.root:
            self._announce(upto=self._k)

    # ------------------------------------------------------------------ root
    def _serve(self) -> None:
        torch.cuda.set_device(self.device)
        stream = torch.cuda.Stream(device=self.device)
        pending = {src: 0 for src in self.ring}
        msg = torch.zeros(4, dtype=torch.int64)
        while any(pending[src] < self.chunks_per_rank[src] for src in pending):
            for src in list(pending):
                k = pending[src]
                if k >= self.chunks_per_rank[src]:
                    continue
                slot = k % self.slots
                dist.recv(msg, src=src, group=self.ctrl, tag=self.slots + slot)
                n_frames = int(msg[1])
                if self._consume is not None:
                    with torch.cuda.stream(stream):
                        self._consume(src, k, self.ring[src][slot], n_frames, int(msg[2]), int(msg[3]))
                    stream.synchronize()
                self.consumed += n_frames
                pending[src] = k + 1
                if k + self.slots < self.chunks_per_rank[src]:
                    dist.send(torch.ones(1, dtype=torch.int64), dst=src, group=self.ctrl, tag=slot)

    def close(self) -> None:
        self.flush()
        if self._thread is not None:
            self._thread.join()
        torch.cuda.synchronize(self.device)
        dist.barrier()
        if self.rank != self.root:
            del self.views
        dist.barrier()
