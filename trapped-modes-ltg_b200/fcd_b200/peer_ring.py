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
import time
from typing import Callable, Optional

import torch
import torch.distributed as dist


class _NoEvent:
    """Stand-in for torch.cuda.Event in the CPU (test) mode: copies are synchronous."""
    def synchronize(self) -> None:
        pass


class PeerRing:
    """root: owns `slots` buffers of `chunk` maps per source rank (symmetric allocation: (world * slots) chunks on
    every rank, only the root's are written -- size the chunk accordingly: 8 ranks x 4 slots x 128 maps of 2048^2 are
    69 GB per GPU) and a consumer thread that calls
    `consume(src, chunk_index, slot_tensor, n_frames, tag0, tag1)` for every filled slot, then frees it.
    other ranks: `push(local_chunk_tensor, n_frames)` copies a finished chunk into the next slot on a side stream.

    All ranks must construct it collectively (it exchanges IPC handles) and call `close()` collectively."""

    def __init__(self, chunk_shape, chunks_per_rank, root: int = 0, slots: int = 2, device=None,
                 consume: Optional[Callable] = None, dtype=torch.float32):
        self.rank, self.world, self.root, self.slots = dist.get_rank(), dist.get_world_size(), root, slots
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.cpu = self.device.type == "cpu"                    # CPU mode (tests): ring in POSIX shared memory
        self.chunks_per_rank = list(chunks_per_rank)            # number of chunks every rank will push
        self.ctrl = dist.new_group(backend="gloo")              # control plane: CPU tensors, no GPU kernels
        self.copy_stream = None if self.cpu else torch.cuda.Stream(device=self.device)
        self._k = 0
        self._copied = [None] * slots                           # events: copy out of the local buffer finished
        self._thread = None
        self.consumed = 0
        self.wait_free_s = 0.0       # sender: time blocked on "slot free" messages
        self.wait_copy_s = 0.0       # sender: time blocked on its own copies before announcing them
        shape = (self.world, slots) + tuple(chunk_shape)
        if self.cpu:
            import math
            import os
            import uuid
            path = [f"/dev/shm/fcd_ring_{os.getpid()}_{uuid.uuid4().hex}" if self.rank == root else None]
            dist.broadcast_object_list(path, src=root, group=self.ctrl)
            self._shm_path = path[0]
            numel = math.prod(shape)
            if self.rank == root:
                self._symm = torch.from_file(self._shm_path, shared=True, size=numel, dtype=dtype).view(shape)
            dist.barrier(group=self.ctrl)
            peer = self._symm if self.rank == root else \
                torch.from_file(self._shm_path, shared=True, size=numel, dtype=dtype).view(shape)
        else:
            import torch.distributed._symmetric_memory as symm_mem
            # symmetric allocation: every rank allocates the ring, only the root's copy is used
            self._symm = symm_mem.empty(shape, dtype=dtype, device=self.device)
            self._hdl = symm_mem.rendezvous(self._symm, dist.group.WORLD)
            peer = self._symm if self.rank == root else self._hdl.get_buffer(root, shape, dtype)
        if self.rank != root:
            self.views = [peer[self.rank, slot] for slot in range(slots)]       # the root's ring, mapped into this process
        else:
            self.ring = {src: [self._symm[src, slot] for slot in range(slots)]
                         for src in range(self.world) if src != root}
            self._consume = consume
            self._thread = threading.Thread(target=self._serve, daemon=True)
            self._thread.start()
        dist.barrier(group=self.ctrl)

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
            t0 = time.perf_counter()
            dist.recv(free, src=self.root, group=self.ctrl, tag=slot)
            self.wait_free_s += time.perf_counter() - t0
        if self.cpu:
            self.views[slot][:n_frames].copy_(chunk[:n_frames])
            done = _NoEvent()
        else:
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
                tw = time.perf_counter()
                done.synchronize()
                self.wait_copy_s += time.perf_counter() - tw
                self._post(torch.tensor([k, n_frames, t0, t1], dtype=torch.int64), self.root, self.slots + slot)
                self._copied[slot] = None

    def _post(self, msg: torch.Tensor, dst: int, tag: int) -> None:
        """Non-blocking control message.  gloo's send is a rendezvous with the matching recv; a blocking send from
        the root ("slot free") and one from a sender ("slot filled") can wait for each other as soon as the ring
        has more than two slots, so both directions post and move on."""
        self._posted = [(w, m) for w, m in getattr(self, "_posted", []) if not w.is_completed()]
        self._posted.append((dist.isend(msg, dst=dst, group=self.ctrl, tag=tag), msg))

    def _drain_posts(self) -> None:
        for w, _ in getattr(self, "_posted", []):
            w.wait()
        self._posted = []

    def flush(self) -> None:
        """Sender: announce every chunk pushed so far and wait until the root has taken the messages."""
        if self.rank != self.root:
            self._announce(upto=self._k)
            self._drain_posts()

    # ------------------------------------------------------------------ root
    def _serve(self) -> None:
        import contextlib
        stream = None
        if not self.cpu:
            torch.cuda.set_device(self.device)
            stream = torch.cuda.Stream(device=self.device, priority=-1)
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
                    with (contextlib.nullcontext() if stream is None else torch.cuda.stream(stream)):
                        self._consume(src, k, self.ring[src][slot], n_frames, int(msg[2]), int(msg[3]))
                    if stream is not None:
                        stream.synchronize()
                self.consumed += n_frames
                pending[src] = k + 1
                if k + self.slots < self.chunks_per_rank[src]:
                    self._post(torch.ones(1, dtype=torch.int64), src, slot)

    def close(self) -> None:
        # The list of posted messages belongs to ONE thread at a time: on the root that is the consumer thread until
        # it has been joined (draining it from here while the consumer still posts dropped in-flight "slot free"
        # messages, and the senders waited for them forever).
        self.flush()
        if self._thread is not None:
            self._thread.join()
            self._drain_posts()
        if not self.cpu:
            torch.cuda.synchronize(self.device)
        dist.barrier(group=self.ctrl)
        if self.rank != self.root:
            del self.views
        dist.barrier(group=self.ctrl)
        if self.cpu and self.rank == self.root:
            import os
            del self.ring, self._symm
            os.unlink(self._shm_path)
