"""Reference-view sharding across the GPUs of one box.

A dataset item is one (scan, ref_view, src_views) tuple and nothing crosses items
(models/TransMVSNet.py:141-226), so the path is embarrassingly parallel over reference views:
one process per GPU, no collective inside the path.  The only exchange is gathering the
per-view depth / confidence maps on rank 0 (the inference analogue of the reference's
DistributedSampler, train.py:377-381, which exists for training only).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment: view v goes to rank v % world_size (balanced to within one view)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, n_views, world_size))


def views_per_rank(n_views: int, world_size: int) -> List[int]:
    return [len(range(r, n_views, world_size)) for r in range(world_size)]


def gather_maps(local_maps: torch.Tensor, n_views: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather per-view maps onto rank `dst` in global view order.

    local_maps: [n_local, K, H, W] (K = 2: depth, confidence) for this rank's views, in shard order.
    Returns [n_views, K, H, W] on rank dst, None elsewhere.  Ranks with fewer views pad to the
    maximum so the collective has equal message sizes (NCCL/gloo all_gather).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = views_per_rank(n_views, world)
    n_max = max(counts)
    if local_maps.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local_maps.shape[0]} views, expected {counts[rank]}")
    pad = local_maps
    if local_maps.shape[0] < n_max:
        fill = torch.zeros((n_max - local_maps.shape[0],) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype,
                           device=local_maps.device)
        pad = torch.cat([local_maps, fill], 0)
    pad = pad.contiguous()
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    if rank != dst:
        return None
    out = torch.empty((n_views,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
    for r in range(world):
        for j, v in enumerate(range(r, n_views, world)):
            out[v] = bufs[r][j]
    return out


class _RawCuda:
    """A raw device pointer dressed as a CUDA array (zero-copy `torch.as_tensor`)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerMapSink:
    """The gather of the per-view maps, fused into the kernel that produces them (single box, NVLink / NVSwitch).

    Rank `dst` owns one buffer [slots, 2, H, W] (depth, confidence) created by tmvs_peer_buffer_create and ships its
    CUDA IPC handle to the other ranks; each opens it with its own GPU current (tmvs_peer_buffer_open), which maps the
    buffer into that GPU's address space.  `slot(i)` is then a tensor over GPU `dst`'s memory that this rank's kernels
    can write: passing it as the read-out kernel's output (ops.softmax_wta(out_depth=, out_conf=),
    pipeline.run_cascade(out_maps=)) makes the kernel's own stores cross NVLink -- no collective, no staging copy, no
    extra kernel competing with the cost-volume kernels for SMs.  The only synchronisation is the caller's barrier
    before rank `dst` reads `result()`.  gather_maps (NCCL / gloo all_gather) remains the transport across boxes.
    """

    def __init__(self, slots: int, hw, device, dst: int = 0, group=None, sync: bool = True):
        import ctypes
        from . import _lib
        lib = _lib.load()
        self._lib = lib
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.dst = dst
        self.device = torch.device(device)
        shape = (int(slots), 2, int(hw[0]), int(hw[1]))
        nbytes = 4 * shape[0] * shape[1] * shape[2] * shape[3]
        self._ptr = ctypes.c_void_p(0)
        self._owner = self.rank == dst
        payload = [None]
        with torch.cuda.device(self.device):
            if self._owner:
                handle = (ctypes.c_ubyte * 64)()
                rc = lib.tmvs_peer_buffer_create(nbytes, ctypes.byref(self._ptr), handle)
                payload = [bytes(handle) if rc == 0 else rc]          # the broadcast happens either way: nobody hangs
            if self.world > 1:
                dist.broadcast_object_list(payload, src=dst, group=group)
            if not isinstance(payload[0], bytes):
                raise _lib.TmvsError(f"tmvs_peer_buffer_create failed on rank {dst} (code {payload[0]})")
            if not self._owner:
                handle = (ctypes.c_ubyte * 64).from_buffer_copy(payload[0])
                _lib.check(lib.tmvs_peer_buffer_open(handle, ctypes.byref(self._ptr)), "tmvs_peer_buffer_open")
            # zero-copy view; torch reads the tensor's device off the pointer (GPU `dst`, also for the peer mapping)
            self.buffer = torch.as_tensor(_RawCuda(self._ptr.value, shape))
            if self.buffer.data_ptr() != self._ptr.value:
                raise RuntimeError("PeerMapSink: torch copied the peer buffer instead of aliasing it")
        if self.world > 1 and sync:      # sync=False: the caller follows with its own collective (e.g. to agree on a fallback)
            dist.barrier(group=group)

    def slot(self, index: int) -> torch.Tensor:
        """[1, 2, H, W] view of slot `index` (peer memory on every rank but `dst`)."""
        return self.buffer[index:index + 1]

    def push(self, index: int, depth: torch.Tensor, conf: torch.Tensor, stream: torch.cuda.Stream) -> None:
        """Copy-engine transport: slot `index` <- (depth, conf) [1,H,W] each, as two asynchronous DMA copies on `stream`
        (a side stream that has waited for the producing kernel).  No SM is involved and the producing kernel never
        waits on the NVLink port.  The caller keeps `depth` / `conf` alive and unmodified until the stream has passed the
        copies (persistent buffers + an event, as bench.py does)."""
        import ctypes
        from . import _lib
        slot = self.buffer[index]                        # [2,H,W] on rank dst
        for k, t in enumerate((depth, conf)):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != slot[k].numel():
                raise _lib.TmvsError("PeerMapSink.push: maps must be contiguous fp32 [1,H,W]")
            with torch.cuda.device(self.device):
                rc = self._lib.tmvs_peer_copy_async(ctypes.c_void_p(slot[k].data_ptr()), ctypes.c_void_p(t.data_ptr()),
                                                    t.numel() * 4, ctypes.c_void_p(stream.cuda_stream))
            _lib.check(rc, "tmvs_peer_copy_async")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001 -- interpreter shutdown: the driver reclaims the mapping
            pass

    def result(self) -> Optional[torch.Tensor]:
        """The whole buffer on rank `dst`, None elsewhere (read it after a barrier that follows every writer's
        stream synchronisation)."""
        return self.buffer if self._owner else None

    def close(self) -> None:
        if getattr(self, "_ptr", None) is not None and self._ptr.value:
            self.buffer = None
            self._lib.tmvs_peer_buffer_release(self._ptr, int(self._owner))
            self._ptr.value = 0
