"""Multi-GPU layer: one process per GPU (torchrun), shards with no data-path collective except
the single all-reduce(MAX) that effects.normalize needs on a time-sharded buffer.

  * clip batches shard by clip (independent units, no collective)           -- shard_clips()
  * a long buffer shards by contiguous OUTPUT frame range; every shard carries the input
    frames its taps touch (its interpolation halo), computed from the GLOBAL fp64 positions
    of A:666 so a sharded run is bitwise identical to a single pass          -- plan_time_shards()
  * normalize on a sharded buffer = local abs-max, all-reduce MAX of ONE float over
    NCCL/NVLink (gloo on CPU for the tests), local scale+clamp               -- ShardedPreload

torch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

from . import _lib
from ._lib import PipelineDesc

_INTERPS = {"none": 0, "linear": 1, "cubic": 2, "sinc": 3}


@dataclass
class TimeShard:
    rank: int
    out_first: int    # first global output frame (0-based)
    n_out: int
    in_first: int     # first global input frame this shard must hold (halo included)
    in_count: int


def plan_time_shards(n_in_total: int, srcRate: float, dstRate: float, interpolation: str, world: int) -> List[TimeShard]:
    """Contiguous, near-equal output ranges; input windows overlap by the halo only."""
    lib = _lib.load()
    n_out = int(lib.aukit_resample_out_len(n_in_total, float(srcRate), float(dstRate)))
    shards = []
    for r in range(world):
        # interior boundaries on multiples of 4 outputs keep every shard's stores 16-byte aligned
        o0 = n_out * r // world // 4 * 4
        o1 = n_out if r == world - 1 else n_out * (r + 1) // world // 4 * 4
        f, c = C.c_uint64(0), C.c_uint64(0)
        _lib.check(lib.aukit_resample_window(n_in_total, float(srcRate), float(dstRate), _INTERPS[interpolation], o0, o1 - o0,
                                             C.byref(f), C.byref(c)))
        shards.append(TimeShard(r, o0, o1 - o0, int(f.value), int(c.value)))
    return shards


def plan_block_shards(n_blocks: int, world: int):
    """ADPCM files shard by contiguous BLOCK range: every block header carries the full decoder state (A:1310,
    A:1513), so ranges are independent -- no halo, no collective.  Returns [(first_block, count)] per rank
    (aukit_block_shard in the C ABI)."""
    lib = _lib.load()
    out = []
    for r in range(world):
        f, c = C.c_uint64(0), C.c_uint64(0)
        _lib.check(lib.aukit_block_shard(n_blocks, world, r, C.byref(f), C.byref(c)))
        out.append((int(f.value), int(c.value)))
    return out


def shard_clips(n_clips: int, world: int, rank: int, sizes: Optional[List[int]] = None) -> List[int]:
    """Clip indices owned by `rank`: round-robin, or greedy size-balanced when sizes are given."""
    if sizes is None:
        return list(range(rank, n_clips, world))
    order = sorted(range(n_clips), key=lambda i: -sizes[i])
    load = [0] * world
    owner = [0] * n_clips
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += sizes[i]
    return [i for i in range(n_clips) if owner[i] == rank]


def allreduce_max_(t, group=None):
    """The path's only collective: elementwise MAX of the per-shard peaks (exact, order-free)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


class ShardedPreload:
    """auplay's chain (unpack -> resample -> mono -> normalize) on ONE rank's time shard.

    peak pass -> all-reduce MAX over ranks -> apply pass.  All launches go to torch's current
    stream so the collective is ordered with the kernels without host synchronisation."""

    def __init__(self, ctx, shard: TimeShard, n_in_total: int, bitDepth=16, dataType="signed", channels=2,
                 srcRate=44100.0, dstRate=48000.0, interpolation="cubic", mono=True, peak=0.8, bigEndian=False):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.lib = ctx.lib
        self.shard = shard
        self.peak = float(peak)
        self.out_channels = 1 if mono else channels
        self.frame_bytes = channels * bitDepth // 8
        self.desc = PipelineDesc(bitDepth, {"signed": 0, "unsigned": 1, "float": 2}[dataType], channels, int(bool(bigEndian)),
                                 float(srcRate), float(dstRate), _INTERPS[interpolation], int(bool(mono)), n_in_total,
                                 shard.in_first, shard.in_count, shard.out_first, shard.n_out)
        self.stride = (shard.n_out + 31) // 32 * 32
        self.d_max = torch.zeros(1, dtype=torch.float32, device="cuda")
        self.d_out = torch.empty((self.out_channels, self.stride), dtype=torch.float32, device="cuda")
        ctx.use_torch_stream()

    @property
    def in_bytes(self) -> int:
        return self.shard.in_count * self.frame_bytes

    def run_device(self, d_in):
        """d_in: uint8 CUDA tensor holding this shard's packed frames (halo included)."""
        self.d_max.zero_()
        _lib.check(self.lib.aukit_cuda_dev_pipeline_peak(self.ctx.handle, C.byref(self.desc), d_in.data_ptr(), self.d_max.data_ptr()))
        allreduce_max_(self.d_max)
        _lib.check(self.lib.aukit_cuda_dev_pipeline_apply(self.ctx.handle, C.byref(self.desc), d_in.data_ptr(), self.peak,
                                                          self.d_max.data_ptr(), self.d_out.data_ptr(), self.stride))
        return self.d_out

    def run_host(self, h_in, d_in, h_out):
        """End to end from pinned host bytes to pinned host floats (H2D, passes, D2H)."""
        d_in.copy_(h_in, non_blocking=True)
        out = self.run_device(d_in)
        h_out.copy_(out[:, : self.shard.n_out], non_blocking=True)
        return h_out

    # ---- pipelined host path: aukit_cuda_preloader_* (clip i's D2H overlaps clip i+1's H2D)
    def _preloader(self):
        if getattr(self, "_pl", None) is None:
            h = C.c_void_p()
            _lib.check(self.lib.aukit_cuda_preloader_create(self.ctx.handle, self.in_bytes, self.stride * self.out_channels, 2,
                                                            C.byref(h)))
            self._pl = h
            self._pl_stream = self.torch.cuda.ExternalStream(int(self.lib.aukit_cuda_preloader_stream(h)))
            self._pl_peaks = {}
        return self._pl

    def _peak_tensor(self, slot: int):
        t = self._pl_peaks.get(slot)
        if t is None:
            ptr = int(self.lib.aukit_cuda_preloader_peak_ptr(self._pl, slot))

            class _Raw:                                    # one float32 in the preloader's slot
                __cuda_array_interface__ = {"shape": (1,), "typestr": "<f4", "data": (ptr, False), "version": 3}
            t = self._pl_peaks[slot] = self.torch.as_tensor(_Raw(), device="cuda")
        return t

    def submit_host(self, h_in, h_out):
        """Asynchronous run_host: returns once the copies and passes are enqueued; call drain() before
        reading h_out or reusing h_in.  h_in / h_out: pinned torch tensors (uint8 bytes / float32)."""
        pl = self._preloader()
        k = C.c_int(-1)
        _lib.check(self.lib.aukit_cuda_preloader_begin(pl, C.byref(self.desc), h_in.data_ptr(), h_in.numel(), C.byref(k)))
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            with self.torch.cuda.stream(self._pl_stream):
                allreduce_max_(self._peak_tensor(k.value))
        _lib.check(self.lib.aukit_cuda_preloader_finish(pl, k.value, self.peak, h_out.data_ptr()))

    def drain(self):
        if getattr(self, "_pl", None) is not None:
            _lib.check(self.lib.aukit_cuda_preloader_drain(self._pl))

    def close(self):
        if getattr(self, "_pl", None) is not None:
            self.lib.aukit_cuda_preloader_destroy(self._pl)
            self._pl = None
