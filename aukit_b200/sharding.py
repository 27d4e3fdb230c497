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


def shard_alignment(srcRate: float, dstRate: float) -> int:
    """Output frames per warp tile of the fused kernels for this ratio: 32 periods of L = dstRate / gcd outputs
    (5120 for 44.1 -> 48 kHz).  A shard that starts on a multiple of it has no partial head tile, so every rank runs
    the same interior kernel from its first output; 4 (a 16-byte store) when the ratio has no such tile."""
    import math
    if srcRate != int(srcRate) or dstRate != int(dstRate):
        return 4
    g = math.gcd(int(srcRate), int(dstRate))
    L = int(dstRate) // g
    return 32 * L if 1 < L <= 512 and (32 * L) % 4 == 0 else 4


def plan_time_shards(n_in_total: int, srcRate: float, dstRate: float, interpolation: str, world: int,
                     align: Optional[int] = None, pad: int = 8) -> List[TimeShard]:
    """Contiguous, near-equal output ranges; input windows overlap by the halo only.

    Interior boundaries fall on multiples of `align` outputs (default: shard_alignment) and every window is widened by
    up to `pad` frames on each side where the signal has them: the bulk copies of the run-per-lane kernels round their
    start down to 16 bytes and read a few frames past the last tap, and with that slack present a shard's first and
    last full tiles stay on the fast kernel instead of the edge kernels."""
    lib = _lib.load()
    n_out = int(lib.aukit_resample_out_len(n_in_total, float(srcRate), float(dstRate)))
    if align is None:
        align = shard_alignment(srcRate, dstRate)
    if n_out // max(world, 1) < 4 * align:
        align = 4
    shards = []
    for r in range(world):
        o0 = n_out * r // world // align * align
        o1 = n_out if r == world - 1 else n_out * (r + 1) // world // align * align
        first, count = padded_window(n_in_total, srcRate, dstRate, interpolation, o0, o1 - o0, pad)
        shards.append(TimeShard(r, o0, o1 - o0, first, count))
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


def padded_window(n_in_total: int, srcRate: float, dstRate: float, interpolation: str, out_first: int, n_out: int, pad: int = 8):
    """(in_first, in_count) for global output frames [out_first, out_first + n_out): the halo of
    aukit_resample_window plus the slack plan_time_shards adds (see there)."""
    lib = _lib.load()
    f, c = C.c_uint64(0), C.c_uint64(0)
    _lib.check(lib.aukit_resample_window(n_in_total, float(srcRate), float(dstRate), _INTERPS[interpolation], out_first, n_out,
                                         C.byref(f), C.byref(c)))
    first, end = int(f.value), int(f.value) + int(c.value)
    if pad and c.value:
        first = max(0, first - pad) // 4 * 4 if first >= pad else first
        end = min(n_in_total, end + pad)
    return first, end - first


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


class PeerExchange:
    """aukit_comm of this rank (include/aukit_cuda.h): the library's own MAX exchange over peer-mapped device memory.
    The 64-byte IPC handles are swapped once, at construction, through torch.distributed (any backend)."""

    def __init__(self, ctx, group=None):
        import torch
        import torch.distributed as dist
        self.ctx, self.lib = ctx, ctx.lib
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        h = C.c_void_p()
        _lib.check(self.lib.aukit_cuda_comm_create(ctx.handle, self.world, self.rank, C.byref(h)))
        self.handle = h
        nb = int(self.lib.aukit_cuda_comm_handle_bytes())
        mine = (C.c_ubyte * nb)()
        _lib.check(self.lib.aukit_cuda_comm_handle(h, mine))
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
        allh = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(allh, t, group=group)
        blob = b"".join(bytes(x.cpu().tolist()) for x in allh)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        # every rank has mapped every block before the first exchange -- or none keeps its communicator: the MIN over
        # ranks of "my mapping worked" doubles as the barrier, so a failure anywhere raises on ALL ranks together
        rc = self.lib.aukit_cuda_comm_connect(h, buf)
        err = "" if rc == 0 else self.lib.aukit_cuda_last_error().decode("latin-1")
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.close()
            raise _lib.AukitError(err or "aukit_cuda: another rank could not map its peers' exchange blocks")

    def allreduce_max_(self, t):
        _lib.check(self.lib.aukit_cuda_comm_allreduce_max(self.handle, t.data_ptr(), t.numel()))
        return t

    def close(self):
        if self.handle:
            self.lib.aukit_cuda_comm_destroy(self.handle)
            self.handle = None


class ShardedPreload:
    """auplay's chain (unpack -> resample -> mono -> normalize) on ONE rank's time shard.

    peak pass -> MAX over ranks -> apply pass.  All launches go to torch's current stream, nothing synchronises the
    host.  On GPUs the exchange is the library's own peer-memory kernel (PeerExchange / aukit_comm, one tiny launch
    between the passes); `exchange="torch"` keeps it on torch.distributed (NCCL, or gloo for the CPU-side tests)."""

    def __init__(self, ctx, shard: TimeShard, n_in_total: int, bitDepth=16, dataType="signed", channels=2,
                 srcRate=44100.0, dstRate=48000.0, interpolation="cubic", mono=True, peak=0.8, bigEndian=False,
                 exchange="peer"):
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.comm = None
        if exchange == "peer" and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # every rank must end up on the same exchange: if the peer mapping fails anywhere (e.g. ranks that cannot see
            # each other's device: CUDA IPC needs the peer visible), ALL ranks fall back to torch.distributed -- loudly
            try:
                self.comm = PeerExchange(ctx)                # raises on every rank together (see there)
            except _lib.AukitError as e:
                import warnings
                warnings.warn("aukit_b200: peer-memory exchange unavailable (%s); the normalize MAX goes through torch.distributed" % e)
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
        if self.comm is not None:
            self.comm.allreduce_max_(self.d_max)
        else:
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
    slots = 2        # device slots of the pipelined host path (three measured no better: 0.929 of the PCIe ceiling either way)

    def _preloader(self):
        if getattr(self, "_pl", None) is None:
            h = C.c_void_p()
            _lib.check(self.lib.aukit_cuda_preloader_create(self.ctx.handle, self.in_bytes, self.stride * self.out_channels, self.slots,
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
            if self.comm is not None:
                # the exchange kernel goes where the slot's passes go: the preloader's run stream
                old = int(self.lib.aukit_cuda_get_stream(self.ctx.handle) or 0)
                self.ctx.set_stream(int(self.lib.aukit_cuda_preloader_stream(pl)))
                try:
                    _lib.check(self.lib.aukit_cuda_comm_allreduce_max(self.comm.handle, self.lib.aukit_cuda_preloader_peak_ptr(pl, k.value), 1))
                finally:
                    self.ctx.set_stream(old)
            else:
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
        if self.comm is not None:
            self.comm.close()
            self.comm = None
