"""Sub-benchmarks of bench.py: one record per BASELINE.json config beside the headline (config 2), each with its
algorithmic bytes, device-timed ms, GB/s, fraction of the measured HBM copy rate, clocks sampled
during its timed region, and an in-run parity spot check (a slice re-done on the CPU by the oracle: bit-exact for the
ADPCM decoders, <= 2^-20 for the float chains).

  c3      1024 x 60 s clips, s24 big-endian stereo, 22.05 / 44.1 / 96 kHz round robin -> 48 kHz cubic -> amplify(0.5);
          sharded BY CLIP over the ranks (strong scaling: the batch is fixed), no collective.
  c4_ima  10 h, 8 channels, blockAlign 8192, IMA ADPCM -> f32; sharded by contiguous BLOCK RANGE, no collective.
  c4_ms   the same, MS-ADPCM (encoder-like nibbles; uniformly random nibbles reported beside it as the worst case).
  c5      24 h 96 kHz 8-channel f32 in 8 time shards (3 h + halo each) -> 48 kHz -> normalize; rank r of N runs shard
          (8/N)(r+1)-1, the peak is MAX-combined over the N ranks.  With N = 8 this is the whole of config 5.
  c5p     the same -> 44.1 kHz cubic (a non-integer ratio, so the halo arithmetic and the exact fp64 positions beyond
          2^30.5 frames are exercised; SURVEY finding 10).
  c2_noise  config 2 on full-scale white noise (SURVEY 8d's first signal): cubic overshoot makes the channel clamp of
          A:668 act in every tile, i.e. K10's worst case.

The oracle (oracle/) is used here ONLY as the checker of those slices.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

TOL = 2.0 ** -20


class Env:
    def __init__(self, torch, dist, ak, ctx, rank, world, local, peak_gbs, sampler_cls):
        self.torch, self.dist, self.ak, self.ctx, self.lib = torch, dist, ak, ctx, ctx.lib
        self.rank, self.world, self.local, self.peak = rank, world, local, peak_gbs
        self.Sampler = sampler_cls
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """W untimed + K timed calls, device events on the launch stream, barrier + synchronize on both sides,
        MAX over ranks; clocks sampled during the timed region on this rank's GPU."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.ctx.launches
        with self.Sampler(self.local) as clk:
            e0.record(self.stream)
            for _ in range(steps):
                fn()
            e1.record(self.stream)
            self.barrier()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item()), (self.ctx.launches - l0) // max(steps, 1), clk.summary()

    def sum_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=torch_f64(self.torch), device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def all_ok(self, ok):
        t = self.torch.tensor([1.0 if ok else 0.0], dtype=torch_f64(self.torch), device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)


def torch_f64(torch):
    return torch.float64


def record(env, name, workload, ms, total_bytes, my_bytes, units, unit_name, launches, clocks, parity, extra=None):
    gbs = total_bytes / (ms * 1e-3) / 1e9
    rec = {"workload": workload, "n_gpus": env.world, "ms": ms, "algorithmic_bytes": int(total_bytes),
           "GB/s": gbs, "per_gpu_GB/s": my_bytes / (ms * 1e-3) / 1e9, "frac": my_bytes / (ms * 1e-3) / 1e9 / env.peak,
           "frac_of_nominal_8000": my_bytes / (ms * 1e-3) / 1e9 / 8000.0,
           "value": units / (ms * 1e-3) / 1e6, "unit": "M%s/s" % unit_name, "gpu_launches_per_step": launches, "clocks": clocks,
           "parity_ok": parity["ok"], "parity": parity}
    if extra:
        rec.update(extra)
    return rec


# ------------------------------------------------------------------------------------------------ config 3
def bench_c3(env, nclips=1024, seconds=60, steps=5, warmup=3):
    import numpy as np
    from oracle import oracle as O
    from util import ref_resample_window
    torch, ak, lib, ctx = env.torch, env.ak, env.lib, env.ctx
    rates_all = [[22050, 44100, 96000][k % 3] for k in range(nclips)]
    mine = [k for k in range(nclips) if k % env.world == env.rank]          # shard by clip (independent units)
    clips = (ak.Clip * len(mine))()
    off = 0
    for j, k in enumerate(mine):
        fr = rates_all[k] * seconds
        clips[j].in_offset, clips[j].frames, clips[j].srcRate = off, fr, float(rates_all[k])
        off += (fr * 6 + 15) // 16 * 16
    total_out = int(lib.aukit_batch_plan(clips, len(mine), 2, 48000.0))
    d_in = torch.empty(off + 16, dtype=torch.uint8, device="cuda")
    for j, k in enumerate(mine):
        g = torch.Generator(device="cuda")
        g.manual_seed(3_000_000 + k)                                         # per-clip seed 3*10^6 + k
        n = int(clips[j].frames) * 6
        d_in[int(clips[j].in_offset): int(clips[j].in_offset) + n] = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda", generator=g)
    d_out = torch.empty(total_out, dtype=torch.float32, device="cuda")

    def step():
        ak._lib.check(lib.aukit_cuda_dev_batch_resample_amplify(ctx.handle, clips, len(mine), 24, 0, 2, 1, 48000.0, 2, 0.5,
                                                                d_in.data_ptr(), d_out.data_ptr()))
    ms, launches, clocks = env.timed(step, steps, warmup)
    my_in = sum(int(c.frames) * 6 for c in clips)
    my_out = sum(int(c.n_out) * 2 * 4 for c in clips)
    tot_in, tot_out = env.sum_ranks(my_in), env.sum_ranks(my_out)
    # parity: one clip per rate class; head, middle and tail slices against the oracle's pcm -> resample -> amplify
    worst, checked = 0.0, 0
    for want in (22050, 44100, 96000):
        j = next((j for j, k in enumerate(mine) if rates_all[k] == want), None)
        if j is None:
            continue
        c = clips[j]
        n_out, fr = int(c.n_out), int(c.frames)
        for o0 in (0, n_out // 2 // 4 * 4, n_out - 2048):
            cnt = 2048
            f0 = max(0, int(o0 * want / 48000) - 4)
            f1 = min(fr, int((o0 + cnt) * want / 48000) + 8)
            raw = d_in[int(c.in_offset) + f0 * 6: int(c.in_offset) + f1 * 6].cpu().numpy().tobytes()
            win = O.pcm(raw, 24, "signed", 2, True, True)
            ref = O.amplify(ref_resample_window(win, f0, fr, want, 48000, o0, cnt, "cubic"), 0.5)
            for ch in range(2):
                base = int(c.out_offset) + ch * int(c.out_stride) + o0
                got = d_out[base: base + cnt].cpu().numpy().astype(np.float64)
                worst = max(worst, float(np.max(np.abs(got - ref[ch]))))
            checked += 2 * cnt
    ok = env.all_ok(worst <= TOL and checked > 0)
    del d_in, d_out
    torch.cuda.empty_cache()
    return record(env, "c3", "%d x %d s clips, s24 big-endian stereo (assumed), 22.05/44.1/96 kHz round robin -> 48 kHz cubic -> amplify(0.5); "
                  "fused single pass per rate class (K14), clips sharded round robin over %d GPU(s), no collective"
                  % (nclips, seconds, env.world), ms, tot_in + tot_out, my_in + my_out, tot_out / 4, "samples", launches, clocks,
                  {"ok": ok, "max_abs_err": worst, "tolerance": TOL, "samples_checked": checked,
                   "how": "oracle pcm -> resample -> amplify on head / middle / tail slices of one clip per rate class (this rank)"},
                  {"scaling": "strong", "B_in": int(tot_in), "B_out": int(tot_out),
                   "data": "uniform random bytes per clip, torch CUDA generator seeded 3e6 + k (full-scale 24-bit noise)"})


# ------------------------------------------------------------------------------------------------ config 4
def plan_block_shards(nblocks, world):
    from aukit_b200.sharding import plan_block_shards as p
    return p(nblocks, world)


def bench_c4(env, kind, steps=5, warmup=3, nblocks_total=None, wild=False):
    import numpy as np
    from oracle import oracle as O
    from util import ms_blocks
    torch, ak, lib, ctx = env.torch, env.ak, env.lib, env.ctx
    CH, BA = 8, 8192
    if nblocks_total is None:
        nblocks_total = 778_236 if kind == "ima" else 779_765                # 10 h at 44.1 kHz
    first, count = plan_block_shards(nblocks_total, env.world)[env.rank]
    g = torch.Generator(device="cuda")
    g.manual_seed(4_000_000 + env.rank)
    if kind == "ima" or wild:
        d_in = torch.randint(0, 256, (count, BA), dtype=torch.uint8, device="cuda", generator=g)   # uniform nibbles: both clamps act
        if kind == "ima":
            for c in range(CH):
                d_in[:, 4 * c + 2] = torch.randint(0, 89, (count,), dtype=torch.uint8, device="cuda", generator=g)   # step index 0..88
        else:
            d_in[:, :CH] = torch.randint(0, 7, (count, CH), dtype=torch.uint8, device="cuda", generator=g)           # predictor index 0..6
            delta = torch.randint(16, 2048, (count, CH), dtype=torch.int16, device="cuda", generator=g)               # delta 16..2047
            d_in[:, CH:3 * CH] = delta.view(torch.uint8).reshape(count, 2 * CH)
    else:
        proto = torch.from_numpy(ms_blocks(4096, BA, CH, seed=4, tame=True).reshape(4096, BA)).cuda()                # encoder-like nibbles
        reps = (count + 4095) // 4096
        d_in = proto.repeat(reps, 1)[:count].contiguous()
        del proto
    nbytes = count * BA
    if kind == "ima":
        frames = int(lib.aukit_ima_adpcm_wav_frames(nbytes, BA, CH, 1))
    else:
        frames = int(lib.aukit_msadpcm_frames(nbytes, BA, CH))
    spb = frames // count
    stride = (frames + 31) // 32 * 32
    d_out = torch.empty((CH, stride), dtype=torch.float32, device="cuda")

    def step():
        if kind == "ima":
            ak._lib.check(lib.aukit_cuda_dev_ima_adpcm_wav(ctx.handle, d_in.data_ptr(), nbytes, BA, CH, 1, d_out.data_ptr(), stride))
        else:
            ak._lib.check(lib.aukit_cuda_dev_msadpcm(ctx.handle, d_in.data_ptr(), nbytes, BA, CH, None, None, 0, 1, d_out.data_ptr(), stride))
    ms, launches, clocks = env.timed(step, steps, warmup)
    ctx.synchronize()                                                        # surfaces device-side header errors
    # parity: four blocks of this rank's range, bit-exact against the generalised oracle (C = 8; SURVEY finding 7)
    bad, checked = 0, 0
    for b in sorted({0, 1, count // 2, count - 1}):
        blk = d_in[b].cpu().numpy()
        ref = (O.wav_ima(blk, BA, CH, O.GENERAL) if kind == "ima" else O.msadpcm(blk, BA, CH, None, O.GENERAL)).astype(np.float32)
        got = d_out[:, b * spb: (b + 1) * spb].cpu().numpy()
        bad += int(np.sum(got.view(np.uint32) != ref.view(np.uint32)))
        checked += ref.size
    ok = env.all_ok(bad == 0 and checked > 0)
    my_bytes = nbytes + frames * CH * 4
    tot = env.sum_ranks(my_bytes)
    tot_samples = env.sum_ranks(frames * CH)
    del d_in, d_out
    torch.cuda.empty_cache()
    name = "IMA ADPCM" if kind == "ima" else "MS-ADPCM"
    return record(env, "c4_" + kind, "10 h 8-channel 44.1 kHz (assumed) WAV %s, blockAlign 8192, %d blocks -> f32 (GENERAL N-channel dialect: the "
                  "reference rejects > 2 channels, A:1349 / A:1199); contiguous block ranges over %d GPU(s), no collective"
                  % (name, nblocks_total, env.world), ms, tot, my_bytes, tot_samples, "samples", launches, clocks,
                  {"ok": ok, "mismatching_samples": bad, "samples_checked": checked,
                   "how": "bit-exact f32 == (float)oracle on blocks {0, 1, middle, last} of this rank's range"},
                  {"scaling": "strong", "blocks_this_rank": count, "first_block_this_rank": first,
                   "data": ("uniformly random nibbles, valid headers (both clamps act)" if (kind == "ima" or wild) else
                            "4096 distinct encoder-like blocks (tests/util.ms_blocks, tame) tiled")})


# ------------------------------------------------------------------------------------------------ config 5 / 5'
def hashed_f32(torch, first_elem, count):
    """Deterministic function of the GLOBAL element index (frame * 8 + channel): uniform in (-0.75, 0.75), so the halo
    frames two neighbouring shards both hold carry identical values."""
    out = torch.empty(count, dtype=torch.float32, device="cuda")
    step = 1 << 26
    for s in range(0, count, step):
        n = min(step, count - s)
        idx = torch.arange(first_elem + s, first_elem + s + n, dtype=torch.int64, device="cuda")
        h = (idx * -7046029254386353131 & 0x7FFFFFFFFFFFFFFF) >> 23      # 0x9E3779B97F4A7C15 as int64, wrapping
        h = (h ^ (h >> 17)) * 0x2545F491 & 0xFFFFFF
        out[s: s + n] = (h.to(torch.float32) - 8388608.0) * (0.75 / 8388608.0)
        del idx, h
    return out


def bench_c5(env, dst_rate, steps=3, warmup=3, hours=24.0, shards=8, cache=None):
    import numpy as np
    from util import ref_resample_window
    from aukit_b200._lib import PipelineDesc
    from aukit_b200.sharding import PeerExchange, plan_time_shards
    torch, ak, lib, ctx = env.torch, env.ak, env.lib, env.ctx
    CH, SRC = 8, 96000
    n_in_total = int(hours * 3600 * SRC)
    per = max(1, shards // env.world)
    shard_idx = min(shards - 1, per * (env.rank + 1) - 1)
    sh = plan_time_shards(n_in_total, SRC, dst_rate, "cubic", shards)[shard_idx]
    # input window: generated once for both target rates (the windows differ by a few halo frames: take the union)
    key = (shard_idx, n_in_total)
    if cache is not None and cache.get("key") == key:
        w_first, w_count, d_in_all = cache["first"], cache["count"], cache["buf"]
    else:
        a = plan_time_shards(n_in_total, SRC, 48000, "cubic", shards)[shard_idx]
        b = plan_time_shards(n_in_total, SRC, 44100, "cubic", shards)[shard_idx]
        w_first = min(a.in_first, b.in_first)
        w_count = max(a.in_first + a.in_count, b.in_first + b.in_count) - w_first
        d_in_all = hashed_f32(torch, w_first * CH, w_count * CH)
        if cache is not None:
            cache.update(key=key, first=w_first, count=w_count, buf=d_in_all)
    d_in = d_in_all[(sh.in_first - w_first) * CH: (sh.in_first - w_first + sh.in_count) * CH]
    desc = PipelineDesc(32, 2, CH, 0, float(SRC), float(dst_rate), 2, 0, n_in_total, sh.in_first, sh.in_count, sh.out_first, sh.n_out)
    stride = (sh.n_out + 31) // 32 * 32
    d_out = torch.empty((CH, stride), dtype=torch.float32, device="cuda")
    d_max = torch.zeros(1, dtype=torch.float32, device="cuda")
    PEAK = 1.0
    comm = PeerExchange(ctx) if env.world > 1 else None        # the library's own MAX exchange (csrc/comm.cu)

    def step():
        d_max.zero_()
        ak._lib.check(lib.aukit_cuda_dev_pipeline_peak(ctx.handle, C.byref(desc), d_in.data_ptr(), d_max.data_ptr()))
        if comm is not None:
            comm.allreduce_max_(d_max)
        ak._lib.check(lib.aukit_cuda_dev_pipeline_apply(ctx.handle, C.byref(desc), d_in.data_ptr(), PEAK, d_max.data_ptr(), d_out.data_ptr(), stride))
    ms, launches, clocks = env.timed(step, steps, warmup)
    if comm is not None:
        torch.cuda.synchronize()
        comm.close()
    # parity: middle and last slices against the numpy restatement of A:653-673 (global fp64 positions) + A:3444-3455
    mx = float(d_max.item())
    worst, checked = 0.0, 0
    for o_rel in (0, sh.n_out // 2, sh.n_out - 1024):
        cnt = 1024
        o0 = sh.out_first + o_rel
        f = C.c_uint64(0)
        c = C.c_uint64(0)
        ak._lib.check(lib.aukit_resample_window(n_in_total, float(SRC), float(dst_rate), 2, o0, cnt, C.byref(f), C.byref(c)))
        lo = int(f.value) - sh.in_first
        win = d_in[lo * CH: (lo + int(c.value)) * CH].cpu().numpy().astype(np.float64).reshape(-1, CH).T.copy()
        ref = ref_resample_window(win, int(f.value), n_in_total, SRC, dst_rate, o0, cnt, "cubic")
        ref = np.clip(ref * (PEAK / np.float64(mx)), -1, 1)
        got = d_out[:, o_rel: o_rel + cnt].cpu().numpy().astype(np.float64)
        worst = max(worst, float(np.max(np.abs(got - ref))))
        checked += ref.size
    out_peak = float(d_out[:, : sh.n_out].abs().max().item())
    t = torch.tensor([out_peak], dtype=torch.float64, device="cuda")
    if env.world > 1:
        env.dist.all_reduce(t, op=env.dist.ReduceOp.MAX)
    ok = env.all_ok(worst <= TOL) and abs(float(t.item()) - PEAK) <= 2.0 ** -20
    in_bytes, out_bytes = sh.in_count * CH * 4, sh.n_out * CH * 4
    my_bytes = 2 * in_bytes + out_bytes
    tot = env.sum_ranks(my_bytes)
    tot_samples = env.sum_ranks(sh.n_out * CH)
    del d_out
    torch.cuda.empty_cache()
    tag = "c5" if dst_rate == 48000 else "c5p"
    return record(env, tag, "%g h 96 kHz 8-channel f32 single buffer in %d time shards (%.1f h + interpolation halo each) -> %d Hz cubic -> "
                  "normalize(1.0); this run holds shard(s) %s of %d on %d GPU(s); MAX exchange of one float between the passes"
                  % (hours, shards, hours / shards, dst_rate, "all" if env.world == shards else "(8/N)(r+1)-1", shards, env.world),
                  ms, tot, my_bytes, tot_samples, "samples", launches, clocks,
                  {"ok": ok, "max_abs_err": worst, "tolerance": TOL, "samples_checked": checked, "global_output_peak": float(t.item()),
                   "how": "numpy restatement of A:653-673 with GLOBAL fp64 positions + normalize with the exchanged max, on three 1024-frame "
                          "slices of this rank's shard; max |output| over all ranks == peakAmplitude"},
                  {"scaling": "weak", "shard_index_this_rank": shard_idx, "per_gpu_in_bytes": in_bytes, "per_gpu_out_bytes": out_bytes,
                   "bytes_convention": "2*B_in + B_out (peak pass re-reads the input)",
                   "data": "hash of the global element index, uniform in (-0.75, 0.75) (the host cannot hold 265 GB)"})


# ------------------------------------------------------------------------------------------------ K10 worst case
def bench_c2_noise(env, sp_factory, n_in, steps, warmup):
    """Config 2 on default_rng-style full-scale white noise.  Returns (record, device input) -- parity on slices."""
    import numpy as np
    from oracle import oracle as O
    from util import ref_resample_window
    torch = env.torch
    sp, shard = sp_factory()
    g = torch.Generator(device="cuda")
    g.manual_seed(2)
    d_in = torch.randint(-32768, 32768, (shard.in_count, 2), dtype=torch.int16, device="cuda", generator=g).view(torch.uint8).reshape(-1)
    ms, launches, clocks = env.timed(lambda: sp.run_device(d_in), steps, warmup)
    mx = float(sp.d_max.item())
    ak, lib, ctx = env.ak, env.lib, env.ctx
    ms_peak, _, _ = env.timed(lambda: ak._lib.check(lib.aukit_cuda_dev_pipeline_peak(ctx.handle, C.byref(sp.desc), d_in.data_ptr(), sp.d_max.data_ptr())),
                              steps, 2)
    ms_apply, _, _ = env.timed(lambda: ak._lib.check(lib.aukit_cuda_dev_pipeline_apply(ctx.handle, C.byref(sp.desc), d_in.data_ptr(), sp.peak,
                                                                                       sp.d_max.data_ptr(), sp.d_out.data_ptr(), sp.stride)), steps, 2)
    worst = 0.0
    for o_rel in (0, shard.n_out // 2 // 4 * 4, shard.n_out - 4096):
        cnt = 4096
        o0 = shard.out_first + o_rel
        f0 = max(shard.in_first, int(o0 * 44100 / 48000) - 4)
        f1 = min(shard.in_first + shard.in_count, int((o0 + cnt) * 44100 / 48000) + 8)
        raw = d_in[(f0 - shard.in_first) * 4: (f1 - shard.in_first) * 4].cpu().numpy().tobytes()
        win = O.pcm(raw, 16, "signed", 2, True, False)
        r = ref_resample_window(win, f0, n_in, 44100, 48000, o0, cnt, "cubic")
        ref = np.clip(((0.0 + r[0]) + r[1]) / 2 * (0.8 / np.float64(mx)), -1, 1)
        got = sp.d_out[0, o_rel: o_rel + cnt].cpu().numpy().astype(np.float64)
        worst = max(worst, float(np.max(np.abs(got - ref))))
    ok = env.all_ok(worst <= TOL)
    in_bytes, out_bytes = d_in.numel(), shard.n_out * 4
    del d_in
    return {"ms_per_step": ms, "ms_peak_pass": ms_peak, "ms_apply_pass": ms_apply, "value": None, "achieved_GB/s_per_gpu": (2 * in_bytes + out_bytes) / (ms * 1e-3) / 1e9,
            "frac": (2 * in_bytes + out_bytes) / (ms * 1e-3) / 1e9 / env.peak, "gpu_launches_per_step": launches, "clocks": clocks,
            "signal": "torch CUDA generator seed 2, integers in [-32768, 32768): full-scale white noise, the channel clamp of A:668 acts in every tile",
            "parity_ok": ok, "parity": {"ok": ok, "max_abs_err": worst, "tolerance": TOL,
                                        "how": "oracle decode + numpy restatement of A:653-689 + normalize with the device max, 3 x 4096 outputs"}}
