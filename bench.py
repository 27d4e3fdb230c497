#!/usr/bin/env python
"""bench.py -- AUKit preload path on B200: decode + 48 kHz cubic resample + mono + normalize.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): output Msamples/s of the whole job.  Workload at N=1 = BASELINE config 2
("1 h 44.1 kHz stereo s16 PCM -> cubic resample to 48 kHz + mono + normalize").  At N > 1 the
buffer is N hours long and time-sharded, one hour (+ interpolation halo) per GPU, the only
collective being the NCCL all-reduce(MAX) of one float between the two passes (weak scaling).

One step = one pass of the hot path over the whole buffer:
    peak kernel (reads packed PCM) -> [all-reduce MAX] -> apply kernel (re-reads PCM, writes f32).
`value`  : inputs already resident in HBM, CUDA events, max over ranks.
`e2e`    : the same through the host-buffer path: pinned host bytes -> H2D -> passes -> D2H.
`roofline`: dominant kernel (apply pass), algorithmic bytes B_in + B_out per launch over its
           event-timed duration, against MEASURED_PEAKS.json's HBM copy rate.
`cpu_baseline`: the oracle's C restatement of aukit.lua (kind "port": no Lua interpreter exists
           in this image) on one host core over a bounded sample of the same signal.

--impl reference times that same CPU restatement on all host cores (bench.py is one of the few
places allowed to execute oracle/); it is a reported baseline, never part of the product path.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SRC_RATE, DST_RATE, CHANNELS, BITS = 44100, 48000, 2, 16
INTERP, PEAK = "cubic", 0.8
METRIC = "output Msamples/s, decode+48kHz resample+mono+normalize"
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def synth_frames_np(first: int, count: int):
    """Config-1 style signal as a function of the GLOBAL frame index (so shard halos agree):
    440 / 660 Hz tones at half scale plus +-256 of hashed integer noise.  int16 [count, 2]."""
    import numpy as np
    idx = np.arange(first, first + count, dtype=np.int64)
    t = idx.astype(np.float64) / SRC_RATE
    out = np.empty((count, CHANNELS), dtype=np.int16)
    for c in range(CHANNELS):
        noise = (((idx * 2654435761 + c * 40503) >> 13) & 511) - 256
        out[:, c] = np.clip(np.round(0.5 * 32767 * np.sin(2 * np.pi * (440 + 220 * c) * t)) + noise, -32768, 32767)
    return out


def synth_frames_cuda(first: int, count: int, torch):
    out = torch.empty((count, CHANNELS), dtype=torch.int16, device="cuda")
    step = 1 << 24
    for s in range(0, count, step):
        n = min(step, count - s)
        idx = torch.arange(first + s, first + s + n, dtype=torch.int64, device="cuda")
        t = idx.to(torch.float64) / SRC_RATE
        for c in range(CHANNELS):
            noise = (((idx * 2654435761 + c * 40503) >> 13) & 511) - 256
            v = torch.round(0.5 * 32767 * torch.sin(2 * torch.pi * (440 + 220 * c) * t)) + noise
            out[s: s + n, c] = v.clamp_(-32768, 32767).to(torch.int16)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([x.strip() for x in r.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def cpu_port_rate(seconds_audio: float, threads: int, steps: int = 1, warmup: int = 0):
    """Times the oracle's chain (decode -> cubic resample -> mono -> normalize) on `threads`
    host cores, each on its own `seconds_audio`-long clip of the bench signal.  Returns
    (Msamples/s, ms per step, output samples per step)."""
    import numpy as np
    from oracle import oracle as O
    O.lib()
    n = int(seconds_audio * SRC_RATE)
    clips = [synth_frames_np(k * n, n).tobytes() for k in range(threads)]
    n_out = O.resample_len(n, SRC_RATE, DST_RATE)

    def one_step():
        if threads == 1:
            O.chain_s16(clips[0], CHANNELS, SRC_RATE, DST_RATE, INTERP, PEAK)
            return
        ts = [threading.Thread(target=O.chain_s16, args=(c, CHANNELS, SRC_RATE, DST_RATE, INTERP, PEAK)) for c in clips]
        [t.start() for t in ts]
        [t.join() for t in ts]      # ctypes releases the GIL: the C chains run in parallel

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    return n_out * threads / dt / 1e6, dt * 1e3, n_out * threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    threads = os.cpu_count() or 1
    clip_s = args.cpu_clip_seconds
    val, ms, n_step = cpu_port_rate(clip_s, threads, steps=args.steps, warmup=args.warmup)
    sample = "%d independent %.0f s clips per step (one per host thread) of the bench signal" % (threads, clip_s)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "1 h 44.1 kHz stereo s16 PCM -> cubic 48 kHz + mono + normalize(0.8) (BASELINE config 2), bounded sample",
                   "sample": sample, "output_samples_per_step": n_step,
                   "note": "aukit.lua is Lua and no Lua interpreter exists in this image; this is the oracle's literal C restatement of the same lines"},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pin_near_gpu(torch, local):
    """Run this rank on the CPUs next to its GPU (sysfs local_cpulist), so that pinned host buffers are first
    touched on the GPU's own NUMA node -- on a multi-socket host the e2e copies of 8 ranks otherwise all cross
    one socket.  Best effort (aukit_cuda_host_alloc does the same for the C ABI's own allocations)."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dev)
        cpus = set()
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except (OSError, ValueError, AttributeError):
        pass


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    torch.cuda.set_device(local)
    pin_near_gpu(torch, local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import aukit_b200 as ak
    from aukit_b200 import build as akbuild
    from aukit_b200.sharding import ShardedPreload, plan_time_shards
    if not os.path.exists(ak.LIB_PATH):
        if rank == 0:
            akbuild.build()
        if world > 1:
            dist.barrier()
    ctx = ak.context(local)
    lib = ctx.lib

    n_in_total = int(args.seconds * SRC_RATE) * world
    shard = plan_time_shards(n_in_total, SRC_RATE, DST_RATE, INTERP, world)[rank]
    if args.emulate_shard:
        # diagnostics: time ONE GPU on shard r of a w-hour buffer (no collective); not a bench line
        r, w = (int(v) for v in args.emulate_shard.split("/"))
        n_in_total = int(args.seconds * SRC_RATE) * w
        shard = plan_time_shards(n_in_total, SRC_RATE, DST_RATE, INTERP, w)[r]
    n_out_total = int(lib.aukit_resample_out_len(n_in_total, float(SRC_RATE), float(DST_RATE)))
    if args.emulate_shard:
        n_out_total = shard.n_out
    sp = ShardedPreload(ctx, shard, n_in_total, BITS, "signed", CHANNELS, SRC_RATE, DST_RATE, INTERP, True, PEAK)
    d_in = synth_frames_cuda(shard.in_first, shard.in_count, torch).view(torch.uint8).reshape(-1)
    in_bytes, out_bytes = d_in.numel(), shard.n_out * 4
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx.launches - l0

    # ---- device-resident step (value)
    with ClockSampler(local) as clk:
        ms_step, launches = timed(lambda: sp.run_device(d_in), args.steps, args.warmup)
    clocks = clk.summary()
    value = n_out_total / (ms_step * 1e-3) / 1e6

    # ---- N > 1: (1) every rank's shard must equal what ONE GPU computes for the same global output range (rank 0
    # regenerates the input of rank r's first 2^20 outputs from the global frame index and re-runs the apply pass with
    # the exchanged max); (2) strong scaling: ONE hour split over the N GPUs
    multi = None
    if world > 1 and not args.emulate_shard:
        from aukit_b200._lib import PipelineDesc
        from aukit_b200.sharding import padded_window, shard_alignment
        CHK = min(200 * shard_alignment(SRC_RATE, DST_RATE), shard.n_out)       # whole warp tiles, like the shards themselves
        mine = sp.d_out[0, :CHK].clone()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        bits_ok = True
        if rank == 0:
            shards_all = plan_time_shards(n_in_total, SRC_RATE, DST_RATE, INTERP, world)
            for r in range(1, world):
                wf, wc = padded_window(n_in_total, SRC_RATE, DST_RATE, INTERP, shards_all[r].out_first, CHK)
                din_r = synth_frames_cuda(wf, wc, torch).view(torch.uint8).reshape(-1)
                desc_r = PipelineDesc(BITS, 0, CHANNELS, 0, float(SRC_RATE), float(DST_RATE), 2, 1, n_in_total, wf, wc,
                                      shards_all[r].out_first, CHK)
                out_r = torch.empty(CHK, dtype=torch.float32, device="cuda")
                ak._lib.check(lib.aukit_cuda_dev_pipeline_apply(ctx.handle, C.byref(desc_r), din_r.data_ptr(), PEAK, sp.d_max.data_ptr(),
                                                                out_r.data_ptr(), CHK))
                torch.cuda.synchronize()
                bits_ok = bits_ok and bool(torch.equal(out_r, gathered[r]))
                del din_r, out_r
        n_in_1h = int(args.seconds * SRC_RATE)
        sh1 = plan_time_shards(n_in_1h, SRC_RATE, DST_RATE, INTERP, world)[rank]
        sp1 = ShardedPreload(ctx, sh1, n_in_1h, BITS, "signed", CHANNELS, SRC_RATE, DST_RATE, INTERP, True, PEAK)
        d_in1 = synth_frames_cuda(sh1.in_first, sh1.in_count, torch).view(torch.uint8).reshape(-1)
        ms_strong, _ = timed(lambda: sp1.run_device(d_in1), args.steps, args.warmup)
        n_out_1h = int(lib.aukit_resample_out_len(n_in_1h, float(SRC_RATE), float(DST_RATE)))
        sp1.close()
        del d_in1, sp1
        # the same weak-scaling step with the exchange done by NCCL (torch.distributed.all_reduce(MAX) between the passes):
        # what north_star names and what round 1 shipped -- the baseline the in-library exchange is measured against
        ms_nccl = None
        if dist.get_backend() == "nccl":
            spn = ShardedPreload(ctx, shard, n_in_total, BITS, "signed", CHANNELS, SRC_RATE, DST_RATE, INTERP, True, PEAK, exchange="torch")
            ms_nccl, _ = timed(lambda: spn.run_device(d_in), args.steps, args.warmup)
            nccl_same = bool(torch.equal(spn.d_out[0, :4096], sp.d_out[0, :4096]))
            spn.close()
            del spn
        multi = {"shard_bits_equal_single_gpu": bits_ok, "shard_check": "rank 0 recomputed the first %d outputs of every other rank's shard "
                 "(own input window from the global frame index, exchanged max) and compared bits" % CHK,
                 "strong": {"workload": "%g h total split over %d GPUs" % (args.seconds / 3600.0, world), "ms_per_step": ms_strong,
                            "value": n_out_1h / (ms_strong * 1e-3) / 1e6, "scaling": "strong"},
                 "exchange": "aukit_comm: one tagged 8-byte posted write per peer into peer-mapped exchange blocks (CUDA IPC) + a poll on the own block, one kernel per rank in stream order between the passes (csrc/comm.cu)"
                             if sp.comm is not None else "torch.distributed all_reduce(MAX)"}
        if ms_nccl is not None:
            multi["nccl_exchange"] = {"ms_per_step": ms_nccl, "value": n_out_total / (ms_nccl * 1e-3) / 1e6, "same_bits": nccl_same,
                                      "what": "the same step with torch.distributed.all_reduce(MAX) over NCCL between the passes"}

    # ---- per-kernel timing for the roofline (same stream, events around each launch batch)
    desc = sp.desc

    def peak_only():
        ak._lib.check(lib.aukit_cuda_dev_pipeline_peak(ctx.handle, C.byref(desc), d_in.data_ptr(), sp.d_max.data_ptr()))

    def apply_only():
        ak._lib.check(lib.aukit_cuda_dev_pipeline_apply(ctx.handle, C.byref(desc), d_in.data_ptr(), PEAK, sp.d_max.data_ptr(),
                                                        sp.d_out.data_ptr(), sp.stride))

    ms_peak, _ = timed(peak_only, args.steps, 2)
    ms_apply, _ = timed(apply_only, args.steps, 2)
    peak_gbs, peak_src = hbm_peak()
    apply_bytes, peak_bytes = in_bytes + out_bytes, in_bytes
    # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture of this same workload (per launch)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_apply_traffic.json")))
        if world == 1 and not args.emulate_shard and abs(args.seconds - 3600.0) < 1e-6:
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except (OSError, ValueError, KeyError):
        pass
    achieved = apply_bytes / (ms_apply * 1e-3) / 1e9

    # ---- end to end through the host-buffer path (pinned host memory both ways)
    h_in = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    h_in.copy_(d_in)
    h_out = torch.empty((1, shard.n_out), dtype=torch.float32, pin_memory=True)
    d_stage = torch.empty_like(d_in)
    torch.cuda.synchronize()
    e2e_steps = max(3, min(args.steps, 10))
    ms_single, _ = timed(lambda: sp.run_host(h_in, d_stage, h_out), e2e_steps, 3)      # one clip at a time: H2D, passes, D2H in series
    del d_stage
    # e2e proper: the same clips through the pipelined C-ABI preloader (aukit_cuda_preloader_*): PCIe is full
    # duplex, so clip i's download overlaps clip i+1's upload.  Every clip's H2D and D2H is inside the timed
    # region; the closing event is recorded after drain() has seen the last download finish.
    h_out2 = torch.empty((1, shard.n_out), dtype=torch.float32, pin_memory=True)
    outs = [h_out, h_out2]

    def e2e_run(steps):
        for i in range(steps):
            sp.submit_host(h_in, outs[i & 1])
        sp.drain()

    e2e_run(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    e2e_run(e2e_steps)
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    e2e_value = n_out_total / (ms_e2e * 1e-3) / 1e6
    checksum = float(h_out.abs().max())
    same = bool(torch.equal(h_out, h_out2))
    e2e_slots = sp.slots

    # ---- raw PCIe ceiling of the same byte counts: concurrent pinned H2D + D2H on two streams, no kernels
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    d_tmp_in = torch.empty(in_bytes, dtype=torch.uint8, device="cuda")

    def copies(steps):
        ev = torch.cuda.Event()
        ev.record(stream)
        s_up.wait_event(ev)
        s_dn.wait_event(ev)
        for i in range(steps):
            with torch.cuda.stream(s_up):
                d_tmp_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_dn):
                outs[i & 1].copy_(sp.d_out[:, : shard.n_out], non_blocking=True)
        stream.wait_stream(s_up)
        stream.wait_stream(s_dn)

    copies(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    copies(e2e_steps)
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_ceiling = float(t.item())
    del d_tmp_in
    # ---- what a Lua string really costs: the same clip from PAGEABLE host memory through aukit_cuda_pipeline_host
    # (staged through the context's pinned slices), results to pageable memory; host-synchronous call, wall clock
    ms_pageable = None
    if world == 1 and not args.emulate_shard:
        import numpy as np
        pg_in = np.empty(in_bytes, dtype=np.uint8)
        pg_in[:] = h_in.numpy()
        pg_out = np.empty((1, shard.n_out), dtype=np.float32)
        ctx.set_stream(None)
        for it in range(3):
            if it == 1:
                t0 = time.perf_counter()
            ak._lib.check(lib.aukit_cuda_pipeline_host(ctx.handle, C.byref(sp.desc), C.c_void_p(pg_in.ctypes.data), in_bytes, PEAK,
                                                       C.c_void_p(pg_out.ctypes.data)))
        ms_pageable = (time.perf_counter() - t0) / 2 * 1e3
        ctx.use_torch_stream()
        del pg_in, pg_out

    # ---- K10 worst case + the other BASELINE configs + per-kernel table (bench_configs.py)
    import bench_configs as BC
    env = BC.Env(torch, dist, ak, ctx, rank, world, local, peak_gbs, ClockSampler)
    noise = None
    if not args.no_configs:
        noise = BC.bench_c2_noise(env, lambda: (sp, shard), n_in_total, args.steps, args.warmup)
        noise["value"] = n_out_total / (noise["ms_per_step"] * 1e-3) / 1e6

    # ---- sustained figure: the same step back to back for >= 2 s (clocks settle; MEASURED_PEAKS saw 1245 MHz under load)
    sustained = None
    if args.sustained_steps > 0:
        barrier()
        with ClockSampler(local) as clk2:
            ms_sus, _ = timed(lambda: sp.run_device(d_in), args.sustained_steps, 3)
        sustained = {"steps": args.sustained_steps, "ms_per_step": ms_sus, "seconds": ms_sus * args.sustained_steps * 1e-3,
                     "value": n_out_total / (ms_sus * 1e-3) / 1e6, "clocks": clk2.summary(),
                     "whole_step_frac": (2 * in_bytes + out_bytes) / (ms_sus * 1e-3) / 1e9 / peak_gbs}
        sustained["note"] = "runs after the noise figure: the 2.4 s of load drive the board into its power cap, which would otherwise colour the next measurement"
    sp.close()
    del sp, d_in, h_in, h_out, h_out2, outs
    torch.cuda.empty_cache()
    configs, kernels = None, None
    if not args.no_configs:
        configs = {}
        want = [c.strip() for c in args.configs.split(",") if c.strip()]
        if "c3" in want:
            configs["c3"] = BC.bench_c3(env, nclips=args.c3_clips)
        if "c4" in want:
            configs["c4_ima"] = BC.bench_c4(env, "ima")
            configs["c4_ms"] = BC.bench_c4(env, "ms")
            wild = BC.bench_c4(env, "ms", wild=True)
            configs["c4_ms"]["worst_case_random_nibbles"] = {k: wild[k] for k in ("ms", "GB/s", "frac", "value", "parity_ok", "clocks", "data")}
        if "c5" in want:
            cache = {}
            configs["c5"] = BC.bench_c5(env, 48000, cache=cache)
            configs["c5p"] = BC.bench_c5(env, 44100, cache=cache)
            cache.clear()
            torch.cuda.empty_cache()
        if world == 1 and not args.no_kernels:
            import bench_kernels
            kernels = bench_kernels.table(ClockSampler, local, scale=args.kernel_scale)

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": "%g h 44.1 kHz stereo s16 PCM per GPU -> cubic resample to 48 kHz + mono + normalize(0.8) "
                            "(BASELINE config 2%s)" % (args.seconds / 3600.0, "" if world == 1 else ", time-sharded with halo, MAX exchange over peer memory between the passes"),
                "input_frames_total": n_in_total, "output_samples_total": n_out_total, "per_gpu_in_bytes": in_bytes,
                "per_gpu_out_bytes": out_bytes, "l2": "inputs (%.0f MB per pass) larger than the 126 MB L2; no flush needed" % (in_bytes / 1e6),
                "signal": "440/660 Hz tones at half scale + +-256 hashed integer noise (config-1 style), function of the global frame index",
                "parallelism": "time-shard x%d" % world, "output_peak_check": checksum,
            },
            "roofline": {
                "bound": "hbm", "kernel": "fused apply pass: run_static_kernel<APPLY=true> (interior) + poly_kernel edges", "achieved": achieved, "peak": peak_gbs,
                "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic,
                "traffic_source": "profiles/r2_apply_traffic.json: ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed --set full capture of this build (profiles/r2_final_ncu_run_kernel.txt), not measured in this run" if traffic else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": apply_bytes, "ms_per_launch": ms_apply,
                "peak_pass": {"algorithmic_bytes_per_launch": peak_bytes, "ms_per_launch": ms_peak,
                              "achieved": peak_bytes / (ms_peak * 1e-3) / 1e9},
                "whole_step": {"algorithmic_bytes": 2 * in_bytes + out_bytes, "compulsory_bytes": in_bytes + out_bytes,
                               "achieved": (2 * in_bytes + out_bytes) / (ms_step * 1e-3) / 1e9,
                               "frac": (2 * in_bytes + out_bytes) / (ms_step * 1e-3) / 1e9 / peak_gbs,
                               "frac_of_nominal_8000": (2 * in_bytes + out_bytes) / (ms_step * 1e-3) / 1e9 / 8000.0},
            },
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": in_bytes * world, "d2h_bytes_per_step": out_bytes * world,
                    "ms_per_step": ms_e2e, "steps": e2e_steps,
                    "mode": "pipelined preloader (C-ABI aukit_cuda_preloader_*): clip i's D2H overlaps clip i+1's H2D, %d device slots" % e2e_slots,
                    "single_clip_ms": ms_single, "single_clip_value": n_out_total / (ms_single * 1e-3) / 1e6,
                    "outputs_identical_across_slots": same,
                    "pcie_ceiling": {"ms_per_step": ms_ceiling, "value": n_out_total / (ms_ceiling * 1e-3) / 1e6,
                                     "frac_of_ceiling": ms_ceiling / ms_e2e,
                                     "how": "the same pinned H2D + D2H byte counts per clip on two streams, no kernels, max over ranks"},
                    "pageable_single_clip_ms": ms_pageable,
                    "pageable_note": "aukit_cuda_pipeline_host from pageable memory (what a Lua string is): staged through pinned slices"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if multi is not None:
            line["multi_gpu"] = multi
        if noise is not None:
            line["value_noise"] = noise["value"]
            line["noise"] = noise
        if sustained is not None:
            line["sustained"] = sustained
        if configs is not None:
            line["configs"] = configs
        if kernels is not None:
            line["kernels"] = kernels
        if world == 1 and not args.no_cpu:
            val, ms, _ = cpu_port_rate(args.cpu_seconds, 1)
            line["cpu_baseline"] = {"value": val, "unit": "Msamples/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
                                    "sample": "first %.0f s of the bench signal, one pass, oracle C restatement of aukit.lua (no Lua in image)" % args.cpu_seconds}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=float, default=3600.0, help="audio seconds per GPU (config 2 = 3600)")
    ap.add_argument("--cpu-seconds", type=float, default=1200.0, help="audio seconds of the 1-core cpu_baseline sample")
    ap.add_argument("--cpu-clip-seconds", type=float, default=30.0, help="--impl reference: audio seconds per thread per step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only (skip c3/c4/c5/c5p, the noise signal and the kernel table)")
    ap.add_argument("--configs", default="c3,c4,c5", help="which BASELINE configs to run beside the headline")
    ap.add_argument("--c3-clips", type=int, default=1024, help="clips in config 3 (1024 = BASELINE; fewer for profiler captures)")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel table (N = 1 only)")
    ap.add_argument("--kernel-scale", type=float, default=1.0, help="size of the per-kernel table's buffers relative to config 2")
    ap.add_argument("--sustained-steps", type=int, default=6000, help="extra back-to-back steps for the sustained figure (0 = skip)")
    ap.add_argument("--emulate-shard", default="", help="diagnostics: 'r/w' = run shard r of a w-GPU time-sharded buffer on one GPU")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
