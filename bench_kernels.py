#!/usr/bin/env python
"""Per-kernel throughput / roofline table for K1..K9 (auxiliary to bench.py, which times the fused
K10 chain the headline metric is quoted on).  Device-resident inputs, CUDA events, inputs larger
than L2.  One JSON line per kernel: algorithmic bytes (B_in + B_out), achieved GB/s,
fraction of the measured HBM copy rate.

    python bench_kernels.py [--scale 1.0] > profiles/rNN_kernels.jsonl
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3, help="untimed launches before timing (0 for ncu captures)")
    args = ap.parse_args()
    from bench import ClockSampler
    t = table(ClockSampler, 0, args.scale, args.iters, args.warm)
    for rec in t["rows"]:
        print(json.dumps(rec), flush=True)
    print(json.dumps({"clocks": t["clocks"]}), flush=True)


def table(sampler_cls, gpu_index=0, scale=1.0, iters=10, warm=3):
    """Times K1..K13 one by one; returns one record per kernel, each with the clocks sampled over the whole table's run
    attached once under the last record's sibling key (see bench.py: line["kernels"])."""
    class A:
        pass
    args = A()
    args.scale, args.iters, args.warm = scale, iters, warm
    records = []
    import numpy as np
    import torch
    import aukit_b200 as ak
    from bench import hbm_peak
    from util import ima_blocks, ms_blocks

    ctx = ak.context()
    ctx.use_torch_stream()
    lib = ctx.lib
    peak, src = hbm_peak()
    stream = torch.cuda.current_stream()

    def timed(fn):
        for _ in range(args.warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ctx.synchronize()
        return e0.elapsed_time(e1) / args.iters

    def report(name, ms, bytes_, units, unit_name, note=""):
        gbs = bytes_ / (ms * 1e-3) / 1e9
        records.append({"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": int(bytes_), "GB/s": round(gbs, 1),
                        "frac_of_measured_peak": round(gbs / peak, 3), "peak_GB/s": peak, unit_name + "/s": units / (ms * 1e-3),
                        "note": note})

    sampler = sampler_cls(gpu_index)
    sampler.__enter__()

    frames = int(158_760_000 * args.scale) // 64 * 64            # config 2: 1 h at 44.1 kHz

    def rand_bytes(n):
        return torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")

    # ---- K1 pcm_unpack
    for bits, dt, ch, be, label in ((16, 0, 2, 0, "s16le stereo (config 2 input)"), (24, 0, 2, 1, "s24be stereo (config 3 input)"),
                                    (32, 2, 8, 0, "f32le 8ch (config 5 input)"), (8, 1, 1, 0, "u8 mono")):
        n = frames if ch <= 2 else frames // 4
        nbytes = n * ch * bits // 8
        d_in = rand_bytes(nbytes)
        stride = (n + 31) // 32 * 32
        d_out = torch.empty((ch, stride), dtype=torch.float32, device="cuda")
        ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_pcm(ctx.handle, d_in.data_ptr(), nbytes, bits, dt, ch, 1, be, d_out.data_ptr(), stride)))
        report("K1 pcm_unpack " + label, ms, nbytes + n * ch * 4, n * ch, "samples")
        del d_in, d_out
    # ---- K2 g711
    n = frames * 2
    d_in = rand_bytes(n)
    d_out = torch.empty((2, n // 2), dtype=torch.float32, device="cuda")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_g711(ctx.handle, d_in.data_ptr(), n, 1, 2, d_out.data_ptr(), n // 2)))
    report("K2 g711 ulaw stereo", ms, n * 5, n, "samples")
    del d_in, d_out
    # ---- K3 / K4 ADPCM, 8 channels, blockAlign 8192 (config 4), 4/10 of its length by default (20 GB of f32 output)
    nblocks = int(311_296 * args.scale) // 512 * 512
    for kind in ("ima", "ms"):
        proto = (ima_blocks if kind == "ima" else ms_blocks)(512, 8192, 8, seed=4)
        d_in = torch.from_numpy(np.tile(proto, nblocks // 512)).cuda()
        nb = d_in.numel()
        if kind == "ima":
            fr = int(lib.aukit_ima_adpcm_wav_frames(nb, 8192, 8, 1))
        else:
            fr = int(lib.aukit_msadpcm_frames(nb, 8192, 8))
        stride = (fr + 31) // 32 * 32
        d_out = torch.empty((8, stride), dtype=torch.float32, device="cuda")
        if kind == "ima":
            f = lambda: ak._lib.check(lib.aukit_cuda_dev_ima_adpcm_wav(ctx.handle, d_in.data_ptr(), nb, 8192, 8, 1, d_out.data_ptr(), stride))
        else:
            f = lambda: ak._lib.check(lib.aukit_cuda_dev_msadpcm(ctx.handle, d_in.data_ptr(), nb, 8192, 8, None, None, 0, 1, d_out.data_ptr(), stride))
        ms = timed(f)
        report("K%d %s_adpcm 8ch blockAlign 8192" % (3 if kind == "ima" else 4, kind), ms, nb + fr * 8 * 4, fr * 8, "samples",
               "serial chain per (block, channel), 32 chains per warp, 256-byte flushes aligned in absolute address"
               + ("" if kind == "ima" else "; records staged through shared memory, speculated straight-line periods"))
        del d_in, d_out
    # ---- K5 resample f32 stereo 44.1 -> 48 kHz
    n = frames // 2
    x = torch.rand((2, n), device="cuda") * 2 - 1
    for mode, name in ((2, "cubic"), (1, "linear"), (0, "none")):
        n_out = int(lib.aukit_resample_out_len(n, 44100.0, 48000.0))
        stride = (n_out + 31) // 32 * 32
        y = torch.empty((2, stride), dtype=torch.float32, device="cuda")
        ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_resample(ctx.handle, x.data_ptr(), n, 2, n, 0, n, 44100.0, 48000.0, mode, 0, n_out,
                                                                     y.data_ptr(), stride)))
        report("K5 resample %s f32 stereo 44.1->48k" % name, ms, (n + n_out) * 2 * 4, n_out * 2, "samples",
               "polyphase weights per thread, TMA-staged double-buffered tiles (resample_planar.cu); exact fp64 decision for the one output per period that sits on an input frame")
    # sinc (8f rank 4): 21 taps per output; a quarter of the buffer is enough for a stable figure
    ns = n // 4
    ns_out = int(lib.aukit_resample_out_len(ns, 44100.0, 48000.0))
    ys = torch.empty((2, (ns_out + 31) // 32 * 32), dtype=torch.float32, device="cuda")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_resample(ctx.handle, x.data_ptr(), n, 2, ns, 0, ns, 44100.0, 48000.0, 3, 0, ns_out,
                                                                 ys.data_ptr(), ys.shape[1])))
    report("K5 resample sinc f32 stereo 44.1->48k", ms, (ns + ns_out) * 2 * 4, ns_out * 2, "samples", "21 taps per output (A:267-281), weight table per phase + first-order correction")
    del ys
    # ---- K6 mono, K7 amplify, K8 absmax, K9 scale_clamp
    m = torch.empty(n, dtype=torch.float32, device="cuda")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_mono(ctx.handle, x.data_ptr(), n, 2, n, m.data_ptr())))
    report("K6 mono 2ch", ms, n * 12, n, "samples")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_amplify(ctx.handle, x.data_ptr(), n, 2, n, 0.999)))
    report("K7 amplify 2ch", ms, n * 2 * 8, n * 2, "samples")
    dmax = torch.zeros(2, dtype=torch.float32, device="cuda")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_absmax(ctx.handle, x.data_ptr(), n, 2, n, 0, dmax.data_ptr())))
    report("K8 absmax 2ch", ms, n * 2 * 4, n * 2, "samples")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_scale_clamp(ctx.handle, x.data_ptr(), n, 2, n, 1.0, 0, dmax.data_ptr())))
    report("K9 scale_clamp 2ch", ms, n * 2 * 8, n * 2, "samples")
    # ---- K13 remaining in-place effects (8f rank 4): 8 B per touched sample
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_invert(ctx.handle, x.data_ptr(), n, 2, n)))
    report("K13 invert 2ch", ms, n * 2 * 8, n * 2, "samples")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_fade(ctx.handle, x.data_ptr(), n, 2, n, 48000.0, 1.0, 0.5, (n - 1) / 48000.0, 1.0)))
    report("K13 fade 2ch (whole buffer)", ms, n * 2 * 8, n * 2, "samples")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_center(ctx.handle, x.data_ptr(), n, 2, n, 48000.0)))
    report("K13 center 2ch (1 s blocks)", ms, n * 2 * 8, n * 2, "samples", "a block's second read (subtract) is L2-resident: 4 B read + 4 B written from HBM")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_delay(ctx.handle, x.data_ptr(), n, 2, n, 48000.0, 0.25, 0.5)))
    report("K13 delay 2ch", ms, n * 2 * 20, n * 2, "samples", "copy (4+4) + two reads + one write per sample")
    # ---- K12 requantisation (8f rank 2): Audio:pcm values (fp64 out) and packed s16 / u8 bytes
    ev = torch.empty(n * 2, dtype=torch.float64, device="cuda")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_encode_pcm(ctx.handle, x.data_ptr(), n, 2, n, 16, 0, 1, ev.data_ptr())))
    report("K12 encode_pcm values s16 stereo interleaved", ms, n * 2 * 12, n * 2, "samples", "4 B read + 8 B (fp64, un-rounded) written")
    eb = torch.empty(n * 2 * 2, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_encode_pcm_bytes(ctx.handle, x.data_ptr(), n, 2, n, 16, 0, 1, 0, eb.data_ptr())))
    report("K12 encode_pcm bytes s16 stereo interleaved", ms, n * 2 * 6, n * 2, "samples")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_encode_pcm_bytes(ctx.handle, x.data_ptr(), n, 1, n, 8, 0, 1, 1, eb.data_ptr())))
    report("K12 encode_pcm bytes s8 mono (speaker format)", ms, n * 5, n, "samples")
    del ev, eb
    # ---- K11 lowpass (8f rank 1): in place, 4 B read + 4 B written per sample
    for f in (24000.0, 200.0):
        ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_lowpass(ctx.handle, x.data_ptr(), n, 2, n, f, 48000.0)))
        report("K11 lowpass 2ch f=%g Hz" % f, ms, n * 2 * 8, n * 2, "samples", "blocked chunks (state carried per CTA) at this size, fp64 state")
    ms = timed(lambda: ak._lib.check(lib.aukit_cuda_dev_highpass(ctx.handle, x.data_ptr(), n, 2, n, 200.0, 48000.0)))
    report("K11 highpass 2ch f=200 Hz", ms, n * 2 * 8, n * 2, "samples", "same kernel, ratio a; last input sample handed from tile to tile in shared memory")
    sampler.__exit__(None, None, None)
    return {"clocks": sampler.summary(), "scale": scale, "iters": iters, "rows": records}


if __name__ == "__main__":
    main()
