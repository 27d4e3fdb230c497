import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import aukit_b200 as ak
from util import ms_blocks, ima_blocks
ctx = ak.context(); ctx.use_torch_stream(); lib = ctx.lib
stream = torch.cuda.current_stream()
def timed(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(it): fn()
    e1.record(stream); torch.cuda.synchronize(); ctx.synchronize()
    return e0.elapsed_time(e1) / it
for kind in ("ms", "ima"):
    for bA in ((8192,) if kind == "ms" else (8192, 8224, 1024, 2048)):
        nblocks = 311296 // 2
        proto = (ms_blocks if kind == "ms" else ima_blocks)(512, bA, 8, seed=4)
        d_in = torch.from_numpy(np.tile(proto, nblocks // 512)).cuda()
        nb = d_in.numel()
        fr = int(lib.aukit_msadpcm_frames(nb, bA, 8)) if kind == "ms" else int(lib.aukit_ima_adpcm_wav_frames(nb, bA, 8, 1))
        stride = (fr + 31) // 32 * 32
        d_out = torch.empty((8, stride), dtype=torch.float32, device="cuda")
        if kind == "ms":
            f = lambda: ak._lib.check(lib.aukit_cuda_dev_msadpcm(ctx.handle, d_in.data_ptr(), nb, bA, 8, None, None, 0, 1, d_out.data_ptr(), stride))
        else:
            f = lambda: ak._lib.check(lib.aukit_cuda_dev_ima_adpcm_wav(ctx.handle, d_in.data_ptr(), nb, bA, 8, 1, d_out.data_ptr(), stride))
        ms = timed(f)
        spb = fr // nblocks
        print(kind, bA, "spb", spb, "row bytes", spb * 4, "mod32", spb * 4 % 32, "ms %.3f" % ms, "GB/s %.0f" % ((nb + fr * 32) / ms / 1e6), flush=True)
        del d_in, d_out
