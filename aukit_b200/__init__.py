"""aukit_b200 -- AUKit's preload audio path (decode -> resample -> mono -> normalize) on B200.

Hand-written sm_100a CUDA kernels behind a plain C ABI (include/aukit_cuda.h, built into
aukit_b200/lib/libaukit_cuda.so by `python -m aukit_b200.build`), plus this Python host layer
that mirrors the reference's Lua API names, argument order, defaults and error strings.
There is no CPU fallback: without the built library and a B200 every call raises.
"""
from ._lib import AukitError, Clip, ContainerInfo, PipelineDesc, WavInfo, SIGNATURES, LIB_PATH  # noqa: F401
from .aukit import (  # noqa: F401
    Audio, Context, context, effects, pcm, g711, adpcm, msadpcm, wav, wav_info, new, preload, preload_clips, Preloader, Group, au, aiff,
    DIALECT_LITERAL, DIALECT_GENERAL, _VERSION,
)
from . import aukit as _aukit


def __getattr__(name):
    if name == "defaultInterpolation":
        return _aukit.defaultInterpolation
    raise AttributeError(name)


def set_default_interpolation(name: str):
    """aukit.defaultInterpolation = name (A:99; auconvert.lua:187 sets it from --interpolation)."""
    _aukit.defaultInterpolation = name
