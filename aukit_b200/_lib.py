"""ctypes binding of libaukit_cuda.so (include/aukit_cuda.h).  No fallback of any kind:
if the library is missing or no B200 is visible, calls raise."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libaukit_cuda.so")


class AukitError(RuntimeError):
    """Raised with the reference's own error string wherever aukit.lua would call error()."""


class WavTag(C.Structure):
    _fields_ = [("id", C.c_char * 5), ("off", C.c_size_t), ("len", C.c_size_t)]


class WavInfo(C.Structure):
    _fields_ = [("format", C.c_int), ("channels", C.c_int), ("sampleRate", C.c_int), ("blockAlign", C.c_int),
                ("bitDepth", C.c_int), ("have_fmt", C.c_int), ("ncoef", C.c_int), ("coef1", C.c_int * 256),
                ("coef2", C.c_int * 256), ("data_off", C.c_size_t), ("data_size", C.c_size_t),
                ("ntags", C.c_int), ("tags", WavTag * 64)]


class _ContainerMeta(C.Structure):
    _fields_ = [("key", C.c_char * 12), ("off", C.c_size_t), ("len", C.c_size_t)]


class ContainerInfo(C.Structure):
    """aukit_container_info (include/aukit_cuda.h)."""
    _fields_ = [("codec", C.c_int), ("bitDepth", C.c_int), ("dataType", C.c_int), ("bigEndian", C.c_int), ("ulaw", C.c_int),
                ("channels", C.c_int), ("sampleRate", C.c_double), ("data_off", C.c_size_t), ("data_len", C.c_size_t),
                ("nmeta", C.c_int), ("meta", _ContainerMeta * 16)]


class Clip(C.Structure):
    """aukit_clip (include/aukit_cuda.h)."""
    _fields_ = [("in_offset", C.c_uint64), ("frames", C.c_uint64), ("srcRate", C.c_double), ("out_offset", C.c_uint64),
                ("out_stride", C.c_uint64), ("n_out", C.c_uint64)]


class PipelineDesc(C.Structure):
    _fields_ = [("bitDepth", C.c_int), ("dataType", C.c_int), ("channels", C.c_int), ("bigEndian", C.c_int),
                ("srcRate", C.c_double), ("dstRate", C.c_double), ("interpolation", C.c_int), ("mono", C.c_int),
                ("n_in_total", C.c_uint64), ("in_first", C.c_uint64), ("in_avail", C.c_size_t),
                ("out_first", C.c_uint64), ("n_out", C.c_size_t)]


_P, _SZ, _I, _D, _U64 = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_uint64
_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/aukit_cuda.h declares
SIGNATURES = {
    "aukit_cuda_abi_version": (_I, []),
    "aukit_cuda_last_error": (C.c_char_p, []),
    "aukit_cuda_init": (_I, [_I, _PP]),
    "aukit_cuda_shutdown": (None, [_P]),
    "aukit_cuda_set_stream": (_I, [_P, _P]),
    "aukit_cuda_get_stream": (_P, [_P]),
    "aukit_cuda_make_current": (_I, [_P]),
    "aukit_cuda_synchronize": (_I, [_P]),
    "aukit_cuda_launch_count": (_U64, [_P]),
    "aukit_cuda_audio_new": (_I, [_P, _I, _SZ, _D, _PP]),
    "aukit_cuda_audio_wrap": (_I, [_P, _P, _I, _SZ, _SZ, _D, _PP]),
    "aukit_cuda_audio_free": (None, [_P, _P]),
    "aukit_cuda_audio_channels": (_I, [_P]),
    "aukit_cuda_audio_frames": (_SZ, [_P]),
    "aukit_cuda_audio_stride": (_SZ, [_P]),
    "aukit_cuda_audio_sample_rate": (_D, [_P]),
    "aukit_cuda_audio_data": (_P, [_P]),
    "aukit_cuda_audio_set_sample_rate": (_I, [_P, _D]),
    "aukit_cuda_audio_stream_chunk": (_I, [_P, _P, _I, _I, _SZ, _SZ, _P, C.POINTER(_SZ)]),
    "aukit_cuda_audio_channel_frames": (_SZ, [_P, _I]),
    "aukit_cuda_audio_download": (_I, [_P, _P, _I, _SZ, _SZ, _P]),
    "aukit_cuda_audio_upload": (_I, [_P, _P, _I, _SZ, _SZ, _P]),
    "aukit_cuda_pcm": (_I, [_P, _P, _SZ, _I, _I, _I, _D, _I, _I, _PP]),
    "aukit_cuda_g711": (_I, [_P, _P, _SZ, _I, _I, _D, _PP]),
    "aukit_cuda_adpcm": (_I, [_P, _P, _SZ, _I, _D, _I, _I, _P, _P, _PP]),
    "aukit_cuda_ima_adpcm_wav": (_I, [_P, _P, _SZ, _I, _I, _D, _I, _PP]),
    "aukit_cuda_msadpcm": (_I, [_P, _P, _SZ, _I, _I, _D, _P, _P, _I, _I, _PP]),
    "aukit_cuda_wav_parse": (_I, [_P, _SZ, C.POINTER(WavInfo)]),
    "aukit_cuda_wav": (_I, [_P, _P, _SZ, _I, _I, C.POINTER(WavInfo), _PP]),
    "aukit_cuda_resample": (_I, [_P, _P, _D, _I, _PP]),
    "aukit_cuda_mono": (_I, [_P, _P, _PP]),
    "aukit_cuda_concat": (_I, [_P, _PP, _I, _PP]),
    "aukit_cuda_amplify": (_I, [_P, _P, _D]),
    "aukit_cuda_normalize": (_I, [_P, _P, _D, _I]),
    "aukit_cuda_absmax": (_I, [_P, _P, _I, _P]),
    "aukit_cuda_scale_clamp": (_I, [_P, _P, _D, _I, _P]),
    "aukit_cuda_dev_pcm": (_I, [_P, _P, _SZ, _I, _I, _I, _I, _I, _P, _SZ]),
    "aukit_cuda_dev_g711": (_I, [_P, _P, _SZ, _I, _I, _P, _SZ]),
    "aukit_cuda_dev_ima_adpcm_wav": (_I, [_P, _P, _SZ, _I, _I, _I, _P, _SZ]),
    "aukit_cuda_dev_msadpcm": (_I, [_P, _P, _SZ, _I, _I, _P, _P, _I, _I, _P, _SZ]),
    "aukit_cuda_dev_resample": (_I, [_P, _P, _SZ, _I, _U64, _U64, _SZ, _D, _D, _I, _U64, _SZ, _P, _SZ]),
    "aukit_cuda_dev_mono": (_I, [_P, _P, _SZ, _I, _SZ, _P]),
    "aukit_cuda_dev_amplify": (_I, [_P, _P, _SZ, _I, _SZ, _D]),
    "aukit_cuda_dev_absmax": (_I, [_P, _P, _SZ, _I, _SZ, _I, _P]),
    "aukit_cuda_dev_scale_clamp": (_I, [_P, _P, _SZ, _I, _SZ, _D, _I, _P]),
    "aukit_cuda_au_parse": (_I, [_P, _SZ, C.POINTER(ContainerInfo)]),
    "aukit_cuda_aiff_parse": (_I, [_P, _SZ, C.POINTER(ContainerInfo)]),
    "aukit_cuda_au": (_I, [_P, _P, _SZ, C.POINTER(ContainerInfo), C.POINTER(_P)]),
    "aukit_cuda_aiff": (_I, [_P, _P, _SZ, _I, C.POINTER(ContainerInfo), C.POINTER(_P)]),
    "aukit_cuda_invert": (_I, [_P, _P]),
    "aukit_cuda_fade": (_I, [_P, _P, _D, _D, _D, _D]),
    "aukit_cuda_delay": (_I, [_P, _P, _D, _D]),
    "aukit_cuda_center": (_I, [_P, _P]),
    "aukit_cuda_dev_invert": (_I, [_P, _P, _SZ, _I, _SZ]),
    "aukit_cuda_dev_fade": (_I, [_P, _P, _SZ, _I, _SZ, _D, _D, _D, _D, _D]),
    "aukit_cuda_dev_delay": (_I, [_P, _P, _SZ, _I, _SZ, _D, _D, _D]),
    "aukit_cuda_dev_center": (_I, [_P, _P, _SZ, _I, _SZ, _D]),
    "aukit_cuda_lowpass": (_I, [_P, _P, _D]),
    "aukit_cuda_highpass": (_I, [_P, _P, _D]),
    "aukit_cuda_dev_highpass": (_I, [_P, _P, _SZ, _I, _SZ, _D, _D]),
    "aukit_cuda_audio_pcm": (_I, [_P, _P, _I, _I, _I, _P]),
    "aukit_cuda_audio_pcm_bytes": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "aukit_cuda_dev_encode_pcm": (_I, [_P, _P, _SZ, _I, _SZ, _I, _I, _I, _P]),
    "aukit_cuda_dev_encode_pcm_bytes": (_I, [_P, _P, _SZ, _I, _SZ, _I, _I, _I, _I, _P]),
    "aukit_cuda_dev_lowpass": (_I, [_P, _P, _SZ, _I, _SZ, _D, _D]),
    "aukit_cuda_dev_pipeline_peak": (_I, [_P, C.POINTER(PipelineDesc), _P, _P]),
    "aukit_cuda_dev_pipeline_apply": (_I, [_P, C.POINTER(PipelineDesc), _P, _D, _P, _P, _SZ]),
    "aukit_batch_plan": (_U64, [C.POINTER(Clip), _SZ, _I, _D]),
    "aukit_cuda_dev_batch_resample_amplify": (_I, [_P, C.POINTER(Clip), _SZ, _I, _I, _I, _I, _D, _I, _D, _P, _P]),
    "aukit_cuda_batch_resample_amplify": (_I, [_P, C.POINTER(_P), C.POINTER(_SZ), C.POINTER(_D), _SZ, _I, _I, _I, _I, _D, _I, _D,
                                               C.POINTER(_P)]),
    "aukit_cuda_pipeline_host": (_I, [_P, C.POINTER(PipelineDesc), _P, _SZ, _D, _P]),
    "aukit_cuda_preloader_create": (_I, [_P, _SZ, _SZ, _I, C.POINTER(_P)]),
    "aukit_cuda_preloader_destroy": (None, [_P]),
    "aukit_cuda_preloader_begin": (_I, [_P, C.POINTER(PipelineDesc), _P, _SZ, C.POINTER(_I)]),
    "aukit_cuda_preloader_peak_ptr": (_P, [_P, _I]),
    "aukit_cuda_preloader_stream": (_P, [_P]),
    "aukit_cuda_preloader_finish": (_I, [_P, _I, _D, _P]),
    "aukit_cuda_preloader_submit": (_I, [_P, C.POINTER(PipelineDesc), _P, _SZ, _D, _P]),
    "aukit_cuda_preloader_drain": (_I, [_P]),
    "aukit_cuda_host_alloc": (_I, [_SZ, C.POINTER(_P)]),
    "aukit_cuda_host_free": (None, [_P]),
    "aukit_resample_out_len": (_U64, [_U64, _D, _D]),
    "aukit_resample_position": (_D, [_U64, _D, _D]),
    "aukit_resample_window": (_I, [_U64, _D, _D, _I, _U64, _U64, C.POINTER(_U64), C.POINTER(_U64)]),
    "aukit_cuda_comm_create": (_I, [_P, _I, _I, C.POINTER(_P)]),
    "aukit_cuda_comm_handle_bytes": (_SZ, []),
    "aukit_cuda_comm_handle": (_I, [_P, _P]),
    "aukit_cuda_comm_connect": (_I, [_P, _P]),
    "aukit_cuda_comm_connect_local": (_I, [C.POINTER(_P), _I]),
    "aukit_cuda_comm_destroy": (None, [_P]),
    "aukit_cuda_comm_allreduce_max": (_I, [_P, _P, _I]),
    "aukit_cuda_comm_values": (_P, [_P]),
    "aukit_cuda_comm_normalize": (_I, [_P, _P, _D, _I]),
    "aukit_cuda_comm_pipeline": (_I, [_P, C.POINTER(PipelineDesc), _P, _D, _P, _SZ]),
    "aukit_cuda_group_create": (_I, [C.POINTER(_I), _I, C.POINTER(_P)]),
    "aukit_cuda_group_destroy": (None, [_P]),
    "aukit_cuda_group_size": (_I, [_P]),
    "aukit_cuda_group_ctx": (_P, [_P, _I]),
    "aukit_cuda_group_comm": (_P, [_P, _I]),
    "aukit_cuda_group_normalize": (_I, [_P, C.POINTER(_P), _D, _I]),
    "aukit_cuda_group_preload": (_I, [_P, C.POINTER(PipelineDesc), _P, _SZ, _D, _P]),
    "aukit_cuda_group_preload_audio": (_I, [_P, _P, C.POINTER(PipelineDesc), _P, _SZ, _D, _PP]),
    "aukit_cuda_preload_audio": (_I, [_P, C.POINTER(PipelineDesc), _P, _SZ, _D, _PP]),
    "aukit_block_shard": (_I, [_U64, _I, _I, C.POINTER(_U64), C.POINTER(_U64)]),
    "aukit_ima_adpcm_wav_frames": (_SZ, [_SZ, _I, _I, _I]),
    "aukit_msadpcm_frames": (_SZ, [_SZ, _I, _I]),
}

_lib = None


def load():
    """Loads libaukit_cuda.so; raises if it has not been built (python -m aukit_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AukitError(
                "libaukit_cuda.so is not built (%s missing). Run `python -m aukit_b200.build`; "
                "there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)   # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.aukit_cuda_abi_version() != 1:
            raise AukitError("libaukit_cuda.so ABI mismatch")
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise AukitError(load().aukit_cuda_last_error().decode("latin-1"))
