"""ctypes binding of the C ABI in include/jarvis_hybridnet_b200.h.

There is no fallback: if the shared library is missing and cannot be built, or a call returns a
non-zero status, a RuntimeError is raised (reference convention: Python exceptions)."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_size_t, c_ulonglong, c_void_p

from . import build as _build

FP32, BF16 = 0, 1
VOL_NCDHW_F32, VOL_V2V_BF16 = 0, 1
LERP_FMA_FIRST, LERP_FMA_SECOND, LERP_NO_FMA = 0, 1, 2
HM_F32_PLANAR, HM_F16_CL, HM_BF16_CL = 0, 1, 2           # enum jhn_heatmap_format
HM_CL_PITCH, HM_F16_SCALE = 24, 0.0625
ABI_VERSION = _build.ABI_VERSION

# every symbol include/jarvis_hybridnet_b200.h declares: name -> (restype, argtypes)
_P = c_void_p
SYMBOLS = {
    "jhn_last_error": (c_char_p, []),
    "jhn_abi_version": (c_int, []),
    "jhn_check_device": (c_int, [c_int]),
    "jhn_reproject_workspace_bytes": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "jhn_reproject_gather": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float,
                                     c_int, c_float, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "jhn_heatmap_convert": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "jhn_v2v_debug_head_centroid": (c_int, [_P, _P, c_int, c_int, c_float, c_float, _P, _P, _P, _P, _P, c_size_t, _P]),
    "jhn_set_sub_batch": (c_int, [c_int]),
    "jhn_v2v_create": (c_int, [POINTER(_P), c_int, c_int, c_int, _P, POINTER(_P)]),
    "jhn_v2v_destroy": (None, [_P]),
    "jhn_v2v_set_workspace_persistent": (c_int, [_P, c_int]),
    "jhn_v2v_workspace_bytes": (c_int, [_P, c_int, c_int, POINTER(c_size_t)]),
    "jhn_v2v_forward": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "jhn_v2v_debug_layer_workspace_bytes": (c_int, [_P, c_int, c_int, c_int, POINTER(c_size_t)]),
    "jhn_v2v_debug_layer": (c_int, [_P, c_int, _P, c_int, c_int, _P, _P, c_size_t, _P]),
    "jhn_centroid_reduce": (c_int, [_P, c_int, c_int, c_int, c_float, c_float, _P, _P, _P, _P, _P]),
    "jhn_hybrid3d_workspace_bytes": (c_int, [_P, c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "jhn_hybrid3d_forward": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_float,
                                     c_int, _P, _P, _P, _P, c_size_t, _P]),
    "jhn_center_locate": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P,
                                  _P, _P, _P, _P, _P, _P, _P, _P]),
    "jhn_crop_normalize": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, POINTER(c_float), POINTER(c_float), _P, _P]),
    "jhn_heatmap_boxes": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P, _P]),
    "jhn_upload_heatmap_boxes": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, POINTER(c_size_t)]),
    "jhn_pull_heatmap_boxes": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "jhn_ingest_frames": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "jhn_crop_normalize_u8": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, POINTER(c_float), POINTER(c_float), _P, _P]),
    "jhn_efftrack_head": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "jhn_softplus2": (c_int, [_P, ctypes.c_longlong, _P, _P]),
    "jhn_pad_heatmaps": (c_int, [_P, ctypes.c_longlong, c_int, _P, _P]),
    # not part of the drop-in surface: launch counter used by bench.py's `gpu_launches`
    "jhn_launch_count": (c_ulonglong, []),
    "jhn_profile_enable": (None, [c_int]),
    "jhn_profile_collect": (c_int, [c_char_p, c_int]),
    "jhn_debug_set_gather_box_bytes": (c_int, [c_int]),
    "jhn_debug_set_pull_config": (None, [c_int, c_int, c_int]),
    "jhn_pull_small": (c_int, [c_int, _P, _P, _P, _P]),
    "jhn_set_transfer_overlap": (None, [c_int]),
    "jhn_heatmap_spans": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P, c_size_t, _P, _P, _P]),
    "jhn_pull_heatmap_spans": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P]),
}

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load (building first if the .so is absent). Raises if the extension cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        _build.build()                # no-op unless a source / header is newer than the library (a stale library would
    except Exception as e:            # be called with the new argument lists)
        if not os.path.exists(_build.LIB):
            raise RuntimeError(f"jarvis_hybridnet_b200: CUDA extension {_build.LIB} is missing and could not be "
                               f"built ({e}); there is no CPU fallback") from e
        # no compiler on this box (the GPU box ships the prebuilt library): the ABI check below decides
    lib = ctypes.CDLL(_build.LIB)
    lib.jhn_abi_version.restype = c_int
    if lib.jhn_abi_version() != ABI_VERSION:
        raise RuntimeError(f"jarvis_hybridnet_b200: {_build.LIB} has ABI {lib.jhn_abi_version()}, the Python binding "
                           f"expects {ABI_VERSION}; rebuild it (python jarvis-hybridnet_b200/build.py --force)")
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().jhn_last_error()
        raise RuntimeError(f"jarvis_hybridnet_b200 status {rc}: {msg.decode() if msg else ''}")


def launch_count():
    return int(load().jhn_launch_count())


def debug_set_gather_box_bytes(n):
    """Test hook: pixel-box slot capacity of the streaming gather (0 = default); returns the value in effect."""
    return int(load().jhn_debug_set_gather_box_bytes(int(n)))


def profile(on):
    load().jhn_profile_enable(1 if on else 0)


def profile_collect():
    """{kernel name: (launches, total_ms)} recorded since profile(True); synchronises the device."""
    buf = ctypes.create_string_buffer(1 << 16)
    load().jhn_profile_collect(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split("\t")
        out[name] = (int(n), float(ms))
    return out


# ---- torch helpers ---------------------------------------------------------------------------------
def dptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    """All tensors on ONE CUDA device; returns a context that makes it current (streams, handles and the library's
    launches all follow the current device)."""
    import torch
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("jarvis_hybridnet_b200 runs on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"jarvis_hybridnet_b200: inputs live on different devices ({dev} and {t.device})")
    return torch.cuda.device(dev) if dev is not None else None


def set_sub_batch(n):
    """Frame sets per internal pass of jhn_hybrid3d_forward (0 = default); returns the value in effect."""
    return int(load().jhn_set_sub_batch(int(n)))


def heatmap_convert(heatmaps, hs, fmt=HM_F16_CL):
    """fp32 planar heat maps [B,ncam,K,S,S] (S = hs or hs-2) -> channels-last 16-bit [B,ncam,hs,hs,24] with the
    F.pad border (`jhn_heatmap_convert`, the producer side of SURVEY.md section 8 row f2)."""
    import torch
    with require_cuda(heatmaps):
        B, ncam, K, S, S2 = heatmaps.shape
        if S != S2 or S not in (hs, hs - 2):
            raise RuntimeError(f"heat maps are {S}x{S2}; expected {hs} (padded) or {hs - 2} per side")
        hm = heatmaps.contiguous().float()
        out = torch.empty((B, ncam, hs, hs, HM_CL_PITCH), dtype=torch.float16 if fmt == HM_F16_CL else torch.bfloat16,
                          device=hm.device)
        check(load().jhn_heatmap_convert(dptr(hm), int(S == hs), B, ncam, K, hs, int(fmt), dptr(out), stream_ptr()))
    return out
