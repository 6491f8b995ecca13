"""ORACLE — TEST INFRASTRUCTURE ONLY (see hybridnet_oracle.c for the contract).

CPU restatement of the reference's 3D inference hot path, one function per reference stage:

  reproject_indices / gather_mean / repro_layer_forward   jarvis/hybridnet/repro_layer.py:40-119
  v2v_forward                                             jarvis/hybridnet/v2vnet.py:12-102
  centroid_tail                                           jarvis/hybridnet/model.py:72-87
  hybrid3d_forward                                        jarvis/hybridnet/model.py:65-88 (the chain)

Index / gather / reduction work runs in the C library (exact fp32 op order, FMA only where the
reference's libraries use it); the V2V network is floating-point convolution work and is restated
with torch-CPU fp32 functional ops (`F.conv3d`, `F.conv_transpose3d`, `F.instance_norm`) — the
"torch fp32 reference" for the floating-point kernels.

Parity status: PINNED against the reference imported in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhybridnet_oracle.so")
_lib = None


def build(force=False):
    """Compile the C restatement with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "hybridnet_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/libhybridnet_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()                                        # no-op unless the C source is newer than the library
        _lib = ctypes.CDLL(_SO)
    return _lib


NUM_THREADS = os.cpu_count() or 1


def _threads(fn, n):
    """Run fn((begin,end)) over a split of range(n) on NUM_THREADS host threads (ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    t = max(1, min(NUM_THREADS, n))
    cuts = [round(i * n / t) for i in range(t + 1)]
    parts = [(cuts[i], cuts[i + 1]) for i in range(t) if cuts[i + 1] > cuts[i]]
    if len(parts) == 1:
        return [fn(parts[0])]
    with ThreadPoolExecutor(len(parts)) as ex:
        return list(ex.map(fn, parts))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def reproject_indices(center3D, centerHM, cameraMatrices, intrinsicMatrices, distortionCoefficients,
                      G, spacing, hs, lerp_mode=0, return_coarse=False):
    """`res` of repro_layer.py:82-83 for one frame set: int64 [ncam,G,G,G], flat padded-pixel index."""
    cam, intr, dist = _f32(cameraMatrices), _f32(intrinsicMatrices), _f32(distortionCoefficients)
    ncam = cam.shape[0]
    assert cam.shape == (ncam, 4, 3) and intr.shape == (ncam, 3, 3) and dist.shape == (ncam, 1, 5)
    c3, chm = _f32(center3D).reshape(3), _i32(centerHM).reshape(ncam, 2)
    h = G // 2
    idx = np.empty((ncam, G, G, G), np.int32)
    ca = np.empty((ncam, h, h, h), np.float32)
    cb = np.empty((ncam, h, h, h), np.float32)
    def run(rng):
        return lib().jho_reproject_indices(_p(cam), _p(intr), _p(dist), _p(c3), _p(chm), ncam, G,
                                           ctypes.c_float(spacing), hs, lerp_mode, _p(idx), _p(ca), _p(cb),
                                           rng[0], rng[1])
    if any(rc != 0 for rc in _threads(run, ncam)):
        raise ValueError("jho_reproject_indices failed (bad shape)")
    idx = idx.astype(np.int64)
    return (idx, ca, cb) if return_coarse else idx


def pixel_boxes_and_row_spans(coarse_x, coarse_y, hs):
    """What a host-buffer caller may upload instead of whole heat maps (no counterpart in the reference, which uploads everything):
    per camera the pixel box {x0, y0, x1, y1} of the voxel grid and, per pixel row, the column range [lo, hi] (lo > hi: no voxel maps
    to the row).  coarse_x / coarse_y: [ncam,h,h,h] coarse projections (`reproject_indices(..., return_coarse=True)`).

    A fine voxel's (x, y) is a chain of rounded convex combinations of the 8 coarse corners of its cell (ATen's trilinear upsample,
    repro_layer.py:78-81) and (v / 2).int() (:82-83) is monotone, so its pixel lies inside the integer box of those 8 corners: every
    cell adds the columns of its box to the rows its box covers; the camera's box is the union."""
    ca, cb = _f32(coarse_x), _f32(coarse_y)
    ncam, h = ca.shape[0], ca.shape[1]
    px, py = np.trunc(ca * np.float32(0.5)).astype(np.int64), np.trunc(cb * np.float32(0.5)).astype(np.int64)

    def cells(a, f):                                   # reduce over the 8 corners of every cell (a single point: one degenerate cell)
        for ax in (1, 2, 3):
            if a.shape[ax] > 1:
                lo_ = [slice(None)] * 4; hi_ = [slice(None)] * 4
                lo_[ax] = slice(0, -1); hi_[ax] = slice(1, None)
                a = f(a[tuple(lo_)], a[tuple(hi_)])
        return a
    x0, x1, y0, y1 = cells(px, np.minimum), cells(px, np.maximum), cells(py, np.minimum), cells(py, np.maximum)
    boxes = np.stack([px.reshape(ncam, -1).min(1), py.reshape(ncam, -1).min(1), px.reshape(ncam, -1).max(1), py.reshape(ncam, -1).max(1)], 1)
    lo = np.full((ncam, hs), np.iinfo(np.int32).max, np.int64)
    hi = np.full((ncam, hs), -1, np.int64)
    for c in range(ncam):
        for a, b, r0, r1 in zip(x0[c].ravel(), x1[c].ravel(), y0[c].ravel(), y1[c].ravel()):
            r0, r1 = max(r0, 0), min(r1, hs - 1)
            lo[c, r0:r1 + 1] = np.minimum(lo[c, r0:r1 + 1], a)
            hi[c, r0:r1 + 1] = np.maximum(hi[c, r0:r1 + 1], b)
    return boxes.astype(np.int32), lo, hi


def gather_mean(heatmaps_padded, idx):
    """index_select + camera mean (repro_layer.py:97-105). heatmaps_padded [ncam,K,hs,hs] -> [K,G,G,G]."""
    hm = _f32(heatmaps_padded)
    ncam, K, hs, _ = hm.shape
    G = idx.shape[-1]
    out = np.empty((K, G, G, G), np.float32)
    ii = _i32(idx)

    def run(rng):
        return lib().jho_gather_mean(_p(hm), _p(ii), ncam, K, hs, G, _p(out), rng[0], rng[1])
    if any(rc != 0 for rc in _threads(run, K)):
        raise ValueError("jho_gather_mean failed")
    return out


def pad_heatmaps(heatmaps):
    """F.pad(..., [1,1,1,1]) of model.py:65-66 on the last two axes."""
    pad = [(0, 0)] * (heatmaps.ndim - 2) + [(1, 1), (1, 1)]
    return np.pad(_f32(heatmaps), pad)


def repro_layer_forward(heatmaps_padded, center3D, centerHM, cameraMatrices, intrinsicMatrices,
                        distortionCoefficients, G, spacing, lerp_mode=0):
    """ReprojectionLayer.forward for ONE frame set (repro_layer.py:110-119): -> [K,G,G,G] fp32."""
    hs = heatmaps_padded.shape[-1]
    idx = reproject_indices(center3D, centerHM, cameraMatrices, intrinsicMatrices,
                            distortionCoefficients, G, spacing, hs, lerp_mode)
    return gather_mean(heatmaps_padded, idx), idx


def centroid_tail(v, spacing, roi, center3D):
    """model.py:73-87 for one frame set. v [K,h,h,h] -> points [K,3] mm, conf [K], argmax [K] (flat voxel)."""
    v = _f32(v)
    K, h = v.shape[0], v.shape[1]
    pts = np.empty((K, 3), np.float32)
    conf = np.empty((K,), np.float32)
    am = np.empty((K,), np.int32)
    rc = lib().jho_centroid(_p(v), K, h, ctypes.c_float(spacing), ctypes.c_float(roi),
                            _p(_f32(center3D).reshape(3)), _p(pts), _p(conf), _p(am))
    if rc != 0:
        raise ValueError(f"jho_centroid rc={rc}")
    return pts, conf, am


# ----------------------------------------------------------------------------- V2V (torch CPU fp32)
def _t(sd, name):
    import torch
    v = sd[name]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _strip(sd):
    """Accept either bare V2VNet keys or the checkpoint's `v2vNet.`-prefixed keys."""
    if any(k.startswith("v2vNet.") for k in sd):
        return {k[len("v2vNet."):]: v for k, v in sd.items() if k.startswith("v2vNet.")}
    return sd


def v2v_forward(state_dict, x, return_intermediates=False):
    """V2VNet.forward in eval mode (v2vnet.py:98-102). x: [B,K,G,G,G] fp32 -> [B,K,G/2,G/2,G/2]."""
    import torch
    import torch.nn.functional as F
    sd = _strip(state_dict)
    x = x if isinstance(x, torch.Tensor) else torch.from_numpy(_f32(x))
    x = x.float()
    inter = {}

    def conv(p, t, stride=1, pad=0):
        return F.conv3d(t, _t(sd, p + ".weight").float(), _t(sd, p + ".bias").float(), stride=stride, padding=pad)

    def basic(p, t, k, stride):                       # Basic3DBlock, v2vnet.py:12-24
        return F.relu(F.instance_norm(conv(p + ".block.0", t, stride, (k - 1) // 2)))

    def res(p, t):                                    # Res3DBlock, v2vnet.py:27-43
        r = F.relu(F.instance_norm(conv(p + ".res_branch.0", t, 1, 1)))
        r = F.instance_norm(conv(p + ".res_branch.3", r, 1, 1))
        return F.relu(r + t)

    with torch.no_grad():
        x = basic("front_layers.0", x, 3, 2); inter["front0"] = x
        x = res("front_layers.1", x); inter["front1"] = x
        s = res("encoder_decoder.skip_res1", x); inter["skip"] = s              # v2vnet.py:76
        y = basic("encoder_decoder.encoder_pool1", x, 2, 2); inter["pool"] = y  # :77
        y = res("encoder_decoder.mid_res", y); inter["mid"] = y                 # :78
        p = "encoder_decoder.decoder_upsample1.block.0"                         # :79, Upsample3DBlock :46-61
        y = F.conv_transpose3d(y, _t(sd, p + ".weight").float(), _t(sd, p + ".bias").float(), stride=2)
        y = F.relu(F.instance_norm(y)); inter["up"] = y
        y = res("encoder_decoder.decoder_res1", y); inter["dec"] = y            # :80
        y = y + s                                                               # :81
        out = conv("output_layer", y)                                           # v2vnet.py:101
    return (out, inter) if return_intermediates else out


def hybrid3d_forward(state_dict, heatmaps, center3D, centerHM, cameraMatrices, intrinsicMatrices,
                     distortionCoefficients, roi, spacing, lerp_mode=0):
    """model.py:65-88 for ONE frame set, from un-padded heat maps [ncam,K,hm,hm] to key points.
    Returns dict(idx, volume [K,G^3] (before /255), v2v [K,h^3], points [K,3], conf [K], argmax [K])."""
    G = int(roi / spacing)
    hp = pad_heatmaps(heatmaps)
    vol, idx = repro_layer_forward(hp, center3D, centerHM, cameraMatrices, intrinsicMatrices,
                                   distortionCoefficients, G, spacing, lerp_mode)
    v = v2v_forward(state_dict, (vol / np.float32(255.0))[None]).numpy()[0]
    pts, conf, am = centroid_tail(v, spacing, roi, center3D)
    return dict(idx=idx, volume=vol, v2v=v, points=pts, conf=conf, argmax=am)
