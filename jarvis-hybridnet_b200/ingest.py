"""Host side of the data formats either side of the 3D path.

  ingest_frames(frames)        predict3D.py:79  `from_numpy(imgs_orig).cuda().float().permute(0,3,1,2)[:, [2,1,0]] / 255.`
                               on the device, from the decoder's uint8 BGR frames (3 bytes per pixel over PCIe, not 12)
  crop_normalize_u8(...)       jarvis3D.py:168-177 straight from the uint8 frames
  FrameUploader                pinned, double-buffered host staging for the frames of a predict3D-style loop: the upload of
                               frame set i+1 overlaps the kernels of frame set i
  EffTrackHead                 stands in for `effTrack.deconv1` (efficienttrack/model.py:89-95,127): same parameter, same
                               state_dict key, output written in the reprojection gather's layout (jhn_efftrack_head)
  softplus2 / pad_heatmaps     the two returned volumes of HybridNetBackbone.forward (model.py:65-66,73,88)
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib


def ingest_frames(frames):
    """uint8 [..., H, W, 3] BGR (device) -> fp32 [..., 3, H, W] RGB in [0, 1]; bit-identical to predict3D.py:79."""
    lib = _lib.load()
    with _lib.require_cuda(frames):
        if frames.dtype != torch.uint8 or frames.shape[-1] != 3:
            raise RuntimeError(f"ingest_frames expects uint8 [...,H,W,3] frames, got {frames.dtype} {tuple(frames.shape)}")
        lead, (H, W) = frames.shape[:-3], frames.shape[-3:-1]
        f = frames.contiguous()
        N = f.numel() // (H * W * 3)
        out = torch.empty(tuple(lead) + (3, H, W), dtype=torch.float32, device=f.device)
        _lib.check(lib.jhn_ingest_frames(_lib.dptr(f), N, H, W, _lib.dptr(out), _lib.stream_ptr()))
    return out


def crop_normalize_u8(frames, centerHM, valid, bbox, mean, std):
    """frames uint8 [B,ncam,H,W,3] BGR (or [ncam,H,W,3]), centerHM [B,ncam,2] i32, valid [B] i32 ->
    fp32 [B,ncam,3,bbox,bbox] = ((u8/255) - mean) / std around centerHM (RGB planes); zeros where valid == 0."""
    lib = _lib.load()
    with _lib.require_cuda(frames, centerHM, valid):
        if frames.dim() == 4:
            frames = frames[None]
        if frames.dtype != torch.uint8 or frames.shape[-1] != 3:
            raise RuntimeError(f"crop_normalize_u8 expects uint8 [B,ncam,H,W,3] frames, got {frames.dtype} {tuple(frames.shape)}")
        B, ncam, H, W, _ = frames.shape
        f = frames.contiguous()
        out = torch.empty((B, ncam, 3, bbox, bbox), dtype=torch.float32, device=f.device)
        m = (ctypes.c_float * 3)(*[float(v) for v in mean])
        s = (ctypes.c_float * 3)(*[float(v) for v in std])
        _lib.check(lib.jhn_crop_normalize_u8(_lib.dptr(f), B, ncam, H, W, int(bbox),
                                             _lib.dptr(centerHM.contiguous().to(torch.int32)),
                                             _lib.dptr(valid.contiguous().to(torch.int32)), m, s, _lib.dptr(out),
                                             _lib.stream_ptr()))
    return out


def softplus2(v2v_out):
    """heatmap_final of model.py:73,88: softplus(softplus(v2v_out)), fp32, one pass."""
    lib = _lib.load()
    with _lib.require_cuda(v2v_out):
        v = v2v_out.contiguous().float()
        out = torch.empty_like(v)
        _lib.check(lib.jhn_softplus2(_lib.dptr(v), v.numel(), _lib.dptr(out), _lib.stream_ptr()))
    return out


def pad_heatmaps(heatmaps):
    """heatmaps_padded of model.py:65-66: F.pad(heatmaps, [1,1,1,1]) for fp32 [..., S, S]."""
    lib = _lib.load()
    with _lib.require_cuda(heatmaps):
        S = heatmaps.shape[-1]
        if heatmaps.shape[-2] != S:
            raise RuntimeError(f"pad_heatmaps expects square maps, got {tuple(heatmaps.shape)}")
        h = heatmaps.contiguous().float()
        out = torch.empty(tuple(h.shape[:-2]) + (S + 2, S + 2), dtype=torch.float32, device=h.device)
        _lib.check(lib.jhn_pad_heatmaps(_lib.dptr(h), h.numel() // (S * S), S, _lib.dptr(out), _lib.stream_ptr()))
    return out


class EffTrackHead(nn.Module):
    """Drop-in for `EfficientTrackBackbone.deconv1` (ConvTranspose2d(C, K, 4, stride 2, padding 1, bias=False)).

    `weight` has the reference layer's shape [C, K, 4, 4] and state_dict key, so a checkpoint loads with strict=True.
    out_format "planar" returns the reference's fp32 [N, K, 2Hq, 2Wq]; "f16_cl" / "bf16_cl" return the 16-bit channels-last
    tensor [N, 2Hq+2, 2Wq+2, 24] with F.pad's zero border that ReprojectionLayer / HybridNet3D read without a staging pass."""

    FORMATS = {"planar": _lib.HM_F32_PLANAR, "f16_cl": _lib.HM_F16_CL, "bf16_cl": _lib.HM_BF16_CL}

    def __init__(self, in_channels, out_channels, out_format="f16_cl"):
        super().__init__()
        if out_format not in self.FORMATS:
            raise ValueError(f"out_format must be one of {sorted(self.FORMATS)}")
        self.in_channels, self.out_channels, self.out_format = in_channels, out_channels, out_format
        self.weight = nn.Parameter(torch.zeros(in_channels, out_channels, 4, 4), requires_grad=False)

    @classmethod
    def from_deconv(cls, deconv, out_format="f16_cl"):
        if (tuple(deconv.kernel_size), tuple(deconv.stride), tuple(deconv.padding)) != ((4, 4), (2, 2), (1, 1)) or deconv.bias is not None:
            raise RuntimeError("EffTrackHead replaces ConvTranspose2d(kernel 4, stride 2, padding 1, bias=False) only")
        m = cls(deconv.in_channels, deconv.out_channels, out_format)
        m.load_state_dict(deconv.state_dict(), strict=True)
        return m.to(deconv.weight.device)

    def forward(self, features):
        lib = _lib.load()
        with _lib.require_cuda(features, self.weight):
            x = features.contiguous().float()
            N, C, Hq, Wq = x.shape
            K = self.out_channels
            if C != self.in_channels:
                raise RuntimeError(f"feature map has {C} channels, the head expects {self.in_channels}")
            fmt = self.FORMATS[self.out_format]
            if fmt == _lib.HM_F32_PLANAR:
                out = torch.empty((N, K, 2 * Hq, 2 * Wq), dtype=torch.float32, device=x.device)
            else:
                out = torch.empty((N, 2 * Hq + 2, 2 * Wq + 2, _lib.HM_CL_PITCH),
                                  dtype=torch.float16 if fmt == _lib.HM_F16_CL else torch.bfloat16, device=x.device)
            _lib.check(lib.jhn_efftrack_head(_lib.dptr(x), _lib.dptr(self.weight.contiguous().float()), N, C, K, Hq, Wq, fmt,
                                             _lib.dptr(out), _lib.stream_ptr()))
        return out


class FrameUploader:
    """Pinned double buffer between a decoder and the device (predict3D.py:72-80 reads into ONE numpy array and blocks on
    `.cuda()` each frame).  `host(i)` is the numpy view the decoder threads fill (cv2 `cap.read()` into a slice, as
    predict3D.read_images does); `upload(i)` enqueues the H2D copy on a copy stream and returns (device uint8 tensor,
    event); the caller's stream waits on the event before the first kernel that reads the frames.  While frame set i is
    being processed the decoder fills buffer 1 - i."""

    def __init__(self, shape, device=None, slots=2):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._pinned = [torch.empty(shape, dtype=torch.uint8, pin_memory=True) for _ in range(slots)]
        self._dev = [torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in range(slots)]
        self._done = [None] * slots                                 # event: kernels that read _dev[i] have been enqueued
        self._stream = torch.cuda.Stream(device=self.device)
        self.slots = slots

    def host(self, i):
        return self._pinned[i % self.slots].numpy()

    def upload(self, i):
        i %= self.slots
        if self._done[i] is not None:
            self._stream.wait_event(self._done[i])                  # do not overwrite frames a pending kernel still reads
        with torch.cuda.stream(self._stream):
            self._dev[i].copy_(self._pinned[i], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        return self._dev[i], ev

    def release(self, i):
        """Call after enqueueing the last kernel that reads slot i's device frames."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._done[i % self.slots] = ev

    @property
    def bytes_per_upload(self):
        return self._pinned[0].numel()
