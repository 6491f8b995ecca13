"""ReprojectionLayer — drop-in for jarvis.hybridnet.repro_layer.ReprojectionLayer (repro_layer.py:11-119),
backed by the sm_100a reprojection kernels through `jhn_reproject_gather`.

Same constructor `(cfg, num_cameras=None)`, same attributes (`grid`, `grid_size`, `grid_spacing`,
`boxsize`, `heatmap_size`, `num_cameras`) and the same `forward(heatmaps, center, centerHM,
cameraMatrices, intrinsicMatrices, distortionCoefficients) -> [1,K,G,G,G]` contract, including the
reference's behaviour of reading only batch element 0 (repro_layer.py:112-117)."""
import torch
import torch.nn as nn

from . import _lib


class ReprojectionLayer(nn.Module):
    def __init__(self, cfg, num_cameras=None, precision="fp32", lerp_mode=_lib.LERP_FMA_FIRST):
        super().__init__()
        self.cfg = cfg
        self.grid_spacing = cfg.HYBRIDNET.GRID_SPACING
        self.boxsize = cfg.HYBRIDNET.ROI_CUBE_SIZE
        self.grid_size = int(cfg.HYBRIDNET.ROI_CUBE_SIZE / cfg.HYBRIDNET.GRID_SPACING)      # :18-19
        self.num_cameras = num_cameras if num_cameras else cfg.HYBRIDNET.NUM_CAMERAS         # :21-24
        self.heatmap_size = int(cfg.KEYPOINTDETECT.BOUNDING_BOX_SIZE / 2 + 2)                # :37
        self.precision = {"fp32": _lib.FP32, "bf16": _lib.BF16}[precision]
        self.lerp_mode = lerp_mode
        self._grid = None
        self._ws = None

    @property
    def grid(self):
        """Static half-resolution grid [h,h,h,3] in mm (repro_layer.py:26-36), closed form; the kernels
        recompute it in registers and never read this tensor."""
        if self._grid is None:
            h = int(self.grid_size / 2)
            half = int(self.grid_size / 2 / 2)
            a = torch.arange(h, dtype=torch.float32) - half
            g = torch.stack(torch.meshgrid(a, a, a, indexing="ij"), dim=3)
            dev = "cuda" if torch.cuda.is_available() else "cpu"
            self._grid = (g * self.grid_spacing * 2).to(dev)
        return self._grid

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._ws

    def _run(self, heatmaps, center, centerHM, cam, intr, dist, post_divide=1.0, want_index=False, volume_layout="ncdhw"):
        """heatmaps [B,ncam,K,S,S] with S == heatmap_size (padded) or heatmap_size-2 (un-padded), or the channels-last
        16-bit form [B,ncam,hs,hs,24] (torch.float16 scaled by 1/16, or torch.bfloat16; bf16 precision only).
        volume_layout "v2v" (bf16 precision): the volume stays in the tensor-core convolution's input layout and is
        returned as an opaque uint8 buffer."""
        lib = _lib.load()
        with _lib.require_cuda(heatmaps, center, centerHM, cam, intr, dist):
            return self._run_on_device(lib, heatmaps, center, centerHM, cam, intr, dist, post_divide, want_index, volume_layout)

    def _run_on_device(self, lib, heatmaps, center, centerHM, cam, intr, dist, post_divide, want_index, volume_layout):
        hs, G = self.heatmap_size, self.grid_size
        if heatmaps.dtype in (torch.float16, torch.bfloat16):
            B, ncam, S, S2, P = heatmaps.shape
            if (S, S2, P) != (hs, hs, _lib.HM_CL_PITCH):
                raise RuntimeError(f"channels-last heat maps {tuple(heatmaps.shape)} do not match [B,ncam,{hs},{hs},24]")
            fmt = _lib.HM_F16_CL if heatmaps.dtype == torch.float16 else _lib.HM_BF16_CL
            hm, K = heatmaps.contiguous(), self.cfg.KEYPOINTDETECT.NUM_JOINTS
        else:
            B, ncam, K, S, S2 = heatmaps.shape
            if S != S2 or S not in (hs, hs - 2):
                raise RuntimeError(f"heat maps are {S}x{S2}; expected {hs} (padded) or {hs - 2} per side")
            fmt, hm = _lib.HM_F32_PLANAR, heatmaps.contiguous().float()
        cam = cam.contiguous().float(); intr = intr.contiguous().float(); dist = dist.contiguous().float()
        # `self.grid + center[0]` (repro_layer.py:113) is an fp32 add: an int centre (predictor, jarvis3D.py:183) is
        # promoted exactly, a float centre (validation path, hybridnet.py:284-304) is taken as it is — never truncated
        c3 = center.contiguous().float(); chm = centerHM.contiguous().to(torch.int32)
        if cam.shape[:2] != (B, ncam) or chm.shape != (B, ncam, 2) or c3.shape != (B, 3):
            raise RuntimeError("calibration / centre tensors do not match heat maps [B,ncam,...]")
        need = _lib.c_size_t()
        _lib.check(lib.jhn_reproject_workspace_bytes(B, ncam, K, hs, G, self.precision, need))
        ws = self._workspace(need.value, hm.device)
        if volume_layout == "v2v":
            if self.precision != _lib.BF16:
                raise RuntimeError("the V2V volume layout belongs to the bf16 path")
            h2, cj = G // 2 + 2, (K + 15) // 16 * 2
            nbytes = B * 8 * cj * h2 * h2 * h2 * 16
            if getattr(self, "_vol", None) is None or self._vol.numel() != nbytes or self._vol.device != hm.device:
                self._vol = torch.zeros(nbytes, dtype=torch.uint8, device=hm.device)      # borders are (re)written by the call
            vol, layout = self._vol, _lib.VOL_V2V_BF16
        else:
            vol, layout = torch.empty((B, K, G, G, G), dtype=torch.float32, device=hm.device), _lib.VOL_NCDHW_F32
        idx = torch.empty((B, ncam, G, G, G), dtype=torch.int32, device=hm.device) if want_index else None
        _lib.check(lib.jhn_reproject_gather(_lib.dptr(hm), fmt, int(S == hs), _lib.dptr(cam), _lib.dptr(intr),
                                            _lib.dptr(dist), _lib.dptr(c3), _lib.dptr(chm), B, ncam, K, hs, G,
                                            float(self.grid_spacing), self.lerp_mode, float(post_divide),
                                            self.precision, layout, _lib.dptr(vol), _lib.dptr(idx),
                                            _lib.dptr(ws), ws.numel(), _lib.stream_ptr()))
        return vol, idx

    def reprojectPoints(self, x, cameraMatrices, intrinsicMatrices, distortionCoefficients, centerHM):
        """int64 [ncam,G,G,G] flat padded-pixel indices, as repro_layer.py:40-85.  `x` is `self.grid + center`
        like the reference passes; the integer centre is read back from its zero voxel."""
        half = int(self.grid_size / 2 / 2)
        center = x[half, half, half].float()[None]             # grid[half,half,half] == 0, so this is the centre itself
        ncam = cameraMatrices.shape[0]
        dummy = torch.zeros((1, ncam, 1, self.heatmap_size, self.heatmap_size), device=x.device)
        _, idx = self._run(dummy, center, centerHM[None], cameraMatrices[None], intrinsicMatrices[None],
                           distortionCoefficients[None], want_index=True)
        return idx[0].long()

    def forward(self, heatmaps, center, centerHM, cameraMatrices, intrinsicMatrices, distortionCoefficients):
        vol, _ = self._run(heatmaps[:1], center[:1], centerHM[:1], cameraMatrices[:1], intrinsicMatrices[:1],
                           distortionCoefficients[:1])
        return vol

    def forward_batched(self, heatmaps, center, centerHM, cameraMatrices, intrinsicMatrices,
                        distortionCoefficients, post_divide=1.0, want_index=False, volume_layout="ncdhw"):
        """All B frame sets in one launch sequence -> ([B,K,G,G,G] fp32, optional int32 indices)."""
        return self._run(heatmaps, center, centerHM, cameraMatrices, intrinsicMatrices, distortionCoefficients,
                         post_divide, want_index, volume_layout)
