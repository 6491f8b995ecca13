"""Host side of the 3D hot path of HybridNetBackbone.forward (jarvis/hybridnet/model.py:53-90).

  accelerate(backbone)  swaps `reproLayer` / `v2vNet` of a loaded reference HybridNetBackbone for the
                        B200 modules and rebinds `forward`, the same seam the reference's own TensorRT
                        loader uses (jarvis/prediction/jarvis3D.py:53-69).  HybridNet, JarvisPredictor3D,
                        predict3D, the yacs config and the .pth files are untouched.
  HybridNet3D           the three stages behind one call (`jhn_hybrid3d_forward`) for B independent frame
                        sets: heat maps in, key points out.  This is what bench.py times.
  shard_range/gather_results   one replica per GPU over disjoint frame-set ranges; a single gather of
                        [N,K,4] at the end, no collective on the hot path.
"""
import ctypes
import types

import torch
import torch.nn as nn

from . import _lib
from .repro_layer import ReprojectionLayer
from .synth import V2V_LAYERS
from .v2vnet import V2VNet


def centroid_tail(v2v_out, spacing, roi, center3D, want_argmax=False):
    """model.py:73-87 on [B,K,h,h,h] fp32 -> points3D [B,K,3] (mm), confidences [B,K] (, argmax [B,K])."""
    lib = _lib.load()
    with _lib.require_cuda(v2v_out, center3D):
        B, K, h = v2v_out.shape[0], v2v_out.shape[1], v2v_out.shape[2]
        v = v2v_out.contiguous().float()
        c3 = center3D.contiguous().float()               # `+ center3D` of model.py:86-87 promotes an int centre to fp32
        pts = torch.empty((B, K, 3), dtype=torch.float32, device=v.device)
        conf = torch.empty((B, K), dtype=torch.float32, device=v.device)
        am = torch.empty((B, K), dtype=torch.int32, device=v.device)
        _lib.check(lib.jhn_centroid_reduce(_lib.dptr(v), B, K, h, float(spacing), float(roi), _lib.dptr(c3),
                                           _lib.dptr(pts), _lib.dptr(conf), _lib.dptr(am), _lib.stream_ptr()))
    return (pts, conf, am) if want_argmax else (pts, conf)


def _accelerated_forward(self, imgs, img_size, centerHM, center3D, cameraMatrices, intrinsicMatrices,
                         distortionCoefficients):
    """Replacement for HybridNetBackbone.forward with the same signature and return tuple (model.py:53-90).
    effTrack is still the reference module (out of scope) — with `accelerate(head_format=...)` its last layer writes the
    gather's channels-last layout (row f2) — and everything after it runs on the B200 kernels.  Like the reference
    (`center[0]`, `centerHM[0]`, repro_layer.py:109-119) one call serves ONE frame set; `forward_batched` serves B.
    `heatmap_final` / `heatmaps_padded` are only materialised when `self.return_volumes` is True: every inference caller
    discards them (jarvis3D.py:180)."""
    from .ingest import pad_heatmaps, softplus2
    batch_size = imgs.shape[0]
    self.heatmap_size = (img_size / 2).int()
    hm = self.effTrack(imgs.reshape(-1, imgs.shape[2], imgs.shape[3], imgs.shape[4]))[1]
    hm = hm.reshape(batch_size, -1, hm.shape[1], hm.shape[2], hm.shape[3])
    # F.pad is folded into the gather (bounds instead of a padded copy); /255 is folded as post_divide
    vol, _ = self.reproLayer.forward_batched(hm[:1], center3D[:1], centerHM[:1], cameraMatrices[:1],
                                             intrinsicMatrices[:1], distortionCoefficients[:1], post_divide=255.0)
    v = self.v2vNet(vol)
    points3D, confidences = centroid_tail(v, float(self.grid_spacing), float(self.grid_size), center3D[:1])
    heatmap_final = heatmaps_padded = None
    if getattr(self, "return_volumes", False):
        if hm.dtype != torch.float32:
            raise RuntimeError("return_volumes needs the planar fp32 heat maps: use accelerate(..., head_format=None)")
        heatmap_final = softplus2(v)                                               # model.py:73,88 (jhn_softplus2)
        heatmaps_padded = pad_heatmaps(hm)                                         # model.py:65-66 (jhn_pad_heatmaps)
    return heatmap_final, heatmaps_padded, points3D, confidences


def _forward_batched(self, imgs, img_size, centerHM, center3D, cameraMatrices, intrinsicMatrices, distortionCoefficients):
    """B frame sets per call — what the reference's forward leaves as a TODO (model.py:75): crops [B,ncam,3,bb,bb],
    centerHM [B,ncam,2], center3D [B,3], calibration [B,ncam,...] or un-batched [ncam,...] (shared rig) ->
    points3D [B,K,3], confidences [B,K].  The 3D stages run through ONE jhn_hybrid3d_forward call."""
    B = imgs.shape[0]
    hm = self.effTrack(imgs.reshape(-1, imgs.shape[2], imgs.shape[3], imgs.shape[4]))[1]
    hm = hm.reshape(B, -1, hm.shape[1], hm.shape[2], hm.shape[3])
    ex = lambda t, nd: t if t.dim() == nd else t[None].expand(B, *t.shape)
    pts, conf, _ = self._jhn3d(hm, center3D, centerHM, ex(cameraMatrices, 4), ex(intrinsicMatrices, 4),
                               ex(distortionCoefficients, 4))
    return pts, conf


def accelerate(backbone, precision="fp32", lerp_mode=_lib.LERP_FMA_FIRST, return_volumes=False, head_format=None):
    """Swap the 3D stages of a reference HybridNetBackbone (already built and `load_state_dict`-ed by
    jarvis.hybridnet.hybridnet.HybridNet, hybridnet.py:77-90) for the B200 implementation, in place.
    head_format "f16_cl" / "bf16_cl" (bf16 precision) also swaps `effTrack.deconv1` for ingest.EffTrackHead so that the
    2D network emits the gather's layout directly; None keeps the reference layer."""
    cfg = backbone.cfg
    K = cfg.KEYPOINTDETECT.NUM_JOINTS
    new_v2v = V2VNet(K, K, precision=precision)
    new_v2v.load_state_dict(backbone.v2vNet.state_dict(), strict=True)
    dev = next(backbone.v2vNet.parameters()).device
    backbone.v2vNet = new_v2v.to(dev)
    backbone.reproLayer = ReprojectionLayer(cfg, getattr(backbone.reproLayer, "num_cameras", None),
                                            precision=precision, lerp_mode=lerp_mode)
    if head_format is not None:
        if precision != "bf16":
            raise RuntimeError("channels-last 16-bit heat maps belong to the bf16 path")
        from .ingest import EffTrackHead
        backbone.effTrack.deconv1 = EffTrackHead.from_deconv(backbone.effTrack.deconv1, head_format)
    backbone.return_volumes = return_volumes
    h3d = HybridNet3D(K, cfg.KEYPOINTDETECT.BOUNDING_BOX_SIZE, cfg.HYBRIDNET.ROI_CUBE_SIZE, cfg.HYBRIDNET.GRID_SPACING,
                      precision=precision, lerp_mode=lerp_mode, v2vNet=backbone.v2vNet)
    object.__setattr__(backbone, "_jhn3d", h3d)                  # not a registered sub-module: the state_dict keeps the reference's keys
    backbone.forward = types.MethodType(_accelerated_forward, backbone)
    backbone.forward_batched = types.MethodType(_forward_batched, backbone)
    return backbone


class HybridNet3D(nn.Module):
    """ReprojectionLayer -> V2VNet -> centroid for B independent frame sets through ONE C-ABI call.

    forward(heatmaps [B,ncam,K,S,S] fp32 (S = BB/2 un-padded or BB/2+2 padded), center3D [B,3] i32,
            centerHM [B,ncam,2] i32, cameraMatrices [B,ncam,4,3], intrinsicMatrices [B,ncam,3,3],
            distortionCoefficients [B,ncam,1,5]) -> points3D [B,K,3] mm, confidences [B,K], argmax [B,K]"""

    def __init__(self, K, bbox, roi, spacing, state_dict=None, precision="bf16", lerp_mode=_lib.LERP_FMA_FIRST, v2vNet=None):
        super().__init__()
        self.K, self.roi, self.spacing = K, roi, spacing
        self.G = int(roi / spacing)
        self.hs = int(bbox / 2 + 2)
        self.lerp_mode = lerp_mode
        self.v2vNet = V2VNet(K, K, precision=precision) if v2vNet is None else v2vNet     # v2vNet: share a loaded network
        if state_dict is not None:
            sd = {k[len("v2vNet."):] if k.startswith("v2vNet.") else k: torch.as_tensor(v) for k, v in state_dict.items()}
            self.v2vNet.load_state_dict(sd, strict=True)
        self._ws = None
        self._host = None
        self._graphs = {}

    def forward(self, heatmaps, center3D, centerHM, cameraMatrices, intrinsicMatrices, distortionCoefficients, _ws=None):
        """heatmaps: fp32 planar [B,ncam,K,S,S] (the reference's tensor), or the gather-native channels-last form
        [B,ncam,hs,hs,24] in torch.float16 (scaled by 1/16, `_lib.heatmap_convert`) / torch.bfloat16."""
        lib = _lib.load()
        with _lib.require_cuda(heatmaps, center3D, centerHM, cameraMatrices, intrinsicMatrices, distortionCoefficients):
            if heatmaps.dtype in (torch.float16, torch.bfloat16):
                B, ncam, S, S2, P = heatmaps.shape
                if (S, S2, P) != (self.hs, self.hs, _lib.HM_CL_PITCH):
                    raise RuntimeError(f"channels-last heat maps {tuple(heatmaps.shape)} do not match [B,ncam,{self.hs},{self.hs},24]")
                fmt = _lib.HM_F16_CL if heatmaps.dtype == torch.float16 else _lib.HM_BF16_CL
                hm, K = heatmaps.contiguous(), self.K
            else:
                B, ncam, K, S, _ = heatmaps.shape
                if K != self.K or S not in (self.hs, self.hs - 2):
                    raise RuntimeError(f"heat maps {tuple(heatmaps.shape)} do not match K={self.K}, hs={self.hs}")
                fmt, hm = _lib.HM_F32_PLANAR, heatmaps.contiguous().float()
            net = self.v2vNet._get_handle()
            need = _lib.c_size_t()
            _lib.check(lib.jhn_hybrid3d_workspace_bytes(net, B, ncam, self.hs, self.G, need))
            if _ws is not None:                                  # a captured graph owns its workspace
                ws = _ws
                if ws.numel() < need.value:
                    raise RuntimeError("graph workspace too small")
            else:
                if self._ws is None or self._ws.numel() < need.value or self._ws.device != heatmaps.device:
                    self._ws = torch.empty(need.value, dtype=torch.uint8, device=heatmaps.device)
                ws = self._ws
            dev = heatmaps.device
            pts = torch.empty((B, K, 3), dtype=torch.float32, device=dev)
            conf = torch.empty((B, K), dtype=torch.float32, device=dev)
            am = torch.empty((B, K), dtype=torch.int32, device=dev)
            f = lambda t: t.contiguous().float()
            cam, intr, dist, c3 = f(cameraMatrices), f(intrinsicMatrices), f(distortionCoefficients), f(center3D)
            chm = centerHM.contiguous().to(torch.int32)
            _lib.check(lib.jhn_hybrid3d_forward(net, _lib.dptr(hm), fmt, int(S == self.hs), _lib.dptr(cam), _lib.dptr(intr),
                                                _lib.dptr(dist), _lib.dptr(c3), _lib.dptr(chm), B, ncam, self.hs, self.G,
                                                float(self.spacing), float(self.roi), self.lerp_mode, _lib.dptr(pts),
                                                _lib.dptr(conf), _lib.dptr(am), _lib.dptr(ws), ws.numel(),
                                                _lib.stream_ptr()))
        return pts, conf, am

    def workspace_bytes(self, B, ncam):
        need = _lib.c_size_t()
        _lib.check(_lib.load().jhn_hybrid3d_workspace_bytes(self.v2vNet._get_handle(), B, ncam, self.hs, self.G, need))
        return int(need.value)

    def forward_graph(self, heatmaps, center3D, centerHM, cameraMatrices, intrinsicMatrices, distortionCoefficients):
        """forward() replayed from a CUDA graph: the low-latency path for small B (the reference's predictor runs one
        frame set per call, jarvis3D.py:178-186, where the 27 launches of a forward cost more on the host than on
        the device).  The first call per input signature captures the graph over static input / output buffers and
        a workspace it owns; later calls copy the inputs in and replay.  The library is capture-safe by contract
        (stream-ordered, no allocation, no synchronisation: SURVEY.md §8b).  The returned tensors are the graph's
        static outputs: consume them before the next call with the same signature."""
        ins = (heatmaps, center3D, centerHM, cameraMatrices, intrinsicMatrices, distortionCoefficients)
        _lib.require_cuda(*ins)
        dts = (heatmaps.dtype if heatmaps.dtype in (torch.float16, torch.bfloat16) else torch.float32, torch.float32,
               torch.int32, torch.float32, torch.float32, torch.float32)
        key = (tuple(tuple(t.shape) for t in ins), heatmaps.dtype, heatmaps.device)
        ent = self._graphs.get(key)
        if ent is None:
            B, ncam = heatmaps.shape[0], heatmaps.shape[1]
            static = [torch.empty(t.shape, dtype=dt, device=t.device) for t, dt in zip(ins, dts)]
            ws = torch.empty(self.workspace_bytes(B, ncam), dtype=torch.uint8, device=heatmaps.device)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=heatmaps.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                    # warm-up on the graph's workspace: zero borders, attributes
                for d, t in zip(static, ins):
                    d.copy_(t)
                for _ in range(2):
                    self.forward(*static, _ws=ws)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.forward(*static, _ws=ws)
            ent = self._graphs[key] = (graph, static, out, ws)
        graph, static, out, _ = ent
        for d, t in zip(static, ins):
            d.copy_(t)
        graph.replay()
        return out

    def forward_host(self, host_inputs, chunk=8, roi_upload="dma"):
        """End-to-end call with HOST (pinned) tensors: H2D copies of the six inputs, the fused forward, and a D2H read of
        [B,K,4] (x,y,z,confidence) into pinned memory.  Returns (result, h2d_bytes, d2h_bytes), the byte counts being what
        actually crossed the link.  == forward_host_async(...).result(); see there."""
        return self.forward_host_async(host_inputs, chunk, roi_upload).result()

    def forward_host_async(self, host_inputs, chunk=8, roi_upload="dma", slots=2):
        """Enqueue one end-to-end step and return a handle whose .result() waits for it: (pinned [B,K,4], h2d_bytes, d2h_bytes).

        The batch is cut into chunks of `chunk` frame sets: a copy stream uploads chunk i+1 while the compute stream runs
        chunk i.  `slots` sets of device buffers alternate between calls, so a caller that submits step n+1 before it collects
        step n keeps the link busy while step n computes (each step still uploads its own inputs and downloads its own
        result; a step costs max(PCIe time, kernel time) in steady state).

        Channels-last 16-bit heat maps: calibration and centres go first (a few KB), jhn_heatmap_boxes projects the voxel
        grids and returns each camera's pixel box, and only those boxes cross PCIe — the gather never reads a pixel outside
        them.  roi_upload "dma" (default): strided copy-engine transfers, all boxes of a chunk in one
        cudaMemcpy3DBatchAsync (jhn_upload_heatmap_boxes; ~36 GB/s on 5 KB rows, against 55 GB/s for a contiguous copy of
        1.8x the bytes); "pull": a small kernel reads the boxes out of the mapped host tensor (jhn_pull_heatmap_boxes,
        45 GB/s; every compute kernel runs ~20 % slower while it is active); "hybrid:f": both at once — the copy engine takes
        the first 1 - f of every chunk's images, the pull kernel the rest (f = 0.6 with chunk = B is the fastest measured:
        4.28 ms per 32 frame sets against 5.1 - 5.4 for "dma", profiles/r02_e2e_hybrid_upload.txt); None: whole tensors."""
        dev = torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        B = host_inputs[0].shape[0]
        chunk = max(1, min(chunk, B))
        key = tuple((tuple(t.shape), t.dtype) for t in host_inputs) + (int(slots),)
        if self._host is None or self._host["key"] != key:
            ncam = host_inputs[0].shape[1]
            n_slots, slots = max(2, int(slots)), []
            for _ in range(n_slots):
                dbuf = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host_inputs]
                if host_inputs[0].dtype in (torch.float16, torch.bfloat16):
                    dbuf[0].zero_()                          # pixels outside the boxes are never read; keep them finite anyway
                slots.append(dict(dbuf=dbuf, res=torch.empty((B, self.K, 4), dtype=torch.float32, device=dev),
                                  hres=torch.empty((B, self.K, 4), dtype=torch.float32, pin_memory=True),
                                  boxes=torch.empty((B, ncam, 4), dtype=torch.int32, device=dev),
                                  hboxes=torch.empty((B, ncam, 4), dtype=torch.int32, pin_memory=True),
                                  pulled=torch.zeros(1, dtype=torch.int64, device=dev), spans=None, cab=None,
                                  hpulled=torch.zeros(1, dtype=torch.int64).pin_memory(), busy=None, done=None))
            self._host = dict(key=key, slots=slots, n=0, copy=torch.cuda.Stream(device=dev, priority=-1),
                              aux=torch.cuda.Stream(device=dev, priority=-1))
        H = self._host
        S = H["slots"][H["n"] % len(H["slots"])]
        H["n"] += 1
        if S["busy"] is not None:
            S["busy"].result()                           # the step that last used this slot must have been collected
        dbuf, res, hres, boxes, hboxes, pulled, hpulled = (S[k] for k in ("dbuf", "res", "hres", "boxes", "hboxes", "pulled", "hpulled"))
        copy_stream, aux = H["copy"], H["aux"]
        main = torch.cuda.current_stream()
        hm_h = host_inputs[0]
        roi = roi_upload if (roi_upload and hm_h.dtype in (torch.float16, torch.bfloat16) and hm_h.is_pinned()) else None
        pull_frac = 0.0
        use_spans = False
        if isinstance(roi, str) and roi.startswith("hybrid"):          # "hybrid:0.25" = a quarter of every chunk's images by the pull kernel
            pull_frac = float(roi.split(":")[1]) if ":" in roi else 0.6
            # "hybrid-spans:f": the pull kernel moves each pixel row's column span (jhn_heatmap_spans) instead of the whole box:
            # 22 % fewer bytes, but computing the spans costs the GPU more than the link gains (4.6 vs 4.16 ms per step, run 54)
            use_spans = roi.startswith("hybrid-spans")
            roi = "hybrid"
        if roi not in (None, "pull", "dma", "hybrid"):
            raise ValueError("roi_upload must be 'pull', 'dma', 'hybrid[:fraction]', 'hybrid-spans[:fraction]' or None")
        if S["done"] is not None:
            copy_stream.wait_event(S["done"])            # the kernels of the slot's previous step are done with its device buffers
            aux.wait_event(S["done"])
        events = []
        h2d = 0
        done_copy = None
        if roi:
            ncam, hs = hm_h.shape[1], hm_h.shape[2]
            pix = hm_h.shape[4] * hm_h.element_size()
            img = hs * hs * pix
            with torch.cuda.stream(aux):                 # small tensors, boxes: off the copy stream so that it never drains
                small_h = [h.contiguous() for h in host_inputs[1:]]
                if all(h.is_pinned() and (h.numel() * h.element_size()) % 4 == 0 for h in small_h) and len(small_h) <= 8:
                    # one kernel reads them out of mapped host memory: cudaMemcpyAsync copies would queue on the copy
                    # engine behind the previous step's heat maps and hold this step's boxes (and transfer) back
                    n = len(small_h)
                    arr = lambda vals, ty: (ty * n)(*vals)
                    _lib.check(lib.jhn_pull_small(n, arr([h.data_ptr() for h in small_h], ctypes.c_void_p),
                                                  arr([d.data_ptr() for d in dbuf[1:]], ctypes.c_void_p),
                                                  arr([h.numel() * h.element_size() for h in small_h], ctypes.c_size_t),
                                                  ctypes.c_void_p(aux.cuda_stream)))
                    S["keep"] = small_h                  # the kernel reads these after the call returns
                else:
                    for d, h in zip(dbuf[1:], host_inputs[1:]):
                        d.copy_(h, non_blocking=True)
                h2d += sum(h.numel() * h.element_size() for h in small_h)
                f = lambda t: t.contiguous().float()
                c3, chm = f(dbuf[1]), dbuf[2].contiguous().to(torch.int32)
                cam, intr, dist = f(dbuf[3]), f(dbuf[4]), f(dbuf[5])
                if roi == "hybrid" and use_spans:        # boxes for the copy engine, per-row column spans for the pull kernel
                    if S["spans"] is None:
                        S["spans"] = torch.empty((B, ncam, hs, 2), dtype=torch.int32, device=dev)
                        S["cab"] = torch.empty(B * ncam * (self.G // 2) ** 3 * 2, dtype=torch.float32, device=dev)
                    _lib.check(lib.jhn_heatmap_spans(_lib.dptr(cam), _lib.dptr(intr), _lib.dptr(dist), _lib.dptr(c3), _lib.dptr(chm),
                                                     B, ncam, hs, self.G, float(self.spacing), _lib.dptr(S["cab"]), S["cab"].numel() * 4,
                                                     _lib.dptr(boxes), _lib.dptr(S["spans"]), ctypes.c_void_p(aux.cuda_stream)))
                else:
                    _lib.check(lib.jhn_heatmap_boxes(_lib.dptr(cam), _lib.dptr(intr), _lib.dptr(dist), _lib.dptr(c3), _lib.dptr(chm),
                                                     B, ncam, hs, self.G, float(self.spacing), _lib.dptr(boxes),
                                                     ctypes.c_void_p(aux.cuda_stream)))
                hboxes.copy_(boxes, non_blocking=True)
                if roi in ("pull", "hybrid"):
                    pulled.zero_()
                small = torch.cuda.Event()
                small.record(aux)
            if roi == "dma" or (roi == "hybrid" and pull_frac < 1.0):
                small.synchronize()                      # 1.5 KB back: the host needs the boxes to describe the strided copies
            copy_stream.wait_event(small)
            main.wait_event(small)
            ps = copy_stream                             # stream of the pull kernel
            if roi == "hybrid":
                # The copy engine is bound by rows per second (~36 GB/s on these 5 KB rows), the link is not (55 GB/s): the pull
                # kernel moves the LAST pull_frac of every chunk's images over the link while the copy engine walks the rest.
                if "pull" not in H:
                    H["pull"] = torch.cuda.Stream(device=dev)
                ps = H["pull"]
                ps.wait_event(small)
                if S["done"] is not None:
                    ps.wait_event(S["done"])
            sp, spull = ctypes.c_void_p(copy_stream.cuda_stream), ctypes.c_void_p(ps.cuda_stream)
            copied = _lib.c_size_t()
            for lo in range(0, B, chunk):
                n_img = min(chunk, B - lo) * ncam
                n_pull = {"dma": 0, "pull": n_img}.get(roi, min(n_img, max(0, int(round(pull_frac * n_img)))))
                n_dma, i0 = n_img - n_pull, lo * ncam
                evs = []
                if n_dma:
                    _lib.check(lib.jhn_upload_heatmap_boxes(ctypes.c_void_p(hm_h.data_ptr() + i0 * img), ctypes.c_void_p(dbuf[0].data_ptr() + i0 * img),
                                                            ctypes.c_void_p(hboxes.data_ptr() + i0 * 16), n_dma, hs, pix, sp, ctypes.byref(copied)))
                    h2d += copied.value
                    evs.append(torch.cuda.Event())
                    evs[-1].record(copy_stream)
                if n_pull:
                    j0 = i0 + n_dma
                    srcp, dstp = ctypes.c_void_p(hm_h.data_ptr() + j0 * img), ctypes.c_void_p(dbuf[0].data_ptr() + j0 * img)
                    if roi == "hybrid" and use_spans:
                        _lib.check(lib.jhn_pull_heatmap_spans(srcp, dstp, ctypes.c_void_p(S["spans"].data_ptr() + j0 * hs * 8), n_pull, hs, pix,
                                                              _lib.dptr(pulled), spull))
                    else:
                        _lib.check(lib.jhn_pull_heatmap_boxes(srcp, dstp, ctypes.c_void_p(boxes.data_ptr() + j0 * 16), n_pull, hs, pix,
                                                              _lib.dptr(pulled), spull))
                    evs.append(torch.cuda.Event())
                    evs[-1].record(ps)
                events.append(evs)
            if roi in ("pull", "hybrid"):                # bytes the kernel moved: counted on the device, 8 bytes back
                with torch.cuda.stream(ps):
                    hpulled.copy_(pulled, non_blocking=True)
                done_copy = torch.cuda.Event()
                done_copy.record(ps)
        else:
            with torch.cuda.stream(copy_stream):
                for lo in range(0, B, chunk):
                    for d, h in zip(dbuf, host_inputs):
                        d[lo:lo + chunk].copy_(h[lo:lo + chunk], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    events.append([ev])
            h2d = sum(t.numel() * t.element_size() for t in host_inputs)
        overlap = roi in ("pull", "hybrid")              # these forwards run next to pull kernels: see jhn_set_transfer_overlap
        if overlap:
            lib.jhn_set_transfer_overlap(1)
        try:
            for i, lo in enumerate(range(0, B, chunk)):
                for ev in events[i]:
                    main.wait_event(ev)
                pts, conf, _ = self.forward(*[d[lo:lo + chunk] for d in dbuf])
                res[lo:lo + chunk, :, :3] = pts
                res[lo:lo + chunk, :, 3] = conf
        finally:
            if overlap:
                lib.jhn_set_transfer_overlap(0)
        hres.copy_(res, non_blocking=True)
        S["done"] = torch.cuda.Event()
        S["done"].record(main)
        d2h = res.numel() * 4 + (boxes.numel() * 4 if roi else 0) + (8 if roi in ("pull", "hybrid") else 0)
        S["busy"] = _HostStep(S, h2d, d2h, done_copy, hpulled if roi in ("pull", "hybrid") else None)
        return S["busy"]


class _HostStep:
    """Handle of one forward_host_async step."""

    def __init__(self, slot, h2d, d2h, done_copy, hpulled):
        self._slot, self._h2d, self._d2h, self._done_copy, self._hpulled, self._out = slot, h2d, d2h, done_copy, hpulled, None

    def result(self):
        if self._out is None:
            self._slot["done"].synchronize()
            h2d = self._h2d
            if self._done_copy is not None:
                self._done_copy.synchronize()
                h2d += int(self._hpulled.item())
            self._out = (self._slot["hres"], h2d, self._d2h)
            if self._slot["busy"] is self:
                self._slot["busy"] = None
        return self._out


# ---------------------------------------------------------------------------------- multi-GPU sharding
def shard_range(n_items, rank, world_size):
    """Contiguous block of frame-set indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_results(local, n_items, group=None):
    """All-gather per-rank [n_local,K,4] results into [n_items,K,4] in frame-set order.  The only
    communication of a sharded run; uses the default process group (NCCL on GPUs, gloo in CPU tests)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    max_n = max(e - s for s, e in sizes)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: e - s] for p, (s, e) in zip(parts, sizes)], dim=0)
