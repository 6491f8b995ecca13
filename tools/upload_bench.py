"""Microbenchmark of jhn_upload_heatmap_boxes (host -> device copy of only the gather's pixel boxes): host time of the call
and device time of the copies for the three JHN_UPLOAD_MODE variants, against one contiguous copy of the whole tensor.
Run each mode in its own process (the mode is read once):  JHN_UPLOAD_MODE=0 python tools/upload_bench.py"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jarvis_hybridnet_b200.synth as S
from jarvis_hybridnet_b200 import _lib

lib = _lib.load()
sh, B = S.EXAMPLE, 32
cam, intr, dist = S.make_rig(sh.ncam, 0)
sets = [S.make_frameset(sh, cam, intr, dist, s) for s in range(4)]
hs, G = sh.bbox // 2 + 2, int(sh.roi / sh.spacing)
rep = lambda a: np.ascontiguousarray(np.broadcast_to(a[None], (B,) + a.shape))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
c3 = t(np.stack([sets[i % 4][1] for i in range(B)]).astype(np.float32))
chm = t(np.stack([sets[i % 4][2] for i in range(B)]).astype(np.int32))
boxes = torch.empty((B, sh.ncam, 4), dtype=torch.int32, device="cuda")
st = torch.cuda.Stream()
sp = ctypes.c_void_p(st.cuda_stream)
with torch.cuda.stream(st):
    _lib.check(lib.jhn_heatmap_boxes(_lib.dptr(t(rep(cam))), _lib.dptr(t(rep(intr))), _lib.dptr(t(rep(dist))), _lib.dptr(c3), _lib.dptr(chm),
                                     B, sh.ncam, hs, G, float(sh.spacing), _lib.dptr(boxes), sp))
st.synchronize()
hb = boxes.cpu().pin_memory()
host = torch.empty((B, sh.ncam, hs, hs, 24), dtype=torch.float16).pin_memory()
host.view(torch.int16).random_(0, 1000)
dev = torch.zeros_like(host, device="cuda")
copied = ctypes.c_size_t()
res = {}
pulled = torch.zeros(1, dtype=torch.int64, device="cuda")
for name in ("boxes", "pull", "whole"):
    cpu, gpu = [], []
    for it in range(12):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            t0 = time.perf_counter()
            if name == "boxes":
                _lib.check(lib.jhn_upload_heatmap_boxes(ctypes.c_void_p(host.data_ptr()), ctypes.c_void_p(dev.data_ptr()),
                                                        ctypes.c_void_p(hb.data_ptr()), B * sh.ncam, hs, 48, sp, ctypes.byref(copied)))
            elif name == "pull":
                pulled.zero_()
                _lib.check(lib.jhn_pull_heatmap_boxes(ctypes.c_void_p(host.data_ptr()), ctypes.c_void_p(dev.data_ptr()),
                                                      _lib.dptr(boxes), B * sh.ncam, hs, 48, _lib.dptr(pulled), sp))
            else:
                dev.copy_(host, non_blocking=True)
            t1 = time.perf_counter()
            e1.record(st)
        st.synchronize()
        if it >= 2:
            cpu.append(1e3 * (t1 - t0)); gpu.append(e0.elapsed_time(e1))
    nbytes = copied.value if name == "boxes" else (int(pulled.item()) if name == "pull" else host.numel() * 2)
    res[name] = dict(host_ms=float(np.median(cpu)), device_ms=float(np.median(gpu)), bytes=int(nbytes), GB_per_s=nbytes / np.median(gpu) / 1e6)
if copied.value:
    bx = hb.numpy().reshape(-1, 4)
    ok = True
    for i in (0, 100, 383):
        x0, y0, x1, y1 = bx[i, 0], bx[i, 1], -bx[i, 2], -bx[i, 3]
        a = host.view(B * sh.ncam, hs, hs, 24)[i, y0:y1 + 1, x0:x1 + 1]
        b = dev.view(B * sh.ncam, hs, hs, 24)[i, y0:y1 + 1, x0:x1 + 1].cpu()
        ok = ok and bool(torch.equal(a, b))
    res["verified"] = ok
print(json.dumps(dict(mode=os.environ.get("JHN_UPLOAD_MODE", "0"), **res)))
