"""Whole-tensor upload vs box upload (copy engine) of the same small batch: per frame set max |difference| of the results.
Run under compute-sanitizer --tool racecheck to see whether the tool changes the outcome (JHN_UPLOAD_MODE=1: one 2-D copy per image)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jarvis_hybridnet_b200.synth as S
from jarvis_hybridnet_b200 import HybridNet3D
sh = S.SMALL
cam, intr, dist = S.make_rig(sh.ncam, 3)
sets = [S.make_frameset(sh, cam, intr, dist, s) for s in range(5)]
net = HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, S.make_v2v_weights(sh.K, 1, "he"), precision="bf16").cuda()
rep = lambda a: np.broadcast_to(a[None], (5,) + a.shape).copy()
hm = np.stack([s[0] for s in sets])
host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in
        (S.to_cl16(hm), np.stack([s[1] for s in sets]), np.stack([s[2] for s in sets]), rep(cam), rep(intr), rep(dist))]
want = net.forward_host(host, chunk=2, roi_upload=None)[0].clone()
for mode in ("dma", None, "dma", "pull", "hybrid:0.5"):
    res = net.forward_host(host, chunk=2, roi_upload=mode)[0].clone()
    print(mode, [round(float((res[b] - want[b]).abs().max()), 5) for b in range(5)], flush=True)
