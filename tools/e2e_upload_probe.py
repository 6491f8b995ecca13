"""End-to-end step time (host buffers in, [B,K,4] back) under different upload strategies:
dma (copy engine, strided boxes), pull (kernel reading mapped host memory), hybrid:f (both at once, fraction f by the kernel),
each with a launch shape of the pull kernel (threads per CTA, CTAs, parts per image).
Usage: python tools/e2e_upload_probe.py "dma" "pull/32/296/8" "hybrid:0.3/32/296/8/16" ...   (mode/threads/ctas/split/chunk/steps-ahead/slots)"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jarvis_hybridnet_b200.synth as S
from jarvis_hybridnet_b200 import HybridNet3D, _lib

sh = S.EXAMPLE
B = 32
w = S.make_v2v_weights(sh.K, 0, "he")
cam, intr, dist = S.make_rig(sh.ncam, 0)
sets = [S.make_frameset(sh, cam, intr, dist, s) for s in range(4)]
rep = lambda a: np.ascontiguousarray(np.broadcast_to(a[None], (B,) + a.shape))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
def batch(o):
    hm = S.to_cl16(np.stack([sets[(i + o) % 4][0] for i in range(B)]))
    return [t(hm), t(np.stack([sets[(i + o) % 4][1] for i in range(B)])), t(np.stack([sets[(i + o) % 4][2] for i in range(B)])), t(rep(cam)), t(rep(intr)), t(rep(dist))]
host = [batch(0), batch(1)]
net = HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, w, precision="bf16").cuda()
lib = _lib.load()
ref = None
K_steps = 12
for spec in sys.argv[1:]:
    parts = spec.split("/")
    mode = parts[0]
    thr, ctas, split = (int(x) for x in parts[1:4]) if len(parts) >= 4 else (128, 48, 8)
    chunk = int(parts[4]) if len(parts) >= 5 else 8
    ahead = int(parts[5]) if len(parts) >= 6 else 1
    nslots = int(parts[6]) if len(parts) >= 7 else ahead + 1
    lib.jhn_debug_set_pull_config(thr, ctas, split)
    for i in range(3):
        out = net.forward_host_async(host[i % 2], chunk=chunk, roi_upload=mode, slots=nslots).result()[0].clone()
    if ref is None:
        ref = out
    same = float((ref - out).abs().max())
    from collections import deque
    torch.cuda.synchronize(); t0 = time.perf_counter()
    q = deque()
    for i in range(K_steps):
        q.append(net.forward_host_async(host[i % 2], chunk=chunk, roi_upload=mode, slots=nslots))
        if len(q) > ahead:
            res, h2d, d2h = q.popleft().result()
    while q:
        res, h2d, d2h = q.popleft().result()
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / K_steps
    print(json.dumps(dict(spec=spec, ms_per_step=round(ms, 3), frame_sets_per_s=round(B / ms * 1e3), h2d_MB=round(h2d / 1e6, 1), max_abs_diff_vs_first=same)), flush=True)
