"""Timeline of the end-to-end steps (torch.profiler / CUPTI): when do the copy-engine transfers, the pull kernel and the compute
kernels of each step start and end?  Usage: python tools/e2e_timeline.py [mode] [chunk]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jarvis_hybridnet_b200.synth as S
from jarvis_hybridnet_b200 import HybridNet3D, _lib
from torch.profiler import profile, ProfilerActivity

mode = sys.argv[1] if len(sys.argv) > 1 else "hybrid:0.6"
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ahead = int(sys.argv[3]) if len(sys.argv) > 3 else 1
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
for i in range(3):
    net.forward_host_async(host[i % 2], chunk=chunk, roi_upload=mode, slots=ahead + 1).result()
torch.cuda.synchronize()
N = 8
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    from collections import deque
    q = deque()
    for i in range(N):
        q.append(net.forward_host_async(host[i % 2], chunk=chunk, roi_upload=mode, slots=ahead + 1))
        if len(q) > ahead:
            q.popleft().result()
    while q:
        q.popleft().result()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
def cls(name):
    n = name.lower()
    if "memcpy" in n: return "memcpy"
    if "pull_boxes" in n: return "pull"
    if "heatmap_boxes" in n or "boxes_kernel" in n: return "boxes"
    if "gather_stream" in n: return "gather"
    if "centroid_finalize" in n: return "finalize"
    return None
rows = []
for e in ev:
    c = cls(e.name)
    if c:
        rows.append((round((e.time_range.start - t0) / 1e3, 3), round((e.time_range.end - t0) / 1e3, 3), c, e.name[:40]))
# merge consecutive memcpys closer than 20 us
out = []
for r in rows:
    k = next((i for i in range(len(out) - 1, max(len(out) - 4, -1), -1) if out[i][2] == "memcpy"), None)
    if r[2] == "memcpy" and k is not None and r[0] - out[k][1] < 0.05:
        out[k] = (out[k][0], max(out[k][1], r[1]), "memcpy", out[k][3])
    else:
        out.append(r)
for r in out:
    if r[1] - r[0] > 0.02 or r[2] in ("boxes", "finalize", "gather"):
        print(f"{r[0]:9.3f} {r[1]:9.3f}  {r[1]-r[0]:7.3f} ms  {r[2]:8s} {r[3]}")
