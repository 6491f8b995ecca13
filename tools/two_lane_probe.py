"""Feasibility probe for batch lanes: two HybridNet3D instances, 16 frame sets each, on two streams, against one instance with 32.
If convolution CTAs of one lane and normalisation blocks of the other share SMs, the two-lane rate exceeds the single-lane rate."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jarvis_hybridnet_b200.synth as S
from jarvis_hybridnet_b200 import HybridNet3D

sh = S.EXAMPLE
w = S.make_v2v_weights(sh.K, 0, "he")
cam, intr, dist = S.make_rig(sh.ncam, 0)
sets = [S.make_frameset(sh, cam, intr, dist, s) for s in range(4)]
def batch(B):
    rep = lambda a: np.ascontiguousarray(np.broadcast_to(a[None], (B,) + a.shape))
    hm = S.to_cl16(np.stack([sets[i % 4][0] for i in range(B)]))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return [t(hm), t(np.stack([sets[i % 4][1] for i in range(B)])), t(np.stack([sets[i % 4][2] for i in range(B)])), t(rep(cam)), t(rep(intr)), t(rep(dist))]
res = {}
one = HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, w, precision="bf16").cuda()
b32 = batch(32)
for _ in range(3): one(*b32)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): one(*b32)
torch.cuda.synchronize(); res["one_lane_32_ms"] = (time.perf_counter() - t0) * 100
nets = [HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, w, precision="bf16").cuda() for _ in range(2)]
b16 = [batch(16), batch(16)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for _ in range(3):
    for n, b, st in zip(nets, b16, streams):
        with torch.cuda.stream(st): n(*b)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    for n, b, st in zip(nets, b16, streams):
        with torch.cuda.stream(st): n(*b)
torch.cuda.synchronize(); res["two_lanes_16_ms"] = (time.perf_counter() - t0) * 100
for _ in range(3): nets[0](*b16[0])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): nets[0](*b16[0]); nets[0](*b16[0])
torch.cuda.synchronize(); res["one_lane_2x16_ms"] = (time.perf_counter() - t0) * 100
print(json.dumps(res))
