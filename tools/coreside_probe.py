"""Does the pull kernel (jhn_pull_heatmap_boxes) run NEXT to the persistent compute kernels, or does it serialise with them?
Stream A: N forwards of 32 resident frame sets.  Stream B: the boxes of the same batch pulled from mapped host memory, repeatedly.
Reports the forward time alone, the pull rate alone, and both when the two streams run together, per launch shape of the pull kernel."""
import ctypes, json, os, sys, time
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
hm = S.to_cl16(np.stack([sets[i % 4][0] for i in range(B)]))
host = [t(hm), t(np.stack([sets[i % 4][1] for i in range(B)])), t(np.stack([sets[i % 4][2] for i in range(B)])), t(rep(cam)), t(rep(intr)), t(rep(dist))]
dev = [h.cuda() for h in host]
net = HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, w, precision="bf16").cuda()
lib = _lib.load()
ncam, hs = hm.shape[1], hm.shape[2]
pix = hm.shape[4] * 2
boxes = torch.empty((B, ncam, 4), dtype=torch.int32, device="cuda")
f = lambda x: x.contiguous().float()
_lib.check(lib.jhn_heatmap_boxes(_lib.dptr(f(dev[3])), _lib.dptr(f(dev[4])), _lib.dptr(f(dev[5])), _lib.dptr(f(dev[1])), _lib.dptr(dev[2].contiguous().to(torch.int32)),
                                 B, ncam, hs, net.G, float(net.spacing), _lib.dptr(boxes), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
bx = boxes.cpu().view(-1, 4)
nbytes = int(((1 - bx[:, 2] - bx[:, 0]) * (1 - bx[:, 3] - bx[:, 1])).clamp_(min=0).sum()) * pix
dst = torch.zeros_like(dev[0])
sa, sb = torch.cuda.Stream(), torch.cuda.Stream(priority=int(os.environ.get("PULL_PRIO", "0")))
NF, NP = 10, int(os.environ.get('PULL_REPS', '6'))

def fwd():
    with torch.cuda.stream(sa):
        for _ in range(NF): net(*dev)
SRC = dev[0] if os.environ.get("PULL_FROM_DEVICE") else host[0]     # PULL_FROM_DEVICE=1: same kernel, no PCIe — separates SM co-residency from link effects
def pull():
    with torch.cuda.stream(sb):
        for _ in range(NP):
            _lib.check(lib.jhn_pull_heatmap_boxes(ctypes.c_void_p(SRC.data_ptr()), ctypes.c_void_p(dst.data_ptr()), _lib.dptr(boxes), B * ncam, hs, pix, None,
                                                  ctypes.c_void_p(sb.cuda_stream)))
def timed(fa, fb):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    if fa: ev[0].record(sa)
    if fb: ev[2].record(sb)
    if fb: fb()
    if fa: fa()
    if fa: ev[1].record(sa)
    if fb: ev[3].record(sb)
    torch.cuda.synchronize()
    return (ev[0].elapsed_time(ev[1]) / NF if fa else None, nbytes * NP / (ev[2].elapsed_time(ev[3]) * 1e6) if fb else None)

lib.jhn_set_transfer_overlap(int(os.environ.get("PROBE_OVERLAP", "1")))     # 1: the 120-register build of the 3x3x3 layers
fwd(); torch.cuda.synchronize()
print(json.dumps(dict(forward_alone_ms=round(timed(fwd, None)[0], 3), box_MB=round(nbytes / 1e6, 1))), flush=True)
if os.environ.get("PROBE_KERNELS"):
    _lib.profile(True); timed(fwd, None); k = _lib.profile_collect(); _lib.profile(False)
    print("   ", {n: round(v[1] / NF, 3) for n, v in sorted(k.items(), key=lambda kv: -kv[1][1])}, flush=True)
for spec in sys.argv[1:]:
    thr, ctas, split = (int(x) for x in spec.split("/"))
    lib.jhn_debug_set_pull_config(thr, ctas, split)
    pull(); torch.cuda.synchronize()
    alone = timed(None, pull)[1]
    both = timed(fwd, pull)
    if os.environ.get("PROBE_KERNELS"):
        _lib.profile(True); timed(fwd, pull); k = _lib.profile_collect(); _lib.profile(False)
        print("   ", {n: round(v[1] / NF, 3) for n, v in sorted(k.items(), key=lambda kv: -kv[1][1]) if n != "pull_boxes_kernel"}, flush=True)
    print(json.dumps(dict(spec=spec, pull_alone_GBps=round(alone, 1), forward_ms_with_pull=round(both[0], 3), pull_GBps_with_forward=round(both[1], 1))), flush=True)

# ---- what does a CTA that merely sits on an SM cost?  (tools/coreside_dummy.cu, built with nvcc into tools/libcoreside_dummy.so)
so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcoreside_dummy.so")
if os.environ.get("PROBE_DUMMY") and os.path.exists(so):
    dl = ctypes.CDLL(so)
    dl.dummy_launch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    scratch = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda")
    for mode, name, carve in ((0, "sleep", -1), (0, "sleep", 100), (0, "sleep", 50)):
        assert dl.dummy_set_carveout(carve) == 0
        name = f"{name} carveout {carve}"
        for ctas, thr in ((48, 128), (148, 128)):
            def dummy():
                rc = dl.dummy_launch(ctas, thr, mode, 60000.0, ctypes.c_void_p(scratch.data_ptr()), scratch.numel() // 16, ctypes.c_void_p(sb.cuda_stream))
                assert rc == 0, rc
            both = timed(fwd, dummy)
            print(json.dumps(dict(dummy=name, ctas=ctas, threads=thr, forward_ms_next_to_it=round(both[0], 3))), flush=True)
            torch.cuda.synchronize()
            if os.environ.get("PROBE_KERNELS"):
                _lib.profile(True); timed(fwd, dummy); k = _lib.profile_collect(); _lib.profile(False)
                print("   ", {n: round(v[1] / NF, 3) for n, v in sorted(k.items(), key=lambda kv: -kv[1][1])}, flush=True)
