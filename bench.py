#!/usr/bin/env python
"""Benchmark of the B200-native JARVIS-HybridNet 3D hot path (BASELINE.json metric: 3D frame-sets/s).

A step = one pass of the hot path (reprojection gather -> V2V 3D CNN -> centroid; reference:
jarvis/hybridnet/model.py:65-88) over one batch of synthetic frame sets.  The 2D CNNs, video decode and
CSV writing are outside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W]             our arm (N>1: launched under torchrun)
  python bench.py --impl reference ...                            the reference algorithm on host cores

Workload (config.workload): BASELINE.json configs[2] — full HybridNet 3D forward, 32 frame sets per step and
GPU, Example_Project shape (12 cameras, K=23, 128^2 heat maps, 72^3 grid).  Each rank runs an independent
frame stream (weak scaling, no collective in the timed region); results are all-gathered once afterwards.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import jarvis_hybridnet_b200.synth as S  # noqa: E402

WORKLOADS = {
    "c3_full3d_example": dict(shape=S.EXAMPLE, batch=32),
    "c2_micro": dict(shape=S.MICRO, batch=8),
    "c5_stress": dict(shape=S.STRESS, batch=8),
    "tiny": dict(shape=S.TINY, batch=4),
    # the predictor glue (centre localisation + crops) in front of the 3D path
    "f1_predictor_glue": dict(shape=S.EXAMPLE, batch=8),
}


_JSON_FD = None


def claim_stdout():
    """Rank 0 prints ONE JSON line on stdout.  Libraries (NCCL's version banner, torch warnings) write to fd 1 behind
    Python's back, so fd 1 is pointed at stderr for the whole run and the line goes to a private copy of the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def v2v_flops(sh):
    """Exact un-padded 2*MAC of V2VNet per frame set (SURVEY.md §8a layer table), by kernel family."""
    K, h, q = sh.K, sh.h, sh.h // 2
    M, Mq = h ** 3, q ** 3
    f = {
        "front_k3s2": 2 * M * K * 2 * K * 27,
        "res_k3_2C": 6 * 2 * M * (2 * K) ** 2 * 27,
        "pool_k2s2": 2 * Mq * 2 * K * 4 * K * 8,
        "res_k3_4C": 2 * 2 * Mq * (4 * K) ** 2 * 27,
        "up_convT": 2 * Mq * 4 * K * 2 * K * 8,
        "head_1x1": 2 * M * 2 * K * K,
    }
    f["total"] = sum(f.values())
    return f


def repro_bytes(sh, s_in=4, s_out=4):
    """Algorithmic bytes of the reprojection stage per frame set: every un-padded map read once, volume written once."""
    return sh.ncam * sh.K * sh.hm * sh.hm * s_in + sh.K * sh.G ** 3 * s_out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback (B200_PROFILING.md)")


def make_inputs(sh, batch, n_batches, seed=0, distinct=8):
    """Pool of `n_batches` host batches (numpy) of `batch` frame sets each; `distinct` rendered frame sets are
    re-used with fresh noise-free rotation so generation stays cheap.  Pool size exceeds L2 many times over."""
    cam, intr, dist = S.make_rig(sh.ncam, seed)
    sets = [S.make_frameset(sh, cam, intr, dist, seed * 100 + i) for i in range(distinct)]
    batches = []
    for nb in range(n_batches):
        ids = [(nb * batch + i) % distinct for i in range(batch)]
        hm = np.stack([sets[i][0] for i in ids])
        c3 = np.stack([sets[i][1] for i in ids]).astype(np.float32)      # the C ABI takes fp32 centres (int centres are promoted exactly)
        chm = np.stack([sets[i][2] for i in ids])
        rep = lambda a: np.broadcast_to(a[None], (batch,) + a.shape).copy()
        batches.append((hm, c3, chm, rep(cam), rep(intr), rep(dist)))
    return batches, sets, (cam, intr, dist)


def roi_fraction(sh, sets, rig):
    """Share of a padded heat map inside the camera's pixel box of the voxel grid (what the staging copy converts):
    the 8 corners of the coarse grid projected with synth.project, +-1 px, averaged over the frame sets and cameras."""
    cam, intr, dist = rig
    hs, fr = sh.hs, []
    lo, hi = -sh.roi / 2.0, sh.roi / 2.0 - 2 * sh.spacing
    for hm, c3, chm, _ in sets:
        corners = np.array([[x, y, z] for x in (lo, hi) for y in (lo, hi) for z in (lo, hi)], np.float64) + c3.astype(np.float64)
        px = S.project(corners, cam, intr, dist)                                  # [ncam,8,2] full-resolution pixels
        a = np.clip(px - chm[:, None, :] + hs - 1, 0, 2 * hs - 3) / 2.0
        x0, x1 = np.floor(a[..., 0].min(1)) - 1, np.floor(a[..., 0].max(1)) + 1
        y0, y1 = np.floor(a[..., 1].min(1)) - 1, np.floor(a[..., 1].max(1)) + 1
        w = np.clip(x1, 0, hs - 1) - np.clip(x0, 0, hs - 1) + 1
        h = np.clip(y1, 0, hs - 1) - np.clip(y0, 0, hs - 1) + 1
        fr.append(float((w * h).mean()) / (hs * hs))
    return float(np.mean(fr))


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm = [float(r[1]) for r in rows if len(r) >= 9]
        if not sm:
            return out
        out["sm_mhz"] = float(np.median(sm))
        out["sm_max_mhz"] = float(rows[0][2])
        out["samples"] = len(sm)
        out["power_w_max"] = max(float(r[3]) for r in rows if len(r) >= 9)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, n in enumerate(names):
            if any(r[5 + i].strip().lower() == "active" for r in rows if len(r) >= 9):
                out["reasons"].append(n)
        return out


def run_reference(args, wl):
    """Reference arm: the reference's algorithm on the host cores (oracle port of repro_layer.py / v2vnet.py /
    model.py tail), all host threads, one bounded sample of the workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import hybridnet_oracle as O
    O.build()
    sh = wl["shape"]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cam, intr, dist = S.make_rig(sh.ncam, 0)
    sample = max(1, min(wl["batch"], args.ref_sample))
    sets = [S.make_frameset(sh, cam, intr, dist, i) for i in range(sample)]
    w = S.make_v2v_weights(sh.K, 0, "he")

    def step():
        for hm, c3, chm, _ in sets:
            O.hybrid3d_forward(w, hm, c3, chm, cam, intr, dist, sh.roi, sh.spacing)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = dict(metric="frame_sets_per_sec", value=v, unit="frame-sets/s", impl="reference", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=args.workload, ncam=sh.ncam, K=sh.K, heatmap=sh.hm, grid=sh.G,
                            frame_sets_per_step=sample),
                cpu_baseline=dict(value=v, unit="frame-sets/s", cores=cores, kind="port",
                                  sample=f"{sample} frame set(s) of the workload shape per step, {args.steps} steps, "
                                         f"oracle port of the reference (C index/gather/tail + torch-CPU V2V)"),
                e2e=dict(value=v, unit="frame-sets/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def cpu_baseline(sh, budget_s=12.0):
    import torch
    from oracle import hybridnet_oracle as O
    O.build()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cam, intr, dist = S.make_rig(sh.ncam, 0)
    hm, c3, chm, _ = S.make_frameset(sh, cam, intr, dist, 0)
    w = S.make_v2v_weights(sh.K, 0, "he")
    O.hybrid3d_forward(w, hm, c3, chm, cam, intr, dist, sh.roi, sh.spacing)      # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        O.hybrid3d_forward(w, hm, c3, chm, cam, intr, dist, sh.roi, sh.spacing)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 64:
            break
    return dict(value=n / dt, unit="frame-sets/s", cores=cores, kind="port",
                sample=f"{n} frame set(s) of the workload shape in {dt:.1f} s, oracle port of the reference "
                       f"(C index/gather/tail on {cores} threads + torch-CPU V2V)")


def run_glue(args, wl):
    """--workload f1_predictor_glue: jhn_center_locate + jhn_crop_normalize (jarvis3D.py:147-177) for `batch` frame sets
    per step: 12 cameras, 128^2 centre heat maps, 1280x1024 fp32 images, 256^2 crops.  Same JSON contract; the
    roofline is the crop kernel's (HBM: window read + crop written)."""
    import torch
    from jarvis_hybridnet_b200 import _lib, crop_normalize, locate_center
    from oracle import center_oracle as C
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    if int(os.environ.get("WORLD_SIZE", "1")) != 1 or args.gpus != 1:
        raise SystemExit("f1_predictor_glue is a single-GPU microbenchmark")
    torch.cuda.set_device(0)
    sh, B = wl["shape"], wl["batch"]
    cdis, bbox, MEAN, STD = 256, sh.bbox, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    cam, intr, dist = S.make_rig(sh.ncam, 0)
    cases = [S.make_center_case(sh.ncam, cam, intr, dist, s, cdis) for s in range(2)]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    n_pool = 2                                                                   # 2 x B x 189 MB of images >> L2
    hm_h = [torch.stack([t(cases[(i + p) % 2][0][:, 0]) for i in range(B)]).pin_memory() for p in range(n_pool)]
    im_h = [torch.stack([t(cases[(i + p) % 2][1]) for i in range(B)]).pin_memory() for p in range(n_pool)]
    rep = lambda a: t(a)[None].expand(B, *a.shape).contiguous().cuda()
    camd, intrd, distd = rep(cam), rep(intr), rep(dist)
    hm_d, im_d = [x.cuda() for x in hm_h], [x.cuda() for x in im_h]
    scratch = torch.zeros(2 * B, dtype=torch.int32, device="cuda")

    def step(hm, im):
        loc = locate_center(hm, (S.IMG_W, S.IMG_H), cdis, bbox // 2, camd, intrd, distd, scratch=scratch)
        return loc, crop_normalize(im, loc["centerHM"], loc["valid"], bbox, MEAN, STD)

    K_steps, W = args.steps, max(args.warmup, 3)
    clk = ClockSampler(0)
    for i in range(W):
        step(hm_d[i % n_pool], im_d[i % n_pool])
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K_steps):
        loc, crops = step(hm_d[i % n_pool], im_d[i % n_pool])
    e1.record(); torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    ms = e0.elapsed_time(e1)
    # end to end: host images + heat maps in, crops stay on the device for the 3D network, centres read back
    hbuf, ibuf = torch.empty_like(hm_d[0]), torch.empty_like(im_d[0])
    t0 = time.perf_counter()
    for i in range(K_steps):
        hbuf.copy_(hm_h[i % n_pool], non_blocking=True); ibuf.copy_(im_h[i % n_pool], non_blocking=True)
        loc, crops = step(hbuf, ibuf)
        chm = loc["centerHM"].cpu()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = clk.stop()
    _lib.profile(True)
    for i in range(K_steps):
        step(hm_d[i % n_pool], im_d[i % n_pool])
    prof = _lib.profile_collect(); _lib.profile(False)
    kern = {k: dict(launches=n, ms_per_step=msk / K_steps) for k, (n, msk) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    pk = peaks()
    crop_bytes = B * sh.ncam * 3 * bbox * bbox * 4 * 2
    cms = prof["crop_normalize_kernel"][1] / prof["crop_normalize_kernel"][0]
    roof = dict(kernel="crop_normalize_kernel", bound="hbm", achieved=crop_bytes / (cms * 1e-3) / 1e9, peak=pk["hbm"], unit="GB/s",
                frac=crop_bytes / (cms * 1e-3) / 1e9 / pk["hbm"], traffic=None, peak_source=pk["src"],
                algorithmic_bytes_per_launch=crop_bytes, avg_launch_ms=cms)
    cb = None
    if not args.no_cpu_baseline:
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 8.0:
            r = C.locate_center(cases[n % 2][0], S.IMG_W, S.IMG_H, cdis, bbox // 2, cam, intr, dist)
            C.crop_normalize(cases[n % 2][1], r["centerHM"], bbox // 2, MEAN, STD)
            n += 1
        dt = time.perf_counter() - t0
        cb = dict(value=n / dt, unit="frame-sets/s", cores=1, kind="port",
                  sample="%d frame set(s) in %.1f s, numpy oracle port of jarvis3D.py:147-177 (1 thread)" % (n, dt))
    h2d = hm_h[0].numel() * 4 + im_h[0].numel() * 4
    line = dict(metric="frame_sets_per_sec", value=B * K_steps / (ms * 1e-3), unit="frame-sets/s", n_gpus=1, steps=K_steps, warmup=W,
                ms_per_step=ms / K_steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="f1_predictor_glue", ncam=sh.ncam, center_heatmap=cdis // 2, image=[S.IMG_W, S.IMG_H], bbox=bbox,
                            frame_sets_per_step_per_gpu=B, l2="inputs rotate through 2 batches of %.0f MB (>> 126 MB L2)" % (h2d / 1e6),
                            timed_region="centre heat maps + images -> crop centres, centre3D, normalised crops (jarvis3D.py:147-177)"),
                clocks=clocks, e2e=dict(value=B * K_steps / e2e_s, unit="frame-sets/s", h2d_bytes_per_step=int(h2d),
                                        d2h_bytes_per_step=int(chm.numel() * 4), ms_per_step=1e3 * e2e_s / K_steps),
                gpu_launches=int(launches), roofline=roof, cpu_baseline=cb, kernels=kern)
    emit(line)


def torch_reference_gpu(sh, weights, batch, n_frame_sets=8, reps=3):
    """"The kernel to beat" (SURVEY.md section 8d): the reference's own algorithm for stages a2-a10 with the SAME torch
    operators in the same order on the B200 — cuBLAS SGEMM + ATen upsample_trilinear3d + index_select + mean
    (jarvis/hybridnet/repro_layer.py:40-119), cuDNN conv3d / conv_transpose3d + ATen instance_norm
    (jarvis/hybridnet/v2vnet.py:12-102), the softplus centroid (jarvis/hybridnet/model.py:65-88) — one frame set per
    call like the reference (repro_layer.py:112-117), inputs resident.  A baseline leg only: library kernels, nothing of
    this repo's product path; /root/reference itself does not exist on the GPU box, so the modules cannot be imported."""
    import torch
    import torch.nn.functional as F
    dev = torch.device("cuda")
    hm_b, c3_b, chm_b, cam_b, intr_b, dist_b = batch
    G, h, hs, K, ncam = sh.G, sh.h, sh.hs, sh.K, sh.ncam
    W = {k: torch.as_tensor(v).to(dev) for k, v in weights.items()}
    half = G // 2 // 2
    ar = torch.arange(h, dtype=torch.float32, device=dev) - half
    grid0 = torch.stack(torch.meshgrid(ar, ar, ar, indexing="ij"), dim=3) * sh.spacing * 2          # repro_layer.py:26-36
    xx, yy, zz = torch.meshgrid(torch.arange(h, device=dev), torch.arange(h, device=dev), torch.arange(h, device=dev), indexing="ij")

    def conv(p, t, stride=1, pad=0):
        return F.conv3d(t, W[p + ".weight"], W[p + ".bias"], stride=stride, padding=pad)

    def basic(p, t, k, stride):
        return F.relu(F.instance_norm(conv(p + ".block.0", t, stride, (k - 1) // 2)))

    def res(p, t):
        r = F.relu(F.instance_norm(conv(p + ".res_branch.0", t, 1, 1)))
        return F.relu(F.instance_norm(conv(p + ".res_branch.3", r, 1, 1)) + t)

    def one(hm, c3, chm, cam, intr, dist):
        hp = F.pad(hm, [1, 1, 1, 1])                                                                # model.py:65-66
        x = grid0 + c3                                                                              # repro_layer.py:113
        intr_p, dist_p, chm_p = intr.permute(1, 2, 0), dist.permute(1, 2, 0), chm.permute(1, 0)
        x = torch.cat((x, torch.ones(h, h, h, 1, device=dev)), 3)
        pa = torch.matmul(x.view(1, -1, 4), cam).view(-1, h, h, h, 3).permute(1, 2, 3, 4, 0)
        v1 = pa[:, :, :, 0] / pa[:, :, :, 2] - intr_p[2, 0]
        v2 = pa[:, :, :, 1] / pa[:, :, :, 2] - intr_p[2, 1]
        r2 = torch.square(v1 / intr_p[0, 0]) + torch.square(v2 / intr_p[1, 1])
        d = 1 + (dist_p[0, 0] + dist_p[0, 1] * r2) * r2
        v1 = v1 * d + intr_p[2, 0]
        v2 = v2 * d + intr_p[2, 1]
        v1 = torch.clamp(v1, chm_p[0] - (hs - 1), chm_p[0] + hs - 2) - chm_p[0] + hs - 1
        v2 = torch.clamp(v2, chm_p[1] - (hs - 1), chm_p[1] + hs - 2) - chm_p[1] + hs - 1
        up = lambda v: F.interpolate(v.permute(3, 0, 1, 2).reshape(1, ncam, h, h, h), size=(G, G, G),
                                     mode="trilinear").view(ncam, G, G, G).permute(1, 2, 3, 0)
        f1, f2 = up(v1), up(v2)
        rp = ((f2 / 2).int() * hs + (f1 / 2).int()).permute(3, 0, 1, 2).long()                      # :82-83
        off = torch.arange(0, hs * hs * ncam, hs * hs, device=dev)
        hmf = hp.transpose(0, 1).flatten(1)                                                         # [K, ncam*hs*hs]  :97-103
        rp = (rp.flatten(1).transpose(1, 0) + off).transpose(1, 0).flatten()
        vol = torch.mean(torch.index_select(hmf, 1, rp).view(K, ncam, G, G, G), dim=1)[None]
        x = basic("front_layers.0", vol / 255., 3, 2)                                               # v2vnet.py:98-102
        x = res("front_layers.1", x)
        s = res("encoder_decoder.skip_res1", x)
        y = basic("encoder_decoder.encoder_pool1", x, 2, 2)
        y = res("encoder_decoder.mid_res", y)
        p = "encoder_decoder.decoder_upsample1.block.0"
        y = F.relu(F.instance_norm(F.conv_transpose3d(y, W[p + ".weight"], W[p + ".bias"], stride=2)))
        y = res("encoder_decoder.decoder_res1", y) + s
        hf = F.softplus(conv("output_layer", y))                                                    # model.py:72-73
        n = torch.sum(hf, dim=[2, 3, 4])
        px = torch.sum(hf * xx, dim=[2, 3, 4]) / n
        py = torch.sum(hf * yy, dim=[2, 3, 4]) / n
        pz = torch.sum(hf * zz, dim=[2, 3, 4]) / n
        pts = torch.stack([px, py, pz], dim=2) * sh.spacing * 2 - sh.roi / 2. + c3
        conf = torch.clamp(torch.max(hf.view(1, K, -1), dim=2)[0], max=255.) / 255.
        return pts, conf

    out = {}
    n = min(n_frame_sets, hm_b.shape[0])
    for tag, tf32 in (("tf32_on", True), ("tf32_off", False)):
        torch.backends.cudnn.allow_tf32 = tf32                      # the reference runs with torch's default (conv TF32 on)
        with torch.no_grad():
            for b in range(2):
                one(hm_b[b], c3_b[b], chm_b[b], cam_b[b], intr_b[b], dist_b[b])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                for b in range(n):
                    pts, conf = one(hm_b[b], c3_b[b], chm_b[b], cam_b[b], intr_b[b], dist_b[b])
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * n)
        out[tag] = dict(ms_per_frame_set=ms, frame_sets_per_sec=1e3 / ms)
    torch.backends.cudnn.allow_tf32 = True
    out["note"] = ("reference algorithm restated op for op with torch library kernels (cuBLAS, cuDNN, ATen) on this GPU, one frame "
                   "set per call as in repro_layer.py:112-117, %d frame sets x %d repetitions, fp32 inputs resident" % (n, reps))
    return out


def reference_modules_gpu(sh, weights, batch, n_frame_sets=8, reps=3):
    """The UNMODIFIED reference modules on this GPU (baseline leg; needs the pip-installed copy under baseline/_ref, see
    baseline/ref_shim.py): HybridNetBackbone built by its own constructor (jarvis/hybridnet/model.py:20-50), this run's V2V
    weights loaded with strict=True, effTrack replaced by a stub returning the resident heat maps (the 2D CNN is outside
    the timed region for both arms), called one frame set at a time as predict3D does (predict3D.py:75-97).  Returns None
    when the install is absent."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import ref_shim
        if not ref_shim.import_reference():
            return None
        from jarvis.hybridnet.model import HybridNetBackbone
    except Exception as e:                                              # a baseline leg must not take the bench down
        return dict(error=str(e)[:200])
    hm_b, c3_b, chm_b, cam_b, intr_b, dist_b = batch
    bb = HybridNetBackbone(ref_shim.make_cfg(sh.ncam, sh.K, sh.bbox, sh.roi, sh.spacing)).cuda().eval()
    bb.v2vNet.load_state_dict({k: torch.as_tensor(v) for k, v in weights.items()}, strict=True)
    bb = bb.cuda()

    class Stub(torch.nn.Module):
        hm = None
        def forward(self, x):
            return None, self.hm
    bb.effTrack = Stub()
    imgs = torch.zeros(1, sh.ncam, 3, 4, 4, device="cuda")
    size = torch.tensor([1280, 1024], device="cuda")

    def one(b):
        bb.effTrack.hm = hm_b[b]
        return bb(imgs, size, chm_b[b:b + 1], c3_b[b:b + 1].int(), cam_b[b:b + 1], intr_b[b:b + 1], dist_b[b:b + 1])
    out = {}
    n = min(n_frame_sets, hm_b.shape[0])
    for tag, tf32 in (("tf32_on", True), ("tf32_off", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.no_grad():
            for b in range(2):
                one(b)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                for b in range(n):
                    one(b)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * n)
        out[tag] = dict(ms_per_frame_set=ms, frame_sets_per_sec=1e3 / ms)
    torch.backends.cudnn.allow_tf32 = True
    out["note"] = ("unmodified jarvis.hybridnet.model.HybridNetBackbone.forward from baseline/_ref on this GPU (reproLayer + v2vNet + "
                   "tail; effTrack stubbed), one frame set per call, %d frame sets x %d repetitions, inputs resident" % (n, reps))
    return out


def micro_reproject_tail(sh, B, pk, steps=10):
    """BASELINE.json configs[1] / metric (ii): ReprojectionLayer (+ /255) and the centroid tail on their own, fp32 and bf16,
    as achieved HBM GB/s over the algorithmic bytes of SURVEY.md section 8d."""
    import torch
    from types import SimpleNamespace as NS
    from jarvis_hybridnet_b200 import ReprojectionLayer, centroid_tail
    cfg = NS(HYBRIDNET=NS(GRID_SPACING=sh.spacing, ROI_CUBE_SIZE=sh.roi, NUM_CAMERAS=sh.ncam),
             KEYPOINTDETECT=NS(BOUNDING_BOX_SIZE=sh.bbox, NUM_JOINTS=sh.K))
    batches, _, _ = make_inputs(sh, B, 2, seed=5, distinct=4)
    devb = [[torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in b] for b in batches]
    out = {}

    def timed(fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for prec, s_in, s_out in (("fp32", 4, 4), ("bf16", 4, 2)):
        L = ReprojectionLayer(cfg, precision=prec)
        layout = "ncdhw" if prec == "fp32" else "v2v"
        ms = timed(lambda i: L.forward_batched(*devb[i % 2], post_divide=255.0, volume_layout=layout))
        by = B * repro_bytes(sh, s_in, s_out)
        out["reproject_" + prec] = dict(ms_per_step=ms, algorithmic_bytes_per_step=by, GB_per_s=by / (ms * 1e-3) / 1e9,
                                        frac_of_hbm=by / (ms * 1e-3) / 1e9 / pk["hbm"], input="fp32 planar (reference tensor)",
                                        volume="fp32 NCDHW" if prec == "fp32" else "bf16 V2V layout")
    # gather-native fp16 channels-last input (row f2 layout): no staging pass, 2 bytes per input element
    L = ReprojectionLayer(cfg, precision="bf16")
    cl = [torch.from_numpy(S.to_cl16(b[0])).cuda() for b in batches]
    ms = timed(lambda i: L.forward_batched(cl[i % 2], *devb[i % 2][1:], post_divide=255.0, volume_layout="v2v"))
    by = B * repro_bytes(sh, 2, 2)
    out["reproject_bf16_f16cl_input"] = dict(ms_per_step=ms, algorithmic_bytes_per_step=by, GB_per_s=by / (ms * 1e-3) / 1e9,
                                             frac_of_hbm=by / (ms * 1e-3) / 1e9 / pk["hbm"], input="fp16 channels-last", volume="bf16 V2V layout")
    v = [torch.randn((B, sh.K, sh.h, sh.h, sh.h), device="cuda") * 20 for _ in range(max(2, int(600e6 // (B * sh.K * sh.h ** 3 * 4)) + 1))]
    ms = timed(lambda i: centroid_tail(v[i % len(v)], sh.spacing, sh.roi, devb[0][1]))
    by = B * sh.K * sh.h ** 3 * 4
    out["tail_fp32"] = dict(ms_per_step=ms, algorithmic_bytes_per_step=by, GB_per_s=by / (ms * 1e-3) / 1e9,
                            frac_of_hbm=by / (ms * 1e-3) / 1e9 / pk["hbm"],
                            note="stand-alone fp32 tail; on the bf16 path the tail is the output layer's epilogue (zero extra traffic)")
    out["config"] = dict(ncam=sh.ncam, K=sh.K, heatmap=sh.hm, grid=sh.G, frame_sets_per_step=B, l2="two input batches alternate")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_full3d_example", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="frame sets per step and GPU (default: workload's)")
    ap.add_argument("--precision", default="auto", choices=["auto", "bf16", "fp32"])
    ap.add_argument("--ref-sample", type=int, default=2, help="frame sets per reference-arm step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--job-frame-sets", type=int, default=0,
                    help="also run one sharded job of this many frame sets with a single gather at the end (BASELINE configs[3]: 100000)")
    ap.add_argument("--no-latency", action="store_true", help="skip the B=1 eager / CUDA-graph latency measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the torch GPU baseline and the configs 2 / 5 side measurements")
    ap.add_argument("--sub-batch", type=int, default=0, help="frame sets per internal pass of jhn_hybrid3d_forward (0 = library default)")
    args = ap.parse_args()
    claim_stdout()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
    if args.workload == "f1_predictor_glue":
        return run_glue(args, wl)
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    from jarvis_hybridnet_b200 import HybridNet3D, _lib, gather_results

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 under torch.distributed.run")
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"               # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sh, B = wl["shape"], wl["batch"]
    K_steps, W = args.steps, max(args.warmup, 0)
    sub_batch = _lib.set_sub_batch(args.sub_batch)

    precision = args.precision
    weights = S.make_v2v_weights(sh.K, 0, "he")
    if precision == "auto":
        precision = "bf16"          # BASELINE.json configs[2] names bf16; a failure to build the tensor-core handle is an error, not a downgrade
    net = HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, weights, precision=precision).cuda()

    n_pool = 2
    batches, sets, rig = make_inputs(sh, B, n_pool, seed=rank)
    to_t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    host = [[to_t(a).pin_memory() for a in b] for b in batches]
    devb = [[t.cuda(non_blocking=True) for t in b] for b in host]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput --------------------------------------------------------------
    clk = ClockSampler(local)          # samples every 50 ms across warm-up, timed region and e2e region
    for i in range(W):
        net(*devb[i % n_pool])
    barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K_steps):
        out = net(*devb[i % n_pool])
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * K_steps / (ms * 1e-3)

    # ---- the same with the heat maps resident in the gather-native layout (fp16 channels-last, SURVEY.md section 8 row f2:
    # what a 2D head that writes for the gather emits; jhn_heatmap_convert / jhn_efftrack_head produce it on the device) ----
    value_cl = None
    if precision == "bf16":
        host_cl = [[to_t(S.to_cl16(b[0])).pin_memory()] + h[1:] for b, h in zip(batches, host)]
        dev_cl = [[h[0].cuda(non_blocking=True)] + d[1:] for h, d in zip(host_cl, devb)]
        for i in range(min(W, 3)):
            net(*dev_cl[i % n_pool])
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(K_steps):
            net(*dev_cl[i % n_pool])
        c1.record()
        barrier()
        t = torch.tensor([c0.elapsed_time(c1)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        value_cl = dict(value=world * B * K_steps / (float(t.item()) * 1e-3), unit="frame-sets/s", ms_per_step=float(t.item()) / K_steps,
                        input="heat maps resident as fp16 channels-last [B,ncam,hs,hs,24] (JHN_HM_F16_CL): no staging pass")
    else:
        host_cl = host

    # ---- end to end through the public API with host buffers: pinned host tensors in, [B,K,4] back in pinned memory.
    # On the bf16 path the heat maps cross PCIe in the gather-native fp16 channels-last form (9.7 MB instead of 18.1 MB per
    # frame set at the Example shape); everything else is the reference's fp32 / int32 tensors. --------------------------
    # Upload strategy (HybridNet3D.forward_host_async): "hybrid:0.6" = the copy engine moves 40 % of every chunk's per-camera pixel
    # boxes (strided DMA, bound by rows per second), the pull kernel reads the other 60 % straight out of the pinned host tensor
    # (bound by the link), both at once; chunk = the whole batch.  JHN_E2E_UPLOAD=dma JHN_E2E_CHUNK=8 is the copy-engine-only path.
    # Up to 4 ranks per host.  With 8 the ranks' transfers together hit the host's memory / PCIe fabric (~190 GB/s on this pool's
    # boxes): bytes decide there, so every image is pulled as its per-row column spans (jhn_heatmap_spans, 22 % fewer bytes than the
    # boxes; "hybrid-spans:1.0").  8 GPUs end to end: 40.8 k frame-sets/s, against 35.1 k with the copy engine alone and 31.8 k with the
    # box hybrid; 4 GPUs: 27.5 k with the box hybrid, 25.5 k with the copy engine (profiles/r02_run64_8gpu_*.json,
    # r02_run66_4gpu_*.json, r02_run78_8gpu_spans.txt).  On one GPU computing the spans costs more than the link gains.
    e2e_upload = os.environ.get("JHN_E2E_UPLOAD", "dma" if precision != "bf16" else "hybrid:0.6" if world <= 4 else "hybrid-spans:1.0")
    e2e_chunk = int(os.environ.get("JHN_E2E_CHUNK", str(B) if e2e_upload.startswith("hybrid") else "8"))
    # Steps are pipelined, as a prediction loop with a prefetching loader runs them: up to `e2e_ahead` later steps are submitted
    # (forward_host_async: their uploads queue behind the running step's on the copy / pull streams) before a step's result is
    # collected, so the link stays busy while that step computes.  Every step uploads its own inputs and downloads its own result
    # inside the region.  Two steps ahead (three sets of device buffers) make the step time insensitive to the copy-engine /
    # pull-kernel split (4.15 - 4.27 ms for fractions 0.5 - 0.7; one step ahead: 4.0 - 4.8; profiles/r02_e2e_hybrid_upload.txt).
    from collections import deque
    e2e_ahead = int(os.environ.get("JHN_E2E_AHEAD", "2" if e2e_upload.startswith("hybrid") else "1"))
    submit = lambda i: net.forward_host_async(host_cl[i % n_pool], chunk=e2e_chunk, roi_upload=e2e_upload, slots=e2e_ahead + 1)
    for i in range(max(min(W, 3), 2)):
        submit(i).result()
    barrier()
    t0 = time.perf_counter()
    inflight = deque()
    for i in range(K_steps):
        inflight.append(submit(i))
        if len(inflight) > e2e_ahead:
            res, h2d, d2h = inflight.popleft().result()
    while inflight:
        res, h2d, d2h = inflight.popleft().result()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = clk.stop()
    e2e = dict(value=world * B * K_steps / e2e_s, unit="frame-sets/s", h2d_bytes_per_step=int(h2d),
               d2h_bytes_per_step=int(d2h), ms_per_step=1e3 * e2e_s / K_steps,
               heatmap_format="fp16 channels-last (JHN_HM_F16_CL)" if precision == "bf16" else "fp32 planar",
               upload="only the pixels of each camera's map that the voxel grid projects to (%s; dma = jhn_heatmap_boxes + jhn_upload_heatmap_boxes, "
                      "copy engine; hybrid:f = a fraction f of the images by jhn_pull_heatmap_boxes, a kernel reading the pinned host tensor; "
                      "hybrid-spans:f = those by jhn_heatmap_spans + jhn_pull_heatmap_spans, per-row column spans), chunk %d, up to %d steps "
                      "submitted ahead" % (e2e_upload, e2e_chunk, e2e_ahead))

    # ---- B=1 latency (the reference's predictor runs one frame set per call): eager launches vs CUDA-graph replay ----
    latency = None
    if rank == 0 and not args.no_latency:
        one = [t[:1].contiguous() for t in devb[0]]
        def timed(fn, n=50):
            for _ in range(5):
                fn(*one)
            torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); a.record()
            for _ in range(n):
                fn(*one)
            b.record(); torch.cuda.synchronize()
            return dict(device_ms=a.elapsed_time(b) / n, wall_ms=1e3 * (time.perf_counter() - t0) / n)
        latency = dict(frame_sets=1, eager=timed(net.forward), cuda_graph=timed(net.forward_graph),
                       note="one frame set per call, inputs resident; graph replay includes the copy into its static inputs")

    # ---- the one collective of the job: gather [N,K,4] at the end ---------------------------------
    gather_ms = None
    if world > 1:
        pts, conf, _ = out
        local_res = torch.cat([pts, conf[..., None]], 2)
        torch.cuda.synchronize(); g0 = time.perf_counter()
        full = gather_results(local_res, world * B)
        torch.cuda.synchronize(); gather_ms = 1e3 * (time.perf_counter() - g0)
        assert full.shape[0] == world * B

    # ---- BASELINE.json configs[3]: one sharded job of N frame sets, outputs gathered once at the end -----------
    job = None
    if args.job_frame_sets > 0:
        from jarvis_hybridnet_b200 import shard_range
        N = args.job_frame_sets
        lo, hi = shard_range(N, rank, world)
        local_res = torch.empty((hi - lo, sh.K, 4), dtype=torch.float32, device="cuda")
        barrier()
        j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        j0.record()
        done, i = 0, 0
        while done < hi - lo:
            nb = min(B, hi - lo - done)
            pts, conf, _ = net(*[t[:nb] for t in devb[i % n_pool]])
            local_res[done:done + nb, :, :3] = pts
            local_res[done:done + nb, :, 3] = conf
            done += nb; i += 1
        full = gather_results(local_res, N) if world > 1 else local_res
        j1.record()
        barrier()
        tj = torch.tensor([j0.elapsed_time(j1)], device="cuda")
        if world > 1:
            dist.all_reduce(tj, op=dist.ReduceOp.MAX)
        assert full.shape[0] == N and bool(torch.isfinite(full).all())
        job = dict(frame_sets=N, ms=float(tj.item()), frame_sets_per_sec=N / (float(tj.item()) * 1e-3), steps_per_rank=i,
                   note="contiguous shard per rank, %d frame sets per step, one all-gather of [N,K,4] inside the timed region" % B)

    # ---- per-kernel device time (CUDA events inside the library) over the same steps ---------------
    _lib.profile(True)
    for i in range(K_steps):
        net(*devb[i % n_pool])
    prof = _lib.profile_collect()
    _lib.profile(False)
    pk = peaks()
    fl = v2v_flops(sh)
    kern = {k: dict(launches=n, ms_per_step=msk / K_steps) for k, (n, msk) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    s_in, s_vol = 4, (4 if precision == "fp32" else 2)
    roi = roi_fraction(sh, sets, rig)
    REPRO = ("coarse_project_kernel", "fine_index_kernel", "relayout_kernel", "gather_mean_kernel", "gather_fused_kernel",
             "gather_staged_kernel", "gather_stream_kernel")
    rp_ms = sum(v["ms_per_step"] for k, v in kern.items() if k in REPRO)
    conv_ms = sum(v["ms_per_step"] for k, v in kern.items() if "conv" in k)
    tail_ms = sum(v["ms_per_step"] for k, v in kern.items() if "centroid" in k)
    stages = dict(
        reproject=dict(ms_per_step=rp_ms, algorithmic_GB_per_s=B * repro_bytes(sh, s_in, s_vol) / (rp_ms * 1e-3) / 1e9 if rp_ms else None),
        v2v_convs=dict(ms_per_step=conv_ms, algorithmic_TFLOP_per_s=B * fl["total"] / (conv_ms * 1e-3) / 1e12 if conv_ms else None,
                       note="includes the fused output-layer + centroid kernel on the bf16 path"),
        tail=dict(ms_per_step=tail_ms),
    )
    # algorithmic work per STEP of each kernel family: FLOP for the GEMM kernels, bytes for the rest
    act = lambda c, d: B * c * d ** 3 * 2                                    # one bf16 activation tensor, un-padded
    h, q, K = sh.h, sh.h // 2, sh.K
    flop_fam = {"tc_conv3_stacked": fl["res_k3_2C"], "tc_conv_k3_resident": fl["res_k3_2C"], "tc_conv_k3_streamed": fl["res_k3_4C"],
                "tc_conv_front_k3s2": fl["front_k3s2"], "tc_conv_pool_k2s2": fl["pool_k2s2"], "tc_conv_up_convT": fl["up_convT"],
                "tc_conv_head_1x1": fl["head_1x1"], "conv3d_f32_kernel<3>": fl["front_k3s2"] + fl["res_k3_2C"] + fl["res_k3_4C"]}
    byte_fam = {"gather_staged_kernel": B * repro_bytes(sh, 2, s_vol), "gather_stream_kernel": B * repro_bytes(sh, 2, s_vol), "gather_fused_kernel": B * repro_bytes(sh, s_vol, s_vol),
                "gather_mean_kernel": B * repro_bytes(sh, s_vol, s_vol),
                # the bf16 path converts only each camera's pixel box of the voxel grid (share `roi` of the padded map)
                "relayout_kernel": B * sh.ncam * K * (sh.hm ** 2 if precision == "fp32" else roi * sh.hs ** 2) * (s_in + s_vol),
                "coarse_project_kernel": B * sh.ncam * h ** 3 * 8,
                "tc_head_centroid_kernel": act(2 * K, h), "centroid_kernel": B * K * h ** 3 * 4,
                # 5 plain + 1 residual + 1 residual/PS-copy + 1 residual/skip on the h grid, 2 plain + 1 residual on the q grid
                "tc_norm_act_kernel": (5 * 2 + 3 + 4 + 4) * act(2 * K, h) + (2 * 2 + 3) * act(4 * K, q)}
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")                 # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp) and args.workload == "c3_full3d_example" and B == 32:
        traffic = json.load(open(tp)).get("dram_bytes_per_launch", {})
    # tensor peak: the timed region is a few tens of ms; when the clock record shows the SMs at >= 1.9 GHz with no power cap it
    # ran in the burst regime of MEASURED_PEAKS.json (cuBLAS best-of-10), otherwise against the sustained figure
    burst = bool(clocks and (clocks.get("sm_mhz") or 0) >= 1900 and "sw_power_cap" not in (clocks.get("reasons") or []))
    t_peak = pk["bf16"] if burst else pk["bf16_sustained"]
    t_src = pk["src"] + (", burst bf16 (SM clock %.0f MHz, no power cap)" % clocks["sm_mhz"] if burst else ", sustained bf16")
    rooflines = {}
    for k, v in kern.items():
        n_per_step = prof[k][0] / K_steps
        per_launch_ms = prof[k][1] / prof[k][0]
        if k in flop_fam:
            work = B * flop_fam[k] / n_per_step
            ach = work / (per_launch_ms * 1e-3) / 1e12
            rooflines[k] = dict(kernel=k, bound="tensor", achieved=ach, peak=t_peak, unit="TFLOP/s",
                                frac=ach / t_peak, traffic=traffic.get(k), peak_source=t_src,
                                algorithmic_flop_per_launch=work, avg_launch_ms=per_launch_ms)
        elif k in byte_fam:
            work = byte_fam[k] / n_per_step
            ach = work / (per_launch_ms * 1e-3) / 1e9
            rooflines[k] = dict(kernel=k, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                                traffic=traffic.get(k), peak_source=pk["src"], algorithmic_bytes_per_launch=work,
                                avg_launch_ms=per_launch_ms)
    top = next(iter(kern)) if kern else None                                # dominant kernel = most device time per step
    roofline = rooflines.get(top)

    extra = None
    if rank == 0 and world == 1 and not args.no_extras and args.workload == "c3_full3d_example":
        extra = dict(torch_gpu_baseline=torch_reference_gpu(sh, weights, devb[0]),
                     reference_gpu_modules=reference_modules_gpu(sh, weights, devb[0]),
                     c2_micro=micro_reproject_tail(S.MICRO, WORKLOADS["c2_micro"]["batch"], pk))
        try:                                                        # BASELINE.json configs[4]: 16 cameras, 96^3 grid, bf16 (this GPU's share)
            st = S.STRESS
            Bs = WORKLOADS["c5_stress"]["batch"]
            bs, _, _ = make_inputs(st, Bs, 2, seed=3, distinct=2)
            nets = HybridNet3D(st.K, st.bbox, st.roi, st.spacing, S.make_v2v_weights(st.K, 0, "he"), precision=precision).cuda()
            dbs = [[to_t(a).cuda() for a in b] for b in bs]
            for i in range(3):
                nets(*dbs[i % 2])
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(6):
                nets(*dbs[i % 2])
            s1.record(); torch.cuda.synchronize()
            msx = s0.elapsed_time(s1) / 6
            extra["c5_stress"] = dict(frame_sets_per_sec=Bs / (msx * 1e-3), ms_per_step=msx,
                                      config=dict(ncam=st.ncam, K=st.K, heatmap=st.hm, grid=st.G, frame_sets_per_step=Bs, dtype=precision),
                                      v2v_algorithmic_TFLOP_per_step=Bs * v2v_flops(st)["total"] / 1e12,
                                      note="one GPU's share of BASELINE.json configs[4]; `bench.py --gpus 8 --workload c5_stress` is the 8-GPU run")
            del nets, dbs
        except RuntimeError as e:
            extra["c5_stress"] = dict(error=str(e)[:200])
    if rank == 0:
        cb = None if args.no_cpu_baseline or world > 1 else cpu_baseline(sh)
        line = dict(metric="frame_sets_per_sec", value=value, unit="frame-sets/s", n_gpus=world, steps=K_steps, warmup=W,
                    ms_per_step=ms / K_steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="bf16" if precision == "bf16" else "f32", data="synthetic",
                    config=dict(workload=args.workload, ncam=sh.ncam, K=sh.K, heatmap=sh.hm, grid=sh.G,
                                frame_sets_per_step_per_gpu=B, frame_sets_per_internal_pass=min(sub_batch, B),
                                weights="random-init he (seed 0)",
                                l2="inputs rotate through %d batches of %.0f MB (>> 126 MB L2)" % (n_pool, B * sh.ncam * sh.K * sh.hm ** 2 * 4 / 1e6),
                                timed_region="heat maps -> key points (stages a2-a10); 2D CNN, decode, CSV excluded",
                                roi_share_of_heatmap=roi),
                    clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cb,
                    stages=stages, kernels=kern, rooflines=rooflines, gather_ms=gather_ms, latency_b1=latency, job=job,
                    value_f16cl_input=value_cl, extra=extra)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
