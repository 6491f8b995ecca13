"""bf16 tensor-core path (tcgen05 implicit GEMM): every convolution of V2VNet on its own against torch's
fp32 convolution of the same bf16-rounded operands, then the whole network and the whole hot path against
the CPU oracle at the bf16 bars (volume 2e-2, key points 0.5 mm)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT, V2V_CASES, case_weights, load_case
from test_gpu_parity import dev, repro_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def make_net(K, weights, precision="bf16"):
    from jarvis_hybridnet_b200 import V2VNet
    net = V2VNet(K, K, precision=precision)
    net.load_state_dict({k: torch.as_tensor(v) for k, v in weights.items()}, strict=True)
    return net.to(DEV)


LAYER_CASES = [(K, h, B) for K, h, B in [(23, 12, 1), (5, 12, 2), (23, 36, 1), (23, 20, 3)]]


@pytest.mark.parametrize("K,h,B", LAYER_CASES)
@pytest.mark.parametrize("layer", range(12))
def test_tc_layer_vs_torch(layer, K, h, B):
    import jarvis_hybridnet_b200.synth as S
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    name, kind, cim, com, k = S.V2V_LAYERS[layer]
    q = h // 2
    w = S.make_v2v_weights(K, 11, "he")
    net = make_net(K, w)
    g = torch.Generator(device="cpu").manual_seed(100 + layer)
    on_q = name.startswith("encoder_decoder.mid_res") or kind == "convT"
    if layer == 0:
        D, Din = h, 2 * h
    elif name.startswith("encoder_decoder.encoder_pool1"):
        D, Din = q, h
    elif on_q:
        D, Din = q, q
    else:
        D, Din = h, h
    x = bf16_round(torch.randn((B, cim * K, Din, Din, Din), generator=g)).to(DEV)
    got = net.debug_layer(layer, x, D)
    wt = bf16_round(torch.as_tensor(w[name + ".weight"])).to(DEV)
    bs = torch.as_tensor(w[name + ".bias"]).to(DEV)
    if kind == "convT":
        want = F.conv_transpose3d(x, wt, bs, stride=2)
    else:
        stride = 2 if (layer == 0 or "encoder_pool1" in name) else 1
        want = F.conv3d(x, wt, bs, stride=stride, padding=(k - 1) // 2)
    assert got.shape == want.shape
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    tol = 1e-5 if layer == 11 else 6e-3          # head writes fp32; the others store bf16 (2^-8 relative)
    assert err <= tol * scale + 1e-6, f"layer {layer} {name}: max err {err:.4g} vs scale {scale:.4g}"


def test_forward_host_pipelined_equals_forward():
    """Chunked H2D/compute overlap returns what the resident call returns; repeated calls (zero borders cached in
    the persistent workspace, different batch sizes through the same handle) agree."""
    import jarvis_hybridnet_b200.synth as S
    from jarvis_hybridnet_b200 import HybridNet3D
    sh = S.SMALL
    cam, intr, dist = S.make_rig(sh.ncam, 3)
    sets = [S.make_frameset(sh, cam, intr, dist, s) for s in range(5)]
    net = HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, S.make_v2v_weights(sh.K, 1, "he"), precision="bf16").to(DEV)
    rep = lambda a: np.broadcast_to(a[None], (5,) + a.shape).copy()
    host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in
            (np.stack([s[0] for s in sets]), np.stack([s[1] for s in sets]), np.stack([s[2] for s in sets]),
             rep(cam), rep(intr), rep(dist))]
    devt = [t.to(DEV) for t in host]
    pts, conf, _ = net(*devt)
    want = torch.cat([pts, conf[..., None]], 2).cpu()
    for chunk in (2, 5, 1, 2):
        res, h2d, d2h = net.forward_host(host, chunk=chunk)
        assert h2d == sum(t.numel() * t.element_size() for t in host) and d2h == want.numel() * 4
        # InstanceNorm statistics are accumulated with atomics, so runs agree to bf16 rounding noise (a tenth of the 0.5 mm bf16 bar), not bit for bit
        assert torch.allclose(res, want, rtol=0, atol=5e-2), (chunk, (res - want).abs().max())
    pts2, conf2, _ = net(*devt)
    assert torch.allclose(pts2, pts, rtol=0, atol=5e-2) and torch.allclose(conf2, conf, rtol=0, atol=1e-3)


@pytest.mark.parametrize("name", V2V_CASES)
def test_v2v_bf16_vs_oracle(oracle, name):
    sh, x, g = load_case(name)
    w = case_weights(name, sh.K)
    vol, _ = oracle.repro_layer_forward(oracle.pad_heatmaps(x["hm"]), x["c3"], x["chm"], x["cam"], x["intr"], x["dist"],
                                        sh.G, sh.spacing)
    xin = (vol / np.float32(255.0))[None]
    want = oracle.v2v_forward(w, xin).numpy()
    got = make_net(sh.K, w)(dev(xin)).cpu().numpy()
    scale = np.abs(want).max()
    err = np.abs(got - want)
    rel = float(err.max() / scale)
    rms = float(np.sqrt((err ** 2).mean()) / scale)
    print(f"{name}: bf16 V2V max err {rel:.4f} rms {rms:.5f} of scale {scale:.3f}")
    # tiny_clamp feeds a near-empty volume (every camera clamped out of its crop, max 2.5/255): InstanceNorm
    # then amplifies the bf16 rounding of a ~constant input, so the raw V2V output is ill-conditioned there;
    # its key points (the north-star bar) are still checked at 0.5 mm below.
    lim = (0.25, 0.04) if name == "tiny_clamp" else (6e-2, 1e-2)
    assert rel < lim[0] and rms < lim[1], (rel, rms)


@pytest.mark.parametrize("name", V2V_CASES)
def test_hybrid3d_bf16_end_to_end(oracle, name):
    from jarvis_hybridnet_b200 import HybridNet3D
    sh, x, g = load_case(name)
    net = HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, case_weights(name, sh.K), precision="bf16").to(DEV)
    pts, conf, am = net(*repro_inputs(x))
    err = np.abs(pts[0].cpu().numpy() - g["points3D"]).max()
    print(f"{name}: bf16 end-to-end max key-point error {err:.4f} mm")
    assert err < 0.5                                                   # mm, bf16 bar
    np.testing.assert_allclose(conf[0].cpu().numpy(), g["confidences"], rtol=5e-2, atol=5e-3)


def test_bf16_batch_matches_single(oracle):
    """B=3 frame sets through one bf16 call == three B=1 calls (InstanceNorm statistics stay per sample)."""
    from jarvis_hybridnet_b200 import HybridNet3D
    import jarvis_hybridnet_b200.synth as S
    sh = S.SMALL
    cam, intr, dist = S.make_rig(sh.ncam, 9)
    sets = [S.make_frameset(sh, cam, intr, dist, s) for s in range(3)]
    w = S.make_v2v_weights(sh.K, 5, "he")
    net = HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, w, precision="bf16").to(DEV)
    stack = lambda i: torch.stack([dev(s[i]) for s in sets])
    rep = lambda a: dev(a)[None].expand(3, *a.shape).contiguous()
    args = (stack(0), stack(1), stack(2), rep(cam), rep(intr), rep(dist))
    pts, conf, _ = net(*args)
    for b in range(3):
        p1, c1, _ = net(*[a[b:b + 1] for a in args])
        assert (p1[0] - pts[b]).abs().max().item() < 2e-2              # atomics reorder the fp32 statistics sums
        want = oracle.hybrid3d_forward(w, sets[b][0], sets[b][1], sets[b][2], cam, intr, dist, sh.roi, sh.spacing)
        assert np.abs(pts[b].cpu().numpy() - want["points"]).max() < 0.5


def test_forward_graph_replay_equals_forward():
    """CUDA-graph replay (the B=1 latency path) returns what the eager call returns, for changing inputs (the
    InstanceNorm statistics are summed with atomics, so runs agree to bf16 rounding noise, not bit for bit)."""
    import jarvis_hybridnet_b200.synth as S
    from jarvis_hybridnet_b200 import HybridNet3D
    sh = S.SMALL
    cam, intr, dist = S.make_rig(sh.ncam, 3)
    net = HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, S.make_v2v_weights(sh.K, 0, "he"), precision="bf16").cuda()
    d = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()[None]
    for seed in (0, 1, 2):
        hm, c3, chm, _ = S.make_frameset(sh, cam, intr, dist, seed)
        args = (d(hm), d(c3), d(chm), d(cam), d(intr), d(dist))
        want = [t.clone() for t in net(*args)]
        got = net.forward_graph(*args)
        torch.cuda.synchronize()
        assert torch.allclose(got[0], want[0], rtol=0, atol=5e-2), (seed, (got[0] - want[0]).abs().max())
        assert torch.allclose(got[1], want[1], rtol=0, atol=1e-3)
    assert len(net._graphs) == 1
