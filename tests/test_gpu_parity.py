"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed golden fixtures.

Bars (BASELINE.json north_star): voxel->pixel indices and argmax voxel bit-exact; fp32 feature volume
rtol 1e-4 (bf16 staging 2e-2); 3D key points within 0.05 mm at fp32 (0.5 mm at bf16)."""
import numpy as np
import pytest
import torch

from conftest import ALL_CASES, V2V_CASES, case_weights, decisive_argmax, load_case, oracle_forward, sha
from test_host import cfg_of

pytestmark = pytest.mark.gpu
DEV = "cuda"


def dev(a, dt=None):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(DEV)


def repro_inputs(x, padded=False, oracle=None):
    hm = oracle.pad_heatmaps(x["hm"]) if padded else x["hm"]
    return (dev(hm)[None], dev(x["c3"])[None], dev(x["chm"])[None], dev(x["cam"])[None], dev(x["intr"])[None],
            dev(x["dist"])[None])


@pytest.mark.parametrize("name", ALL_CASES)
def test_indices_bit_exact(oracle, name):
    from jarvis_hybridnet_b200 import ReprojectionLayer
    sh, x, g = load_case(name)
    L = ReprojectionLayer(cfg_of(sh))
    _, idx = L.forward_batched(*repro_inputs(x), want_index=True)
    idx = idx[0].cpu().numpy()
    assert sha(idx.astype(np.int32)) == str(g["idx_sha"]), "differs from the reference's own indices"
    want = oracle.reproject_indices(x["c3"], x["chm"], x["cam"], x["intr"], x["dist"], sh.G, sh.spacing, sh.hs)
    assert np.array_equal(idx, want)


@pytest.mark.parametrize("mode", [1, 2])
def test_indices_other_lerp_modes(oracle, mode):
    from jarvis_hybridnet_b200 import ReprojectionLayer
    sh, x, g = load_case("example_mh")
    L = ReprojectionLayer(cfg_of(sh), lerp_mode=mode)
    _, idx = L.forward_batched(*repro_inputs(x), want_index=True)
    want = oracle.reproject_indices(x["c3"], x["chm"], x["cam"], x["intr"], x["dist"], sh.G, sh.spacing, sh.hs,
                                    lerp_mode=mode)
    assert np.array_equal(idx[0].cpu().numpy(), want)


def test_reprojectPoints_reference_signature(oracle):
    from jarvis_hybridnet_b200 import ReprojectionLayer
    sh, x, g = load_case("tiny_s0")
    L = ReprojectionLayer(cfg_of(sh))
    res = L.reprojectPoints(L.grid + dev(x["c3"]), dev(x["cam"]), dev(x["intr"]), dev(x["dist"]), dev(x["chm"]))
    assert res.dtype == torch.int64 and tuple(res.shape) == (sh.ncam, sh.G, sh.G, sh.G)
    assert np.array_equal(res.cpu().numpy(), g["idx"])


@pytest.mark.parametrize("padded", [False, True])
@pytest.mark.parametrize("name", ALL_CASES)
def test_volume_fp32(oracle, name, padded):
    from jarvis_hybridnet_b200 import ReprojectionLayer
    sh, x, g = load_case(name)
    if padded and name in ("micro_idx", "stress_idx"):
        pytest.skip("padded variant covered on the smaller shapes")
    L = ReprojectionLayer(cfg_of(sh))
    vol = L(*repro_inputs(x, padded, oracle))            # reference-style forward: [1,K,G,G,G]
    assert tuple(vol.shape) == (1, sh.K, sh.G, sh.G, sh.G)
    want, _ = oracle.repro_layer_forward(oracle.pad_heatmaps(x["hm"]), x["c3"], x["chm"], x["cam"], x["intr"],
                                         x["dist"], sh.G, sh.spacing)
    np.testing.assert_allclose(vol[0].cpu().numpy(), want, rtol=1e-4, atol=1e-4)
    ref = g["volume"] if "volume" in g else None
    if ref is not None:
        np.testing.assert_allclose(vol[0].cpu().numpy(), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["tiny_s0", "small_mh", "example_mh"])
def test_volume_bf16_staging(oracle, name):
    from jarvis_hybridnet_b200 import ReprojectionLayer
    sh, x, g = load_case(name)
    L = ReprojectionLayer(cfg_of(sh), precision="bf16")
    vol, idx = L.forward_batched(*repro_inputs(x), post_divide=255.0, want_index=True)
    want, widx = oracle.repro_layer_forward(oracle.pad_heatmaps(x["hm"]), x["c3"], x["chm"], x["cam"], x["intr"],
                                            x["dist"], sh.G, sh.spacing)
    assert np.array_equal(idx[0].cpu().numpy(), widx)              # indices stay exact in bf16 mode
    np.testing.assert_allclose(vol[0].cpu().numpy(), want / 255.0, rtol=2e-2, atol=2e-2 * 4 / 255)


@pytest.mark.parametrize("name", ["tiny_s0", "tiny_clamp", "small_mh", "example_mh", "example_he"])
def test_volume_bf16_staged_gather(oracle, name):
    """Throughput gather (no index dump): fp16 staging copy, TMA-staged pixel boxes, fp16 camera sum."""
    from jarvis_hybridnet_b200 import ReprojectionLayer
    sh, x, g = load_case(name)
    L = ReprojectionLayer(cfg_of(sh), precision="bf16")
    vol, idx = L.forward_batched(*repro_inputs(x), post_divide=255.0, want_index=False)
    assert idx is None
    want, _ = oracle.repro_layer_forward(oracle.pad_heatmaps(x["hm"]), x["c3"], x["chm"], x["cam"], x["intr"],
                                         x["dist"], sh.G, sh.spacing)
    got = vol[0].cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, want / 255.0, rtol=2e-2, atol=2e-2 * 4 / 255)
    # fp16 staging (11-bit mantissa) + fp16 camera sum: typically 10x tighter than the bf16 bar
    err = np.abs(got - want / 255.0)
    assert np.sqrt((err ** 2).mean()) <= 2e-3 * np.abs(want / 255.0).max()


@pytest.mark.parametrize("name", ["micro_idx", "stress_idx"])
def test_streaming_gather_large_shapes(oracle, name):
    """BASELINE configs 2 and 5 (12 x 256^2 maps on a 64^3 grid; 16 cameras on a 96^3 grid) through the streaming
    gather, against the oracle's fp32 gather + camera mean."""
    from jarvis_hybridnet_b200 import ReprojectionLayer
    sh, x, g = load_case(name)
    L = ReprojectionLayer(cfg_of(sh), precision="bf16")
    vol, _ = L.forward_batched(*repro_inputs(x), post_divide=255.0, want_index=False)
    want, _ = oracle.repro_layer_forward(oracle.pad_heatmaps(x["hm"]), x["c3"], x["chm"], x["cam"], x["intr"],
                                         x["dist"], sh.G, sh.spacing)
    got = vol[0].cpu().numpy()
    np.testing.assert_allclose(got, want / 255.0, rtol=2e-2, atol=2e-2 * 4 / 255)
    err = np.abs(got - want / 255.0)
    assert np.sqrt((err ** 2).mean()) <= 2e-3 * np.abs(want / 255.0).max()


@pytest.mark.parametrize("name", ["tiny_clamp", "small_mh", "example_mh"])
def test_streaming_gather_global_path_is_identical(name):
    """Pixel boxes that do not fit a shared-memory slot are gathered from global memory with the same arithmetic:
    shrinking the slot (test hook) must not change a single bit of the volume."""
    from jarvis_hybridnet_b200 import ReprojectionLayer, _lib
    sh, x, g = load_case(name)
    L = ReprojectionLayer(cfg_of(sh), precision="bf16")
    ref, _ = L.forward_batched(*repro_inputs(x), post_divide=255.0, want_index=False)
    try:
        for limit in (48, 1536, 4096):            # nothing fits / a few boxes fit / most boxes fit
            assert _lib.debug_set_gather_box_bytes(limit) == limit
            got, _ = L.forward_batched(*repro_inputs(x), post_divide=255.0, want_index=False)
            assert torch.equal(got, ref), limit
    finally:
        _lib.debug_set_gather_box_bytes(0)


def test_streaming_gather_batched_equals_single():
    """Persistent CTAs walk tiles of all frame sets in one sequence; the result per frame set must not depend on it
    (B = 5 is not a multiple of anything in the tile schedule)."""
    from jarvis_hybridnet_b200 import ReprojectionLayer
    import jarvis_hybridnet_b200.synth as S
    sh = S.SMALL
    cam, intr, dist = S.make_rig(sh.ncam, 11)
    sets = [S.make_frameset(sh, cam, intr, dist, s) for s in range(5)]
    L = ReprojectionLayer(cfg_of(sh), precision="bf16")
    stack = lambda i, dt=None: torch.stack([dev(s[i], dt) for s in sets])
    rep = lambda a: dev(a)[None].expand(5, *a.shape).contiguous()
    args = (stack(0), stack(1), stack(2), rep(cam), rep(intr), rep(dist))
    vol, _ = L.forward_batched(*args, post_divide=255.0, want_index=False)
    for b in range(5):
        one, _ = L.forward_batched(*[a[b:b + 1] for a in args], post_divide=255.0, want_index=False)
        assert torch.equal(one[0], vol[b]), b


@pytest.mark.parametrize("ncam", [1, 3, 5])
def test_streaming_gather_odd_camera_counts(oracle, ncam):
    """Camera counts below / not a multiple of the number of producer warps (items = tile x camera are dealt round
    robin to four producers): no producer may be left waiting, results as the oracle's."""
    from jarvis_hybridnet_b200 import ReprojectionLayer
    import jarvis_hybridnet_b200.synth as S
    sh = S.Shape3D(ncam, 5, 64, 48, 2)
    cam, intr, dist = S.make_rig(ncam, 21)
    hm, c3, chm, _ = S.make_frameset(sh, cam, intr, dist, 4)
    L = ReprojectionLayer(cfg_of(sh), precision="bf16")
    x = dict(hm=hm, c3=c3, chm=chm, cam=cam, intr=intr, dist=dist)
    vol, _ = L.forward_batched(*repro_inputs(x), post_divide=255.0, want_index=False)
    want, _ = oracle.repro_layer_forward(oracle.pad_heatmaps(hm), c3, chm, cam, intr, dist, sh.G, sh.spacing)
    np.testing.assert_allclose(vol[0].cpu().numpy(), want / 255.0, rtol=2e-2, atol=2e-2 * 4 / 255)


@pytest.mark.parametrize("seed", range(8))
def test_streaming_gather_random_shapes(oracle, seed):
    """Random small shapes (camera count, key points, map size, grid side, spacing, frame sets per call) and crop
    centres shifted so that part of the grid clamps at the map border: streaming gather vs the oracle's fp32 gather."""
    from jarvis_hybridnet_b200 import ReprojectionLayer
    import jarvis_hybridnet_b200.synth as S
    rng = np.random.default_rng(100 + seed)
    ncam, K = int(rng.integers(1, 9)), int(rng.integers(1, 25))
    bbox = int(rng.choice([32, 64, 96, 128]))
    spacing = int(rng.choice([1, 2, 3]))
    G = int(rng.choice([8, 16, 24, 32, 40]))
    B = int(rng.integers(1, 4))
    sh = S.Shape3D(ncam, K, bbox, G * spacing, spacing)
    cam, intr, dist = S.make_rig(ncam, 50 + seed)
    sets = [S.make_frameset(sh, cam, intr, dist, 10 * seed + b) for b in range(B)]
    shift = rng.integers(-bbox // 3, bbox // 3 + 1, (B, ncam, 2)).astype(np.int32) if seed % 2 else np.zeros((B, ncam, 2), np.int32)
    chm = np.stack([s[2] for s in sets]) + shift
    L = ReprojectionLayer(cfg_of(sh), precision="bf16")
    rep = lambda a: dev(a)[None].expand(B, *a.shape).contiguous()
    vol, _ = L.forward_batched(torch.stack([dev(s[0]) for s in sets]), torch.stack([dev(s[1]) for s in sets]), dev(chm),
                               rep(cam), rep(intr), rep(dist), post_divide=255.0, want_index=False)
    for b in range(B):
        want, _ = oracle.repro_layer_forward(oracle.pad_heatmaps(sets[b][0]), sets[b][1], chm[b], cam, intr, dist, sh.G, sh.spacing)
        np.testing.assert_allclose(vol[b].cpu().numpy(), want / 255.0, rtol=2e-2, atol=2e-2 * 4 / 255, err_msg=str((seed, b, sh)))


def test_batched_equals_single(oracle):
    """B frame sets in one launch == B reference-style B=1 forwards."""
    from jarvis_hybridnet_b200 import ReprojectionLayer
    import jarvis_hybridnet_b200.synth as S
    sh = S.SMALL
    cam, intr, dist = S.make_rig(sh.ncam, 9)
    sets = [S.make_frameset(sh, cam, intr, dist, s) for s in range(3)]
    L = ReprojectionLayer(cfg_of(sh))
    stack = lambda i, dt=None: torch.stack([dev(s[i], dt) for s in sets])
    rep = lambda a: dev(a)[None].expand(3, *a.shape).contiguous()
    args = (stack(0), stack(1), stack(2), rep(cam), rep(intr), rep(dist))
    vol, idx = L.forward_batched(*args, want_index=True)
    for b in range(3):
        one, oidx = L.forward_batched(*[a[b:b + 1] for a in args], want_index=True)
        assert torch.equal(one[0], vol[b]) and torch.equal(oidx[0], idx[b])
        want = oracle.reproject_indices(sets[b][1], sets[b][2], cam, intr, dist, sh.G, sh.spacing, sh.hs)
        assert np.array_equal(idx[b].cpu().numpy(), want)


@pytest.mark.parametrize("name", V2V_CASES)
def test_v2v_fp32(oracle, name):
    from jarvis_hybridnet_b200 import V2VNet
    sh, x, g = load_case(name)
    w = case_weights(name, sh.K)
    vol, _ = oracle.repro_layer_forward(oracle.pad_heatmaps(x["hm"]), x["c3"], x["chm"], x["cam"], x["intr"],
                                        x["dist"], sh.G, sh.spacing)
    xin = (vol / np.float32(255.0))[None]
    want = oracle_forward(name)["v2v"][None]
    net = V2VNet(sh.K, sh.K, precision="fp32")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    net = net.to(DEV)
    got = net(dev(xin)).cpu().numpy()
    assert got.shape == want.shape == (1, sh.K, sh.h, sh.h, sh.h)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 2e-4 * scale + 1e-6, np.abs(got - want).max() / scale
    ref = g["v2v"]
    sub = got[0].reshape(-1)[::int(g["v2v_stride"])].reshape(ref.shape)
    assert np.abs(sub - ref).max() <= 2e-4 * scale + 1e-6


@pytest.mark.parametrize("name", V2V_CASES)
def test_tail(oracle, name):
    from jarvis_hybridnet_b200 import centroid_tail
    sh, x, g = load_case(name)
    out = oracle_forward(name)
    pts, conf, am = centroid_tail(dev(out["v2v"])[None], sh.spacing, sh.roi, dev(x["c3"])[None], want_argmax=True)
    assert np.array_equal(am[0].cpu().numpy(), out["argmax"])                  # bit-exact argmax voxel
    assert np.abs(pts[0].cpu().numpy() - out["points"]).max() < 0.05           # mm
    np.testing.assert_allclose(conf[0].cpu().numpy(), out["conf"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", V2V_CASES)
def test_hybrid3d_fp32_end_to_end(oracle, name):
    """Heat maps in, key points out through jhn_hybrid3d_forward, against the reference's own key points."""
    from jarvis_hybridnet_b200 import HybridNet3D
    sh, x, g = load_case(name)
    net = HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, case_weights(name, sh.K), precision="fp32").to(DEV)
    pts, conf, am = net(*repro_inputs(x))
    assert np.abs(pts[0].cpu().numpy() - g["points3D"]).max() < 0.05           # mm, fp32 bar
    np.testing.assert_allclose(conf[0].cpu().numpy(), g["confidences"], rtol=1e-4, atol=1e-5)
    # argmax voxel: bit-exact wherever the maximum is decisive (top-2 gap above fp32 re-association noise of the network)
    dec = decisive_argmax(oracle_forward(name)["v2v"])
    assert dec.mean() >= 0.5, "fixture has too few decisive maxima to test the argmax"
    assert np.array_equal(am[0].cpu().numpy()[dec], g["argmax"][dec])


def test_accelerate_seam():
    """accelerate() on an object shaped like HybridNetBackbone (model.py:20-50) keeps the forward contract."""
    import torch.nn as nn
    from jarvis_hybridnet_b200 import V2VNet, accelerate
    sh, x, g = load_case("tiny_s0")

    class FakeTrack(nn.Module):
        def forward(self, imgs):
            return None, dev(x["hm"])

    bb = nn.Module()
    bb.cfg = cfg_of(sh)
    bb.grid_spacing, bb.grid_size = torch.tensor(sh.spacing), torch.tensor(sh.roi)
    bb.effTrack = FakeTrack()
    bb.reproLayer = nn.Module()
    bb.v2vNet = V2VNet(sh.K, sh.K)
    bb.v2vNet.load_state_dict({k: torch.from_numpy(v) for k, v in case_weights("tiny_s0", sh.K).items()})
    bb = accelerate(bb.to(DEV), precision="fp32", return_volumes=True)
    hf, hp, p3, conf = bb(torch.zeros(1, sh.ncam, 3, 4, 4, device=DEV), torch.tensor([1280, 1024], device=DEV),
                          dev(x["chm"])[None], dev(x["c3"])[None], dev(x["cam"])[None], dev(x["intr"])[None],
                          dev(x["dist"])[None])
    assert tuple(p3.shape) == (1, sh.K, 3) and tuple(conf.shape) == (1, sh.K)
    assert tuple(hp.shape) == (1, sh.ncam, sh.K, sh.hs, sh.hs) and tuple(hf.shape) == (1, sh.K, sh.h, sh.h, sh.h)
    assert np.abs(p3[0].cpu().numpy() - g["points3D"]).max() < 0.05
    assert abs(hf.double().sum().item() - float(g["hf_sum"])) <= 1e-4 * abs(float(g["hf_sum"]))


def test_errors_are_loud():
    from jarvis_hybridnet_b200 import ReprojectionLayer
    import jarvis_hybridnet_b200.synth as S
    sh = S.TINY
    L = ReprojectionLayer(cfg_of(sh))
    z = lambda *s: torch.zeros(*s, device=DEV)
    with pytest.raises(RuntimeError, match="expected"):
        L(z(1, sh.ncam, sh.K, 20, 20), z(1, 3), z(1, sh.ncam, 2), z(1, sh.ncam, 4, 3), z(1, sh.ncam, 3, 3), z(1, sh.ncam, 1, 5))
    bad = ReprojectionLayer(cfg_of(S.Shape3D(4, 5, 64, 44, 2)))            # G = 22, not a multiple of 4
    with pytest.raises(RuntimeError, match="multiple of 4"):
        bad(z(1, 4, 5, 34, 34), z(1, 3), z(1, 4, 2), z(1, 4, 4, 3), z(1, 4, 3, 3), z(1, 4, 1, 5))
