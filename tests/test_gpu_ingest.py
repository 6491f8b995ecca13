"""SURVEY.md §8 rows f4 (frame ingest), f2 (effTrack head in the gather's layout) and a11 (returned volumes) on the B200,
through the C ABI, against oracle/ingest_oracle.py and the torch statements the reference executes."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_case
from oracle import ingest_oracle as IO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def frames_u8(shape, seed=0):
    return np.random.default_rng(seed).integers(0, 256, shape, dtype=np.uint8)


@pytest.mark.parametrize("shape", [(3, 20, 24, 3), (2, 12, 1024, 1280, 3), (1, 5, 8, 3)])
def test_ingest_frames_bit_exact(shape):
    from jarvis_hybridnet_b200.ingest import ingest_frames
    fr = frames_u8(shape)
    got = ingest_frames(torch.from_numpy(fr).to(DEV))
    flat = fr.reshape((-1,) + shape[-3:])
    if flat.shape[0] * flat.shape[1] * flat.shape[2] < 1 << 20:
        assert np.array_equal(got.reshape((-1, 3) + shape[-3:-1]).cpu().numpy(), IO.ingest_frames(flat))
    # predict3D.py:79, the reference's own statement on this GPU
    want = torch.from_numpy(flat).to(DEV).float().permute(0, 3, 1, 2)[:, [2, 1, 0]] / 255.
    assert torch.equal(got.reshape(want.shape), want)
    assert got.shape == shape[:-3] + (3,) + shape[-3:-1]


def test_ingest_rejects_bad_input():
    from jarvis_hybridnet_b200.ingest import ingest_frames
    with pytest.raises(RuntimeError):
        ingest_frames(torch.zeros((2, 8, 8, 3), dtype=torch.float32, device=DEV))
    with pytest.raises(RuntimeError):
        ingest_frames(torch.zeros((2, 8, 6, 3), dtype=torch.uint8, device=DEV))          # W % 4
    with pytest.raises(RuntimeError):
        ingest_frames(torch.zeros((2, 8, 8, 3), dtype=torch.uint8))                       # host tensor: no CPU fallback


def test_crop_normalize_u8_bit_exact():
    from jarvis_hybridnet_b200.ingest import crop_normalize_u8, ingest_frames
    from jarvis_hybridnet_b200.predictor import crop_normalize
    B, ncam, H, W, bbox = 3, 4, 96, 128, 32
    fr = frames_u8((B, ncam, H, W, 3), 5)
    rng = np.random.default_rng(6)
    chm = np.stack([rng.integers(16, W - 16, (B, ncam)), rng.integers(16, H - 16, (B, ncam))], -1).astype(np.int32)
    chm[0, 0] = (16, 16); chm[0, 1] = (W - 16, H - 16)                                     # windows touching the borders
    valid = np.array([1, 0, 1], np.int32)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    d = lambda a: torch.from_numpy(a).to(DEV)
    got = crop_normalize_u8(d(fr), d(chm), d(valid), bbox, mean, std)
    assert np.array_equal(got.cpu().numpy(), IO.crop_normalize_u8(fr, chm, valid, bbox, mean, std))
    # == the fp32 route (jhn_ingest_frames -> jhn_crop_normalize), which the f1 tests pin on the reference's predictor
    via = crop_normalize(ingest_frames(d(fr)), d(chm), d(valid), bbox, mean, std)
    assert torch.equal(got, via)


@pytest.mark.parametrize("C,K,Hq,Wq", [(64, 23, 64, 64), (11, 5, 6, 7), (88, 24, 17, 33), (160, 1, 32, 32)])
def test_efftrack_head_planar(C, K, Hq, Wq):
    from jarvis_hybridnet_b200.ingest import EffTrackHead
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(C)
    x = torch.randn(3, C, Hq, Wq, generator=g).to(DEV)
    deconv = torch.nn.ConvTranspose2d(C, K, kernel_size=4, stride=2, padding=1, bias=False).to(DEV)   # efficienttrack/model.py:89-95
    head = EffTrackHead.from_deconv(deconv, "planar")
    assert list(head.state_dict().keys()) == list(deconv.state_dict().keys())
    with torch.no_grad():
        want = deconv(x)
        got = head(x)
    assert got.shape == want.shape
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() < 2e-6 * scale + 1e-6
    if C * Hq * Wq < 1 << 16:
        o = IO.efftrack_head(x.cpu().numpy(), deconv.weight.detach().cpu().numpy())
        assert np.abs(got.cpu().numpy() - o).max() < 2e-6 * scale + 1e-6


@pytest.mark.parametrize("fmt", ["f16_cl", "bf16_cl"])
def test_efftrack_head_channels_last(fmt):
    """The head's channels-last output == jhn_heatmap_convert of its own planar output (same rounding, same border, same
    channel padding), so the gather reads identical bytes with and without the staging pass."""
    from jarvis_hybridnet_b200 import _lib
    from jarvis_hybridnet_b200.ingest import EffTrackHead
    C, K, Hq = 64, 23, 64
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(12, C, Hq, Hq, generator=g) * 3).to(DEV)
    w = (torch.randn(C, K, 4, 4, generator=g) * 0.4).to(DEV)
    planar, cl = EffTrackHead(C, K, "planar").to(DEV), EffTrackHead(C, K, fmt).to(DEV)
    planar.weight.data.copy_(w); cl.weight.data.copy_(w)
    p = planar(x)                                                                         # [12,K,128,128]
    c = cl(x)                                                                             # [12,130,130,24]
    hs = 2 * Hq + 2
    assert tuple(c.shape) == (12, hs, hs, 24) and c.dtype == (torch.float16 if fmt == "f16_cl" else torch.bfloat16)
    want = _lib.heatmap_convert(p[None], hs, _lib.HM_F16_CL if fmt == "f16_cl" else _lib.HM_BF16_CL)[0]
    assert torch.equal(c.view(torch.int16), want.view(torch.int16))
    o = IO.to_channels_last(p[:2].cpu().numpy(), bf16=(fmt == "bf16_cl"))
    assert np.array_equal(c[:2].float().cpu().numpy(), o)


def test_head_to_gather_without_staging():
    """Row f2 end to end: features -> EffTrackHead(f16_cl) -> HybridNet3D == features -> reference layer (planar fp32) ->
    HybridNet3D, bit for bit (the fp32-planar entry converts to the same fp16 values internally)."""
    import jarvis_hybridnet_b200.synth as S
    from jarvis_hybridnet_b200 import HybridNet3D
    from jarvis_hybridnet_b200.ingest import EffTrackHead
    sh, x, _ = load_case("example_he")
    d = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(DEV)
    # features whose transposed convolution has heat-map-like magnitudes: features = the map at half resolution, weights
    # = a bilinear-ish kernel on the diagonal plus random cross-talk
    hm = d(x["hm"])                                                                        # [12,23,128,128]
    C = 64
    g = torch.Generator().manual_seed(2)
    feats = torch.zeros(sh.ncam, C, 64, 64, device=DEV)
    feats[:, :sh.K] = F.avg_pool2d(hm, 2)
    w = torch.randn(C, sh.K, 4, 4, generator=g).to(DEV) * 0.02
    k1 = torch.tensor([0.25, 0.75, 0.75, 0.25], device=DEV)
    for k in range(sh.K):
        w[k, k] += torch.outer(k1, k1)
    net = HybridNet3D(sh.K, sh.bbox, sh.roi, sh.spacing, state_dict=S.make_v2v_weights(sh.K, 4, "he"), precision="bf16").to(DEV)
    args = [d(x["c3"])[None], d(x["chm"])[None], d(x["cam"])[None], d(x["intr"])[None], d(x["dist"])[None]]
    planar, cl = EffTrackHead(C, sh.K, "planar").to(DEV), EffTrackHead(C, sh.K, "f16_cl").to(DEV)
    planar.weight.data.copy_(w); cl.weight.data.copy_(w)
    p0, c0, a0 = net(planar(feats)[None], *args)
    p0, c0, a0 = p0.clone(), c0.clone(), a0.clone()
    p1, c1, a1 = net(cl(feats)[None], *args)
    assert torch.equal(p0, p1) and torch.equal(c0, c1) and torch.equal(a0, a1)
    assert torch.isfinite(p1).all()


def test_returned_volumes():
    from jarvis_hybridnet_b200.ingest import pad_heatmaps, softplus2
    g = torch.Generator().manual_seed(3)
    hm = torch.randn(2, 5, 7, 33, 33, generator=g).to(DEV)
    assert torch.equal(pad_heatmaps(hm), F.pad(hm, [1, 1, 1, 1]))                         # hybridnet/model.py:65-66
    v = torch.cat([torch.randn(100003, generator=g) * 8, torch.tensor([19.9, 20.0, 20.1, 60., -60., 0.])]).to(DEV)
    want = F.softplus(F.softplus(v))                                                       # model.py:73,88
    got = softplus2(v)
    assert torch.allclose(got, want, rtol=3e-7, atol=1e-30)
    assert np.allclose(got.cpu().numpy(), IO.softplus2(v.cpu().numpy()), rtol=3e-7, atol=1e-30)


def test_frame_uploader_round_trip():
    from jarvis_hybridnet_b200.ingest import FrameUploader
    up = FrameUploader((2, 3, 16, 16, 3))
    for i in range(5):
        fr = frames_u8((2, 3, 16, 16, 3), i)
        up.host(i)[...] = fr
        dev, ev = up.upload(i)
        torch.cuda.current_stream().wait_event(ev)
        got = dev.clone()
        up.release(i)
        ev.synchronize()
        assert np.array_equal(got.cpu().numpy(), fr)
    assert up.bytes_per_upload == 2 * 3 * 16 * 16 * 3


@pytest.mark.parametrize("hs,source", [(40, "host"), (256, "device"), (300, "host")])
def test_pull_heatmap_spans_moves_exactly_the_row_spans(hs, source):
    """jhn_pull_heatmap_spans (flat kernel for maps of <= 256 rows, warp-per-row kernel above) and jhn_pull_small:
    arbitrary spans, rows without a span, a pinned host or a device tensor as the source; everything else stays untouched."""
    import ctypes
    from jarvis_hybridnet_b200 import _lib
    lib = _lib.load()
    n_img = 5
    rng = np.random.default_rng(hs)
    src = torch.from_numpy(rng.integers(1, 2 ** 15, size=(n_img, hs, hs, 24), dtype=np.int16))
    src = src.pin_memory() if source == "host" else src.to(DEV)
    dst = torch.full(src.shape, -7, dtype=torch.int16, device=DEV)
    a, b = rng.integers(0, hs, size=(n_img, hs)), rng.integers(0, hs, size=(n_img, hs))
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    empty = rng.random((n_img, hs)) < 0.3
    spans = np.stack([np.where(empty, 0x7f7f7f7f, lo), np.where(empty, 0x7f7f7f7f, -hi)], -1).astype(np.int32)
    counter = torch.zeros(1, dtype=torch.int64, device=DEV)
    for shape in ((128, 48, 8), (32, 3, 5)):
        lib.jhn_debug_set_pull_config(*shape)
        dst.fill_(-7); counter.zero_()
        _lib.check(lib.jhn_pull_heatmap_spans(ctypes.c_void_p(src.data_ptr()), _lib.dptr(dst), _lib.dptr(torch.from_numpy(spans).to(DEV)),
                                              n_img, hs, 48, _lib.dptr(counter), _lib.stream_ptr()))
        want = np.full(src.shape, -7, np.int16)
        s = src.cpu().numpy()
        for i in range(n_img):
            for y in range(hs):
                if not empty[i, y]:
                    want[i, y, lo[i, y]:hi[i, y] + 1] = s[i, y, lo[i, y]:hi[i, y] + 1]
        assert np.array_equal(dst.cpu().numpy(), want)
        assert int(counter.item()) == int(((hi - lo + 1) * ~empty).sum()) * 48
    lib.jhn_debug_set_pull_config(128, 48, 8)
    # jhn_pull_small: a few small pinned tensors by one kernel
    hts = [torch.from_numpy(rng.standard_normal(n).astype(np.float32)).pin_memory() for n in (3, 96, 1, 4321)]
    dts = [torch.zeros_like(t, device=DEV) for t in hts]
    n = len(hts)
    _lib.check(lib.jhn_pull_small(n, (ctypes.c_void_p * n)(*[t.data_ptr() for t in hts]), (ctypes.c_void_p * n)(*[t.data_ptr() for t in dts]),
                                  (ctypes.c_size_t * n)(*[t.numel() * 4 for t in hts]), _lib.stream_ptr()))
    for h, d in zip(hts, dts):
        assert torch.equal(d.cpu(), h)
    not_pinned = torch.zeros(4)
    rc = lib.jhn_pull_small(1, (ctypes.c_void_p * 1)(not_pinned.data_ptr()), (ctypes.c_void_p * 1)(dts[0].data_ptr()), (ctypes.c_size_t * 1)(12), _lib.stream_ptr())
    assert rc != 0 and b"not pinned" in lib.jhn_last_error()
