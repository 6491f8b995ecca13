"""GPU parity of the predictor glue through the C ABI: jhn_center_locate / jhn_crop_normalize
against the fixtures of the reference's JarvisPredictor3D.forward and against the CPU oracle."""
import numpy as np
import pytest
import torch

from test_center_oracle import CENTER_CASES, MEAN, STD, center_case

pytestmark = pytest.mark.gpu
DEV = "cuda"
d = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("name", CENTER_CASES)
def test_center_locate_matches_reference(name):
    from jarvis_hybridnet_b200 import crop_normalize, locate_center
    from oracle import center_oracle as C
    x, g = center_case(name)
    r = locate_center(d(x["hm"]), (x["W"], x["H"]), x["cdis"], x["bbox"] // 2, d(x["cam"]), d(x["intr"]), d(x["dist"]))
    want = C.locate_center(x["hm"], x["W"], x["H"], x["cdis"], x["bbox"] // 2, x["cam"], x["intr"], x["dist"])
    assert int(r["valid"][0]) == int(bool(g["valid"]))
    assert np.array_equal(r["preds"][0].cpu().numpy(), want["preds"])                  # argmax: bit-exact
    assert np.array_equal(r["maxvals"][0].cpu().numpy(), want["maxvals"])
    crops = crop_normalize(d(x["imgs"]), r["centerHM"], r["valid"], x["bbox"], MEAN, STD)[0].cpu().numpy()
    if not bool(g["valid"]):
        assert not crops.any()
        return
    # fp64 Jacobi on the normal matrix vs LAPACK's fp32 SVD: the centre (~100 mm) agrees to a few 1e-4 mm
    assert np.abs(r["center3D"][0].cpu().numpy() - g["center3D"]).max() < 2e-3
    assert np.array_equal(r["center3D_int"][0].cpu().numpy(), g["center3D_int"])
    assert np.array_equal(r["centerHM"][0].cpu().numpy(), g["centerHM"])
    assert np.array_equal(crops.reshape(-1)[::997], g["crops_sample"])                 # bit-exact
    s = np.array([crops.astype(np.float64).sum(), (crops.astype(np.float64) ** 2).sum()])
    np.testing.assert_allclose(s, g["crops_sum"], rtol=1e-12)


def test_center_locate_batched_and_relaunch():
    """B frame sets in one launch == B single launches; the ticket scratch is left zeroed, so the same scratch
    serves consecutive launches (as the accelerated predictor uses it)."""
    from jarvis_hybridnet_b200 import locate_center
    import jarvis_hybridnet_b200.synth as S
    ncam = 12
    cam, intr, dist = S.make_rig(ncam, 7)
    cases = [S.make_center_case(ncam, cam, intr, dist, s, 256, n_weak=(11 if s == 2 else 0)) for s in range(4)]
    hm = d(np.stack([c[0][:, 0] for c in cases]))
    rep = lambda a: d(a)[None].expand(4, *a.shape).contiguous()
    scratch = torch.zeros(8, dtype=torch.int32, device=DEV)
    for _ in range(2):
        r = locate_center(hm, (S.IMG_W, S.IMG_H), 256, 128, rep(cam), rep(intr), rep(dist), scratch=scratch)
        torch.cuda.synchronize()
        assert not scratch.any()
    assert r["valid"].tolist() == [1, 1, 0, 1]
    for b in range(4):
        one = locate_center(hm[b:b + 1], (S.IMG_W, S.IMG_H), 256, 128, rep(cam)[:1], rep(intr)[:1], rep(dist)[:1])
        for k in ("preds", "maxvals", "center3D", "center3D_int", "centerHM", "valid"):
            assert torch.equal(one[k][0], r[k][b]), (b, k)
        if b != 2:
            assert np.abs(r["center3D"][b].cpu().numpy() - cases[b][2]).max() < 12.0   # heat-map pixel = 10 px: cm-level


def test_accelerate_predictor_seam():
    """accelerate_predictor on a stand-in for a loaded JarvisPredictor3D (CNN stubs): same call, same return contract."""
    import torch.nn as nn
    from jarvis_hybridnet_b200 import accelerate_predictor
    from jarvis_hybridnet_b200.predictor import _accelerated_predict  # noqa: F401
    x, g = center_case("c12_s0")
    seen = {}

    class Detect(nn.Module):
        def forward(self, t):
            return None, d(x["hm"])

    class Hybrid(nn.Module):
        def forward(self, crops, img_size, centerHM, center3D, cam, intr, dist):
            seen.update(crops=crops, centerHM=centerHM, center3D=center3D)
            return None, None, torch.ones(1, 23, 3, device=DEV), torch.ones(1, 23, device=DEV)

    p = nn.Module()
    p.centerDetect, p.hybridNet = Detect(), Hybrid()
    p.transform_mean = torch.tensor(MEAN, device=DEV).view(3, 1, 1)
    p.transform_std = torch.tensor(STD, device=DEV).view(3, 1, 1)
    p.bbox_hw, p.bounding_box_size, p.num_cameras, p.center_detect_img_size = x["bbox"] // 2, x["bbox"], 12, x["cdis"]
    import jarvis_hybridnet_b200.model as M
    orig = M.accelerate
    M.accelerate = lambda bb, precision="fp32", **kw: bb            # the 3D seam has its own test (test_accelerate_seam)
    try:
        accelerate_predictor(p)
    finally:
        M.accelerate = orig
    pts, conf = p(d(x["imgs"]), d(x["cam"]), d(x["intr"]), d(x["dist"]))
    assert tuple(pts.shape) == (1, 23, 3) and tuple(conf.shape) == (1, 23)
    assert np.array_equal(seen["centerHM"][0].cpu().numpy(), g["centerHM"])
    assert np.array_equal(seen["center3D"][0].cpu().numpy(), g["center3D_int"])
    assert np.array_equal(seen["crops"][0].cpu().numpy().reshape(-1)[::997], g["crops_sample"])
    x2, g2 = center_case("c6_undetected")
    p.num_cameras, p.bbox_hw, p.bounding_box_size, p.center_detect_img_size = 6, x2["bbox"] // 2, x2["bbox"], x2["cdis"]
    p.centerDetect.forward = lambda t: (None, d(x2["hm"]))
    assert p(d(x2["imgs"]), d(x2["cam"]), d(x2["intr"]), d(x2["dist"])) == (None, None)
