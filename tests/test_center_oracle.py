"""The centre-localisation oracle (oracle/center_oracle.py) against fixtures produced by the reference's own
JarvisPredictor3D.forward / ReprojectionTool (tests/golden/make_golden_center.py).  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def center_case(name):
    import jarvis_hybridnet_b200.synth as S
    g = np.load(os.path.join(GOLDEN, "center_cases.npz"))
    ncam, rig_seed, seed, cdis, bbox, n_weak, small = [int(v) for v in g[name + "/cfg"]]
    cam, intr, dist = S.make_rig(ncam, rig_seed)
    if small:
        cam = cam.copy(); intr = intr.copy()
        cam[:, :, :2] *= 0.25; intr[:, :, :2] *= 0.25; intr[:, 2, 2] = 1.0
    hm, imgs, centre = S.make_center_case(ncam, cam, intr, dist, seed, cdis, n_weak=n_weak, small_images=bool(small))
    gold = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}
    return dict(hm=hm, imgs=imgs, cam=cam, intr=intr, dist=dist, cdis=cdis, bbox=bbox, W=imgs.shape[3], H=imgs.shape[2]), gold


CENTER_CASES = ["c12_s0", "c12_s1", "c12_weak3", "c4_small", "c16_s4", "c6_undetected"]


@pytest.mark.parametrize("name", CENTER_CASES)
def test_center_oracle_matches_reference(name):
    from oracle import center_oracle as C
    x, g = center_case(name)
    r = C.locate_center(x["hm"], x["W"], x["H"], x["cdis"], x["bbox"] // 2, x["cam"], x["intr"], x["dist"])
    assert r["valid"] == bool(g["valid"])
    if not r["valid"]:
        return
    assert np.array_equal(r["preds"], g["preds"])
    assert np.array_equal(r["maxvals"], g["maxvals"])                       # one fp32 division: bit-exact
    # SVD of a [2*ncam,4] fp32 matrix: LAPACK builds differ in the last bits; the centre is ~100 mm
    assert np.abs(r["center3D"] - g["center3D"]).max() < 2e-3
    assert np.array_equal(r["center3D_int"], g["center3D_int"])
    assert np.abs(C.reproject_point(g["center3D"], x["cam"], x["intr"], x["dist"]) - g["repro"]).max() < 1e-3
    assert np.array_equal(r["centerHM"], g["centerHM"])
    crops = C.crop_normalize(x["imgs"], g["centerHM"], x["bbox"] // 2, MEAN, STD)
    assert np.array_equal(crops.reshape(-1)[::997], g["crops_sample"])       # copy, subtract, divide: bit-exact
    s = np.array([crops.astype(np.float64).sum(), (crops.astype(np.float64) ** 2).sum()])
    np.testing.assert_allclose(s, g["crops_sum"], rtol=1e-12)
