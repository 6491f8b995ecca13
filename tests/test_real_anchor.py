"""Real-data anchor (SURVEY.md §8c pin (2)): key-point heat maps of the reference's own EfficientTrack network on three
validation frame sets of its Example_Dataset, the 12 real calibration files and the bundled MonkeyHand weights —
fixtures produced by the UNMODIFIED reference (tests/golden/make_golden_real.py; whole split: 3.079 mm mean error,
tests/golden/real_val_summary.json).

CPU: the oracle against the reference's indices / volume / key points on these inputs.
GPU: the CUDA path (fp32 and bf16) against the reference's key points at the north-star bars (0.05 mm / 0.5 mm), indices
bit-exact, and the error against the annotated ground truth unchanged."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, sha

ROI, SPACING, BBOX, K, NCAM = 144, 2, 256, 23, 12


def load_real():
    d = np.load(os.path.join(GOLDEN, "real_example.npz"), allow_pickle=False)
    q = float(d["q"])
    out = []
    for i in range(len(d["names"])):
        f = {k[len(f"fs{i}_"):]: d[k] for k in d.files if k.startswith(f"fs{i}_")}
        f["hm"] = f.pop("hm_q").astype(np.float32) / np.float32(q)          # exact: multiples of 1/8 below 2^12
        f["c3"] = f.pop("center3D").astype(np.int32)
        f["chm"] = f.pop("centerHM").astype(np.int32)
        f["name"] = str(d["names"][i])
        out.append(f)
    cal = dict(cam=d["cameraMatrices"], intr=d["intrinsicMatrices"], dist=d["distortionCoefficients"])
    return out, cal


def camera_names():
    d = np.load(os.path.join(GOLDEN, "real_example.npz"), allow_pickle=False)
    return [str(c) for c in d["cameras"]]


def weights():
    return dict(np.load(os.path.join(GOLDEN, "monkeyhand_v2v_small.npz")))


def test_summary_is_the_survey_anchor():
    s = json.load(open(os.path.join(GOLDEN, "real_val_summary.json")))
    assert s["framesets"] == 30 and s["detected"] == 30
    assert abs(s["mean_err_mm"] - 3.08) < 0.01                                  # SURVEY.md §8c / BASELINE.md §5.5


@pytest.mark.parametrize("i", range(3))
def test_oracle_on_real_data(oracle, i):
    sets, cal = load_real()
    f = sets[i]
    G, hs = ROI // SPACING, BBOX // 2 + 2
    idx = oracle.reproject_indices(f["c3"], f["chm"], cal["cam"], cal["intr"], cal["dist"], G, SPACING, hs)
    assert sha(idx.astype(np.int32)) == str(f["idx_sha"])                       # the reference's own indices, real calibration
    out = oracle.hybrid3d_forward(weights(), f["hm"], f["c3"], f["chm"], cal["cam"], cal["intr"], cal["dist"], ROI, SPACING)
    np.testing.assert_allclose(out["volume"].reshape(-1)[::512], f["vol_sample"], rtol=1e-6, atol=1e-5)
    assert np.abs(out["points"] - f["points3D"]).max() < 5e-3
    np.testing.assert_allclose(out["conf"], f["confidences"], rtol=1e-5, atol=1e-6)
    assert np.abs(f["points3D"] - f["points3D_unq"]).max() < 0.02              # 1/8 quantisation of the maps is immaterial


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda")


def _inputs(f, cal):
    return (_dev(f["hm"])[None], _dev(f["c3"])[None], _dev(f["chm"])[None], _dev(cal["cam"])[None], _dev(cal["intr"])[None],
            _dev(cal["dist"])[None])


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(3))
def test_gpu_indices_real_calibration(i):
    from jarvis_hybridnet_b200 import ReprojectionLayer
    from test_host import cfg_of
    import jarvis_hybridnet_b200.synth as S
    sets, cal = load_real()
    f = sets[i]
    L = ReprojectionLayer(cfg_of(S.Shape3D(NCAM, K, BBOX, ROI, SPACING)))
    _, idx = L.forward_batched(*_inputs(f, cal), want_index=True)
    assert sha(idx[0].cpu().numpy().astype(np.int32)) == str(f["idx_sha"])


@pytest.mark.gpu
@pytest.mark.parametrize("precision,bar,conf_tol", [("fp32", 0.05, 1e-4), ("bf16", 0.5, 5e-2)])
def test_gpu_key_points_real_data(precision, bar, conf_tol):
    from jarvis_hybridnet_b200 import HybridNet3D
    sets, cal = load_real()
    net = HybridNet3D(K, BBOX, ROI, SPACING, weights(), precision=precision).to("cuda")
    # the three frame sets as ONE batch (per-sample InstanceNorm, per-sample centres) and one at a time
    cat = [torch.cat([_inputs(f, cal)[j] for f in sets]) for j in range(6)]
    pts, conf, _ = net(*cat)
    ref_err, our_err = [], []
    for i, f in enumerate(sets):
        p = pts[i].cpu().numpy()
        err = np.abs(p - f["points3D"]).max()
        print(f"{f['name']}: {precision} max key-point difference to the reference {err:.4f} mm")
        assert err < bar
        np.testing.assert_allclose(conf[i].cpu().numpy(), f["confidences"], rtol=conf_tol, atol=conf_tol / 10)
        one, _, _ = net(*_inputs(f, cal))
        assert np.abs(one[0].cpu().numpy() - p).max() < (1e-3 if precision == "fp32" else 5e-2)
        ok = np.abs(f["kps_gt"]).sum(1) > 0
        ref_err.append(np.linalg.norm(f["points3D"][ok] - f["kps_gt"][ok], axis=1))
        our_err.append(np.linalg.norm(p[ok] - f["kps_gt"][ok], axis=1))
    r, o = np.concatenate(ref_err).mean(), np.concatenate(our_err).mean()
    print(f"mean error against the annotated ground truth: reference {r:.3f} mm, {precision} CUDA path {o:.3f} mm")
    assert abs(r - o) < (0.01 if precision == "fp32" else 0.1)
