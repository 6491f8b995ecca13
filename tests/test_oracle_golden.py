"""The CPU oracle against fixtures produced by the reference itself (tests/golden/make_golden.py).

This is what PINS the oracle: indices bit-exact (sha256 of the whole tensor), feature volume / V2V
output / key points within fp32 re-association noise of the reference's own CPU run."""
import numpy as np
import pytest

from conftest import ALL_CASES, V2V_CASES, case_weights, load_case, oracle_forward, sha


@pytest.mark.parametrize("name", ALL_CASES)
def test_indices_bit_exact(oracle, name):
    sh, x, g = load_case(name)
    idx = oracle.reproject_indices(x["c3"], x["chm"], x["cam"], x["intr"], x["dist"], sh.G, sh.spacing, sh.hs)
    assert idx.dtype == np.int64 and idx.shape == (sh.ncam, sh.G, sh.G, sh.G)
    assert sha(idx.astype(np.int32)) == str(g["idx_sha"])
    if "idx" in g:
        assert np.array_equal(idx, g["idx"])
    else:
        assert np.array_equal(idx.reshape(-1)[::97], g["idx_sample"])
    assert idx.min() >= 0 and (idx % sh.hs).max() <= sh.hs - 2 and (idx // sh.hs).max() <= sh.hs - 2


@pytest.mark.parametrize("name", ALL_CASES)
def test_volume(oracle, name):
    sh, x, g = load_case(name)
    vol, _ = oracle.repro_layer_forward(oracle.pad_heatmaps(x["hm"]), x["c3"], x["chm"], x["cam"], x["intr"],
                                        x["dist"], sh.G, sh.spacing)
    if "volume" in g:
        ref, mine = g["volume"], vol
    else:
        ref, mine = g["volume_sample"], vol.reshape(-1)[::int(g["volume_stride"])]
    # same gathers, same camera order; only the final mean may round differently (sum/12 vs sum*(1/12))
    np.testing.assert_allclose(mine, ref, rtol=1e-6, atol=1e-5)
    s = np.array([vol.astype(np.float64).sum(), (vol.astype(np.float64) ** 2).sum()])
    np.testing.assert_allclose(s, g["vol_sum"], rtol=1e-6)


@pytest.mark.parametrize("name", V2V_CASES)
def test_v2v_and_tail(oracle, name):
    sh, x, g = load_case(name)
    out = oracle_forward(name)
    v = out["v2v"].reshape(-1)[::int(g["v2v_stride"])].reshape(g["v2v"].shape)
    scale = np.abs(g["v2v"]).max()
    assert np.abs(v - g["v2v"]).max() <= 2e-5 * scale + 1e-6
    # key points: 0.05 mm is the fp32 bar of the north star; the oracle itself sits far inside it
    assert np.abs(out["points"] - g["points3D"]).max() < 5e-3
    np.testing.assert_allclose(out["conf"], g["confidences"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(out["argmax"], g["argmax"])


def test_lerp_modes_differ_only_on_ties(oracle):
    """lerp_mode is the single knob for the library-defined trilinear rounding; the three candidates
    agree except on a handful of truncation ties (SURVEY.md §7 'bit-exact indices')."""
    sh, x, g = load_case("small_mh")
    base = oracle.reproject_indices(x["c3"], x["chm"], x["cam"], x["intr"], x["dist"], sh.G, sh.spacing, sh.hs)
    for mode in (1, 2):
        alt = oracle.reproject_indices(x["c3"], x["chm"], x["cam"], x["intr"], x["dist"], sh.G, sh.spacing,
                                       sh.hs, lerp_mode=mode)
        assert (alt != base).mean() < 1e-4


def test_centroid_known_answer(oracle):
    """A single hot voxel: centroid lands on it, confidence = min(v,255)/255, argmax = its flat index."""
    K, h = 2, 8
    v = np.full((K, h, h, h), -200.0, np.float32)       # softplus(-200) == 0 in fp32
    v[0, 2, 5, 7] = 300.0
    v[1, 6, 1, 0] = 100.0
    pts, conf, am = oracle.centroid_tail(v, spacing=2, roi=32, center3D=[10, -20, 30])
    np.testing.assert_allclose(pts[0], np.array([2, 5, 7]) * 4 - 16 + np.array([10, -20, 30]), atol=1e-4)
    np.testing.assert_allclose(pts[1], np.array([6, 1, 0]) * 4 - 16 + np.array([10, -20, 30]), atol=1e-4)
    np.testing.assert_allclose(conf, [1.0, 100 / 255], rtol=1e-6)
    assert list(am) == [(2 * h + 5) * h + 7, (6 * h + 1) * h + 0]


def test_pad_matches_reference_pad(oracle):
    a = np.arange(2 * 3 * 4 * 4, dtype=np.float32).reshape(2, 3, 4, 4)
    p = oracle.pad_heatmaps(a)
    assert p.shape == (2, 3, 6, 6) and p[:, :, 0].sum() == 0 and p[:, :, :, -1].sum() == 0
    assert np.array_equal(p[:, :, 1:-1, 1:-1], a)


@pytest.mark.parametrize("name", ALL_CASES)
def test_row_spans_hold_every_reference_index(oracle, name):
    """oracle.pixel_boxes_and_row_spans (what jhn_heatmap_boxes / jhn_heatmap_spans compute on the GPU): the indices pinned above
    by the reference's own sha all fall inside the per-row column spans, the spans are no wider than the boxes, and the boxes
    are the spans' hull."""
    sh, x, g = load_case(name)
    idx, ca, cb = oracle.reproject_indices(x["c3"], x["chm"], x["cam"], x["intr"], x["dist"], sh.G, sh.spacing, sh.hs,
                                           return_coarse=True)
    boxes, lo, hi = oracle.pixel_boxes_and_row_spans(ca, cb, sh.hs)
    span_px = box_px = 0
    for c in range(sh.ncam):
        xx, yy = idx[c].ravel() % sh.hs, idx[c].ravel() // sh.hs
        assert (xx >= lo[c][yy]).all() and (xx <= hi[c][yy]).all()
        rows = hi[c] >= lo[c]
        x0, y0, x1, y1 = boxes[c]
        assert np.flatnonzero(rows).min() == y0 and np.flatnonzero(rows).max() == y1 and rows[y0:y1 + 1].all()
        assert lo[c][rows].min() == x0 and hi[c][rows].max() == x1
        assert xx.min() >= x0 and xx.max() <= x1 and yy.min() >= y0 and yy.max() <= y1
        span_px += int((hi[c] - lo[c] + 1)[rows].sum())
        box_px += int((x1 - x0 + 1) * (y1 - y0 + 1))
    assert span_px <= box_px
