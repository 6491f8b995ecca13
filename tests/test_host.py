"""Host-side mirror of the reference interface: names, shapes, error behaviour (CPU only)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, sha
from types import SimpleNamespace as NS


def cfg_of(sh):
    return NS(HYBRIDNET=NS(GRID_SPACING=sh.spacing, ROI_CUBE_SIZE=sh.roi, NUM_CAMERAS=sh.ncam),
              KEYPOINTDETECT=NS(BOUNDING_BOX_SIZE=sh.bbox, NUM_JOINTS=sh.K))


def test_v2v_state_dict_matches_checkpoint_layout():
    """strict=True load of the bundled MonkeyHand block (hybridnet.py:89-90): 24 tensors, 897,851 params."""
    from jarvis_hybridnet_b200 import V2VNet
    mh = dict(np.load(os.path.join(GOLDEN, "monkeyhand_v2v_small.npz")))
    net = V2VNet(23, 23)
    sd = net.state_dict()
    assert list(sd) == list(mh) and len(sd) == 24
    assert sum(v.numel() for v in sd.values()) == 897851
    net.load_state_dict({k: torch.from_numpy(v) for k, v in mh.items()}, strict=True)
    assert tuple(sd["encoder_decoder.decoder_upsample1.block.0.weight"].shape) == (92, 46, 2, 2, 2)
    with pytest.raises(RuntimeError):
        net.load_state_dict({k: torch.from_numpy(v) for k, v in list(mh.items())[:-1]}, strict=True)


def test_v2v_reference_init_statistics():
    from jarvis_hybridnet_b200 import V2VNet
    torch.manual_seed(0)
    net = V2VNet(23, 23)
    w = net.state_dict()["front_layers.1.res_branch.0.weight"]
    assert abs(float(w.std()) - 0.001) < 1e-4 and float(net.state_dict()["output_layer.bias"].abs().max()) == 0.0


def test_repro_layer_attributes_and_grid():
    import jarvis_hybridnet_b200.synth as S
    from jarvis_hybridnet_b200 import ReprojectionLayer
    L = ReprojectionLayer(cfg_of(S.EXAMPLE))
    assert (L.grid_size, L.heatmap_size, L.num_cameras, L.boxsize, L.grid_spacing) == (72, 130, 12, 144, 2)
    if not torch.cuda.is_available():
        g = L.grid
        assert tuple(g.shape) == (36, 36, 36, 3)
        assert g[0, 0, 0].tolist() == [-72.0, -72.0, -72.0] and g[35, 18, 1].tolist() == [68.0, 0.0, -68.0]
    assert ReprojectionLayer(cfg_of(S.EXAMPLE), num_cameras=5).num_cameras == 5


def test_cpu_tensors_raise_not_fallback():
    import jarvis_hybridnet_b200.synth as S
    from jarvis_hybridnet_b200 import ReprojectionLayer, V2VNet, centroid_tail
    sh = S.TINY
    L = ReprojectionLayer(cfg_of(sh))
    z = torch.zeros
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L(z(1, sh.ncam, sh.K, sh.hs, sh.hs), z(1, 3), z(1, sh.ncam, 2), z(1, sh.ncam, 4, 3), z(1, sh.ncam, 3, 3),
          z(1, sh.ncam, 1, 5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        V2VNet(sh.K, sh.K)(z(1, sh.K, sh.G, sh.G, sh.G))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        centroid_tail(z(1, sh.K, sh.h, sh.h, sh.h), 2, 48, z(1, 3))


def test_shard_range_partitions():
    from jarvis_hybridnet_b200 import shard_range
    for n in (0, 1, 7, 8, 100000):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in r) - min(e - s for s, e in r) <= 1


def test_synth_is_deterministic_and_in_view():
    import jarvis_hybridnet_b200.synth as S
    sh = S.EXAMPLE
    cam, intr, dist = S.make_rig(sh.ncam, 0)
    cam2, _, _ = S.make_rig(sh.ncam, 0)
    assert sha(cam) == sha(cam2) and cam.shape == (12, 4, 3) and cam.dtype == np.float32
    hm, c3, chm, kps = S.make_frameset(sh, cam, intr, dist, 0)
    assert hm.shape == (12, 23, 128, 128) and c3.dtype == np.int32 and chm.shape == (12, 2)
    px = S.project(kps, cam, intr, dist)
    assert (px[..., 0] > 0).all() and (px[..., 0] < S.IMG_W).all() and (px[..., 1] > 0).all() and (px[..., 1] < S.IMG_H).all()
    assert hm.max() > 200          # every key point is rendered inside its crop
    w = S.make_v2v_weights(23, 0, "ref")
    assert abs(w["front_layers.0.block.0.weight"].std() - 0.001) < 2e-5
