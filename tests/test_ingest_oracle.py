"""Oracle of rows f4 / f2 / a11 (oracle/ingest_oracle.py) pinned against the reference's own statements: those rows ARE
torch calls in the reference (predict3D.py:79, jarvis3D.py:168-177, efficienttrack/model.py:89-95,127, hybridnet/model.py:
65-66,73,88), so the statements are executed here on CPU tensors with seeded inputs."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from conftest import ROOT
from oracle import ingest_oracle as IO


def frames_u8(B, ncam, H, W, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (B, ncam, H, W, 3), dtype=np.uint8)


def test_ingest_matches_predict3D_line_79():
    fr = frames_u8(1, 3, 20, 24)[0]
    want = (torch.from_numpy(fr).float().permute(0, 3, 1, 2)[:, [2, 1, 0]] / 255.).numpy()
    assert np.array_equal(IO.ingest_frames(fr, cuda_scalar_division=False), want)       # ATen CPU: IEEE quotient
    # ATen CUDA (the reference's path, `.cuda()` is hard-wired): multiplication by the fp32 reciprocal; tests/test_gpu_ingest.py
    # holds the oracle's default to torch's CUDA result on the B200
    assert np.array_equal(IO.ingest_frames(fr), (torch.from_numpy(fr).float().permute(0, 3, 1, 2)[:, [2, 1, 0]] * (torch.tensor(1.) / 255.)).numpy())
    assert np.abs(IO.ingest_frames(fr) - want).max() <= 2 ** -24
    allv = np.arange(256, dtype=np.uint8).reshape(1, 1, 256, 1).repeat(3, 3)        # every byte value: [1,1,256,3]
    assert np.array_equal(IO.ingest_frames(allv, cuda_scalar_division=False)[0, 0, 0, :], (torch.arange(256).float() / 255.).numpy())


def test_crop_matches_jarvis3D_lines_168_177():
    B, ncam, H, W, bbox = 2, 3, 40, 48, 16
    fr = frames_u8(B, ncam, H, W, 1)
    rng = np.random.default_rng(2)
    chm = np.stack([rng.integers(8, W - 8, (B, ncam)), rng.integers(8, H - 8, (B, ncam))], -1).astype(np.int32)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    got = IO.crop_normalize_u8(fr, chm, np.array([1, 0]), bbox, mean, std)
    imgs = torch.from_numpy(fr[0]).float().permute(0, 3, 1, 2)[:, [2, 1, 0]] * (torch.tensor(1.) / 255.)    # CUDA semantics of `/ 255.`
    tm, ts = torch.tensor(mean).view(3, 1, 1), torch.tensor(std).view(3, 1, 1)
    hw = bbox // 2
    want = torch.zeros(ncam, 3, bbox, bbox)
    for i in range(ncam):                                                                       # jarvis3D.py:171-177
        want[i] = imgs[i, :, chm[0, i, 1] - hw:chm[0, i, 1] + hw, chm[0, i, 0] - hw:chm[0, i, 0] + hw]
    want = (want - tm) / ts
    assert np.array_equal(got[0], want.numpy())
    assert not got[1].any()


def test_head_matches_conv_transpose2d():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 11, 6, 7, generator=g)
    deconv = torch.nn.ConvTranspose2d(11, 5, kernel_size=4, stride=2, padding=1, bias=False)     # efficienttrack/model.py:89-95
    with torch.no_grad():
        want = deconv(x).numpy()
    got = IO.efftrack_head(x.numpy(), deconv.weight.detach().numpy())
    assert got.shape == want.shape and np.abs(got - want).max() < 2e-6 * np.abs(want).max() + 1e-6


def test_head_on_the_bundled_checkpoint():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import pytest
    import ref_shim
    path = os.path.join(ref_shim.WEIGHTS, "HybridNet-small.pth")
    if not os.path.exists(path):
        pytest.skip("baseline/_ref weights absent")
    sd = torch.load(path, map_location="cpu")
    w = sd["effTrack.deconv1.weight"]
    assert tuple(w.shape[2:]) == (4, 4)
    x = torch.randn(1, w.shape[0], 8, 8, generator=torch.Generator().manual_seed(1))
    want = F.conv_transpose2d(x, w, stride=2, padding=1).numpy()
    got = IO.efftrack_head(x.numpy(), w.numpy())
    assert np.abs(got - want).max() < 2e-6 * np.abs(want).max() + 1e-6


def test_pad_and_softplus2():
    g = torch.Generator().manual_seed(3)
    hm = torch.randn(2, 3, 4, 9, 9, generator=g)
    assert np.array_equal(IO.pad_heatmaps(hm.numpy()), F.pad(hm, [1, 1, 1, 1]).numpy())          # hybridnet/model.py:65-66
    v = torch.cat([torch.randn(1000, generator=g) * 8, torch.tensor([19.9, 20.0, 20.1, 60., -60., 0.])])
    want = F.softplus(F.softplus(v)).numpy()                                                      # model.py:73,88
    got = IO.softplus2(v.numpy())
    assert np.allclose(got, want, rtol=3e-7, atol=1e-30)


def test_channels_last_rounding():
    hm = (np.random.default_rng(4).random((1, 5, 6, 6)).astype(np.float32) * 255)
    cl = IO.to_channels_last(hm)
    assert cl.shape == (1, 8, 8, 24) and not cl[:, 0].any() and not cl[..., 5:].any()
    want = (torch.from_numpy(hm) * 0.0625).half().float().permute(0, 2, 3, 1).numpy()
    assert np.array_equal(cl[:, 1:-1, 1:-1, :5], want)
    clb = IO.to_channels_last(hm, bf16=True)
    assert np.array_equal(clb[:, 1:-1, 1:-1, :5], torch.from_numpy(hm).bfloat16().float().permute(0, 2, 3, 1).numpy())


def test_info_yaml_matches_a_yaml_dump(tmp_path):
    """None must be written as an empty value, awkward strings quoted — the file must parse back to the dict."""
    import yaml
    from jarvis_hybridnet_b200 import create_info_file
    for rec, name in [("/data/rec1", None), ("C:\\rec: 1 #x", "Example_Dataset"), ("- lead", "12"), ("it's", "true")]:
        p = create_info_file(str(tmp_path), rec, name, 0, 100)
        assert yaml.safe_load(open(p)) == dict(recording_path=rec, dataset_name=name, frame_start=0, number_frames=100)
    p = create_info_file(str(tmp_path), "/data/rec1", None, 5, 7)
    assert open(p).read() == "recording_path: /data/rec1\ndataset_name:\nframe_start: 5\nnumber_frames: 7\n"
