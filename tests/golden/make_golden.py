#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

How the reference is made to run here: empty stub modules for the packages its
*unused* imports pull in, a TorchFunctionMode that rewrites device='cuda' -> 'cpu' and makes
Tensor.cuda() the identity, nn.Module.cuda -> identity, a torch.cuda.IntTensor shim and
torch.load(map_location='cpu').  Nothing under /root/reference is modified or copied.

For every case the INPUTS come from jarvis_hybridnet_b200.synth (seeded numpy), so only their sha256
is stored; the OUTPUTS stored are produced by the reference's own classes:
  idx      ReprojectionLayer.reprojectPoints        (repro_layer.py:40-85)    int32, full or sha256+sample
  volume   ReprojectionLayer.forward                (repro_layer.py:110-119)  fp32, full or strided sample
  v2v      V2VNet.forward                           (v2vnet.py:98-102)        fp32
  points3D, confidences, heatmap_final  HybridNetBackbone.forward (model.py:53-90) with effTrack replaced
                                                    by a stub that returns the synthetic heat maps
The script also re-checks the two library-defined roundings the oracle relies on (SGEMM K=4 FMA chain,
trilinear lerp contraction) and records the result in tests/golden/MANIFEST.json.
"""
import hashlib
import importlib
import json
import os
import sys
import types
from types import SimpleNamespace as NS

import numpy as np
import torch
from torch.overrides import TorchFunctionMode

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

import jarvis_hybridnet_b200.synth as S  # noqa: E402
from oracle import hybridnet_oracle as O  # noqa: E402


# ------------------------------------------------------------------ reference import shim
class _Any:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, n): return _Any()


def _stub_getattr(n):
    if n.startswith("__"):
        raise AttributeError(n)
    return _Any


def install_stubs():
    for name in ["matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "imgaug",
                 "imgaug.augmenters", "imgaug.augmentables", "imgaug.augmentables.kps", "streamlit", "yacs",
                 "yacs.config", "ruamel", "ruamel.yaml", "seaborn", "inquirer", "streamlit_option_menu"]:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            m.__path__ = []
            m.__getattr__ = _stub_getattr
            sys.modules[name] = m


class CpuRedirect(TorchFunctionMode):
    def __torch_function__(self, func, types_, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        d = kwargs.get("device")
        if d is not None and "cuda" in str(d):
            kwargs["device"] = torch.device("cpu")
        if func is torch.Tensor.cuda:
            return args[0]
        return func(*args, **kwargs)


def patch_torch():
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.IntTensor = lambda x: torch.tensor(x, dtype=torch.int32)
    _load = torch.load

    def load(f, *a, **k):
        k.setdefault("map_location", "cpu")
        k.setdefault("weights_only", False)
        return _load(f, *a, **k)
    torch.load = load


def make_cfg(sh):
    return NS(DATASET=NS(DATASET_ROOT_DIR="", DATASET_3D="", DATASET_2D=""),
              HYBRIDNET=NS(GRID_SPACING=sh.spacing, ROI_CUBE_SIZE=sh.roi, NUM_CAMERAS=sh.ncam),
              KEYPOINTDETECT=NS(BOUNDING_BOX_SIZE=sh.bbox, NUM_JOINTS=sh.K, MODEL_SIZE="small"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class StubTrack(torch.nn.Module):
    """Stands in for EfficientTrackBackbone (out of scope): returns the synthetic heat maps at [1]."""
    def __init__(self, hm):
        super().__init__()
        self.hm = hm

    def forward(self, imgs):
        return None, self.hm


def build_backbone(sh, v2v_sd):
    """The reference's own HybridNetBackbone (model.py:20-50), built by its own constructor; only the
    V2V parameters are then loaded (strict) and effTrack is swapped for the heat-map stub."""
    from jarvis.hybridnet.model import HybridNetBackbone
    bb = HybridNetBackbone(make_cfg(sh))
    bb.v2vNet.load_state_dict({k: torch.from_numpy(v) for k, v in v2v_sd.items()}, strict=True)
    bb.eval()
    return bb


def run_case(name, sh, rig_seed, fs_seed, weights, full=True, chm_shift=None, sample=64, with_v2v=True,
             idx_full=False, float_centre=False):
    cam, intr, dist = S.make_rig(sh.ncam, rig_seed)
    hm, c3, chm, kps = S.make_frameset(sh, cam, intr, dist, fs_seed)
    if float_centre:
        # the validation path hands the 3D network float centres int(c / spacing) * spacing (dataset3D via
        # hybridnet.py:284-304): non-integer whenever GRID_SPACING is fractional
        c3 = (np.trunc(c3.astype(np.float64) / sh.spacing) * sh.spacing + (0.5 * sh.spacing if float_centre == "half" else 0.0)).astype(np.float32)
    if chm_shift is not None:
        chm = (chm + np.asarray(chm_shift, np.int32)).astype(np.int32)
    out = dict(shape=np.array([sh.ncam, sh.K, sh.bbox, sh.roi, sh.spacing], np.float64),
               rig_seed=rig_seed, fs_seed=fs_seed, chm=chm, c3=c3,
               in_sha=np.array([sha(hm), sha(cam), sha(intr), sha(dist), sha(chm), sha(c3)]))
    with CpuRedirect(), torch.no_grad():
        bb = build_backbone(sh, weights)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        hp = torch.nn.functional.pad(t(hm)[None], [1, 1, 1, 1])
        # stage 1: indices (reference internals, same call forward() makes)
        grid = bb.reproLayer.grid + t(c3)[None][0]
        idx = bb.reproLayer.reprojectPoints(grid, t(cam), t(intr), t(dist), t(chm)).numpy().astype(np.int32)
        vol = bb.reproLayer(hp, t(c3)[None], t(chm)[None], t(cam)[None], t(intr)[None], t(dist)[None])[0].numpy()
        out["idx_sha"] = sha(idx)
        out["vol_sum"] = np.array([vol.astype(np.float64).sum(), (vol.astype(np.float64) ** 2).sum()])
        if full:
            out["idx"] = idx
            out["volume"] = vol
        else:
            if idx_full:
                out["idx"] = idx
            out["idx_sample"] = idx.reshape(-1)[::97].copy()
            out["volume_sample"] = vol.reshape(-1)[::sample].copy()
            out["volume_stride"] = sample
        if with_v2v:
            v = bb.v2vNet(torch.from_numpy(vol)[None] / 255.)[0].numpy()
            bb.effTrack = StubTrack(t(hm))
            imgs = torch.zeros(1, sh.ncam, 3, 4, 4)
            hf, hpad, p3, conf = bb(imgs, torch.tensor([S.IMG_W, S.IMG_H]), t(chm)[None], t(c3)[None],
                                    t(cam)[None], t(intr)[None], t(dist)[None])
            vs = 1 if v.size <= 200_000 else (4 if v.size <= 1_200_000 else 16)
            out["v2v"] = v if vs == 1 else v.reshape(-1)[::vs].copy()
            out["v2v_stride"] = vs
            out["points3D"] = p3[0].numpy()
            out["confidences"] = conf[0].numpy()
            out["kps_true"] = kps.astype(np.float32)
            sp = torch.nn.functional.softplus(torch.from_numpy(v))
            out["argmax"] = sp.view(sh.K, -1).argmax(1).numpy().astype(np.int32)
            out["hf_sum"] = hf.double().sum().item()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    err = np.linalg.norm(out["points3D"] - kps, axis=1).mean() if with_v2v else float("nan")
    print(f"{name}: idx {idx.shape} vol max {vol.max():.2f}  mean |p-kps| {err:.3f} mm")
    return out


def check_library_roundings():
    """Re-verify what the oracle assumes about torch CPU (see oracle/hybridnet_oracle.c header)."""
    sh = S.EXAMPLE
    cam, intr, dist = S.make_rig(sh.ncam, 3)
    _, c3, chm, _ = S.make_frameset(sh, cam, intr, dist, 3)
    from jarvis.hybridnet.repro_layer import ReprojectionLayer
    with CpuRedirect():
        L = ReprojectionLayer(make_cfg(sh))
        t = torch.from_numpy
        ref = L.reprojectPoints(L.grid + t(c3), t(cam), t(intr), t(dist), t(chm)).numpy()
    res = {}
    for mode in (0, 1, 2):
        mine = O.reproject_indices(c3, chm, cam, intr, dist, sh.G, sh.spacing, sh.hs, lerp_mode=mode)
        res[f"lerp_mode_{mode}_mismatches"] = int((mine != ref).sum())
    return res


def main():
    install_stubs()
    patch_torch()
    sys.path.insert(0, REF)
    torch.manual_seed(0)
    manifest = {"torch": torch.__version__, "reference": REF, "library_roundings": None, "cases": {}}

    # bundled MonkeyHand V2V block (24 tensors, 897,851 params) as a weights fixture
    sd = torch.load(os.path.join(REF, "pretrained/MonkeyHand/HybridNet-small.pth"))
    mh = {k[len("v2vNet."):]: v.numpy().astype(np.float32) for k, v in sd.items() if k.startswith("v2vNet.")}
    assert len(mh) == 24
    np.savez_compressed(os.path.join(HERE, "monkeyhand_v2v_small.npz"), **mh)

    with CpuRedirect():
        manifest["library_roundings"] = check_library_roundings()
    print(manifest["library_roundings"])

    T, SM, EX = S.TINY, S.SMALL, S.EXAMPLE
    cases = [
        ("tiny_s0", T, 0, 0, S.make_v2v_weights(T.K, 0, "he"), dict()),
        ("tiny_s1", T, 1, 1, S.make_v2v_weights(T.K, 1, "he"), dict()),
        ("tiny_refinit", T, 0, 2, S.make_v2v_weights(T.K, 2, "ref"), dict()),
        ("tiny_clamp", T, 2, 3, S.make_v2v_weights(T.K, 0, "he"), dict(chm_shift=[[40, -25]] * T.ncam)),
        ("tiny_sp15", S.Shape3D(ncam=3, K=4, bbox=64, roi=36, spacing=1.5), 4, 4,
         S.make_v2v_weights(4, 3, "he"), dict()),
        ("tiny_fc", S.Shape3D(ncam=3, K=4, bbox=64, roi=36, spacing=1.5), 4, 6,
         S.make_v2v_weights(4, 3, "he"), dict(float_centre="half")),
        ("small_mh", SM, 5, 5, mh, dict(full=False, sample=4, idx_full=True)),
        ("example_mh", EX, 0, 0, mh, dict(full=False)),
        ("example_he", EX, 1, 1, S.make_v2v_weights(EX.K, 4, "he"), dict(full=False)),
        # BASELINE configs 2 and 5 end to end (V2V at h=32/q=16 and h=48/q=24): MonkeyHand weights on the micro shape,
        # seeded "he" weights on the stress shape
        ("micro_idx", S.MICRO, 0, 0, mh, dict(full=False, sample=128)),
        ("stress_idx", S.STRESS, 0, 0, S.make_v2v_weights(S.STRESS.K, 6, "he"), dict(full=False, sample=256)),
    ]
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only and os.path.exists(os.path.join(HERE, "MANIFEST.json")):
        manifest = json.load(open(os.path.join(HERE, "MANIFEST.json")))
    for name, sh, rs, fs, w, kw in cases:
        if only and name not in only:
            continue
        if w is None:
            w = S.make_v2v_weights(sh.K, 0, "ref")
        o = run_case(name, sh, rs, fs, w, **kw)
        manifest["cases"][name] = {"idx_sha": str(o["idx_sha"])}
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
