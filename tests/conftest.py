import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_case(name):
    """Golden fixture + the synthetic inputs it was produced from (regenerated, sha-checked)."""
    import jarvis_hybridnet_b200.synth as S
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    ncam, K, bbox, roi, spacing = g["shape"]
    sh = S.Shape3D(int(ncam), int(K), int(bbox), float(roi) if roi != int(roi) else int(roi),
                   float(spacing) if spacing != int(spacing) else int(spacing))
    cam, intr, dist = S.make_rig(sh.ncam, int(g["rig_seed"]))
    hm, c3, chm, kps = S.make_frameset(sh, cam, intr, dist, int(g["fs_seed"]))
    chm = g["chm"].astype(np.int32)        # some cases shift the crop centre after generation
    c3 = g["c3"]                           # int32 (predictor path) or float32 (validation path: non-integer centres)
    got = [sha(hm), sha(cam), sha(intr), sha(dist), sha(chm), sha(c3)]
    assert got == list(g["in_sha"]), f"synthetic inputs of {name} drifted from what the reference saw"
    return sh, dict(hm=hm, c3=c3, chm=chm, cam=cam, intr=intr, dist=dist, kps=kps), g


def case_weights(name, K):
    import jarvis_hybridnet_b200.synth as S
    if name in ("small_mh", "example_mh"):
        return dict(np.load(os.path.join(GOLDEN, "monkeyhand_v2v_small.npz")))
    if name == "micro_idx":
        return dict(np.load(os.path.join(GOLDEN, "monkeyhand_v2v_small.npz")))
    table = {"tiny_s0": (0, "he"), "tiny_s1": (1, "he"), "tiny_refinit": (2, "ref"), "tiny_clamp": (0, "he"),
             "tiny_sp15": (3, "he"), "tiny_fc": (3, "he"), "example_he": (4, "he"), "stress_idx": (6, "he")}
    seed, scale = table[name]
    return S.make_v2v_weights(K, seed, scale)


FULL_CASES = ["tiny_s0", "tiny_s1", "tiny_refinit", "tiny_clamp", "tiny_sp15", "tiny_fc"]
# micro_idx / stress_idx: BASELINE.json configs 2 and 5 (12 cameras, 256^2 maps, 64^3 grid; 16 cameras, 96^3 grid) end to end
V2V_CASES = FULL_CASES + ["small_mh", "example_mh", "example_he", "micro_idx", "stress_idx"]
ALL_CASES = V2V_CASES


@pytest.fixture(scope="session")
def oracle():
    from oracle import hybridnet_oracle as O
    O.build()
    return O


_ORACLE_OUT = {}


def oracle_forward(name):
    """oracle.hybrid3d_forward of a golden case, computed once per session (the 96^3 case is ~100 GFLOP on the CPU)."""
    if name not in _ORACLE_OUT:
        from oracle import hybridnet_oracle as O
        O.build()
        sh, x, g = load_case(name)
        _ORACLE_OUT[name] = O.hybrid3d_forward(case_weights(name, sh.K), x["hm"], x["c3"], x["chm"], x["cam"], x["intr"],
                                               x["dist"], sh.roi, sh.spacing)
    return _ORACLE_OUT[name]


def decisive_argmax(v, rel=1e-4):
    """Key points whose maximum beats the runner-up by more than `rel` of the volume's scale: there the argmax voxel
    is a property of the data, not of the summation order, and must match bit for bit."""
    K = v.shape[0]
    flat = np.asarray(v, np.float64).reshape(K, -1)
    top2 = np.partition(flat, -2, axis=1)[:, -2:]
    return (top2[:, 1] - top2[:, 0]) > rel * np.abs(flat).max()
