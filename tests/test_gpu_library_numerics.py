"""Which roundings do the reference's GPU libraries (cuBLAS SGEMM K=4, ATen CUDA trilinear) use on this
B200?  The torch-op restatement of the reference chain runs on the device and is compared with the oracle
in each lerp mode; the kernel is then checked in the mode that reproduces the reference's GPU path."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, load_case
from ref_chain import torch_chain_indices
from test_host import cfg_of

pytestmark = pytest.mark.gpu


def test_gpu_reference_chain_vs_lerp_modes(oracle):
    import torch
    from jarvis_hybridnet_b200 import ReprojectionLayer
    torch.backends.cuda.matmul.allow_tf32 = False
    report = {}
    for name in ("small_mh", "example_mh", "micro_idx"):
        sh, x, g = load_case(name)
        ref = torch_chain_indices(x["c3"], x["chm"], x["cam"], x["intr"], x["dist"], sh.G, sh.spacing, sh.hs,
                                  "cuda").cpu().numpy()
        mism = {}
        for mode in (0, 1, 2):
            mine = oracle.reproject_indices(x["c3"], x["chm"], x["cam"], x["intr"], x["dist"], sh.G, sh.spacing,
                                            sh.hs, lerp_mode=mode)
            mism[mode] = int((mine != ref).sum())
        report[name] = dict(total=int(ref.size), mismatches_by_lerp_mode=mism)
        best = min(mism, key=mism.get)
        L = ReprojectionLayer(cfg_of(sh), lerp_mode=best)
        d = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()[None]
        _, idx = L.forward_batched(d(x["hm"]), d(x["c3"]), d(x["chm"]), d(x["cam"]), d(x["intr"]), d(x["dist"]),
                                   want_index=True)
        report[name]["kernel_vs_gpu_reference_chain"] = int((idx[0].cpu().numpy() != ref).sum())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_library_numerics.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report))
    for name, r in report.items():
        # at most a handful of truncation ties may separate the GPU libraries from any single candidate
        assert min(r["mismatches_by_lerp_mode"].values()) <= 1e-5 * r["total"], report
        assert r["kernel_vs_gpu_reference_chain"] <= 1e-5 * r["total"], report
