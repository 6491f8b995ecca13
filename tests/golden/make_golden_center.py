#!/usr/bin/env python
"""Generate tests/golden/center_cases.npz by running the UNMODIFIED JarvisPredictor3D.forward of the reference
(jarvis/prediction/jarvis3D.py:131-194) on CPU in the build container (needs /root/reference):

    python tests/golden/make_golden_center.py

The predictor object is created without its constructor (which would load the CNN checkpoints): the attributes
forward() reads are set by hand, `centerDetect` is a stub returning the synthetic centre heat maps and `hybridNet`
a stub that records the arguments the predictor passes on — imgs_cropped, centerHMs, center3D.int() — which are
exactly the outputs of the glue this repo rebuilds.  ReprojectionTool is the reference's own
(reprojection.py:16-90).  Import shim as in make_golden.py; nothing under /root/reference is modified or copied.
Inputs come from jarvis_hybridnet_b200.synth (seeded), so only the seeds are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (reference import shim)
import jarvis_hybridnet_b200.synth as S  # noqa: E402

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]

# name: (ncam, rig seed, case seed, centre-detect image size, bounding box, weak cameras, small images)
CASES = {
    "c12_s0": (12, 0, 0, 256, 256, 0, False),
    "c12_s1": (12, 1, 1, 256, 256, 0, False),
    "c12_weak3": (12, 0, 2, 256, 256, 3, False),
    "c4_small": (4, 2, 3, 128, 64, 0, True),
    "c16_s4": (16, 3, 4, 320, 256, 0, False),
    "c6_undetected": (6, 4, 5, 128, 64, 5, True),
}


class Recorder(torch.nn.Module):
    def forward(self, imgs, img_size, centerHM, center3D, cam, intr, dist):
        self.seen = dict(crops=imgs[0].clone(), img_size=img_size.clone(), centerHM=centerHM[0].clone(),
                         center3D_int=center3D[0].clone())
        return None, None, torch.zeros(1, 1, 3), torch.zeros(1, 1)


class StubDetect(torch.nn.Module):
    def __init__(self, hm):
        super().__init__()
        self.hm = hm

    def forward(self, x):
        return None, self.hm


def main():
    MG.install_stubs()
    MG.patch_torch()
    sys.path.insert(0, MG.REF)
    from jarvis.prediction.jarvis3D import JarvisPredictor3D
    from jarvis.utils.reprojection import ReprojectionTool
    out = {}
    for name, (ncam, rig_seed, seed, cdis, bbox, n_weak, small) in CASES.items():
        cam, intr, dist = S.make_rig(ncam, rig_seed)
        if small:                                   # quarter-size images: scale the pixel geometry with them
            cam = cam.copy(); intr = intr.copy()
            cam[:, :, :2] *= 0.25; intr[:, :, :2] *= 0.25; intr[:, 2, 2] = 1.0
        hm, imgs, centre = S.make_center_case(ncam, cam, intr, dist, seed, cdis, n_weak=n_weak, small_images=small)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        with MG.CpuRedirect(), torch.no_grad():
            pred = object.__new__(JarvisPredictor3D)
            torch.nn.Module.__init__(pred)
            pred.centerDetect = StubDetect(t(hm))
            pred.hybridNet = Recorder()
            pred.transform_mean = torch.tensor(MEAN).view(3, 1, 1)
            pred.transform_std = torch.tensor(STD).view(3, 1, 1)
            pred.bbox_hw = bbox // 2
            pred.num_cameras = ncam
            pred.bounding_box_size = bbox
            pred.reproTool = ReprojectionTool(device="cpu")
            pred.center_detect_img_size = cdis
            p3, conf = pred.forward(t(imgs), t(cam), t(intr), t(dist))
            valid = p3 is not None
            out[name + "/valid"] = np.array(valid)
            if valid:
                seen = pred.hybridNet.seen
                out[name + "/centerHM"] = seen["centerHM"].numpy().astype(np.int32)
                out[name + "/center3D_int"] = seen["center3D_int"].numpy().astype(np.int32)
                crops = seen["crops"].numpy()
                out[name + "/crops_sample"] = crops.reshape(-1)[::997].copy()
                out[name + "/crops_sum"] = np.array([crops.astype(np.float64).sum(), (crops.astype(np.float64) ** 2).sum()])
                # the float centre, from the reference's own reconstructPoint on the same inputs
                tool = ReprojectionTool(device="cpu")
                tool.cameraMatrices, tool.intrinsicMatrices, tool.distortionCoefficients = t(cam), t(intr), t(dist)
                hg = t(hm).view(ncam, 1, -1)
                m = hg.argmax(2).view(ncam, 1, 1)
                preds = torch.cat((m % hm.shape[2], m // hm.shape[3]), dim=2)
                maxvals = hg.gather(2, m) / 255.
                H, W = imgs.shape[2], imgs.shape[3]
                ds = torch.tensor([W / float(cdis), H / float(cdis)]).float()
                X = tool.reconstructPoint((preds.reshape(ncam, 2) * (ds * 2)).transpose(0, 1), maxvals)
                out[name + "/center3D"] = X.numpy().astype(np.float32)
                out[name + "/preds"] = preds.reshape(ncam, 2).numpy().astype(np.int32)
                out[name + "/maxvals"] = maxvals.reshape(ncam).numpy().astype(np.float32)
                out[name + "/repro"] = tool.reprojectPoint(X.unsqueeze(0)).numpy().astype(np.float32)
        out[name + "/cfg"] = np.array([ncam, rig_seed, seed, cdis, bbox, n_weak, int(small)], np.int64)
        print(name, "valid" if valid else "no detection", out.get(name + "/center3D"), "true", centre)
    np.savez_compressed(os.path.join(HERE, "center_cases.npz"), **out)


if __name__ == "__main__":
    main()
