"""Host side of the predictor glue: what JarvisPredictor3D.forward does between the centre-detect
CNN and the 3D network (jarvis/prediction/jarvis3D.py:143-178), on the B200 kernels and without host syncs.

  locate_center(...)        argmax / detection count / triangulation / reprojection / clamp   (jarvis3D.py:147-166,
                            jarvis/utils/reprojection.py:45-90) for B frame sets in one launch
  crop_normalize(...)       the crop loop + normalisation                                      (jarvis3D.py:168-177)
  accelerate_predictor(p)   rebinds `forward` of a loaded reference JarvisPredictor3D: same signature, same return
                            contract ((points3D, confidences) or (None, None)); the two CNNs stay the reference's, the
                            3D stages are swapped by model.accelerate().  The only device->host read is the `valid`
                            flag, after everything has been enqueued.
"""
import ctypes
import types

import torch

from . import _lib


def locate_center(center_heatmaps, img_size, center_detect_img_size, bbox_hw, cameraMatrices, intrinsicMatrices,
                  distortionCoefficients, threshold=50.0, scratch=None):
    """center_heatmaps [B,ncam,Hc,Wc] (or the CNN's [ncam,1,Hc,Wc] for B=1), img_size (W, H), matrices [B,ncam,...]
    (or un-batched) -> dict(preds [B,ncam,2] i32, maxvals [B,ncam], center3D [B,3] fp32, center3D_int [B,3] i32,
    centerHM [B,ncam,2] i32, valid [B] i32), all on the device."""
    _lib.require_cuda(center_heatmaps, cameraMatrices, intrinsicMatrices, distortionCoefficients)
    lib = _lib.load()
    hm = center_heatmaps
    if cameraMatrices.dim() == 3:                       # un-batched, as the reference predictor holds them
        hm = hm.reshape(1, hm.shape[0], hm.shape[-2], hm.shape[-1])
        cameraMatrices, intrinsicMatrices, distortionCoefficients = (t[None] for t in (cameraMatrices, intrinsicMatrices,
                                                                                       distortionCoefficients))
    B, ncam, Hc, Wc = hm.shape
    f = lambda t: t.contiguous().float()
    hm, cam, intr, dist = f(hm), f(cameraMatrices), f(intrinsicMatrices), f(distortionCoefficients)
    dev = hm.device
    out = dict(preds=torch.empty((B, ncam, 2), dtype=torch.int32, device=dev),
               maxvals=torch.empty((B, ncam), dtype=torch.float32, device=dev),
               center3D=torch.empty((B, 3), dtype=torch.float32, device=dev),
               center3D_int=torch.empty((B, 3), dtype=torch.int32, device=dev),
               centerHM=torch.empty((B, ncam, 2), dtype=torch.int32, device=dev),
               valid=torch.empty((B,), dtype=torch.int32, device=dev))
    if scratch is None:
        scratch = torch.zeros(2 * B, dtype=torch.int32, device=dev)
    W, H = int(img_size[0]), int(img_size[1])
    _lib.check(lib.jhn_center_locate(_lib.dptr(hm), B, ncam, Hc, Wc, W, H, int(center_detect_img_size), int(bbox_hw),
                                     float(threshold), _lib.dptr(cam), _lib.dptr(intr), _lib.dptr(dist),
                                     _lib.dptr(out["preds"]), _lib.dptr(out["maxvals"]), _lib.dptr(out["center3D"]),
                                     _lib.dptr(out["center3D_int"]), _lib.dptr(out["centerHM"]), _lib.dptr(out["valid"]),
                                     _lib.dptr(scratch), _lib.stream_ptr()))
    return out


def crop_normalize(imgs, centerHM, valid, bbox, mean, std):
    """imgs [B,ncam,3,H,W] fp32 (or [ncam,3,H,W]), centerHM [B,ncam,2] i32, valid [B] i32 -> [B,ncam,3,bbox,bbox]."""
    _lib.require_cuda(imgs, centerHM, valid)
    lib = _lib.load()
    if imgs.dim() == 4:
        imgs = imgs[None]
    B, ncam, _, H, W = imgs.shape
    imgs = imgs.contiguous().float()
    out = torch.empty((B, ncam, 3, bbox, bbox), dtype=torch.float32, device=imgs.device)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s = (ctypes.c_float * 3)(*[float(v) for v in std])
    _lib.check(lib.jhn_crop_normalize(_lib.dptr(imgs), B, ncam, H, W, int(bbox), _lib.dptr(centerHM.contiguous()),
                                      _lib.dptr(valid.contiguous()), m, s, _lib.dptr(out), _lib.stream_ptr()))
    return out


def _accelerated_predict(self, imgs, cameraMatrices, intrinsicMatrices, distortionCoefficients):
    """Replacement for JarvisPredictor3D.forward (jarvis3D.py:131-194): same arguments and return values."""
    from torchvision import transforms
    H, W = imgs.shape[2], imgs.shape[3]
    cdis = self.center_detect_img_size
    small = transforms.functional.resize(imgs, [cdis, cdis])                         # jarvis3D.py:140-142 (reference CNN input)
    small = (small - self.transform_mean) / self.transform_std
    hm = self.centerDetect(small)[1]
    loc = locate_center(hm, (W, H), cdis, self.bbox_hw, cameraMatrices, intrinsicMatrices, distortionCoefficients,
                        scratch=self._jhn_scratch)
    crops = crop_normalize(imgs, loc["centerHM"], loc["valid"], self.bounding_box_size, self._jhn_mean, self._jhn_std)
    img_size = torch.tensor([W, H], device=imgs.device)
    _, _, points3D, confidences = self.hybridNet(crops, img_size, loc["centerHM"], loc["center3D_int"],
                                                 cameraMatrices[None], intrinsicMatrices[None], distortionCoefficients[None])
    # the one host read, after everything is enqueued: an undetected frame still costs the (already enqueued) CNN +
    # 3D launches, which the reference skips — the price of a predictor without a host sync in front of them
    if int(loc["valid"][0].item()) == 0:
        return None, None
    return points3D, confidences


def accelerate_predictor(predictor, precision="fp32"):
    """Swap the 3D stages (model.accelerate) and the glue of a loaded reference JarvisPredictor3D, in place."""
    from .model import accelerate
    accelerate(predictor.hybridNet, precision=precision)
    dev = predictor.transform_mean.device
    predictor._jhn_scratch = torch.zeros(2, dtype=torch.int32, device=dev)
    # host copies of the normalisation constants, read once here instead of two device->host syncs per frame
    predictor._jhn_mean = [float(v) for v in predictor.transform_mean.flatten().tolist()]
    predictor._jhn_std = [float(v) for v in predictor.transform_std.flatten().tolist()]
    predictor.forward = types.MethodType(_accelerated_predict, predictor)
    return predictor
