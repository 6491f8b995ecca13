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


def _predict_frames(self, frames, cameraMatrices, intrinsicMatrices, distortionCoefficients):
    """B frame sets per call from the decoder's frames: uint8 [B,ncam,H,W,3] BGR on the device (what
    `cv2.VideoCapture.read()` wrote, predict3D.py:72-78) -> points3D [B,K,3] mm, confidences [B,K], valid [B] i32 —
    all on the device, no host sync.  Rows of undetected frame sets (valid == 0: fewer than two cameras above the
    threshold, jarvis3D.py:149-151) are meaningless; write them as NaN (output.write_data3D_csv does).

    The same stages as JarvisPredictor3D.forward (jarvis3D.py:131-194) with B as a real batch dimension: u8 -> fp32 RGB
    (jhn_ingest_frames, row f4), resize + normalise + centre-detect CNN (reference modules), centre localisation for all
    B in one launch (jhn_center_locate), crops straight from the uint8 frames (jhn_crop_normalize_u8), key-point CNN
    (reference module; last layer emits the gather's layout when accelerate_predictor(head_format=...) was given, row
    f2), reprojection + V2V + soft-argmax through one jhn_hybrid3d_forward call."""
    from torchvision import transforms
    from .ingest import crop_normalize_u8, ingest_frames
    B, ncam, H, W, _ = frames.shape
    cdis = self.center_detect_img_size
    imgs = ingest_frames(frames.reshape(B * ncam, H, W, 3))                          # predict3D.py:79
    small = transforms.functional.resize(imgs, [cdis, cdis])                         # jarvis3D.py:140-142
    small = (small - self.transform_mean) / self.transform_std
    del imgs
    hm = self.centerDetect(small)[1]
    ex = lambda t: t if t.dim() == 4 else t[None].expand(B, *t.shape)
    cam, intr, dist = ex(cameraMatrices), ex(intrinsicMatrices), ex(distortionCoefficients)
    loc = locate_center(hm.reshape(B, ncam, hm.shape[-2], hm.shape[-1]), (W, H), cdis, self.bbox_hw, cam, intr, dist)
    crops = crop_normalize_u8(frames, loc["centerHM"], loc["valid"], self.bounding_box_size, self._jhn_mean, self._jhn_std)
    img_size = torch.tensor([W, H], device=frames.device)
    points3D, confidences = self.hybridNet.forward_batched(crops, img_size, loc["centerHM"], loc["center3D_int"], cam, intr, dist)
    return points3D, confidences, loc["valid"]


def accelerate_predictor(predictor, precision="fp32", head_format=None):
    """Swap the 3D stages (model.accelerate) and the glue of a loaded reference JarvisPredictor3D, in place.
    `predictor(imgs, ...)` keeps the reference's signature and return contract; `predictor.predict_frames(frames, ...)` is
    the batched entry a predict3D-style loop uses to reach the B = 32 rate."""
    from .model import accelerate
    accelerate(predictor.hybridNet, precision=precision, head_format=head_format)
    dev = predictor.transform_mean.device
    predictor._jhn_scratch = torch.zeros(2, dtype=torch.int32, device=dev)
    # host copies of the normalisation constants, read once here instead of two device->host syncs per frame
    predictor._jhn_mean = [float(v) for v in predictor.transform_mean.flatten().tolist()]
    predictor._jhn_std = [float(v) for v in predictor.transform_std.flatten().tolist()]
    predictor.forward = types.MethodType(_accelerated_predict, predictor)
    predictor.predict_frames = types.MethodType(_predict_frames, predictor)
    return predictor


def predict3D_frames(predictor, read_frame_set, n_frame_sets, calibration, csv_path, ncam, img_size, batch=32, keypoint_names=None):
    """The loop of jarvis/prediction/predict3D.py:72-103 with frame sets batched: `read_frame_set(dst)` fills a uint8
    [ncam,H,W,3] numpy view with the next frame of every camera (predict3D.read_images does exactly that into
    `imgs_orig`) and returns False at the end of the videos.  Frames are decoded into pinned memory, uploaded as bytes
    while the previous batch computes (ingest.FrameUploader), predicted `batch` frame sets at a time, and the rows of
    data3D.csv are written from ONE device->host copy per batch (output.write_data3D_csv: the reference's bytes)."""
    import csv
    from .ingest import FrameUploader
    from .output import create_header, format_rows
    cam, intr, dist = calibration
    W, H = int(img_size[0]), int(img_size[1])
    up = FrameUploader((batch, ncam, H, W, 3))
    pending, uploads, done, slot, total = None, [None, None], 0, 0, 0
    with open(csv_path, "w", newline="") as fh:
        writer = csv.writer(fh, delimiter=",", quotechar='"', quoting=csv.QUOTE_MINIMAL)
        if keypoint_names:
            create_header(writer, keypoint_names)

        def flush(p):
            res, n = p                                                       # [n,K,4] + valid, one D2H
            pts, conf, valid = (t.cpu() for t in res)
            for row in format_rows(pts[:n], conf[:n], valid[:n]):
                writer.writerow(row)

        while done < n_frame_sets:
            if uploads[slot] is not None:
                uploads[slot].synchronize()                                  # the pinned slot is free again
            host = up.host(slot)
            n = 0
            while n < batch and done + n < n_frame_sets and read_frame_set(host[n]):
                n += 1
            if n == 0:
                break
            frames, ev = up.upload(slot)
            uploads[slot] = ev
            torch.cuda.current_stream().wait_event(ev)
            with torch.no_grad():
                res = predictor.predict_frames(frames[:n] if n < batch else frames, cam, intr, dist)
            up.release(slot)
            if pending is not None:
                flush(pending)                                               # the previous batch: its kernels finished long ago
            pending = (res, n)
            done += n
            total += n
            slot ^= 1
        if pending is not None:
            flush(pending)
    return total
