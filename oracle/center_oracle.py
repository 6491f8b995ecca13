"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, every line a separately rounded fp32 op) of the glue between
the centre-detect CNN and the 3D network in the reference's predictor.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU legs may import this; the product path never does.

Follows, line by line:
  jarvis/prediction/jarvis3D.py:143-178     argmax / threshold / scaling / .int() / clamp / crop / normalise
  jarvis/utils/reprojection.py:45-66        ReprojectionTool.reprojectPoint
  jarvis/utils/reprojection.py:69-90        ReprojectionTool.reconstructPoint (weighted DLT, torch.linalg.svd)
Pinned against tests/golden/center_cases.npz, which tests/golden/make_golden_center.py produced by running the
UNMODIFIED JarvisPredictor3D.forward of /root/reference (CNNs replaced by stubs) in the build container.
"""
import numpy as np

f32 = np.float32


def center_argmax(hm):
    """hm [ncam,1,Hc,Wc] fp32 -> preds [ncam,2] int64 (x, y), maxvals [ncam] fp32 (raw).  jarvis3D.py:147-153
    (the reference takes `% shape[2]` and `// shape[3]`: square maps)."""
    n, _, Hc, Wc = hm.shape
    flat = hm.reshape(n, -1)
    m = flat.argmax(1)
    return np.stack([m % Hc, m // Wc], 1), flat[np.arange(n), m].astype(f32)


def reproject_point(X, cam, intr, dist):
    """reprojection.py:45-66 for one point X [3] fp32 -> [ncam,2] fp32 (full-resolution pixels)."""
    P = np.concatenate([X.astype(f32), np.ones(1, f32)])
    # torch.matmul([1,1,4],[ncam,4,3]): K=4 dot product as the FMA chain of the GEMM libraries (see hybridnet_oracle.c)
    uvw = np.zeros((cam.shape[0], 3), f32)
    for q in range(3):
        s = f32(P[0]) * cam[:, 0, q]
        for k in range(1, 4):
            s = (s.astype(np.float64) + np.float64(P[k]) * cam[:, k, q].astype(np.float64)).astype(f32)   # fma: one rounding
        uvw[:, q] = s
    cx, cy, fx, fy = intr[:, 2, 0], intr[:, 2, 1], intr[:, 0, 0], intr[:, 1, 1]
    a = (uvw[:, 0] / uvw[:, 2] - cx).astype(f32)
    b = (uvw[:, 1] / uvw[:, 2] - cy).astype(f32)
    r2 = (np.square((a / fx).astype(f32)) + np.square((b / fy).astype(f32))).astype(f32)
    d = (f32(1) + ((dist[:, 0, 0] + (dist[:, 0, 1] * r2).astype(f32)).astype(f32) * r2).astype(f32)).astype(f32)
    return np.stack([((a * d).astype(f32) + cx).astype(f32), ((b * d).astype(f32) + cy).astype(f32)], 1)


def dlt_rows(points, maxvals, cam, intr, dist):
    """reprojection.py:69-84: the weighted [ncam,2,4] DLT rows.  points [2,ncam] fp32 (full-res pixels),
    maxvals [ncam] fp32 (already / 255)."""
    cx, cy, fx, fy = intr[:, 2, 0], intr[:, 2, 1], intr[:, 0, 0], intr[:, 1, 1]
    x = (points[0] - cx).astype(f32)
    y = (points[1] - cy).astype(f32)
    r2 = (np.square((x / fx).astype(f32)) + np.square((y / fy).astype(f32))).astype(f32)
    d = (f32(1) + ((dist[:, 0, 0] + (dist[:, 0, 1] * r2).astype(f32)).astype(f32) * r2).astype(f32)).astype(f32)
    x = ((x / d).astype(f32) + cx).astype(f32)
    y = ((y / d).astype(f32) + cy).astype(f32)
    Pt = np.transpose(cam, (0, 2, 1))                       # [ncam,3,4]
    A = np.stack([(x[:, None] * Pt[:, 2]).astype(f32) - Pt[:, 0], (y[:, None] * Pt[:, 2]).astype(f32) - Pt[:, 1]], 1).astype(f32)
    return (A * maxvals[:, None, None]).astype(f32)


def reconstruct_point(points, maxvals, cam, intr, dist):
    """reprojection.py:69-90 -> X [3] fp32 (mm)."""
    A = dlt_rows(points, maxvals, cam, intr, dist).reshape(-1, 4)
    _, _, vh = np.linalg.svd(A.astype(f32))
    X = vh[-1]
    return (X / X[-1])[:3].astype(f32)


def locate_center(center_hm, img_w, img_h, cdis, bbox_hw, cam, intr, dist, threshold=50.0):
    """jarvis3D.py:143-166 -> dict(preds, maxvals, num_detect, valid, center3D fp32, center3D_int, centerHM int32)."""
    preds, raw = center_argmax(center_hm)
    num = int((raw > f32(threshold)).sum())
    maxvals = (raw / f32(255.)).astype(f32)
    scale2 = (np.array([img_w / float(cdis), img_h / float(cdis)], f32) * f32(2)).astype(f32)
    pts = (preds.astype(f32) * scale2[None]).astype(f32).T.copy()        # [2,ncam]
    out = dict(preds=preds.astype(np.int32), maxvals=maxvals, num_detect=num, valid=num >= 2)
    if num >= 2:
        X = reconstruct_point(pts, maxvals, cam, intr, dist)
        chm = reproject_point(X, cam, intr, dist).astype(np.int32)       # .int(): truncation
        chm[:, 0] = np.clip(chm[:, 0], bbox_hw, img_w - bbox_hw)
        chm[:, 1] = np.clip(chm[:, 1], bbox_hw, img_h - bbox_hw)
        out.update(center3D=X, center3D_int=X.astype(np.int32), centerHM=chm)
    return out


def crop_normalize(imgs, centerHM, bbox_hw, mean, std):
    """jarvis3D.py:168-177: imgs [ncam,3,H,W] fp32 -> [ncam,3,2*bbox_hw,2*bbox_hw] fp32, (x - mean) / std."""
    n = imgs.shape[0]
    out = np.zeros((n, 3, 2 * bbox_hw, 2 * bbox_hw), f32)
    for i in range(n):
        cx, cy = int(centerHM[i, 0]), int(centerHM[i, 1])
        out[i] = imgs[i, :, cy - bbox_hw:cy + bbox_hw, cx - bbox_hw:cx + bbox_hw]
    m = np.asarray(mean, f32).reshape(1, 3, 1, 1)
    s = np.asarray(std, f32).reshape(1, 3, 1, 1)
    return ((out - m).astype(f32) / s).astype(f32)
