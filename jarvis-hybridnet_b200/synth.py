"""Synthetic camera rigs, frame sets, heat maps and V2V weights of the shapes BASELINE.json names.

Recipe: SURVEY.md §8(d).  Pure numpy (seeded `default_rng`) so the same seed gives identical arrays
in the build container (where the goldens are produced from the reference) and on the GPU box.
Conventions follow the reference: camera matrix `P = [R;T] @ Kt` of shape [4,3] acting on row
vectors (jarvis/utils/reprojection.py:105-107, :27-39), intrinsics stored transposed
(`intr[2,0]=cx`, `intr[2,1]=cy`), distortion `[1,5]` of which only k1,k2 are read
(jarvis/hybridnet/repro_layer.py:42-43,55-61).
"""
from dataclasses import dataclass
import numpy as np


@dataclass(frozen=True)
class Shape3D:
    """The four config scalars the hot path reads (jarvis/hybridnet/repro_layer.py:16-24,37)."""
    ncam: int
    K: int
    bbox: int          # KEYPOINTDETECT.BOUNDING_BOX_SIZE (full-res crop, px)
    roi: float         # HYBRIDNET.ROI_CUBE_SIZE (mm)
    spacing: float     # HYBRIDNET.GRID_SPACING (mm)

    @property
    def G(self):       # fine grid side
        return int(self.roi / self.spacing)

    @property
    def h(self):       # coarse grid side == V2V output side
        return int(self.G / 2)

    @property
    def hm(self):      # un-padded heat-map side
        return self.bbox // 2

    @property
    def hs(self):      # padded heat-map side (repro_layer.py:37)
        return int(self.bbox / 2 + 2)


# BASELINE.json configs (SURVEY.md §8 shorthand)
EXAMPLE = Shape3D(ncam=12, K=23, bbox=256, roi=144, spacing=2)      # Ex : hs=130, G=72
MICRO = Shape3D(ncam=12, K=23, bbox=512, roi=128, spacing=2)        # µB : hs=258, G=64
STRESS = Shape3D(ncam=16, K=23, bbox=512, roi=96, spacing=1)        # St : hs=258, G=96
TINY = Shape3D(ncam=4, K=5, bbox=64, roi=48, spacing=2)             # test size: hs=34, G=24
SMALL = Shape3D(ncam=6, K=23, bbox=128, roi=80, spacing=2)          # test size: hs=66, G=40

IMG_W, IMG_H = 1280, 1024


def make_rig(ncam, seed=0):
    """Cameras on a sphere of radius 1.0-1.3 m looking at the origin.

    Returns float32 arrays cameraMatrices [ncam,4,3], intrinsicMatrices [ncam,3,3],
    distortionCoefficients [ncam,1,5]."""
    rng = np.random.default_rng(seed)
    cam = np.zeros((ncam, 4, 3), np.float32)
    intr = np.zeros((ncam, 3, 3), np.float32)
    dist = np.zeros((ncam, 1, 5), np.float32)
    for c in range(ncam):
        az = 2 * np.pi * (c + 0.3 * rng.random()) / ncam
        el = np.deg2rad(rng.uniform(15, 60)) * (1 if c % 2 == 0 else -0.4)
        rad = rng.uniform(1000.0, 1300.0)
        C = rad * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
        zc = -C / np.linalg.norm(C)
        up = np.array([0.0, 0.0, 1.0])
        xc = np.cross(zc, up)
        xc /= np.linalg.norm(xc)
        yc = np.cross(zc, xc)
        roll = rng.uniform(-0.2, 0.2)
        xr = np.cos(roll) * xc + np.sin(roll) * yc
        yr = -np.sin(roll) * xc + np.cos(roll) * yc
        R = np.stack([xr, yr, zc], axis=1)            # x_cam = (X - C) @ R   (row vectors)
        T = -C @ R
        f = rng.uniform(750.0, 1400.0)
        cx = IMG_W / 2 + rng.uniform(-20, 20)
        cy = IMG_H / 2 + rng.uniform(-20, 20)
        Kt = np.array([[f, 0, 0], [0, f, 0], [cx, cy, 1.0]])
        intr[c] = Kt.astype(np.float32)
        RT = np.concatenate([R, T[None]], 0).astype(np.float32)
        cam[c] = RT @ intr[c]                          # float32 product, like the reference
        dist[c, 0, 0] = rng.uniform(-0.08, 0.02)
        dist[c, 0, 1] = rng.uniform(-0.05, 0.06)
        dist[c, 0, 2:] = rng.uniform(-1e-3, 1e-3, 3)   # present but ignored by the path
    return cam, intr, dist


def project(points, cam, intr, dist):
    """Full-resolution pixel coordinates [ncam,N,2] of world points [N,3] (float64 math;
    same model as jarvis/utils/reprojection.py:49-66)."""
    P = np.concatenate([points, np.ones((len(points), 1))], 1)
    uvw = np.einsum("nk,ckj->cnj", P, cam.astype(np.float64))
    cx = intr[:, 2, 0, None].astype(np.float64)
    cy = intr[:, 2, 1, None].astype(np.float64)
    fx = intr[:, 0, 0, None].astype(np.float64)
    fy = intr[:, 1, 1, None].astype(np.float64)
    a = uvw[..., 0] / uvw[..., 2] - cx
    b = uvw[..., 1] / uvw[..., 2] - cy
    r2 = (a / fx) ** 2 + (b / fy) ** 2
    d = 1 + (dist[:, 0, 0, None] + dist[:, 0, 1, None] * r2) * r2
    return np.stack([a * d + cx, b * d + cy], -1)


def make_frameset(shape, cam, intr, dist, seed=0, noise=2.0, dtype=np.float32):
    """One frame set: un-padded heat maps [ncam,K,hm,hm], center3D [3] i32, centerHM [ncam,2] i32,
    and the ground-truth key points [K,3] (mm) the maps were rendered from."""
    rng = np.random.default_rng(1000 + seed)
    centre = rng.uniform(-100.0, 100.0, 3)
    kps = centre + rng.uniform(-0.3, 0.3, (shape.K, 3)) * shape.roi
    center3D = centre.astype(np.int32)                       # trunc, like .int() (jarvis3D.py:183)
    chm = project(centre[None], cam, intr, dist)[:, 0, :].astype(np.int32)   # jarvis3D.py:161-162
    half = shape.bbox // 2
    chm[:, 0] = np.clip(chm[:, 0], half, IMG_W - half)       # jarvis3D.py:163-166
    chm[:, 1] = np.clip(chm[:, 1], half, IMG_H - half)
    px = project(kps, cam, intr, dist)                       # [ncam,K,2] full-res
    hm = shape.hm
    sigma = 1.5 * hm / 64.0                                  # dataset2D.py:289-300 scaled to map size
    loc = (px - (chm[:, None, :] - half)) / 2.0              # heat-map pixel units
    ys = np.arange(hm)[None, None, :, None]
    xs = np.arange(hm)[None, None, None, :]
    g = 255.0 * np.exp(-((xs - loc[..., 0, None, None]) ** 2 + (ys - loc[..., 1, None, None]) ** 2)
                       / (2 * sigma ** 2))
    g = g + rng.normal(0.0, noise, g.shape)
    return g.astype(dtype), center3D, chm.astype(np.int32), kps


def to_cl16(hm, scale=0.0625, dtype=np.float16, pitch=24):
    """Host-side producer of the gather-native heat-map layout (include/jarvis_hybridnet_b200.h, JHN_HM_F16_CL): un-padded
    planar maps [..., ncam, K, S, S] -> channels-last [..., ncam, S+2, S+2, 24] with the F.pad border
    (jarvis/hybridnet/model.py:65-66) materialised and the values scaled by 2^-4."""
    hm = np.asarray(hm)
    K, S = hm.shape[-3], hm.shape[-1]
    out = np.zeros(hm.shape[:-3] + (S + 2, S + 2, pitch), dtype)
    out[..., 1:-1, 1:-1, :K] = (np.moveaxis(hm, -3, -1) * np.float32(scale)).astype(dtype)
    return out


V2V_LAYERS = (
    # (state_dict prefix under v2vNet., kind, cin_mult, cout_mult, kernel)   v2vnet.py:62-96
    ("front_layers.0.block.0", "conv", 1, 2, 3),
    ("front_layers.1.res_branch.0", "conv", 2, 2, 3),
    ("front_layers.1.res_branch.3", "conv", 2, 2, 3),
    ("encoder_decoder.encoder_pool1.block.0", "conv", 2, 4, 2),
    ("encoder_decoder.mid_res.res_branch.0", "conv", 4, 4, 3),
    ("encoder_decoder.mid_res.res_branch.3", "conv", 4, 4, 3),
    ("encoder_decoder.decoder_upsample1.block.0", "convT", 4, 2, 2),
    ("encoder_decoder.decoder_res1.res_branch.0", "conv", 2, 2, 3),
    ("encoder_decoder.decoder_res1.res_branch.3", "conv", 2, 2, 3),
    ("encoder_decoder.skip_res1.res_branch.0", "conv", 2, 2, 3),
    ("encoder_decoder.skip_res1.res_branch.3", "conv", 2, 2, 3),
    ("output_layer", "conv", 2, 1, 1),
)


def make_v2v_weights(K, seed=0, scale="he"):
    """Random V2VNet parameters with the reference's names/shapes (v2vnet.py:86-112; checkpoint layout
    SURVEY.md §9.2): Conv3d [Cout,Cin,k,k,k], ConvTranspose3d [Cin,Cout,k,k,k].
    scale="ref" reproduces the reference init N(0, 0.001) with zero bias; scale="he" uses
    N(0, sqrt(2/fan_in)) and N(0, 0.1) biases so activations have a realistic dynamic range."""
    rng = np.random.default_rng(7000 + seed)
    sd = {}
    for name, kind, cim, com, k in V2V_LAYERS:
        cin, cout = cim * K, com * K
        shp = (cout, cin, k, k, k) if kind == "conv" else (cin, cout, k, k, k)
        if scale == "ref":
            w = rng.normal(0.0, 0.001, shp)
            b = np.zeros(cout)
        else:
            w = rng.normal(0.0, np.sqrt(2.0 / (cin * k ** 3)), shp)
            b = rng.normal(0.0, 0.1, cout)
        sd[name + ".weight"] = w.astype(np.float32)
        sd[name + ".bias"] = b.astype(np.float32)
    return sd


def make_center_case(ncam, cam, intr, dist, seed=0, cdis=256, noise=2.0, n_weak=0, small_images=False):
    """Inputs of the predictor glue (jarvis/prediction/jarvis3D.py:143-178) for one frame set:
    centre-detect heat maps [ncam,1,cdis/2,cdis/2] (a Gaussian of amplitude 255, sigma 2 px at the projected
    centre, plus N(0, noise); the last `n_weak` cameras only reach amplitude 30, i.e. stay under the reference's
    detection threshold of 50), full images [ncam,3,H,W] fp32 in [0,1] and the true centre [3] (mm).
    `small_images`: 320 x 256 images instead of 1280 x 1024 (same geometry scaled by 1/4) for cheap fixtures."""
    rng = np.random.default_rng(5000 + seed)
    centre = rng.uniform(-100.0, 100.0, 3)
    W, H = (IMG_W // 4, IMG_H // 4) if small_images else (IMG_W, IMG_H)
    px = project(centre[None], cam, intr, dist)[:, 0, :]             # full-resolution pixels
    hc = cdis // 2
    loc = px / (np.array([W / cdis, H / cdis]) * 2.0)                # jarvis3D.py:157-159 inverted
    ys = np.arange(hc)[None, :, None]
    xs = np.arange(hc)[None, None, :]
    amp = np.full(ncam, 255.0)
    if n_weak:
        amp[ncam - n_weak:] = 30.0
    g = amp[:, None, None] * np.exp(-((xs - loc[:, 0, None, None]) ** 2 + (ys - loc[:, 1, None, None]) ** 2) / (2 * 2.0 ** 2))
    g = g + rng.normal(0.0, noise, g.shape)
    imgs = rng.random((ncam, 3, H, W), dtype=np.float32)
    return g[:, None].astype(np.float32), imgs, centre
