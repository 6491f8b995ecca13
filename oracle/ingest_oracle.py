"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, separately rounded fp32 ops) of the data formats either side of the
3D path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

Follows:
  jarvis/prediction/predict3D.py:79          from_numpy(imgs_orig).cuda().float().permute(0,3,1,2)[:, [2,1,0]] / 255.
  jarvis/prediction/jarvis3D.py:168-177      crop around centerHM, (img - mean) / std
  jarvis/efficienttrack/model.py:89-95,127   res2 = deconv1(res1): ConvTranspose2d(C, K, 4, stride 2, padding 1, bias=False)
  jarvis/hybridnet/model.py:65-66            heatmaps_padded = F.pad(heatmaps, [1,1,1,1])
  jarvis/hybridnet/model.py:73,88            heatmap_final = softplus(softplus(v2v))
Pinned in tests/test_ingest_oracle.py against the reference's own statements (torch CPU ops: the reference IS those torch
calls) on seeded inputs, and against the reference's real `deconv1` weights of the bundled MonkeyHand checkpoint when
baseline/_ref is present.
"""
import numpy as np

f32 = np.float32


def ingest_frames(frames, cuda_scalar_division=True):
    """uint8 [N,H,W,3] BGR -> fp32 [N,3,H,W] RGB in [0,1].  predict3D.py:79 runs on the GPU, where ATen divides by a Python
    scalar as `x * (1.f / 255.f)` (aten/src/ATen/native/cuda/BinaryDivTrueKernel.cu); cuda_scalar_division=False gives the
    IEEE quotient ATen's CPU kernel returns (the two differ in the last place for some byte values)."""
    x = frames.astype(f32).transpose(0, 3, 1, 2)[:, [2, 1, 0]]
    if cuda_scalar_division:
        return (x * (f32(1.) / f32(255.))).astype(f32)
    return (x / f32(255.)).astype(f32)


def crop_normalize_u8(frames, centerHM, valid, bbox, mean, std):
    """frames uint8 [B,ncam,H,W,3]; centerHM [B,ncam,2] (x, y); valid [B] -> fp32 [B,ncam,3,bbox,bbox]."""
    B, ncam, H, W, _ = frames.shape
    hw = bbox // 2
    out = np.zeros((B, ncam, 3, bbox, bbox), f32)
    m = np.asarray(mean, f32).reshape(3, 1, 1)
    s = np.asarray(std, f32).reshape(3, 1, 1)
    for b in range(B):
        if not valid[b]:
            continue
        img = ingest_frames(frames[b])
        for c in range(ncam):
            cx, cy = int(centerHM[b, c, 0]), int(centerHM[b, c, 1])
            crop = img[c, :, cy - hw:cy + hw, cx - hw:cx + hw]
            out[b, c] = (((crop - m).astype(f32)) / s).astype(f32)
    return out


def efftrack_head(features, weight):
    """ConvTranspose2d(k=4, s=2, p=1, no bias): features [N,C,Hq,Wq], weight [C,K,4,4] -> [N,K,2Hq,2Wq].
    out[n,k,2*iy-1+ky,2*ix-1+kx] += in[n,c,iy,ix] * w[c,k,ky,kx]; accumulated in float64 and rounded once (the GPU kernel
    and cuDNN differ in summation order; the tests hold both to this within fp32 accumulation error)."""
    N, C, Hq, Wq = features.shape
    K = weight.shape[1]
    full = np.zeros((N, K, 2 * Hq + 2, 2 * Wq + 2), np.float64)           # index = output + 1 (padding 1 cropped below)
    x = features.astype(np.float64)
    w = weight.astype(np.float64)
    for ky in range(4):
        for kx in range(4):
            contrib = np.einsum("nchw,ck->nkhw", x, w[:, :, ky, kx])
            full[:, :, ky:ky + 2 * Hq:2, kx:kx + 2 * Wq:2] += contrib
    return full[:, :, 1:-1, 1:-1].astype(f32)


def to_channels_last(hm, bf16=False, pitch=24, scale=0.0625):
    """fp32 [N,K,S,S] -> the gather's channels-last 16-bit layout [N,S+2,S+2,24] with the F.pad border, returned as fp32
    VALUES of the stored numbers (fp16 of hm/16, or bf16 of hm), round to nearest even."""
    N, K, S, _ = hm.shape
    out = np.zeros((N, S + 2, S + 2, pitch), f32)
    v = hm.transpose(0, 2, 3, 1)
    if bf16:
        u = v.astype(f32).view(np.uint32).astype(np.uint64)
        r = ((u + 0x7fff + ((u >> 16) & 1)) >> 16 << 16).astype(np.uint32)
        q = r.view(f32)
    else:
        q = (v * f32(scale)).astype(f32).astype(np.float16).astype(f32)
    out[:, 1:-1, 1:-1, :K] = q
    return out


def pad_heatmaps(hm):
    out = np.zeros(hm.shape[:-2] + (hm.shape[-2] + 2, hm.shape[-1] + 2), f32)
    out[..., 1:-1, 1:-1] = hm
    return out


def softplus2(v):
    """torch.nn.Softplus (beta 1, threshold 20) twice, in float64 rounded once per application."""
    def sp(x):
        x64 = x.astype(np.float64)
        return np.where(x > f32(20.), x, np.log1p(np.exp(np.minimum(x64, 50.))).astype(f32)).astype(f32)
    return sp(sp(v.astype(f32)))
