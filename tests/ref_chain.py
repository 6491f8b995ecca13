"""Test helper: the reference's index chain (jarvis/hybridnet/repro_layer.py:40-85) restated with the same
torch operators in the same order, so it can run on ANY device.  On CPU it re-pins the oracle on the
machine the tests run on (MKL kernel selection can differ between hosts); on the B200 it exercises the
reference's real GPU libraries (cuBLAS SGEMM K=4, ATen upsample_trilinear3d CUDA kernel), which is how the
library-defined roundings of the GPU path are determined without /root/reference being present."""
import torch
import torch.nn.functional as F


def torch_chain_indices(center3D, centerHM, cam, intr, dist, G, spacing, hs, device):
    t = lambda a, dt=None: torch.as_tensor(a, dtype=dt).to(device)
    h = G // 2
    half = G // 2 // 2
    ar = torch.arange(h, dtype=torch.float32, device=device) - half
    grid = torch.stack(torch.meshgrid(ar, ar, ar, indexing="ij"), dim=3) * spacing * 2
    x = grid + t(center3D)                                               # int32 + fp32 -> fp32
    cam, intr, dist, chm = t(cam), t(intr).permute(1, 2, 0), t(dist).permute(1, 2, 0), t(centerHM).permute(1, 0)
    ncam = cam.shape[0]
    x = torch.cat((x, torch.ones(h, h, h, 1, device=device)), 3)
    pa = torch.matmul(x.view(1, -1, 4), cam).view(-1, h, h, h, 3).permute(1, 2, 3, 4, 0)
    v1 = pa[:, :, :, 0] / pa[:, :, :, 2] - intr[2, 0]
    v2 = pa[:, :, :, 1] / pa[:, :, :, 2] - intr[2, 1]
    r2 = torch.square(v1 / intr[0, 0]) + torch.square(v2 / intr[1, 1])
    d = 1 + (dist[0, 0] + dist[0, 1] * r2) * r2
    v1 = v1 * d + intr[2, 0]
    v2 = v2 * d + intr[2, 1]
    v1 = torch.clamp(v1, chm[0] - (hs - 1), chm[0] + hs - 2) - chm[0] + hs - 1
    v2 = torch.clamp(v2, chm[1] - (hs - 1), chm[1] + hs - 2) - chm[1] + hs - 1
    up = lambda v: F.interpolate(v.permute(3, 0, 1, 2).reshape(1, ncam, h, h, h), size=(G, G, G),
                                 mode="trilinear").view(ncam, G, G, G)
    f1, f2 = up(v1), up(v2)
    return ((f2 / 2).int() * hs + (f1 / 2).int()).long()
