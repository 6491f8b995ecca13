"""data3D.csv: byte-for-byte what the reference's prediction loop writes.  The expected bytes are
produced here with the reference's own row-building statements (jarvis/prediction/predict3D.py:68-70, 87-97, 141-146,
quoted in the comments) applied to CPU tensors — they are three lines of csv / torch code, not an algorithm."""
import csv
import io
import itertools
import os

import numpy as np
import torch


def reference_bytes(points, confs, valid, names):
    """The reference loop, verbatim statements on CPU tensors."""
    buf = io.StringIO(newline='')
    writer = csv.writer(buf, delimiter=',', quotechar='"', quoting=csv.QUOTE_MINIMAL)           # predict3D.py:66-67
    K = points.shape[1]
    if len(names) == K:                                                                          # :69-70
        joints = list(itertools.chain.from_iterable(itertools.repeat(x, 4) for x in names))     # :142-146
        coords = ['x', 'y', 'z', 'confidence'] * len(names)
        writer.writerow(joints)
        writer.writerow(coords)
    for n in range(points.shape[0]):
        points3D_net = points[n:n + 1] if valid[n] else None
        confidences = confs[n:n + 1]
        if points3D_net != None:                                                                 # noqa: E711  (:87)
            row = []
            for point, conf in zip(points3D_net.squeeze(), confidences.squeeze().cpu().numpy()):  # :89-90
                row = row + point.tolist() + [conf]
            writer.writerow(row)
        else:
            row = []
            for i in range(K * 4):                                                               # :93-96
                row = row + ['NaN']
            writer.writerow(row)
    return buf.getvalue()


def test_data3D_csv_matches_reference_loop(tmp_path):
    from jarvis_hybridnet_b200 import write_data3D_csv
    g = torch.Generator().manual_seed(0)
    N, K = 7, 23
    pts = (torch.randn(N, K, 3, generator=g) * 80).float()
    conf = torch.rand(N, K, generator=g).float()
    pts[0, 0] = torch.tensor([1.0, -0.0, 1e-7]); conf[0, 0] = 1.0          # integers, negative zero, tiny values
    valid = np.array([1, 1, 0, 1, 1, 0, 1])
    names = ["kp%d" % i for i in range(K)]
    res = torch.cat([pts, conf[..., None]], 2)
    path = write_data3D_csv(str(tmp_path), res, valid, names)
    assert open(path, newline='').read() == reference_bytes(pts, conf, valid, names)
    # no names configured -> no header (predict3D.py:69)
    path = write_data3D_csv(str(tmp_path), res.numpy(), None, [])
    assert open(path, newline='').read() == reference_bytes(pts, conf, np.ones(N), [])


def test_info_yaml(tmp_path):
    from jarvis_hybridnet_b200 import create_info_file
    p = create_info_file(str(tmp_path), "/data/rec1", "Example_Dataset", 0, 100)
    assert open(p).read() == "recording_path: /data/rec1\ndataset_name: Example_Dataset\nframe_start: 0\nnumber_frames: 100\n"
