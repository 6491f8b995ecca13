"""Output format of `jarvis predict predict3D`: `data3D.csv` as
jarvis/prediction/predict3D.py:64-70 (header), :87-97 (rows) and :141-146 (create_header) write it, and `info.yaml`
(:148-155).  The reference formats and writes one row per frame inside the prediction loop, with a device->host read
(`.tolist()`, `.cpu()`) per frame; here the [N,K,4] result tensor of a whole run (model.gather_results) is read back
once and formatted afterwards, off the GPU's critical path.  The bytes are the reference's:

  * x, y, z come from `point.tolist()` of an fp32 tensor -> Python floats -> csv writes repr(float), i.e. the
    shortest round-trip decimal of the fp32 value widened to double (e.g. 12.345678329467773);
  * the confidence is a numpy float32 scalar -> csv writes str(np.float32), the shortest round-trip decimal of the
    fp32 value itself (e.g. 0.98765);
  * a frame whose centre was seen by fewer than two cameras is a row of 4*K 'NaN' strings.
"""
import csv
import itertools
import os

import numpy as np


def create_header(writer, keypoint_names):
    """predict3D.py:141-146."""
    joints = list(itertools.chain.from_iterable(itertools.repeat(x, 4) for x in keypoint_names))
    coords = ['x', 'y', 'z', 'confidence'] * len(keypoint_names)
    writer.writerow(joints)
    writer.writerow(coords)


def write_data3D_csv(output_dir, results, valid=None, keypoint_names=None):
    """results: [N,K,4] (x, y, z, confidence) fp32 — torch tensor (any device) or numpy; valid: optional [N] (0 = the
    predictor returned None for that frame).  Writes <output_dir>/data3D.csv and returns its path."""
    if hasattr(results, "detach"):
        results = results.detach().float().cpu().numpy()          # the one device->host read of the run
    results = np.ascontiguousarray(results, dtype=np.float32)
    N, K, _ = results.shape
    if valid is None:
        valid = np.ones(N, bool)
    elif hasattr(valid, "detach"):
        valid = valid.detach().cpu().numpy()
    valid = np.asarray(valid).astype(bool)
    path = os.path.join(output_dir, 'data3D.csv')
    with open(path, 'w', newline='') as f:
        writer = csv.writer(f, delimiter=',', quotechar='"', quoting=csv.QUOTE_MINIMAL)
        if keypoint_names is not None and len(keypoint_names) == K:      # predict3D.py:68-70
            create_header(writer, keypoint_names)
        nan_row = ['NaN'] * (K * 4)
        xyz = results[:, :, :3].astype(np.float64)                       # fp32 widened exactly, as Tensor.tolist() does
        for n in range(N):
            if not valid[n]:
                writer.writerow(nan_row)
                continue
            row = []
            for k in range(K):
                row += xyz[n, k].tolist() + [results[n, k, 3]]
            writer.writerow(row)
    return path


def create_info_file(output_dir, recording_path, dataset_name, frame_start, number_frames):
    """predict3D.py:148-155 (ruamel round-trip dump of a flat dict == these four `key: value` lines)."""
    path = os.path.join(output_dir, 'info.yaml')
    with open(path, 'w') as f:
        for k, v in (('recording_path', recording_path), ('dataset_name', dataset_name), ('frame_start', frame_start),
                     ('number_frames', number_frames)):
            f.write(f"{k}: {v}\n")
    return path
