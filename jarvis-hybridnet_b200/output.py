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


def format_rows(xyz, conf, valid=None):
    """Rows of data3D.csv (predict3D.py:87-97) for [N,K,3] points and [N,K] confidences (fp32, host)."""
    xyz32 = np.ascontiguousarray(np.asarray(xyz), dtype=np.float32)
    conf = np.ascontiguousarray(np.asarray(conf), dtype=np.float32)
    N, K = conf.shape
    valid = np.ones(N, bool) if valid is None else np.asarray(valid).astype(bool)
    nan_row = ['NaN'] * (K * 4)
    xyz64 = xyz32.astype(np.float64)                                     # fp32 widened exactly, as Tensor.tolist() does
    for n in range(N):
        if not valid[n]:
            yield nan_row
            continue
        row = []
        for k in range(K):
            row += xyz64[n, k].tolist() + [conf[n, k]]
        yield row


def write_data3D_csv(output_dir, results, valid=None, keypoint_names=None):
    """results: [N,K,4] (x, y, z, confidence) fp32 — torch tensor (any device) or numpy; valid: optional [N] (0 = the
    predictor returned None for that frame).  Writes <output_dir>/data3D.csv and returns its path."""
    if hasattr(results, "detach"):
        results = results.detach().float().cpu().numpy()          # the one device->host read of the run
    results = np.ascontiguousarray(results, dtype=np.float32)
    N, K, _ = results.shape
    if valid is not None and hasattr(valid, "detach"):
        valid = valid.detach().cpu().numpy()
    path = os.path.join(output_dir, 'data3D.csv')
    with open(path, 'w', newline='') as f:
        writer = csv.writer(f, delimiter=',', quotechar='"', quoting=csv.QUOTE_MINIMAL)
        if keypoint_names is not None and len(keypoint_names) == K:      # predict3D.py:68-70
            create_header(writer, keypoint_names)
        for row in format_rows(results[:, :, :3], results[:, :, 3], valid):
            writer.writerow(row)
    return path


_YAML_PLAIN_FIRST = set("-?:,[]{}#&*!|>'\"%@`")
_YAML_WORDS = {"", "~", "null", "Null", "NULL", "true", "True", "TRUE", "false", "False", "FALSE", "yes", "Yes", "YES", "no",
               "No", "NO", "on", "On", "ON", "off", "Off", "OFF"}


def _yaml_scalar(v):
    """One flat-mapping value as ruamel.yaml's round-trip dumper writes it: None -> empty, bool -> true/false, numbers
    plain, strings plain unless a plain scalar would be read back as something else (then single-quoted)."""
    if v is None:
        return ""
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, float)):
        return repr(v)
    s = str(v)
    plain = (s not in _YAML_WORDS and s[0] not in _YAML_PLAIN_FIRST and s == s.strip() and ": " not in s and " #" not in s
             and not s.endswith(":") and "\n" not in s)
    if plain:
        try:
            float(s)
            plain = False                                                # a numeric-looking string must stay a string
        except ValueError:
            pass
    return s if plain else "'" + s.replace("'", "''") + "'"


def create_info_file(output_dir, recording_path, dataset_name, frame_start, number_frames):
    """predict3D.py:148-155: ruamel's round-trip dump of the flat dict {recording_path, dataset_name, frame_start,
    number_frames} — one `key: value` line each, None as an empty value."""
    path = os.path.join(output_dir, 'info.yaml')
    with open(path, 'w') as f:
        for k, v in (('recording_path', recording_path), ('dataset_name', dataset_name), ('frame_start', frame_start),
                     ('number_frames', number_frames)):
            sv = _yaml_scalar(v)
            f.write(f"{k}: {sv}\n" if sv != "" else f"{k}:\n")
    return path
