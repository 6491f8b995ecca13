"""B200-native (sm_100a) implementation of JARVIS-HybridNet's 3D inference hot path:
ReprojectionLayer -> V2VNet -> softplus centroid (reference: jarvis/hybridnet/model.py:65-88).

Sub-modules are imported lazily so that `synth` (numpy only) is usable without torch/CUDA."""
__version__ = "0.1.0"


def __getattr__(name):
    import importlib
    table = {"ReprojectionLayer": "repro_layer", "V2VNet": "v2vnet", "HybridNet3D": "model", "accelerate": "model",
             "centroid_tail": "model", "shard_range": "model", "gather_results": "model",
             "write_data3D_csv": "output", "create_info_file": "output",
             "locate_center": "predictor", "crop_normalize": "predictor", "accelerate_predictor": "predictor",
             "predict3D_frames": "predictor", "ingest_frames": "ingest", "crop_normalize_u8": "ingest", "EffTrackHead": "ingest",
             "FrameUploader": "ingest", "softplus2": "ingest", "pad_heatmaps": "ingest", "format_rows": "output"}
    if name in table:
        return getattr(importlib.import_module("." + table[name], __name__), name)
    raise AttributeError(name)
