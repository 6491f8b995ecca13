"""B200-native (sm_100a) implementation of JARVIS-HybridNet's 3D inference hot path:
ReprojectionLayer -> V2VNet -> softplus centroid (reference: jarvis/hybridnet/model.py:65-88).

Sub-modules are imported lazily so that `synth` (numpy only) is usable without torch/CUDA."""
__version__ = "0.1.0"
