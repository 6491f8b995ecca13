"""V2VNet — drop-in for jarvis.hybridnet.v2vnet.V2VNet (v2vnet.py:86-112), backed by `jhn_v2v_forward`.

The module owns parameters under exactly the checkpoint's names (`front_layers.0.block.0.weight`, ...,
`output_layer.bias`: 24 tensors) so `load_state_dict(strict=True)` of a reference `.pth` works
(hybridnet.py:89-90), but it has no torch.nn compute: `forward` hands the raw parameter pointers to the
C ABI, which packs them once per parameter version."""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from .synth import V2V_LAYERS


def _attach(root, dotted, param):
    mod = root
    parts = dotted.split(".")
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, nn.Module())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], param)


class V2VNet(nn.Module):
    def __init__(self, input_channels, output_channels, precision="fp32"):
        super().__init__()
        if input_channels != output_channels:
            raise RuntimeError("V2VNet: the hot path is built for input_channels == output_channels "
                               "(HybridNetBackbone passes NUM_JOINTS for both, model.py:39-40)")
        self.K = input_channels
        self.precision = {"fp32": _lib.FP32, "bf16": _lib.BF16}[precision]
        for name, kind, cim, com, k in V2V_LAYERS:
            cin, cout = cim * self.K, com * self.K
            shape = (cout, cin, k, k, k) if kind == "conv" else (cin, cout, k, k, k)
            w = nn.Parameter(torch.empty(shape).normal_(0, 0.001), requires_grad=False)     # v2vnet.py:105-112
            b = nn.Parameter(torch.zeros(cout), requires_grad=False)
            _attach(self, name + ".weight", w)
            _attach(self, name + ".bias", b)
        self._handle = None
        self._handle_key = None
        self._ws = None

    # ---- packed-weights handle ---------------------------------------------------------------------
    def _params_in_order(self):
        sd = dict(self.named_parameters())
        return [sd[name + suffix] for name, *_ in V2V_LAYERS for suffix in (".weight", ".bias")]

    def _get_handle(self):
        ps = self._params_in_order()
        _lib.require_cuda(*ps)
        key = tuple((p.data_ptr(), p._version) for p in ps) + (self.precision,)
        if self._handle is None or key != self._handle_key:
            self.release()
            lib = _lib.load()
            keep = [p.detach().contiguous().float() for p in ps]
            arr = (ctypes.c_void_p * len(keep))(*[p.data_ptr() for p in keep])
            out = ctypes.c_void_p()
            _lib.check(lib.jhn_v2v_create(arr, len(keep), self.K, self.precision, _lib.stream_ptr(),
                                          ctypes.byref(out)))
            torch.cuda.current_stream().synchronize()      # packing reads `keep`
            # the workspaces handed to this handle are private tensors of the wrapper modules (V2VNet._ws,
            # HybridNet3D._ws): nobody else writes them, so the zero borders may be cached between calls
            _lib.check(lib.jhn_v2v_set_workspace_persistent(out, 1))
            self._handle, self._handle_key = out, key
        return self._handle

    def release(self):
        if self._handle is not None:
            _lib.load().jhn_v2v_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._ws

    def debug_layer(self, layer, x, D):
        """Test aid: run convolution `layer` (index into V2V_LAYERS) alone on the tensor-core path.
        x: fp32 NCDHW input of that layer; D: output grid side (input side for the transposed conv).
        Returns the raw fp32 NCDHW output (bias added, no InstanceNorm)."""
        _lib.require_cuda(x)
        lib = _lib.load()
        h = self._get_handle()
        name, kind, cim, com, k = V2V_LAYERS[layer]
        B = x.shape[0]
        Dout = 2 * D if kind == "convT" else D
        need = _lib.c_size_t()
        _lib.check(lib.jhn_v2v_debug_layer_workspace_bytes(h, layer, B, D, need))
        ws = torch.empty(need.value, dtype=torch.uint8, device=x.device)
        out = torch.zeros((B, com * self.K, Dout, Dout, Dout), dtype=torch.float32, device=x.device)
        xin = x.contiguous().float()
        _lib.check(lib.jhn_v2v_debug_layer(h, layer, _lib.dptr(xin), B, D, _lib.dptr(out), _lib.dptr(ws), ws.numel(),
                                           _lib.stream_ptr()))
        return out

    def forward(self, x):
        """x [B,K,G,G,G] fp32 (already /255, model.py:72) -> [B,K,G/2,G/2,G/2] fp32."""
        _lib.require_cuda(x)
        lib = _lib.load()
        B, K, G = x.shape[0], x.shape[1], x.shape[2]
        if K != self.K or x.shape[3] != G or x.shape[4] != G:
            raise RuntimeError(f"V2VNet expects [B,{self.K},G,G,G], got {tuple(x.shape)}")
        h = self._get_handle()
        need = _lib.c_size_t()
        _lib.check(lib.jhn_v2v_workspace_bytes(h, B, G, need))
        ws = self._workspace(need.value, x.device)
        xin = x.contiguous().float()
        out = torch.empty((B, K, G // 2, G // 2, G // 2), dtype=torch.float32, device=x.device)
        _lib.check(lib.jhn_v2v_forward(h, _lib.dptr(xin), _lib.VOL_NCDHW_F32, B, G, _lib.dptr(out), _lib.dptr(ws),
                                       ws.numel(), _lib.stream_ptr()))
        return out
