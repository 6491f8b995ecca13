// Packed V2VNet weights (opaque to callers as `jhn_v2v`) and the two forward paths.
#pragma once
#include "common.cuh"

namespace jhn {

enum LayerId {   // order of jarvis_hybridnet_b200.synth.V2V_LAYERS == checkpoint order (v2vnet.py:86-96)
    L_FRONT0 = 0, L_FRONT1A, L_FRONT1B, L_POOL, L_MIDA, L_MIDB, L_UP, L_DECA, L_DECB, L_SKIPA, L_SKIPB, L_HEAD,
    NUM_LAYERS
};

struct LayerDesc { int cin, cout, ks, stride, pad, transposed; };

struct LayerF32 {            // [ceil(cout/8)][cin][taps][8] fp32, zero-padded output channels
    float *w, *bias;
};

struct TcNet;                // tensor-core (bf16) side, conv_tc.cu

}  // namespace jhn

struct jhn_v2v {
    int K, precision, device;
    jhn::LayerDesc desc[jhn::NUM_LAYERS];
    jhn::LayerF32 f32[jhn::NUM_LAYERS];
    float *blob;             // one allocation behind all fp32 packed tensors
    jhn::TcNet *tc;          // non-null iff precision == JHN_BF16
};

namespace jhn {
void layer_table(int K, LayerDesc *d);
size_t v2v_f32_workspace(const jhn_v2v *net, int B, int G);
int v2v_f32_forward(const jhn_v2v *net, const float *x, int B, int G, float *out, void *ws, size_t ws_bytes,
                    cudaStream_t st);
int v2v_f32_pack(jhn_v2v *net, const float *const *tensors, cudaStream_t st);
}  // namespace jhn
