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
    // Zero-border cache (jhn_v2v_set_workspace_persistent): the padded bf16 tensors keep their zero borders from
    // one forward to the next because every kernel writes zeros (or nothing) there, so they are cleared only
    // when the workspace pointer or the shape changes.  Off by default: the caller must promise that nobody
    // else writes the workspace between calls.
    mutable int ws_persistent;
    mutable const void *z_ws; mutable int z_B, z_G, z_kind;       // tensors of tc_forward
    mutable const void *zv_ptr; mutable int zv_B, zv_G;           // V2V-layout volume written by the reprojection stage
};

namespace jhn {
// centroid tail fused into the bf16 output layer (head_tc.cu); `acc` is head_acc_bytes(B, K) of scratch
struct TailArgs {
    float spacing, roi;
    const int32_t *center3D;
    float *points, *conf;
    int32_t *argmax;
    void *acc;
};
size_t head_acc_bytes(int B, int K);
void layer_table(int K, LayerDesc *d);
size_t v2v_f32_workspace(const jhn_v2v *net, int B, int G);
int v2v_f32_forward(const jhn_v2v *net, const float *x, int B, int G, float *out, void *ws, size_t ws_bytes,
                    cudaStream_t st);
int v2v_f32_pack(jhn_v2v *net, const float *const *tensors, cudaStream_t st);
}  // namespace jhn
