// Packed V2VNet weights (opaque to callers as `jhn_v2v`) and the two forward paths.
#pragma once
#include <mutex>

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
    // when a workspace is first seen or its shape changes.  Off by default: the caller must promise that nobody
    // else writes the workspace between calls.  The cache is a small table keyed by workspace pointer behind a
    // mutex, so one handle may serve several streams / threads as long as each has its own workspace.
    mutable std::mutex mu;
    mutable int ws_persistent;
    struct BorderKey { const void *base; unsigned long long sig; };
    mutable BorderKey zc[16];
    mutable int zc_n;
    // One entry per workspace BASE pointer; `sig` encodes everything that decides where tensors lie inside it.  True iff
    // the workspace was last used with the same carving (its zero borders are intact); records the carving otherwise —
    // a different carving of the same memory wrote activations over the old borders, so everything is re-zeroed.
    static unsigned long long border_sig(int kind, int B, int G, int ncam, int hs)
    {
        return ((unsigned long long)(kind & 0xf) << 60) | ((unsigned long long)(B & 0xfffff) << 40) |
               ((unsigned long long)(G & 0x3ff) << 30) | ((unsigned long long)(ncam & 0xfff) << 18) | (unsigned long long)(hs & 0x3ffff);
    }
    bool borders_cached(const void *base, unsigned long long sig) const
    {
        std::lock_guard<std::mutex> g(mu);
        if (!ws_persistent) return false;
        for (int i = 0; i < zc_n; ++i)
            if (zc[i].base == base) {
                const bool ok = zc[i].sig == sig;
                zc[i].sig = sig;
                return ok;
            }
        if (zc_n == 16) zc_n = 0;                                // table full: forget everything (costs one re-zeroing each)
        zc[zc_n++] = BorderKey{base, sig};
        return false;
    }
    void borders_forget() const { std::lock_guard<std::mutex> g(mu); zc_n = 0; }
};

namespace jhn {
// centroid tail fused into the bf16 output layer (head_tc.cu); `acc` is head_acc_bytes(B, K) of scratch
struct TailArgs {
    float spacing, roi;
    const float *center3D;
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
