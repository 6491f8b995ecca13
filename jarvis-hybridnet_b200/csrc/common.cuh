// Shared host/device helpers for the sm_100a kernels of the JARVIS-HybridNet 3D hot path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/jarvis_hybridnet_b200.h"

namespace jhn {

// thread-local last-error string + status helpers (api.cu)
int fail(int status, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void count_launch(int n = 1);
struct ProfScope {
    int slot; cudaStream_t st;
    ProfScope(const char *name, cudaStream_t s);
    void end();
};

#define JHN_CUDA(expr)                                                      \
    do {                                                                    \
        cudaError_t e__ = (expr);                                           \
        if (e__ != cudaSuccess) return ::jhn::cuda_fail(e__, #expr);        \
    } while (0)

// Every kernel launch goes through this: counts it (bench.py `gpu_launches`), checks the launch, and —
// only while jhn_profile_enable(1) is set — brackets it with CUDA events on the launching stream so
// bench.py can report per-kernel device time without a profiler attached.
#define JHN_LAUNCH(name, st, ...)                                           \
    do {                                                                    \
        ::jhn::ProfScope prof__(name, st);                                  \
        __VA_ARGS__;                                                        \
        prof__.end();                                                       \
        cudaError_t e__ = cudaGetLastError();                               \
        if (e__ != cudaSuccess) return ::jhn::cuda_fail(e__, name);         \
        ::jhn::count_launch();                                              \
    } while (0)

#define JHN_TRY(expr)                                                       \
    do {                                                                    \
        int s__ = (expr);                                                   \
        if (s__ != JHN_OK) return s__;                                      \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// bump allocator over the caller's workspace (256-byte aligned slices)
struct Arena {
    char *base;
    size_t size, off;
    Arena(void *p, size_t n) : base((char *)p), size(n), off(0) {}
    template <typename T>
    T *take(size_t count)
    {
        size_t bytes = align_up(count * sizeof(T), 256);
        T *r = (T *)(base ? base + off : nullptr);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= size; }
};

constexpr int KP = 24;   // channel pitch of the channels-last heat-map staging copy (K <= 24)

// ---- stage launchers (each returns a jhn_status) -------------------------------------------------
struct ReprojectArgs {
    const void *heatmaps; int hm_format; int padded;
    const float *cam, *intr, *dist;
    const float *center3D; const int32_t *centerHM;
    int B, ncam, K, hs, G;
    float spacing; int lerp_mode; float post_divide;
    int precision, layout;
    void *volume_out; int32_t *index_out;
    int borders_valid;     // V2V-layout volume: zero border already in place (hybrid path, persistent workspace)
};
size_t reproject_workspace(int B, int ncam, int K, int hs, int G, int precision);
int reproject_launch(const ReprojectArgs &a, void *ws, size_t ws_bytes, cudaStream_t st);
int gather_set_box_bytes(int bytes);   // test hook (jhn_debug_set_gather_box_bytes)
int heatmap_convert_launch(const float *hm, int padded, int B, int ncam, int K, int hs, int dst_format, void *dst, cudaStream_t st);

// zero the never-written border of a blocked+padded (BP/PS) bf16 tensor of `chunks_total` chunk volumes (conv_tc.cu)
int tc_zero_border_launch(void *tensor, int chunks_total, int D, cudaStream_t st);

int centroid_launch(const float *v, int B, int K, int h, float spacing, float roi, const float *center3D,
                    float *points, float *conf, int32_t *argmax, cudaStream_t st);

int center_locate_launch(const float *hm, int B, int ncam, int Hc, int Wc, int img_w, int img_h, int cdis, int bbox_hw,
                         float threshold, const float *cam, const float *intr, const float *dist, int32_t *preds,
                         float *maxvals, float *center3D, int32_t *center3D_int, int32_t *centerHM, int32_t *valid,
                         void *scratch, cudaStream_t st);
int crop_normalize_launch(const float *imgs, int B, int ncam, int H, int W, int bbox, const int32_t *centerHM,
                          const int32_t *valid, const float *mean, const float *std, float *out, cudaStream_t st);

int heatmap_boxes_launch(const float *cam, const float *intr, const float *dist, const float *center3D, const int32_t *centerHM,
                         int B, int ncam, int hs, int G, float spacing, int32_t *boxes, cudaStream_t st);

void pull_set_config(int threads, int ctas, int split);
void c3_set_transfer_overlap(int on);   // conv3_tc.cu: thread-local, read when a 3x3x3 layer is launched
int pull_spans_launch(const void *host_mapped, void *dev, const int32_t *spans, int n_images, int hs, int pixel_bytes,
                      unsigned long long *bytes_out, cudaStream_t st);
int heatmap_spans_launch(const float *cam, const float *intr, const float *dist, const float *center3D, const int32_t *centerHM,
                         int B, int ncam, int hs, int G, float spacing, void *scratch, int32_t *boxes, int32_t *spans, cudaStream_t st);
int pull_segments_launch(int n, const void *const *src_mapped, void *const *dst, const size_t *bytes, cudaStream_t st);
int pull_boxes_launch(const void *host_mapped, void *dev, const int32_t *boxes, int n_images, int hs, int pixel_bytes,
                      unsigned long long *bytes_out, cudaStream_t st);
// ingest.cu (rows f4 / a11) and head2d.cu (row f2)
int ingest_frames_launch(const uint8_t *frames, int N, int H, int W, float *out, cudaStream_t st);
int crop_normalize_u8_launch(const uint8_t *frames, int B, int ncam, int H, int W, int bbox, const int32_t *centerHM,
                             const int32_t *valid, const float *mean, const float *std, float *out, cudaStream_t st);
int softplus2_launch(const float *v, long long n, float *out, cudaStream_t st);
int pad_border_launch(const float *in, long long N, int S, float *out, cudaStream_t st);
int efftrack_head_launch(const float *features, const float *weight, int N, int C, int K, int Hq, int Wq, int out_format,
                         void *heatmaps, cudaStream_t st);

}  // namespace jhn
