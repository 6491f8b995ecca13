// extern "C" surface declared in include/jarvis_hybridnet_b200.h: argument validation, workspace
// carving and stage sequencing.  No computation lives here.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "v2v.cuh"

namespace jhn {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int fail(int status, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

int cuda_fail(cudaError_t e, const char *what)
{
    return fail(JHN_ERR_CUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// ---- optional CUDA-event profiler (off by default; never active during graph capture) -------------
struct ProfRec { const char *name; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::atomic<int> g_prof_on{0};

ProfScope::ProfScope(const char *name, cudaStream_t s) : slot(-1), st(s)
{
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, s);
    std::lock_guard<std::mutex> g(g_prof_mu);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}
void ProfScope::end()
{
    if (slot < 0) return;
    std::lock_guard<std::mutex> g(g_prof_mu);
    cudaEventRecord(g_prof[slot].b, st);
}

// tensor-core side (conv_tc.cu)
int tc_create(jhn_v2v *net, const float *const *tensors, cudaStream_t st);
void tc_destroy(jhn_v2v *net);
size_t tc_workspace(const jhn_v2v *net, int B, int G);
int tc_forward(const jhn_v2v *net, const void *volume_in, int in_layout, int B, int G, float *out, void *ws,
               size_t ws_bytes, cudaStream_t st, const TailArgs *tail, int carveB, int borders);
int tc_debug_head(const jhn_v2v *net, const float *in, int B, int h, const TailArgs &tail, void *ws, size_t ws_bytes, cudaStream_t st);
bool head_supported(int K, int cin_pad, int cout_pad, int D);

int tc_debug_layer(const jhn_v2v *net, int l, const float *in, int B, int D, float *out, void *ws, size_t ws_bytes, cudaStream_t st);
size_t tc_debug_workspace(const jhn_v2v *net, int l, int B, int D);

static int check_repro_shape(int B, int ncam, int K, int hs, int G)
{
    if (B < 1 || ncam < 1 || K < 1 || K > KP) return fail(JHN_ERR_SHAPE, "need B>=1, ncam>=1, 1<=K<=%d (got B=%d ncam=%d K=%d)", KP, B, ncam, K);
    if (hs < 4 || hs > 1024) return fail(JHN_ERR_SHAPE, "padded heat-map side %d out of range [4,1024]", hs);
    if (G < 4 || (G % 4) != 0 || G > 256) return fail(JHN_ERR_SHAPE, "grid side %d must be a multiple of 4 in [4,256]", G);
    return JHN_OK;
}

}  // namespace jhn

using namespace jhn;

extern "C" {

const char *jhn_last_error(void) { return g_err; }
int jhn_abi_version(void) { return 6; }
unsigned long long jhn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void jhn_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }
int jhn_debug_set_gather_box_bytes(int bytes) { return gather_set_box_bytes(bytes); }
void jhn_debug_set_pull_config(int threads, int ctas, int split) { pull_set_config(threads, ctas, split); }
void jhn_set_transfer_overlap(int on) { c3_set_transfer_overlap(on); }

// Frame sets per internal pass of jhn_hybrid3d_forward.  Default: the whole batch in one pass.  Passes of 8 keep a
// convolution's input + output (2 x 42 MB at the Example shape) inside the 126 MB L2, but measured on the B200 that buys
// nothing (profiles/r02_run1_subbatch.txt: 32 -> 3.36 ms, 16 -> 3.58, 8 -> 4.02, 4 -> 4.91 ms per 32 frame sets): the
// persistent kernels lose more to their shorter tile ranges than the normalisation passes gain.  The knob stays for
// callers that must bound the workspace.  Frame sets are independent; on the bf16 path the split changes how a sample's
// InstanceNorm partial sums are grouped (fp32 re-association, <= 5e-3 mm on the key points; bit-identical for equal splits).
static std::atomic<int> g_sub_batch{1 << 20};
int jhn_set_sub_batch(int n)
{
    if (n <= 0) n = 1 << 20;
    g_sub_batch.store(n, std::memory_order_relaxed);
    return n;
}

// Synchronises the device, then writes one line per kernel name: "<name>\t<launches>\t<total_ms>\n".
// Returns the number of bytes written (truncated to `cap`), and clears the records.
int jhn_profile_collect(char *buf, int cap)
{
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> g(g_prof_mu);
    std::map<std::string, std::pair<int, double>> acc;
    for (auto &r : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { auto &e = acc[r.name]; e.first += 1; e.second += ms; }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
    std::string out;
    for (auto &kv : acc) { char line[256]; snprintf(line, sizeof(line), "%s\t%d\t%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second); out += line; }
    int n = (int)out.size() < cap - 1 ? (int)out.size() : (cap > 0 ? cap - 1 : 0);
    if (buf && cap > 0) { memcpy(buf, out.data(), n); buf[n] = 0; }
    return n;
}

int jhn_check_device(int device)
{
    cudaDeviceProp p;
    JHN_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) return fail(JHN_ERR_ARCH, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, p.major, p.minor);
    return JHN_OK;
}

int jhn_reproject_workspace_bytes(int B, int ncam, int K, int hs, int G, int precision, size_t *bytes)
{
    if (!bytes) return fail(JHN_ERR_ARG, "bytes is null");
    JHN_TRY(check_repro_shape(B, ncam, K, hs, G));
    *bytes = reproject_workspace(B, ncam, K, hs, G, precision);
    return JHN_OK;
}

int jhn_reproject_gather(const void *heatmaps, int hm_format, int heatmaps_padded, const float *cameraMatrices,
                         const float *intrinsicMatrices, const float *distortionCoefficients,
                         const float *center3D, const int32_t *centerHM, int B, int ncam, int K, int hs, int G,
                         float spacing, int lerp_mode, float post_divide, int precision, int layout,
                         void *volume_out, int32_t *index_out, void *workspace, size_t workspace_bytes,
                         jhn_stream_t stream)
{
    if (!heatmaps || !cameraMatrices || !intrinsicMatrices || !distortionCoefficients || !center3D || !centerHM ||
        !volume_out || !workspace)
        return fail(JHN_ERR_ARG, "jhn_reproject_gather: null pointer argument");
    JHN_TRY(check_repro_shape(B, ncam, K, hs, G));
    if (lerp_mode < 0 || lerp_mode > 2) return fail(JHN_ERR_ARG, "lerp_mode %d not in {0,1,2}", lerp_mode);
    if (hm_format < JHN_HM_F32_PLANAR || hm_format > JHN_HM_BF16_CL) return fail(JHN_ERR_ARG, "hm_format %d unknown", hm_format);
    if (precision != JHN_FP32 && precision != JHN_BF16) return fail(JHN_ERR_ARG, "precision %d unknown", precision);
    if (layout != JHN_VOL_NCDHW_F32 && layout != JHN_VOL_V2V_BF16) return fail(JHN_ERR_ARG, "layout %d unknown", layout);
    if (!(post_divide > 0.f)) return fail(JHN_ERR_ARG, "post_divide must be > 0");
    if (((uintptr_t)workspace & 255) != 0) return fail(JHN_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    ReprojectArgs a{heatmaps, hm_format, heatmaps_padded ? 1 : 0, cameraMatrices, intrinsicMatrices, distortionCoefficients,
                    center3D, centerHM, B, ncam, K, hs, G, spacing, lerp_mode, post_divide, precision, layout,
                    volume_out, index_out, 0};
    return reproject_launch(a, workspace, workspace_bytes, (cudaStream_t)stream);
}

int jhn_heatmap_convert(const float *heatmaps, int heatmaps_padded, int B, int ncam, int K, int hs, int dst_format,
                        void *dst, jhn_stream_t stream)
{
    if (!heatmaps || !dst) return fail(JHN_ERR_ARG, "jhn_heatmap_convert: null pointer argument");
    if (B < 1 || ncam < 1 || K < 1 || K > KP || hs < 4 || hs > 1024)
        return fail(JHN_ERR_SHAPE, "need B>=1, ncam>=1, 1<=K<=%d, 4<=hs<=1024 (got B=%d ncam=%d K=%d hs=%d)", KP, B, ncam, K, hs);
    return heatmap_convert_launch(heatmaps, heatmaps_padded ? 1 : 0, B, ncam, K, hs, dst_format, dst, (cudaStream_t)stream);
}

int jhn_v2v_create(const float *const *tensors, int num_tensors, int K, int precision, jhn_stream_t stream,
                   jhn_v2v **out)
{
    if (!tensors || !out) return fail(JHN_ERR_ARG, "jhn_v2v_create: null pointer argument");
    if (num_tensors != 2 * NUM_LAYERS) return fail(JHN_ERR_SHAPE, "expected %d tensors (weight,bias x %d layers), got %d", 2 * NUM_LAYERS, NUM_LAYERS, num_tensors);
    if (K < 1 || K > KP) return fail(JHN_ERR_SHAPE, "K=%d out of range [1,%d]", K, KP);
    if (precision != JHN_FP32 && precision != JHN_BF16) return fail(JHN_ERR_ARG, "precision %d unknown", precision);
    for (int i = 0; i < num_tensors; ++i)
        if (!tensors[i]) return fail(JHN_ERR_ARG, "tensor %d is null", i);
    int dev = 0;
    JHN_CUDA(cudaGetDevice(&dev));
    JHN_TRY(jhn_check_device(dev));
    jhn_v2v *net = new (std::nothrow) jhn_v2v();
    if (!net) return fail(JHN_ERR_CUDA, "out of host memory");
    net->K = K; net->precision = precision; net->device = dev; net->blob = nullptr; net->tc = nullptr;
    net->ws_persistent = 0; net->zc_n = 0;
    layer_table(K, net->desc);
    int s = v2v_f32_pack(net, tensors, (cudaStream_t)stream);
    if (s == JHN_OK && precision == JHN_BF16) s = tc_create(net, tensors, (cudaStream_t)stream);
    if (s != JHN_OK) { jhn_v2v_destroy(net); return s; }
    *out = net;
    return JHN_OK;
}

int jhn_v2v_set_workspace_persistent(jhn_v2v *net, int on)
{
    if (!net) return fail(JHN_ERR_ARG, "net is null");
    { std::lock_guard<std::mutex> g(net->mu); net->ws_persistent = on ? 1 : 0; }
    net->borders_forget();
    return JHN_OK;
}

void jhn_v2v_destroy(jhn_v2v *net)
{
    if (!net) return;
    if (net->tc) tc_destroy(net);
    if (net->blob) cudaFree(net->blob);
    delete net;
}

static int check_v2v_shape(const jhn_v2v *net, int B, int G)
{
    if (!net) return fail(JHN_ERR_ARG, "net is null");
    if (B < 1) return fail(JHN_ERR_SHAPE, "B=%d must be >= 1", B);
    if (G < 4 || (G % 4) != 0 || G > 256) return fail(JHN_ERR_SHAPE, "grid side %d must be a multiple of 4 in [4,256]", G);
    return JHN_OK;
}

int jhn_v2v_workspace_bytes(const jhn_v2v *net, int B, int G, size_t *bytes)
{
    if (!bytes) return fail(JHN_ERR_ARG, "bytes is null");
    JHN_TRY(check_v2v_shape(net, B, G));
    *bytes = net->precision == JHN_FP32 ? v2v_f32_workspace(net, B, G) : tc_workspace(net, B, G);
    return JHN_OK;
}

int jhn_v2v_forward(const jhn_v2v *net, const void *volume_in, int in_layout, int B, int G, float *out,
                    void *workspace, size_t workspace_bytes, jhn_stream_t stream)
{
    JHN_TRY(check_v2v_shape(net, B, G));
    if (!volume_in || !out || !workspace) return fail(JHN_ERR_ARG, "jhn_v2v_forward: null pointer argument");
    if (((uintptr_t)workspace & 255) != 0) return fail(JHN_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    if (net->precision == JHN_FP32) {
        if (in_layout != JHN_VOL_NCDHW_F32) return fail(JHN_ERR_ARG, "fp32 V2V takes the NCDHW fp32 volume");
        return v2v_f32_forward(net, (const float *)volume_in, B, G, out, workspace, workspace_bytes, (cudaStream_t)stream);
    }
    return tc_forward(net, volume_in, in_layout, B, G, out, workspace, workspace_bytes, (cudaStream_t)stream, nullptr, 0, -1);
}

int jhn_v2v_debug_layer_workspace_bytes(const jhn_v2v *net, int layer, int B, int D, size_t *bytes)
{
    if (!net || !bytes || !net->tc) return fail(JHN_ERR_ARG, "debug layer needs a JHN_BF16 network");
    if (layer < 0 || layer >= NUM_LAYERS || B < 1 || D < 2 || D > 61) return fail(JHN_ERR_SHAPE, "bad layer/B/D");
    *bytes = tc_debug_workspace(net, layer, B, D);
    return JHN_OK;
}

int jhn_v2v_debug_layer(const jhn_v2v *net, int layer, const float *in, int B, int D, float *out, void *workspace,
                        size_t workspace_bytes, jhn_stream_t stream)
{
    if (!net || !in || !out || !workspace) return fail(JHN_ERR_ARG, "jhn_v2v_debug_layer: null pointer argument");
    if (B < 1 || D < 2 || D > 61) return fail(JHN_ERR_SHAPE, "bad B/D");
    return tc_debug_layer(net, layer, in, B, D, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int jhn_v2v_debug_head_centroid(const jhn_v2v *net, const float *in, int B, int h, float spacing, float roi,
                                const float *center3D, float *points, float *conf, int32_t *argmax, void *workspace,
                                size_t workspace_bytes, jhn_stream_t stream)
{
    if (!net || !net->tc || !in || !center3D || !points || !conf || !workspace)
        return fail(JHN_ERR_ARG, "jhn_v2v_debug_head_centroid: null pointer argument / not a JHN_BF16 network");
    if (B < 1 || h < 2 || h > 61) return fail(JHN_ERR_SHAPE, "bad B/h");
    TailArgs tail{spacing, roi, center3D, points, conf, argmax, nullptr};
    return tc_debug_head(net, in, B, h, tail, workspace, workspace_bytes, (cudaStream_t)stream);
}

int jhn_centroid_reduce(const float *v2v_out, int B, int K, int h, float spacing, float roi, const float *center3D,
                        float *points, float *conf, int32_t *argmax, jhn_stream_t stream)
{
    if (!v2v_out || !center3D || !points || !conf) return fail(JHN_ERR_ARG, "jhn_centroid_reduce: null pointer argument");
    if (B < 1 || K < 1 || h < 1 || h > 256) return fail(JHN_ERR_SHAPE, "bad shape B=%d K=%d h=%d", B, K, h);
    return centroid_launch(v2v_out, B, K, h, spacing, roi, center3D, points, conf, argmax, (cudaStream_t)stream);
}

static int hybrid_layout(const jhn_v2v *net) { return net->precision == JHN_BF16 ? JHN_VOL_V2V_BF16 : JHN_VOL_NCDHW_F32; }
static size_t hybrid_volume_bytes(const jhn_v2v *net, int B, int G);

static int sub_batch_of(int B)
{
    const int n = g_sub_batch.load(std::memory_order_relaxed);
    return B < n ? B : n;
}

int jhn_hybrid3d_workspace_bytes(const jhn_v2v *net, int B, int ncam, int hs, int G, size_t *bytes)
{
    if (!bytes) return fail(JHN_ERR_ARG, "bytes is null");
    JHN_TRY(check_v2v_shape(net, B, G));
    JHN_TRY(check_repro_shape(B, ncam, net->K, hs, G));
    const int SB = sub_batch_of(B);                     // the batch is walked in passes of SB frame sets over ONE workspace
    size_t v2v = 0;
    JHN_TRY(jhn_v2v_workspace_bytes(net, SB, G, &v2v));
    const int h = G / 2;
    size_t tail = (size_t)SB * net->K * h * h * h * sizeof(float);            // fp32 V2V output, or the fused head's accumulators
    if (head_acc_bytes(SB, net->K) > tail) tail = head_acc_bytes(SB, net->K);
    *bytes = align_up(reproject_workspace(SB, ncam, net->K, hs, G, net->precision), 256) +
             align_up(hybrid_volume_bytes(net, SB, G), 256) + align_up(v2v, 256) + align_up(tail, 256);
    return JHN_OK;
}

int jhn_hybrid3d_forward(const jhn_v2v *net, const void *heatmaps, int hm_format, int heatmaps_padded, const float *cameraMatrices,
                         const float *intrinsicMatrices, const float *distortionCoefficients,
                         const float *center3D, const int32_t *centerHM, int B, int ncam, int hs, int G,
                         float spacing, float roi, int lerp_mode, float *points, float *conf, int32_t *argmax,
                         void *workspace, size_t workspace_bytes, jhn_stream_t stream)
{
    size_t need = 0;
    JHN_TRY(jhn_hybrid3d_workspace_bytes(net, B, ncam, hs, G, &need));
    if (!workspace || workspace_bytes < need) return fail(JHN_ERR_WORKSPACE, "hybrid3d workspace: need %zu bytes, got %zu", need, workspace_bytes);
    if (((uintptr_t)workspace & 255) != 0) return fail(JHN_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    if (!heatmaps || !cameraMatrices || !intrinsicMatrices || !distortionCoefficients || !center3D || !centerHM || !points || !conf)
        return fail(JHN_ERR_ARG, "jhn_hybrid3d_forward: null pointer argument");
    if (lerp_mode < 0 || lerp_mode > 2) return fail(JHN_ERR_ARG, "lerp_mode %d not in {0,1,2}", lerp_mode);
    if (hm_format < JHN_HM_F32_PLANAR || hm_format > JHN_HM_BF16_CL) return fail(JHN_ERR_ARG, "hm_format %d unknown", hm_format);
    const int h = G / 2, K = net->K;
    const int SB = sub_batch_of(B);
    char *p = (char *)workspace;
    const size_t rws = align_up(reproject_workspace(SB, ncam, K, hs, G, net->precision), 256);
    void *ws_r = p; p += rws;
    void *vol = p; p += align_up(hybrid_volume_bytes(net, SB, G), 256);
    size_t v2v = 0;
    JHN_TRY(jhn_v2v_workspace_bytes(net, SB, G, &v2v));
    void *ws_v = p; p += align_up(v2v, 256);
    float *vout = (float *)p;
    const int S = heatmaps_padded ? hs : hs - 2;
    const size_t hm_stride = hm_format == JHN_HM_F32_PLANAR ? (size_t)ncam * K * S * S * sizeof(float)
                                                            : (size_t)ncam * hs * hs * KP * 2;       // bytes per frame set
    const int c2 = (2 * K + 15) / 16 * 16, c1 = (K + 15) / 16 * 16;
    const bool fused_head = net->precision == JHN_BF16 && head_supported(K, c2, c1, h);
    // zero borders of every padded tensor in the workspace: intact iff this workspace was last carved the same way
    const int borders = net->precision == JHN_BF16 && net->borders_cached(workspace, jhn_v2v::border_sig(0, SB, G, ncam, hs)) ? 1 : 0;
    for (int b0 = 0; b0 < B; b0 += SB) {
        const int nb = B - b0 < SB ? B - b0 : SB;
        const size_t bc = (size_t)b0 * ncam;
        ReprojectArgs ra{(const char *)heatmaps + (size_t)b0 * hm_stride, hm_format, heatmaps_padded ? 1 : 0, cameraMatrices + bc * 12,
                         intrinsicMatrices + bc * 9, distortionCoefficients + bc * 5, center3D + (size_t)b0 * 3, centerHM + bc * 2,
                         nb, ncam, K, hs, G, spacing, lerp_mode, 255.f, net->precision, hybrid_layout(net), vol, nullptr, 0};
        // per-sample offsets of the padded layouts do not depend on the batch size, and the first pass always has nb == SB:
        // a shorter last pass finds the borders of its samples already zero
        ra.borders_valid = (borders || b0 > 0) ? 1 : 0;
        JHN_TRY(reproject_launch(ra, ws_r, rws, (cudaStream_t)stream));
        float *pts = points + (size_t)b0 * K * 3, *cf = conf + (size_t)b0 * K;
        int32_t *am = argmax ? argmax + (size_t)b0 * K : nullptr;
        if (fused_head) {
            // bf16 path: the output layer's epilogue is the centroid tail; the [B,K,h^3] volume is never materialised
            TailArgs tail{spacing, roi, center3D + (size_t)b0 * 3, pts, cf, am, vout};
            JHN_TRY(tc_forward(net, vol, hybrid_layout(net), nb, G, nullptr, ws_v, align_up(v2v, 256), (cudaStream_t)stream, &tail, SB,
                               (borders || b0 > 0) ? 1 : 0));
            continue;
        }
        if (net->precision == JHN_BF16)
            JHN_TRY(tc_forward(net, vol, hybrid_layout(net), nb, G, vout, ws_v, align_up(v2v, 256), (cudaStream_t)stream, nullptr, SB,
                               (borders || b0 > 0) ? 1 : 0));
        else
            JHN_TRY(jhn_v2v_forward(net, vol, hybrid_layout(net), nb, G, vout, ws_v, align_up(v2v, 256), stream));
        JHN_TRY(jhn_centroid_reduce(vout, nb, K, h, spacing, roi, center3D + (size_t)b0 * 3, pts, cf, am, stream));
    }
    return JHN_OK;
}

int jhn_center_locate(const float *center_heatmaps, int B, int ncam, int Hc, int Wc, int img_w, int img_h,
                      int center_detect_img_size, int bbox_hw, float threshold, const float *cameraMatrices,
                      const float *intrinsicMatrices, const float *distortionCoefficients, int32_t *preds,
                      float *maxvals, float *center3D, int32_t *center3D_int, int32_t *centerHM, int32_t *valid,
                      void *scratch, jhn_stream_t stream)
{
    if (!center_heatmaps || !cameraMatrices || !intrinsicMatrices || !distortionCoefficients || !preds || !maxvals ||
        !center3D || !center3D_int || !centerHM || !valid || !scratch)
        return fail(JHN_ERR_ARG, "jhn_center_locate: null pointer argument");
    if (B < 1 || B > 65535 || ncam < 1 || ncam > 64) return fail(JHN_ERR_SHAPE, "need 1<=B<=65535, 1<=ncam<=64 (got B=%d ncam=%d)", B, ncam);
    if (Hc < 1 || Wc < 1 || (long long)Hc * Wc > (1 << 26)) return fail(JHN_ERR_SHAPE, "centre heat map %dx%d out of range", Hc, Wc);
    if (center_detect_img_size < 1 || bbox_hw < 1 || img_w < 2 * bbox_hw || img_h < 2 * bbox_hw)
        return fail(JHN_ERR_SHAPE, "image %dx%d smaller than the bounding box (half width %d)", img_w, img_h, bbox_hw);
    return center_locate_launch(center_heatmaps, B, ncam, Hc, Wc, img_w, img_h, center_detect_img_size, bbox_hw, threshold,
                                cameraMatrices, intrinsicMatrices, distortionCoefficients, preds, maxvals, center3D,
                                center3D_int, centerHM, valid, scratch, (cudaStream_t)stream);
}

int jhn_crop_normalize(const float *imgs, int B, int ncam, int H, int W, int bbox, const int32_t *centerHM,
                       const int32_t *valid, const float *mean, const float *std, float *crops, jhn_stream_t stream)
{
    if (!imgs || !centerHM || !valid || !mean || !std || !crops) return fail(JHN_ERR_ARG, "jhn_crop_normalize: null pointer argument");
    if (B < 1 || ncam < 1 || (long long)B * ncam > 65535) return fail(JHN_ERR_SHAPE, "need B*ncam in [1,65535] (got B=%d ncam=%d)", B, ncam);
    if (bbox < 4 || (bbox % 4) != 0 || bbox > H || bbox > W) return fail(JHN_ERR_SHAPE, "bounding box %d must be a multiple of 4 and fit the %dx%d image", bbox, W, H);
    for (int i = 0; i < 3; ++i)
        if (!(std[i] > 0.f)) return fail(JHN_ERR_ARG, "std[%d] must be > 0", i);
    return crop_normalize_launch(imgs, B, ncam, H, W, bbox, centerHM, valid, mean, std, crops, (cudaStream_t)stream);
}

int jhn_heatmap_boxes(const float *cameraMatrices, const float *intrinsicMatrices, const float *distortionCoefficients,
                      const float *center3D, const int32_t *centerHM, int B, int ncam, int hs, int G, float spacing,
                      int32_t *boxes, jhn_stream_t stream)
{
    if (!cameraMatrices || !intrinsicMatrices || !distortionCoefficients || !center3D || !centerHM || !boxes)
        return fail(JHN_ERR_ARG, "jhn_heatmap_boxes: null pointer argument");
    if (B < 1 || B > 65535 || ncam < 1 || ncam > 64) return fail(JHN_ERR_SHAPE, "need 1<=B<=65535, 1<=ncam<=64 (got B=%d ncam=%d)", B, ncam);
    if (hs < 4 || G < 2 || (G & 1)) return fail(JHN_ERR_SHAPE, "need hs>=4 and an even grid side (got hs=%d G=%d)", hs, G);
    return heatmap_boxes_launch(cameraMatrices, intrinsicMatrices, distortionCoefficients, center3D, centerHM, B, ncam, hs, G, spacing,
                                boxes, (cudaStream_t)stream);
}

int jhn_heatmap_spans(const float *cameraMatrices, const float *intrinsicMatrices, const float *distortionCoefficients,
                      const float *center3D, const int32_t *centerHM, int B, int ncam, int hs, int G, float spacing,
                      void *scratch, size_t scratch_bytes, int32_t *boxes, int32_t *spans, jhn_stream_t stream)
{
    if (!cameraMatrices || !intrinsicMatrices || !distortionCoefficients || !center3D || !centerHM || !boxes || !spans || !scratch)
        return fail(JHN_ERR_ARG, "jhn_heatmap_spans: null pointer argument");
    if (B < 1 || B > 65535 || ncam < 1 || ncam > 64) return fail(JHN_ERR_SHAPE, "need 1<=B<=65535, 1<=ncam<=64 (got B=%d ncam=%d)", B, ncam);
    if (hs < 4 || G < 2 || (G & 1)) return fail(JHN_ERR_SHAPE, "need hs>=4 and an even grid side (got hs=%d G=%d)", hs, G);
    const size_t h = (size_t)G / 2, need = (size_t)B * ncam * h * h * h * sizeof(float2);
    if (scratch_bytes < need) return fail(JHN_ERR_WORKSPACE, "jhn_heatmap_spans scratch: need %zu bytes, got %zu", need, scratch_bytes);
    if ((uintptr_t)scratch & 7) return fail(JHN_ERR_WORKSPACE, "jhn_heatmap_spans scratch must be 8-byte aligned");
    return heatmap_spans_launch(cameraMatrices, intrinsicMatrices, distortionCoefficients, center3D, centerHM, B, ncam, hs, G, spacing,
                                scratch, boxes, spans, (cudaStream_t)stream);
}

int jhn_upload_heatmap_boxes(const void *host_heatmaps, void *device_heatmaps, const int32_t *boxes_host, int n_images, int hs,
                             int pixel_bytes, jhn_stream_t stream, size_t *bytes_copied)
{
    if (!host_heatmaps || !device_heatmaps || !boxes_host) return fail(JHN_ERR_ARG, "jhn_upload_heatmap_boxes: null pointer argument");
    if (n_images < 1 || hs < 1 || pixel_bytes < 1) return fail(JHN_ERR_SHAPE, "need n_images>=1, hs>=1, pixel_bytes>=1");
    const size_t pitch = (size_t)hs * pixel_bytes, img = pitch * hs;
    size_t total = 0;
    // mode 0 (default): ONE cudaMemcpy3DBatchAsync for all boxes (a cudaMemcpy2DAsync per image costs the calling thread
    // ~10 us each — 384 images per 32 frame sets made the host the bottleneck); 1: one 2-D copy per image; 2: one contiguous
    // copy of the box's rows per image (more bytes, cheapest descriptors)
    static const int mode = [] { const char *e = getenv("JHN_UPLOAD_MODE"); return e ? atoi(e) : 0; }();
    std::vector<cudaMemcpy3DBatchOp> ops;
    if (mode == 0) ops.reserve(n_images);
    for (int i = 0; i < n_images; ++i) {
        const int x0 = boxes_host[4 * i + 0], y0 = boxes_host[4 * i + 1], x1 = -boxes_host[4 * i + 2], y1 = -boxes_host[4 * i + 3];
        if (x0 < 0 || y0 < 0 || x1 >= hs || y1 >= hs || x1 < x0 || y1 < y0)
            return fail(JHN_ERR_ARG, "image %d: box [%d,%d]x[%d,%d] is not inside the %dx%d map (boxes not computed yet?)", i, x0, x1, y0, y1, hs, hs);
        const size_t off = (size_t)i * img + (size_t)y0 * pitch + (size_t)x0 * pixel_bytes;
        const size_t w = (size_t)(x1 - x0 + 1) * pixel_bytes, hgt = (size_t)(y1 - y0 + 1);
        if (mode == 2) {
            const size_t roff = (size_t)i * img + (size_t)y0 * pitch;
            JHN_CUDA(cudaMemcpyAsync((char *)device_heatmaps + roff, (const char *)host_heatmaps + roff, pitch * hgt, cudaMemcpyHostToDevice,
                                     (cudaStream_t)stream));
            total += pitch * hgt;
            continue;
        }
        if (mode == 1) {
            JHN_CUDA(cudaMemcpy2DAsync((char *)device_heatmaps + off, pitch, (const char *)host_heatmaps + off, pitch, w, hgt,
                                       cudaMemcpyHostToDevice, (cudaStream_t)stream));
        } else {
            cudaMemcpy3DBatchOp op{};
            op.src.type = cudaMemcpyOperandTypePointer;
            op.src.op.ptr.ptr = (void *)((const char *)host_heatmaps + off);
            op.src.op.ptr.rowLength = pitch; op.src.op.ptr.layerHeight = 0;
            op.dst.type = cudaMemcpyOperandTypePointer;
            op.dst.op.ptr.ptr = (char *)device_heatmaps + off;
            op.dst.op.ptr.rowLength = pitch; op.dst.op.ptr.layerHeight = 0;
            op.extent = make_cudaExtent(w, hgt, 1);
            op.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
            op.flags = 0;
            ops.push_back(op);
        }
        total += w * hgt;
    }
    if (mode == 0) {
        if (!stream) return fail(JHN_ERR_ARG, "jhn_upload_heatmap_boxes needs a non-default stream (cudaMemcpy3DBatchAsync)");
        // A copy engine spends ~0.16 us per row of a strided copy, i.e. ~35 GB/s on 5 KB rows — below the PCIe link.  The
        // boxes are therefore split over UP_LANES internal streams (fork / join on the caller's stream with events) so that
        // several copy engines walk rows concurrently.
        static const int lanes_env = [] { const char *e = getenv("JHN_UPLOAD_LANES"); return e ? atoi(e) : 0; }();
        constexpr int UP_LANES_MAX = 8;
        const int lanes = lanes_env > 0 ? (lanes_env > UP_LANES_MAX ? UP_LANES_MAX : lanes_env) : 4;
        struct Lanes { cudaStream_t s[UP_LANES_MAX]; cudaEvent_t fork, join[UP_LANES_MAX]; int dev; };
        static thread_local std::map<int, Lanes> cache;                 // per (thread, device)
        int dev = 0;
        JHN_CUDA(cudaGetDevice(&dev));
        auto it = cache.find(dev);
        if (it == cache.end()) {
            Lanes L{};
            L.dev = dev;
            for (int i = 0; i < UP_LANES_MAX; ++i) {
                JHN_CUDA(cudaStreamCreateWithFlags(&L.s[i], cudaStreamNonBlocking));
                JHN_CUDA(cudaEventCreateWithFlags(&L.join[i], cudaEventDisableTiming));
            }
            JHN_CUDA(cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming));
            it = cache.emplace(dev, L).first;
        }
        Lanes &L = it->second;
        if (lanes <= 1 || ops.size() < 2 * (size_t)lanes) {
            size_t fail_idx = 0;
            JHN_CUDA(cudaMemcpy3DBatchAsync(ops.size(), ops.data(), &fail_idx, 0, (cudaStream_t)stream));
        } else {
            JHN_CUDA(cudaEventRecord(L.fork, (cudaStream_t)stream));
            const size_t per = (ops.size() + lanes - 1) / lanes;
            for (int l = 0; l < lanes; ++l) {
                const size_t lo = l * per, hi = std::min(ops.size(), lo + per);
                if (lo >= hi) break;
                JHN_CUDA(cudaStreamWaitEvent(L.s[l], L.fork, 0));
                size_t fail_idx = 0;
                JHN_CUDA(cudaMemcpy3DBatchAsync(hi - lo, ops.data() + lo, &fail_idx, 0, L.s[l]));
                JHN_CUDA(cudaEventRecord(L.join[l], L.s[l]));
                JHN_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, L.join[l], 0));
            }
        }
    }
    if (bytes_copied) *bytes_copied = total;
    return JHN_OK;
}

int jhn_pull_heatmap_boxes(const void *host_heatmaps, void *device_heatmaps, const int32_t *boxes, int n_images, int hs,
                           int pixel_bytes, unsigned long long *bytes_pulled, jhn_stream_t stream)
{
    if (!host_heatmaps || !device_heatmaps || !boxes) return fail(JHN_ERR_ARG, "jhn_pull_heatmap_boxes: null pointer argument");
    if (n_images < 1 || n_images > 65535 || hs < 1 || pixel_bytes < 16 || (pixel_bytes % 16) != 0)
        return fail(JHN_ERR_SHAPE, "need 1<=n_images<=65535, hs>=1, pixel_bytes a multiple of 16 (got %d, %d, %d)", n_images, hs, pixel_bytes);
    if (((uintptr_t)host_heatmaps | (uintptr_t)device_heatmaps) & 15) return fail(JHN_ERR_ARG, "jhn_pull_heatmap_boxes: tensors must be 16-byte aligned");
    void *mapped = nullptr;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, host_heatmaps) == cudaSuccess && at.type == cudaMemoryTypeDevice) {
        mapped = const_cast<void *>(host_heatmaps);           // a device tensor as the source: a box-only device-to-device copy
    } else if (cudaHostGetDevicePointer(&mapped, const_cast<void *>(host_heatmaps), 0) != cudaSuccess || !mapped) {
        cudaGetLastError();
        return fail(JHN_ERR_ARG, "jhn_pull_heatmap_boxes: host_heatmaps is not pinned, device-mapped host memory (cudaHostAlloc / cudaHostRegister)");
    }
    return pull_boxes_launch(mapped, device_heatmaps, boxes, n_images, hs, pixel_bytes, bytes_pulled, (cudaStream_t)stream);
}

int jhn_pull_heatmap_spans(const void *host_heatmaps, void *device_heatmaps, const int32_t *spans, int n_images, int hs,
                           int pixel_bytes, unsigned long long *bytes_pulled, jhn_stream_t stream)
{
    if (!host_heatmaps || !device_heatmaps || !spans) return fail(JHN_ERR_ARG, "jhn_pull_heatmap_spans: null pointer argument");
    if (n_images < 1 || n_images > 65535 || hs < 1 || pixel_bytes < 16 || (pixel_bytes % 16) != 0)
        return fail(JHN_ERR_SHAPE, "need 1<=n_images<=65535, hs>=1, pixel_bytes a multiple of 16 (got %d, %d, %d)", n_images, hs, pixel_bytes);
    if (((uintptr_t)host_heatmaps | (uintptr_t)device_heatmaps) & 15) return fail(JHN_ERR_ARG, "jhn_pull_heatmap_spans: tensors must be 16-byte aligned");
    void *mapped = nullptr;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, host_heatmaps) == cudaSuccess && at.type == cudaMemoryTypeDevice) {
        mapped = const_cast<void *>(host_heatmaps);
    } else if (cudaHostGetDevicePointer(&mapped, const_cast<void *>(host_heatmaps), 0) != cudaSuccess || !mapped) {
        cudaGetLastError();
        return fail(JHN_ERR_ARG, "jhn_pull_heatmap_spans: host_heatmaps is not pinned, device-mapped host memory (cudaHostAlloc / cudaHostRegister)");
    }
    return pull_spans_launch(mapped, device_heatmaps, spans, n_images, hs, pixel_bytes, bytes_pulled, (cudaStream_t)stream);
}

int jhn_pull_small(int n, const void *const *host_tensors, void *const *device_tensors, const size_t *bytes, jhn_stream_t stream)
{
    if (n < 1 || n > 8) return fail(JHN_ERR_ARG, "jhn_pull_small: 1 to 8 tensors per call (got %d)", n);
    if (!host_tensors || !device_tensors || !bytes) return fail(JHN_ERR_ARG, "jhn_pull_small: null pointer argument");
    const void *mapped[8];
    for (int k = 0; k < n; ++k) {
        if (!host_tensors[k] || !device_tensors[k]) return fail(JHN_ERR_ARG, "jhn_pull_small: tensor %d is null", k);
        if ((bytes[k] & 3) || bytes[k] > (64u << 20) || (((uintptr_t)host_tensors[k] | (uintptr_t)device_tensors[k]) & 3))
            return fail(JHN_ERR_ARG, "jhn_pull_small: tensor %d must be 4-byte aligned, a multiple of 4 bytes and <= 64 MB (got %zu bytes)", k, bytes[k]);
        void *m = nullptr;
        if (cudaHostGetDevicePointer(&m, const_cast<void *>(host_tensors[k]), 0) != cudaSuccess || !m) {
            cudaGetLastError();
            return fail(JHN_ERR_ARG, "jhn_pull_small: tensor %d is not pinned, device-mapped host memory (cudaHostAlloc / cudaHostRegister)", k);
        }
        mapped[k] = m;
    }
    return pull_segments_launch(n, mapped, device_tensors, bytes, (cudaStream_t)stream);
}

int jhn_ingest_frames(const uint8_t *frames, int N, int H, int W, float *imgs, jhn_stream_t stream)
{
    if (!frames || !imgs) return fail(JHN_ERR_ARG, "jhn_ingest_frames: null pointer argument");
    if (N < 1 || H < 1 || W < 4 || (W % 4) != 0) return fail(JHN_ERR_SHAPE, "need N>=1, H>=1, W a multiple of 4 (got N=%d H=%d W=%d)", N, H, W);
    if ((long long)N * H * (W / 4) > 0x7fffffffLL * 256) return fail(JHN_ERR_SHAPE, "too many pixels for one call");
    return ingest_frames_launch(frames, N, H, W, imgs, (cudaStream_t)stream);
}

int jhn_crop_normalize_u8(const uint8_t *frames, int B, int ncam, int H, int W, int bbox, const int32_t *centerHM,
                          const int32_t *valid, const float *mean, const float *std, float *crops, jhn_stream_t stream)
{
    if (!frames || !centerHM || !valid || !mean || !std || !crops) return fail(JHN_ERR_ARG, "jhn_crop_normalize_u8: null pointer argument");
    if (B < 1 || ncam < 1 || (long long)B * ncam > 65535) return fail(JHN_ERR_SHAPE, "need B*ncam in [1,65535] (got B=%d ncam=%d)", B, ncam);
    if (bbox < 4 || (bbox % 4) != 0 || bbox > H || bbox > W) return fail(JHN_ERR_SHAPE, "bounding box %d must be a multiple of 4 and fit the %dx%d image", bbox, W, H);
    for (int i = 0; i < 3; ++i)
        if (!(std[i] > 0.f)) return fail(JHN_ERR_ARG, "std[%d] must be > 0", i);
    return crop_normalize_u8_launch(frames, B, ncam, H, W, bbox, centerHM, valid, mean, std, crops, (cudaStream_t)stream);
}

int jhn_efftrack_head(const float *features, const float *weight, int N, int C, int K, int Hq, int Wq, int out_format,
                      void *heatmaps, jhn_stream_t stream)
{
    if (!features || !weight || !heatmaps) return fail(JHN_ERR_ARG, "jhn_efftrack_head: null pointer argument");
    if (N < 1 || N > 65535 || C < 1 || K < 1 || K > KP) return fail(JHN_ERR_SHAPE, "need 1<=N<=65535, C>=1, 1<=K<=%d (got N=%d C=%d K=%d)", KP, N, C, K);
    if (Hq < 1 || Wq < 1 || Hq > 4096 || Wq > 4096) return fail(JHN_ERR_SHAPE, "feature map %dx%d out of range", Hq, Wq);
    return efftrack_head_launch(features, weight, N, C, K, Hq, Wq, out_format, heatmaps, (cudaStream_t)stream);
}

int jhn_softplus2(const float *v2v_out, long long n, float *heatmap_final, jhn_stream_t stream)
{
    if (!v2v_out || !heatmap_final) return fail(JHN_ERR_ARG, "jhn_softplus2: null pointer argument");
    if (n < 1 || n > (1LL << 40)) return fail(JHN_ERR_SHAPE, "element count %lld out of range", n);
    if (((uintptr_t)v2v_out | (uintptr_t)heatmap_final) & 15) return fail(JHN_ERR_ARG, "jhn_softplus2: buffers must be 16-byte aligned");
    return softplus2_launch(v2v_out, n, heatmap_final, (cudaStream_t)stream);
}

int jhn_pad_heatmaps(const float *heatmaps, long long N, int S, float *heatmaps_padded, jhn_stream_t stream)
{
    if (!heatmaps || !heatmaps_padded) return fail(JHN_ERR_ARG, "jhn_pad_heatmaps: null pointer argument");
    if (N < 1 || S < 1 || S > 8192 || N * (S + 2) * (S + 2) > (1LL << 40)) return fail(JHN_ERR_SHAPE, "need N>=1, 1<=S<=8192");
    return pad_border_launch(heatmaps, N, S, heatmaps_padded, (cudaStream_t)stream);
}

}  // extern "C"

namespace jhn {
size_t tc_volume_bytes(const jhn_v2v *net, int B, int G);
}
static size_t hybrid_volume_bytes(const jhn_v2v *net, int B, int G)
{
    if (net->precision == JHN_BF16) return jhn::tc_volume_bytes(net, B, G);
    return (size_t)B * net->K * G * G * G * sizeof(float);
}
