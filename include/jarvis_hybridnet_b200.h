/*
 * jarvis_hybridnet_b200.h — C ABI of the B200-native (sm_100a) 3D inference hot path of
 * JARVIS-HybridNet:  ReprojectionLayer -> V2VNet -> softplus centroid.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own for this path;
 * its plugin seam is attribute replacement on HybridNetBackbone after torch.ops.load_library
 * (jarvis/prediction/jarvis3D.py:53-69).  Each entry point below names the reference code it replaces.
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer owned by the caller (PyTorch allocates);
 *     the library never frees caller memory and allocates nothing except the packed-weights handle;
 *   - every call is asynchronous and ordered on `stream` (pass torch.cuda.current_stream()); no call
 *     synchronises, so all of them are CUDA-graph capturable, and re-entrant across streams/devices;
 *   - return value: 0 on success, negative jhn_status on failure; jhn_last_error() returns a
 *     thread-local message.  There is NO fallback: an unsupported device or shape is an error;
 *   - B is the number of independent frame sets processed by one call (the reference always runs B=1,
 *     jarvis/hybridnet/repro_layer.py:112-117).
 */
#ifndef JARVIS_HYBRIDNET_B200_H
#define JARVIS_HYBRIDNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *jhn_stream_t;          /* == cudaStream_t */
typedef struct jhn_v2v jhn_v2v;                    /* opaque packed V2VNet weights */

enum jhn_status {
    JHN_OK = 0,
    JHN_ERR_ARG = -1,          /* null pointer / bad enum                      */
    JHN_ERR_SHAPE = -2,        /* unsupported or inconsistent dimensions       */
    JHN_ERR_ARCH = -3,         /* device is not sm_100 (B200)                  */
    JHN_ERR_CUDA = -4,         /* a CUDA call failed (message has the detail)  */
    JHN_ERR_WORKSPACE = -5     /* caller workspace too small or misaligned     */
};

enum jhn_precision {
    JHN_FP32 = 0,              /* fp32 storage and FFMA arithmetic: the parity path            */
    JHN_BF16 = 1               /* bf16 storage, tcgen05 tensor-core convolutions, fp32 accum   */
};

enum jhn_lerp_mode {           /* rounding of ATen's trilinear lerp           */
    JHN_LERP_FMA_FIRST = 0,    /* fma(w0, a, w1*b)  — ATen CPU                                 */
    JHN_LERP_FMA_SECOND = 1,   /* fma(w1, b, w0*a)                                             */
    JHN_LERP_NO_FMA = 2        /* w0*a + w1*b, both products rounded                           */
};

/* Storage of the 2D key-point heat maps handed to the reprojection stage (SURVEY.md section 8b(1) "layout/dtype enum").
 * The reference's effTrack head emits JHN_HM_F32_PLANAR (jarvis/efficienttrack/model.py:127,130 -> model.py:57-66);
 * the channels-last 16-bit forms are the gather's native staging layout: one (camera, pixel) is ONE contiguous
 * JHN_HM_CL_PITCH-channel vector, the 1-pixel zero border of F.pad (model.py:65-66) is materialised, channels
 * K .. JHN_HM_CL_PITCH-1 are zero.  A producer that writes JHN_HM_F16_CL directly (jhn_heatmap_convert, or
 * jhn_efftrack_head for the effTrack deconvolution) spares the staging pass and half of the input bytes. */
enum jhn_heatmap_format {
    JHN_HM_F32_PLANAR = 0,     /* fp32 [B][ncam][K][S][S], S = hs (padded) or hs-2 (un-padded)               */
    JHN_HM_F16_CL = 1,         /* fp16 [B][ncam][hs][hs][24], values scaled by JHN_HM_F16_SCALE (exact 2^-4) */
    JHN_HM_BF16_CL = 2         /* bf16 [B][ncam][hs][hs][24], unscaled                                       */
};
#define JHN_HM_CL_PITCH 24
#define JHN_HM_F16_SCALE 0.0625f

/* Layout of the feature volume handed from the reprojection stage to V2VNet. */
enum jhn_volume_layout {
    JHN_VOL_NCDHW_F32 = 0,     /* [B][K][G][G][G] fp32 — the reference tensor (repro_layer.py:119) */
    JHN_VOL_V2V_BF16 = 1       /* bf16, parity-split channel-blocked, consumed by the tensor-core
                                  front convolution (DESIGN.md "HBM layouts")                      */
};

const char *jhn_last_error(void);
int jhn_abi_version(void);
/* JHN_OK iff `device` is a compute-capability 10.x part this library was built for. */
int jhn_check_device(int device);

/* Measurement aids (not part of the drop-in surface; used by bench.py):
 *   jhn_launch_count     kernels launched by this library since load (bench.py "gpu_launches");
 *   jhn_profile_enable   while on, every launch is bracketed by CUDA events on its stream;
 *   jhn_profile_collect  device-sync, then "<kernel>\t<launches>\t<total_ms>\n" per kernel name into buf;
 *   jhn_debug_set_gather_box_bytes  test hook: capacity of one shared-memory pixel-box slot of the streaming gather
 *                        (0 = default).  Boxes larger than it take the kernel's global-memory path, so a small
 *                        value lets tests drive that path on ordinary inputs.  Returns the value in effect. */
unsigned long long jhn_launch_count(void);
void jhn_profile_enable(int on);
int jhn_profile_collect(char *buf, int cap);
int jhn_debug_set_gather_box_bytes(int bytes);
/* Frame sets per internal pass of jhn_hybrid3d_forward (0 = library default, which keeps the activations of one pass
 * inside the 126 MB L2).  Frame sets are independent; on the bf16 path the split changes how a sample's InstanceNorm partial sums
 * are grouped (fp32 re-association, <= 5e-3 mm on the key points; bit-identical for equal splits).  Returns the value in effect. */
int jhn_set_sub_batch(int frame_sets);

/* ------------------------------------------------------------------------------------------------
 * Stage 1 — replaces F.pad (jarvis/hybridnet/model.py:65-66) + ReprojectionLayer.forward
 * (jarvis/hybridnet/repro_layer.py:110-119, :88-107, :40-85) + the /255 of model.py:72.
 *
 *   heatmaps        device, storage `hm_format` (enum jhn_heatmap_format).  JHN_HM_F32_PLANAR: fp32 [B][ncam][K][S][S],
 *                   S = hs if heatmaps_padded else hs-2 (heatmaps_padded=1 is the tensor the reference passes to
 *                   reproLayer).  The channels-last forms are always padded and need precision JHN_BF16.
 *   cameraMatrices  device fp32 [B][ncam][4][3]      intrinsicMatrices device fp32 [B][ncam][3][3]
 *   distortion      device fp32 [B][ncam][1][5]
 *   center3D        device fp32 [B][3] (mm): added to the grid in fp32 exactly as `self.grid + center[0]`
 *                   (repro_layer.py:113); the predictor passes integers (jarvis3D.py:183), the validation path
 *                   multiples of GRID_SPACING (hybridnet.py:284-304) — both are taken as they are
 *   centerHM        device i32 [B][ncam][2]
 *   hs              padded heat-map side = BOUNDING_BOX_SIZE/2 + 2 (repro_layer.py:37)
 *   G, spacing      grid side ROI_CUBE_SIZE/GRID_SPACING (even) and GRID_SPACING in mm
 *   post_divide     1.0f -> reference ReprojectionLayer output; 255.0f folds model.py:72
 *   volume_out      device, layout `layout`
 *   index_out       optional device i32 [B][ncam][G][G][G]: flat padded-pixel index y*hs+x, the
 *                   int64 `res` of repro_layer.py:82-83 (bit-exact parity target); may be NULL
 * ------------------------------------------------------------------------------------------------ */
int jhn_reproject_workspace_bytes(int B, int ncam, int K, int hs, int G, int precision, size_t *bytes);
int jhn_reproject_gather(const void *heatmaps, int hm_format, int heatmaps_padded,
                         const float *cameraMatrices, const float *intrinsicMatrices,
                         const float *distortionCoefficients,
                         const float *center3D, const int32_t *centerHM,
                         int B, int ncam, int K, int hs, int G, float spacing,
                         int lerp_mode, float post_divide, int precision, int layout,
                         void *volume_out, int32_t *index_out,
                         void *workspace, size_t workspace_bytes, jhn_stream_t stream);

/* Producer side of the channels-last formats: fp32 planar heat maps
 * [B][ncam][K][S][S] -> `dst_format` (JHN_HM_F16_CL or JHN_HM_BF16_CL) [B][ncam][hs][hs][24] with the F.pad border.
 * One HBM-bound pass; jhn_reproject_gather does the same internally when it is handed JHN_HM_F32_PLANAR. */
int jhn_heatmap_convert(const float *heatmaps, int heatmaps_padded, int B, int ncam, int K, int hs,
                        int dst_format, void *dst, jhn_stream_t stream);

/* Host-buffer callers (the end-to-end path of HybridNet3D.forward_host): the gather of one frame set reads, per camera,
 * only the pixels inside the projection of the voxel grid — 57 % of a 130x130 map at the Example shape, 12 % of a 258x258
 * map at the micro-benchmark shape.
 *   jhn_heatmap_boxes         runs the projection of repro_layer.py:46-68 for the coarse grid and returns, per (frame set,
 *                             camera), the pixel box {x0, y0, -x1, -y1} (padded heat-map pixels, inclusive) that bounds every
 *                             index jhn_reproject_gather / jhn_hybrid3d_forward will read for these centres and calibration.
 *                             boxes: device i32 [B][ncam][4].
 *   jhn_upload_heatmap_boxes  host code: one cudaMemcpy2DAsync per image that copies just that box of a channels-last host
 *                             tensor [n_images][hs][hs][pixel_bytes] (pinned) to the same place of the device tensor; the
 *                             pixels outside the boxes are never read by the gather and may hold anything.
 *                             boxes_host: HOST copy of jhn_heatmap_boxes' output; bytes_copied (optional) = bytes moved. */
int jhn_heatmap_boxes(const float *cameraMatrices, const float *intrinsicMatrices, const float *distortionCoefficients,
                      const float *center3D, const int32_t *centerHM, int B, int ncam, int hs, int G, float spacing,
                      int32_t *boxes, jhn_stream_t stream);
int jhn_upload_heatmap_boxes(const void *host_heatmaps, void *device_heatmaps, const int32_t *boxes_host, int n_images, int hs,
                             int pixel_bytes, jhn_stream_t stream, size_t *bytes_copied);
/*   jhn_pull_heatmap_boxes    the same transfer executed by a kernel that reads the boxes straight out of pinned,
 *                             device-mapped host memory (cudaHostAlloc / torch pin_memory()): no copy-engine row overhead
 *                             (a strided DMA of 5 KB rows runs at 36 GB/s on a B200, a contiguous one at 55 GB/s), and `boxes`
 *                             is the DEVICE tensor jhn_heatmap_boxes wrote, so nothing synchronises with the host.
 *                             bytes_pulled (optional, device u64) is incremented by the bytes read over the link. */
int jhn_pull_heatmap_boxes(const void *host_heatmaps, void *device_heatmaps, const int32_t *boxes, int n_images, int hs,
                           int pixel_bytes, unsigned long long *bytes_pulled, jhn_stream_t stream);
/*   jhn_heatmap_spans         jhn_heatmap_boxes plus, per (frame set, camera, pixel row), the column range the gather can touch:
 *                             spans = int32 [B][ncam][hs][2] holding {lo, -hi} (rows no voxel maps to have lo > hi).  The spans
 *                             hold ~78 % of the boxes' pixels (the voxel cube projects to a hexagon).  scratch: device,
 *                             B * ncam * (G/2)^3 * 8 bytes (the coarse grid's projections).
 *   jhn_pull_heatmap_spans    jhn_pull_heatmap_boxes over those row spans (device tensor `spans`). */
int jhn_heatmap_spans(const float *cameraMatrices, const float *intrinsicMatrices, const float *distortionCoefficients,
                      const float *center3D, const int32_t *centerHM, int B, int ncam, int hs, int G, float spacing,
                      void *scratch, size_t scratch_bytes, int32_t *boxes, int32_t *spans, jhn_stream_t stream);
int jhn_pull_heatmap_spans(const void *host_heatmaps, void *device_heatmaps, const int32_t *spans, int n_images, int hs,
                           int pixel_bytes, unsigned long long *bytes_pulled, jhn_stream_t stream);
/*   jhn_pull_small            n <= 8 small pinned host tensors (calibration, centres) copied to the device by one kernel that
 *                             reads mapped host memory: unlike cudaMemcpyAsync they do not queue on the host->device copy
 *                             engine behind a large transfer issued earlier on another stream.  bytes[k] % 4 == 0. */
int jhn_pull_small(int n, const void *const *host_tensors, void *const *device_tensors, const size_t *bytes, jhn_stream_t stream);
/*   jhn_set_transfer_overlap  announces (1) / withdraws (0), for the CALLING THREAD, that forwards launched from now on run
 *                             next to jhn_pull_* transfers: the 3x3x3 layers are then launched in their 120-register build,
 *                             which leaves room on every SM for the transfer kernel's CTAs (5 % slower alone, but a forward
 *                             that overlaps a pull otherwise runs those layers in two waves).  Results are identical. */
void jhn_set_transfer_overlap(int on);
/*   jhn_debug_set_pull_config tuning hook: launch shape of the pull kernel — threads per CTA (32, 64, 96 or 128), number of
 *                             CTAs, parts per image; an argument <= 0 leaves that value unchanged. */
void jhn_debug_set_pull_config(int threads, int ctas, int split);

/* ------------------------------------------------------------------------------------------------
 * Stage 2 — replaces V2VNet (jarvis/hybridnet/v2vnet.py:86-102) in eval mode.
 *
 * jhn_v2v_create packs the 24 checkpoint tensors (device fp32, PyTorch layouts: Conv3d
 * [Cout][Cin][k][k][k], ConvTranspose3d [Cin][Cout][k][k][k]) in the order of
 * jarvis_hybridnet_b200.synth.V2V_LAYERS (weight, bias per layer).  C = K key points:
 * channel widths are K, 2K, 4K (v2vnet.py:62-96).
 *
 * jhn_v2v_forward: volume_in has layout `in_layout` ([B][K][G][G][G] fp32, already divided by 255, or
 * the bf16 V2V layout written by jhn_reproject_gather); out is device fp32 [B][K][G/2][G/2][G/2]
 * (the tensor v2vNet returns, model.py:72).
 * ------------------------------------------------------------------------------------------------ */
int jhn_v2v_create(const float *const *tensors, int num_tensors, int K, int precision,
                   jhn_stream_t stream, jhn_v2v **out);
void jhn_v2v_destroy(jhn_v2v *net);
/* Promise (on=1) that the workspace handed to jhn_v2v_forward / jhn_hybrid3d_forward with this network is
 * written by nobody else between calls.  The bf16 path then clears the zero borders of its padded activation
 * tensors only when the workspace pointer or the shape changes instead of on every call.  Default: off. */
int jhn_v2v_set_workspace_persistent(jhn_v2v *net, int on);
int jhn_v2v_workspace_bytes(const jhn_v2v *net, int B, int G, size_t *bytes);
int jhn_v2v_forward(const jhn_v2v *net, const void *volume_in, int in_layout, int B, int G,
                    float *out, void *workspace, size_t workspace_bytes, jhn_stream_t stream);

/* Test aid (bf16 networks only): run ONE convolution of V2VNet on the tensor cores.  `in` is device fp32
 * NCDHW of the layer's input, `out` device fp32 NCDHW of its raw output (bias added, before InstanceNorm).
 * `layer` indexes synth.V2V_LAYERS; D is the layer's output grid side (input side for the transposed conv). */
int jhn_v2v_debug_layer_workspace_bytes(const jhn_v2v *net, int layer, int B, int D, size_t *bytes);
int jhn_v2v_debug_layer(const jhn_v2v *net, int layer, const float *in, int B, int D, float *out,
                        void *workspace, size_t workspace_bytes, jhn_stream_t stream);
/* Test aid (bf16 networks only): the fused output layer + centroid tail (stage 3 inside the last GEMM's epilogue) on
 * given activations.  `in` device fp32 NCDHW [B][2K][h][h][h] (rounded to bf16 on entry); outputs as jhn_centroid_reduce.
 * Workspace: jhn_v2v_debug_layer_workspace_bytes(net, 11, B, h). */
int jhn_v2v_debug_head_centroid(const jhn_v2v *net, const float *in, int B, int h, float spacing, float roi,
                                const float *center3D, float *points, float *conf, int32_t *argmax,
                                void *workspace, size_t workspace_bytes, jhn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 3 — replaces the inline tail of HybridNetBackbone.forward (jarvis/hybridnet/model.py:73-87):
 * softplus, sum-normalised centroid, confidence, voxel -> mm.
 *
 *   v2v_out   device fp32 [B][K][h][h][h]        center3D device fp32 [B][3]
 *   points    device fp32 [B][K][3] (mm)         conf device fp32 [B][K]
 *   argmax    optional device i32 [B][K]: first flat index of the per-key-point maximum
 * ------------------------------------------------------------------------------------------------ */
int jhn_centroid_reduce(const float *v2v_out, int B, int K, int h, float spacing, float roi,
                        const float *center3D, float *points, float *conf, int32_t *argmax,
                        jhn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stages 1-3 fused behind one call (model.py:65-88): heat maps in, key points out.
 * ------------------------------------------------------------------------------------------------ */
int jhn_hybrid3d_workspace_bytes(const jhn_v2v *net, int B, int ncam, int hs, int G, size_t *bytes);
int jhn_hybrid3d_forward(const jhn_v2v *net, const void *heatmaps, int hm_format, int heatmaps_padded,
                         const float *cameraMatrices, const float *intrinsicMatrices,
                         const float *distortionCoefficients,
                         const float *center3D, const int32_t *centerHM,
                         int B, int ncam, int hs, int G, float spacing, float roi, int lerp_mode,
                         float *points, float *conf, int32_t *argmax,
                         void *workspace, size_t workspace_bytes, jhn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Predictor glue — replaces, without host synchronisation, the lines of
 * JarvisPredictor3D.forward between the centre-detect CNN and the 3D network
 * (jarvis/prediction/jarvis3D.py:147-177) and the two ReprojectionTool methods they call
 * (jarvis/utils/reprojection.py:45-66 reprojectPoint, :69-90 reconstructPoint).
 *
 * jhn_center_locate: per-camera argmax of the centre heat maps, detection count (max > threshold, the
 *   reference uses 50), weighted DLT triangulation of the centre from all cameras, projection into every
 *   camera, truncation and clamping of the crop centres.
 *     center_heatmaps device fp32 [B][ncam][Hc][Wc] (centre-detect output [1]); img_w / img_h full image size;
 *     center_detect_img_size cfg.CENTERDETECT.IMAGE_SIZE; bbox_hw = BOUNDING_BOX_SIZE / 2; ncam <= 64
 *     preds     device i32 [B][ncam][2] (x, y) in heat-map pixels      maxvals device fp32 [B][ncam] (max / 255)
 *     center3D  device fp32 [B][3] (mm)     center3D_int device i32 [B][3] (center3D.int())
 *     centerHM  device i32 [B][ncam][2] (full-resolution pixels, clamped to [bbox_hw, size - bbox_hw])
 *     valid     device i32 [B]: 1 iff at least two cameras detected the centre (the reference returns None
 *               otherwise); when 0 the other outputs of that frame set are zeros / bbox_hw
 *     scratch   device, 8 * B bytes, zeroed once by the caller (the kernel leaves it zeroed)
 * jhn_crop_normalize: jarvis3D.py:168-177, imgs device fp32 [B][ncam][3][H][W] -> crops device fp32
 *   [B][ncam][3][bbox][bbox] = (window around centerHM - mean) / std; zeros for frame sets with valid == 0.
 *   mean / std: 3 host floats each (cfg.DATASET.MEAN / STD).  bbox must be a multiple of 4.
 * ------------------------------------------------------------------------------------------------ */
int jhn_center_locate(const float *center_heatmaps, int B, int ncam, int Hc, int Wc, int img_w, int img_h,
                      int center_detect_img_size, int bbox_hw, float threshold,
                      const float *cameraMatrices, const float *intrinsicMatrices,
                      const float *distortionCoefficients, int32_t *preds, float *maxvals, float *center3D,
                      int32_t *center3D_int, int32_t *centerHM, int32_t *valid, void *scratch,
                      jhn_stream_t stream);
int jhn_crop_normalize(const float *imgs, int B, int ncam, int H, int W, int bbox, const int32_t *centerHM,
                       const int32_t *valid, const float *mean, const float *std, float *crops,
                       jhn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * the data formats either side of the path.
 *
 * jhn_ingest_frames: jarvis/prediction/predict3D.py:79
 *     `torch.from_numpy(imgs_orig).cuda().float().permute(0,3,1,2)[:, [2,1,0]] / 255.`
 *   frames  device uint8 [N][H][W][3], BGR as cv2.VideoCapture.read() delivers them (N = cameras x frame sets)
 *   imgs    device fp32  [N][3][H][W], RGB, float(u8) * (1.f / 255.f): the bits of the reference's CUDA tensor (ATen divides by a
 *           Python scalar as a multiplication by its fp32 reciprocal).  W % 4 == 0.
 * jhn_crop_normalize_u8: jarvis/prediction/jarvis3D.py:168-177 straight from the uint8 frames:
 *   crops device fp32 [B][ncam][3][bbox][bbox] = ((u8 / 255) - mean) / std around centerHM; zeros where valid == 0.
 *   Same values as jhn_ingest_frames followed by jhn_crop_normalize, without the fp32 image.
 * jhn_efftrack_head: the last layer of EfficientTrackBackbone, `res2 = self.deconv1(res1)`
 *   (jarvis/efficienttrack/model.py:89-95,127: ConvTranspose2d(C, K, kernel 4, stride 2, padding 1, bias=False)), with the
 *   result written in any jhn_heatmap_format — JHN_HM_F16_CL / JHN_HM_BF16_CL are what jhn_reproject_gather /
 *   jhn_hybrid3d_forward read without a staging pass (F.pad's zero border included, jarvis/hybridnet/model.py:65-66).
 *   features device fp32 [N][C][Hq][Wq]   weight device fp32 [C][K][4][4] (the checkpoint's `deconv1.weight`)
 *   heatmaps JHN_HM_F32_PLANAR: fp32 [N][K][2Hq][2Wq] (the reference's tensor, un-padded)
 *            channels-last:     16-bit [N][2Hq+2][2Wq+2][24]                                    1 <= K <= 24
 * jhn_softplus2: heatmap_final = softplus(softplus(v2v_out)) (jarvis/hybridnet/model.py:73,88), n fp32 elements.
 * jhn_pad_heatmaps: heatmaps_padded = F.pad(heatmaps, [1,1,1,1]) (model.py:65-66): fp32 [N][S][S] -> [N][S+2][S+2].
 * ------------------------------------------------------------------------------------------------ */
int jhn_ingest_frames(const uint8_t *frames, int N, int H, int W, float *imgs, jhn_stream_t stream);
int jhn_crop_normalize_u8(const uint8_t *frames, int B, int ncam, int H, int W, int bbox, const int32_t *centerHM,
                          const int32_t *valid, const float *mean, const float *std, float *crops,
                          jhn_stream_t stream);
int jhn_efftrack_head(const float *features, const float *weight, int N, int C, int K, int Hq, int Wq,
                      int out_format, void *heatmaps, jhn_stream_t stream);
int jhn_softplus2(const float *v2v_out, long long n, float *heatmap_final, jhn_stream_t stream);
int jhn_pad_heatmaps(const float *heatmaps, long long N, int S, float *heatmaps_padded, jhn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* JARVIS_HYBRIDNET_B200_H */
