/* cnl_b200.h - C ABI of the sm_100a CenterNet inference hot path.
 *
 * The reference (gau-nernst/centernet-lightning) has no FFI/plugin registry: its hot path
 * is a Python object boundary.  Each entry point below names the reference
 * function it replaces (paths relative to the reference root).  All pointers are raw
 * DEVICE pointers unless the name ends in _host; all calls are asynchronous on `stream`
 * (a cudaStream_t passed as void*), never synchronise, never allocate tensor memory
 * (the caller owns outputs and workspace) and return 0 on success or a cnl_status.
 * The reference-side binding is the ctypes stub shown in INTEGRATION.md.
 */
#ifndef CNL_B200_H_
#define CNL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

typedef enum {
  CNL_OK = 0,
  CNL_ERR_INVALID_ARGUMENT = 1,   /* shapes / k > H*W / even nms kernel / misaligned pointer */
  CNL_ERR_UNSUPPORTED = 2,        /* configuration the sm_100a kernels do not implement       */
  CNL_ERR_CUDA = 3,               /* a CUDA runtime / driver call failed                      */
  CNL_ERR_WORKSPACE = 4           /* workspace too small                                      */
} cnl_status;

/* Human-readable description of the last error on the calling thread. */
const char* cnl_last_error(void);

/* Library/ABI version (major*1000 + minor) and the SM architecture it was compiled for (100). */
int cnl_version(void);
int cnl_compiled_sm(void);

/* ------------------------------------------------------------------------------------------
 * Decode: heatmap -> top-k detections.
 *
 * Replaces CenterNet.decode_detections / get_topk_from_heatmap / gather_and_decode_boxes
 * (centernet_lightning/models/centernet.py:229-304) and, when `reid` is non-NULL,
 * EmbeddingHead.gather_at_indices / FairMOT.gather_tracking2d
 * (centernet_lightning/models/fairmot.py:63-73, 138-151).
 *
 *   heatmap      (N, C, H, W) float32, contiguous NCHW.
 *                from_logits = 0: probabilities, exactly what the reference's decode receives
 *                                 (validation_step applies .sigmoid() first, centernet.py:205).
 *                from_logits = 1: raw head output; the kernel applies 1/(1+exp(-x)) in fp32 itself
 *                                 (fuses the sigmoid of centernet.py:205 into the same pass).  Results are those of
 *                                 from_logits = 0 on cnl_sigmoid(heatmap), bit for bit: peaks and the class arg-max
 *                                 are decided on the fp32 PROBABILITIES (distinct logits that round to one probability
 *                                 are a plateau, as in the reference), evaluated lazily on near ties only.
 *   box_offsets  (N, 4, H, W) float32 ltrb map, or NULL together with boxes (top-k only).
 *   reid         (N, E, H, W) float32 or NULL (E = reid_dim).
 *   nms_kernel   odd, 1..7 (centernet.py:93 default 3).  k = num_detections <= H*W (torch.topk's limit, centernet.py:259);
 *                k > 1024 takes a slower whole-image sort and needs the larger workspace of cnl_decode_workspace_bytes_k.
 *   Outputs (caller-allocated): boxes (N,k,4) f32 xyxy; scores (N,k) f32 descending;
 *   labels (N,k) int64; indices (N,k) int64 flat y*W+x; embeddings (N,k,E) f32 or NULL.
 *   Tie order: score descending, then flat index ascending (torch.topk leaves it unspecified).
 *   workspace: cnl_decode_workspace_bytes_k(N,H,W,k) bytes (= cnl_decode_workspace_bytes(N,H,W) for k <= 1024),
 *   256-byte aligned; contents are scratch.
 * ------------------------------------------------------------------------------------------ */
size_t cnl_decode_workspace_bytes(int n, int h, int w);
size_t cnl_decode_workspace_bytes_k(int n, int h, int w, int num_detections);

/* CNL_DECODE_WORKSPACE_CLEAN: accepted for ABI compatibility and ignored (the candidate histogram lives in the select
 * kernel's shared memory since ABI 1.1; the workspace needs no initialisation). */
#define CNL_DECODE_WORKSPACE_CLEAN 2
/* Profiling aid, OR-ed into `from_logits`: launch only the streaming (peaks) kernel and skip the per-image select, so that
 * the HBM-bound pass can be timed alone with CUDA events.  Outputs are not written and the workspace histogram is left
 * dirty: the next call on this workspace must not claim CNL_DECODE_WORKSPACE_CLEAN. */
#define CNL_DECODE_PEAKS_ONLY 4
int cnl_decode_detections(const float* heatmap, const float* box_offsets, const float* reid,
                          int n, int c, int h, int w, int reid_dim,
                          int from_logits, int nms_kernel, int num_detections,
                          int normalize_boxes, int box_log, float box_multiplier, int stride,
                          float* boxes, float* scores, int64_t* labels, int64_t* indices,
                          float* embeddings,
                          void* workspace, size_t workspace_bytes, void* stream);

/* The same decode, additionally writing every detection as one packed float32 row of `packed_width` = 8 + reid_dim lanes
 * into `packed` (N, k, packed_width; 16-byte aligned, reid_dim % 4 == 0):
 *   [x1, y1, x2, y2, score, bits(int32 flat index), bits(label low 32), bits(label high 32), embedding...]
 * - the fixed-shape send buffer of the cross-rank collation (the reference gathers pickled per-image dicts with
 * dist.all_gather_object, centernet_lightning/eval/coco.py:10-18); integer lanes travel bit-exactly. */
int cnl_decode_detections_packed(const float* heatmap, const float* box_offsets, const float* reid,
                                 int n, int c, int h, int w, int reid_dim,
                                 int from_logits, int nms_kernel, int num_detections,
                                 int normalize_boxes, int box_log, float box_multiplier, int stride,
                                 float* boxes, float* scores, int64_t* labels, int64_t* indices,
                                 float* embeddings, float* packed, int packed_width,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* Elementwise fp32 logistic 1/(1+exp(-x)) - the `.sigmoid()` the reference applies to the heatmap head
 * (centernet_lightning/models/centernet.py:205; G1 forward(), tests/test_models.py:88-99). */
int cnl_sigmoid(const float* in, float* out, size_t n, void* stream);

/* (N*k,4) xyxy -> xywh, the torchvision.ops.box_convert the reference's validation_step applies before handing
 * detections to the COCO evaluator (centernet_lightning/models/centernet.py:207).  In-place allowed. */
int cnl_boxes_xyxy_to_xywh(const float* boxes_xyxy, float* boxes_xywh, size_t n_boxes, void* stream);

/* Input side of the path: n uint8 RGB images in HWC order (cv2.imread + cvtColor + cv2.resize of the reference's
 * InferenceDataset, centernet_lightning/datasets/inference.py:28-33) -> normalised fp32 NCHW, with the arithmetic
 * of albumentations.Normalize as the reference applies it (README.md:84-87): out = float(double(x) - mean255[c]) *
 * inv_std255[c], mean255 = mean*255 (double), inv_std255 = float(1/(std*255)).  mean255 / inv_std255 are HOST
 * pointers to 3 values; images are device pointers. */
int cnl_normalize_images_u8(const uint8_t* images_hwc, float* images_nchw, int n, int h, int w,
                            const double* mean255, const float* inv_std255, void* stream);

/* Stand-alone box gather for caller-supplied indices: CenterNet.gather_and_decode_boxes
 * (centernet_lightning/models/centernet.py:263-304, a staticmethod the reference also calls from its
 * loss, :162-165).  indices (N,k) int64 device; boxes (N,k,4) f32, 16-byte aligned.  Out-of-range
 * indices (torch.gather would raise) produce NaN rows. */
int cnl_gather_boxes(const float* box_offsets, const int64_t* indices, int n, int h, int w, int k,
                     int normalize_boxes, int box_log, float box_multiplier, int stride,
                     float* boxes, void* stream);

/* Tracker association costs for one frame (the consumer of the tracking head's detections):
 *   reid_cost[i][j] = scipy.spatial.distance.cdist(det_emb, trk_emb, "cosine")   (centernet_lightning/models/tracker.py:61,157)
 *   box_cost[i][j]  = 1 - IoU (giou = 0) or 1 - GIoU (giou = 1) of xyxy boxes    (centernet_lightning/utils/box.py:49-92,
 *                                                                                 tracker.py:63,169)
 * All arrays are device pointers, row-major float64: det_emb (n_det, emb_dim), trk_emb (n_trk, emb_dim), det_box (n_det, 4),
 * trk_box (n_trk, 4), outputs (n_det, n_trk).  Either pair (embeddings + reid_cost / boxes + box_cost) may be NULL.
 * fp64 with the host code's operation order: equal to scipy / numpy float64 results to the last bit.  The Hungarian
 * assignment (tracker.py:28) stays on the host. */
size_t cnl_track_workspace_bytes(int n_det, int n_trk);
int cnl_track_cost_matrices(const double* det_emb, const double* trk_emb, int emb_dim,
                            const double* det_box, const double* trk_box, int n_det, int n_trk, int giou,
                            double* reid_cost, double* box_cost, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Forward: backbone -> neck -> heads as a list of fused convolution launches.
 *
 * Replaces GenericModel.forward (centernet_lightning/models/meta.py:41-47) and the
 * vision_toolbox backbone/neck/ConvBnAct modules it calls (meta.py:9-10, 21-30, 87-96).
 * A cnl_engine is an immutable plan for one (batch, height, width): packed fp16 weights,
 * TMA descriptors and activation buffer offsets.  The caller provides one arena of
 * cnl_engine_arena_bytes() bytes of device memory; the engine never allocates tensors.
 *
 * precision: 0 = CNL_PRECISION_SPLIT  activations and weights are carried as fp16 hi+lo pairs and
 *                every product is accumulated in fp32 from three tcgen05 passes (hi*hi, hi*lo, lo*hi):
 *                fp32-equivalent results (the 1e-3 parity bar of BASELINE.json).
 *                The two correction products accumulate in their own TMEM accumulator (the tensor core truncates
 *                after every accumulate step; this keeps the main sum to one third of the roundings).
 *            2 = CNL_PRECISION_SPLIT_FUSED  same three passes into ONE accumulator (double-buffered, slightly
 *                faster, ~3x the accumulated rounding bias of mode 0).
 *            1 = CNL_PRECISION_FAST   single fp16 pass, fp32 accumulate (reduced precision, reported
 *                separately; does not meet the 1e-3 bar).
 * ------------------------------------------------------------------------------------------ */
typedef struct cnl_engine cnl_engine;

enum { CNL_PRECISION_SPLIT = 0, CNL_PRECISION_FAST = 1, CNL_PRECISION_SPLIT_FUSED = 2 };

/* One fused convolution: out = act(conv(in, w) + bias [+ residual]).  Host-side description. */
typedef struct {
  int kind;            /* 0 = conv (tcgen05 implicit GEMM), 1 = stem (7x7/2 conv + ReLU + 3x3/2 max-pool),
                          2 = depthwise 3x3 conv, pad 1, stride 1/2 (reference models/layers.py:58-62 and
                              MobileNetV2's depthwise stages; weight (C,1,3,3)),
                          3 = fusion node: dst = scale0*src + scale1*src2 [+ scale2*src3], the last source
                              resized first (reference models/layers.py:138-177 `Fuse`),
                          4 = 3x3/2 conv from the fp32 image (MobileNetV2 stem; weight (cout,3,3,3))          */
  int src, dst;        /* buffer ids                                                                      */
  int cin, cout;       /* real channel counts                                                             */
  int ksize, stride, pad;
  int relu;            /* 0 = linear, 1 = ReLU, 2 = ReLU6 (reference models/layers.py:60,64)                             */
  int src_c_off, dst_c_off;
  int residual;        /* buffer id or -1                                                                 */
  int residual_up;     /* 1 or 2 (nearest-upsampled half-resolution residual)                             */
  int kh, kw;          /* explicit window (1..3 per side); 0 = ksize x ksize with symmetric `pad`         */
  int pad_h, pad_w;    /* top / left zero padding of the explicit window                                  */
  int dst_up;          /* 0/1 = dst has the conv's output size; 2 = dst has twice that size               */
  int dst_phase;       /* dst_up == 2: -1 = each output pixel fills its 2x2 block (conv + nearest x2
                          upsample, reference models/layers.py:98-99); 0..3 = only sub-pixel
                          (py, px) = (phase >> 1, phase & 1) is written: one phase of a stride-2
                          ConvTranspose2d (reference models/layers.py:86-96)                              */
  const float* weight_host;   /* (cout, cin, kh, kw) fp32, BatchNorm already folded                       */
  const float* bias_host;     /* (cout,) fp32                                                             */
  /* kind 3 only */
  int src2, src3;             /* further source buffers (src3 = -1: two inputs)                           */
  float scale0, scale1, scale2;
  int resize;                 /* of the LAST source: 0 = none, 1 = nearest x2 up-sample, 2 = MaxPool2d(2,2) */
} cnl_conv_desc;

typedef struct {
  int channels;
  int stride;          /* spatial size = (height/stride, width/stride)            */
  int fp32_nchw;       /* 1: (N,C,H,W) float32 (image input, head outputs)        */
} cnl_buffer_desc;

int cnl_engine_create(cnl_engine** out, const cnl_buffer_desc* buffers, int n_buffers,
                      const cnl_conv_desc* ops, int n_ops,
                      int batch, int height, int width, int precision, int device);
void cnl_engine_destroy(cnl_engine* e);

/* Bytes of device memory the engine needs for weights + activations, and the byte offset of a
 * buffer inside the arena (head outputs are read by the caller from there). */
size_t cnl_engine_arena_bytes(const cnl_engine* e);
size_t cnl_engine_buffer_offset(const cnl_engine* e, int buffer);

/* Kernel form the plan chose for op `op` (for tests and profiles): bit 0 = row-rolling A operand, bit 1 = CTA pair
 * (cta_group::2), bit 2 = separate correction accumulator, bit 3 = concatenated hi*[hi|lo] MMA, bits 4-7 = CTAs per cluster,
 * bits 8-11 = load stages, bits 12-15 = epilogue staging depth, bits 16-19 = input-row slots, bits 20-27 = Cout tile / 16;
 * -1 for a bad index. */
int cnl_engine_op_form(const cnl_engine* e, int op);

/* Upload packed weights into the arena (once, or again after a weight change). */
int cnl_engine_upload(cnl_engine* e, void* arena, void* stream);

/* Run ops [first_op, last_op) on `stream`.  image: (N,3,H,W) fp32 device pointer (buffer 0 is bound
 * to it).  CUDA-graph capturable.  Returns the number of kernels launched in *launches if non-NULL. */
int cnl_engine_forward(cnl_engine* e, void* arena, const float* image, int first_op, int last_op,
                       void* stream, int* launches);

/* The same, with the logistic 1/(1+exp(-x)) (cnl_sigmoid's arithmetic) applied by the convolution that writes fp32 output
 * buffer `sigmoid_buffer` (-1: none) - the `.sigmoid()` the reference applies to the heat-map head before decoding
 * (centernet_lightning/models/centernet.py:205; G1 forward(), tests/test_models.py:88-99), fused into the producing
 * kernel's epilogue so that the decode receives probabilities without a separate pass over the map. */
int cnl_engine_forward_act(cnl_engine* e, void* arena, const float* image, int first_op, int last_op,
                           int sigmoid_buffer, void* stream, int* launches);

/* Debug / test helpers: convert an NHWC-fp16(hi[,lo]) activation buffer to (N,C,H,W) fp32 and back. */
int cnl_engine_read_buffer(cnl_engine* e, void* arena, int buffer, float* out_nchw, void* stream);
int cnl_engine_write_buffer(cnl_engine* e, void* arena, int buffer, const float* in_nchw, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CNL_B200_H_ */
