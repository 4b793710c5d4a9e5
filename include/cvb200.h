/*
 * cvb200.h -- C ABI of libcvb200.so: the B200 (sm_100a) detector forward hot path.
 *
 * The reference (shanglianlm0525/CvPytorch) is pure Python on PyTorch and has no FFI; every entry
 * point below replaces a run of PyTorch/torchvision library calls on the reference's inference
 * path.  The "replaces" notes cite the reference file:line (paths relative to /root/reference).
 *
 * Conventions
 *   - plain pointers + sizes only; no torch / C++ types cross this boundary.
 *   - every device buffer (activations, weights, workspaces, outputs) is OWNED BY THE CALLER; the
 *     library never allocates or frees device memory and keeps no reference after a call returns,
 *     except for the raw pointers baked into a CvbConvPlan (valid while the caller keeps them alive).
 *   - all work is enqueued on the cudaStream_t passed in (as void*); no call synchronises the
 *     device or reads device memory from the host, so every *_run entry point is CUDA-graph
 *     capturable.  Plan creation only encodes TMA descriptors on the host.
 *   - return value: 0 = CVB_OK, negative = error; cvb_last_error_string() gives the thread-local text.
 *
 * Activation layout ("split-NHWC"): two fp16 planes [2][B][H][W][C]; plane 0 = hi = fp16(x),
 * plane 1 = lo = fp16(x - hi).  hi+lo carries ~22 mantissa bits, so three fp16 tensor-core
 * products (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) reproduce fp32 convolution to ~1e-5.
 */
#ifndef CVB200_H_
#define CVB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVB_OK 0
#define CVB_ERR_INVALID (-1)   /* bad argument / unsupported shape            */
#define CVB_ERR_CUDA (-2)      /* CUDA runtime / driver error                 */
#define CVB_ERR_NO_DEVICE (-3) /* no sm_100 device / driver entry point found */

#define CVB_ACT_NONE 0
#define CVB_ACT_SILU 1 /* x*sigmoid(x)   src/models/bricks/swish.py:9-25, nn.SiLU */
#define CVB_ACT_RELU 2

#define CVB_OUT_SPLIT16 0 /* two fp16 planes (hi, lo)          */
#define CVB_OUT_F32 1     /* plain fp32 NHWC (partials / heads) */
#define CVB_OUT_YOLO 2    /* fused YOLOv5 decode: the conv writes z rows + NMS histogram / per-row best scores (CvbConvDesc.yolo) */

/* A channel-slice view of an NHWC tensor. */
typedef struct CvbView {
  void* base;           /* device pointer to element (b=0,h=0,w=0,c=first channel of the view), plane 0 */
  int32_t B, H, W, C;   /* logical extent of the view                                                */
  int32_t c_pitch;      /* channels per pixel of the underlying buffer (>= C)                         */
  int64_t plane_stride; /* bytes from plane 0 (hi) to plane 1 (lo); ignored for fp32 tensors          */
} CvbView;

/*
 * Fused conv2d + folded-BN bias + activation (+ residual) (+ nearest-upsampled fp32 partial).
 *   out = act( conv(in, W) + up2x(partial) + bias ) + residual          (residual_before_act = 0, Darknet bottleneck)
 *   out = act( conv(in, W) + up2x(partial) + bias + residual )          (residual_before_act = 1, ResNet bottleneck:
 *                                                                        torchvision Bottleneck.forward `out += identity; relu`)
 * replaces: ConvModule.forward  src/models/bricks/conv_module.py:201-214 (conv -> BN -> act),
 *           Conv.forward        src/models/modules/yolo11_modules.py:27-39,
 *           the shortcut add of DarknetBottleneck.forward src/models/modules/yolo_modules.py:95-104,
 *           nn.UpsamplingNearest2d + torch.cat of UpsamplingModule.forward
 *           src/models/modules/yolo11_modules.py:388-397 (via conv1x1(up(x)) == up(conv1x1(x))),
 *           BN folding algebra src/utils/fuse.py:33-54 (done by the caller when packing weights).
 * Weights: device fp16 [2 planes][cout_pad][kh*kw*cin] with k = (ky*kw + kx)*cin + c  (K-major).
 */
/*
 * Fused head: 1x1 detect conv + sigmoid + box decode + NMS pre-pass in ONE kernel (out_kind = CVB_OUT_YOLO).  The conv's epilogue turns
 * the accumulator of (anchor a, pixel) into the z row the reference builds with permute + sigmoid + two slice writes + cat
 * (src/models/detects/yolov5_detect.py:42-55) and, when nms_workspace is given, accumulates what cvb_yolo_decode would have (score
 * histogram, per-row best score) -- the fp32 raw tensor [B,ny,nx,na*no] is never written or re-read.  Same arithmetic as
 * CVB_OUT_F32 conv + cvb_yolo_decode: bit-identical z.
 *   weights / bias: packed with ONE ANCHOR PER 128-WIDE N-TILE: row a*128 + o = output channel a*no + o of nn.Conv2d (o < no <= 96, the
 *   rest zero), i.e. cout_pad = na * 128; `out` only carries B/H/W (C = na * 128); base may be the z pointer.
 */
typedef struct CvbYoloDecode {
  int32_t na, no;          /* anchors per level (<= 4), outputs per anchor (5 + classes, <= 96) */
  float anchors_px[8];     /* anchor (w, h) in pixels = anchors * stride                        */
  float stride;
  float* z;                /* [B, z_rows, no] fp32; this level's rows start at z_off (a * ny * nx + y * nx + x inside the level) */
  int64_t z_rows, z_off;
  void* nms_workspace;     /* prepared with cvb_nms_workspace_reset (sized for A = z_rows), or NULL */
  float conf_thres;
  int32_t multi_label;
} CvbYoloDecode;

typedef struct CvbConvDesc {
  CvbView in;           /* split16 input view, C = cin (multiple of 16)              */
  CvbView out;          /* output view, C = cout                                     */
  const void* weights;  /* packed fp16 hi/lo weights, see above                      */
  int32_t cout_pad;     /* rows of the packed weight matrix (>= cout, multiple of 8) */
  const float* bias;    /* fp32 [cout_pad] folded bias (never NULL)                  */
  int32_t kh, kw, stride, pad, dilation;
  int32_t act;          /* CVB_ACT_*  */
  int32_t out_kind;     /* CVB_OUT_*  */
  CvbView residual;     /* base == NULL -> none; split16, same B/H/W/C as out        */
  CvbView up_partial;   /* base == NULL -> none; fp32 [B, ceil(H/2), ceil(W/2), C]   */
  int32_t block_n;      /* 0 = auto, else 32/64/128/256                              */
  int32_t sm_limit;     /* 0 = all SMs; else cap on the persistent grid              */
  int32_t no_resident;  /* 1 = never pin the weights in shared memory (tuning / testing)  */
  int32_t residual_before_act; /* see formula above */
  int32_t w_window;     /* 0 = off.  n > 0 ("row window" mode for tiny cin, used by the stem): `in` describes a
                           tensor that is physically zero-padded along W (1 column left, n-2 right: in.W =
                           out.W + n - 1, c_pitch == C) and the GEMM K chunk of filter row ky is the n adjacent
                           pixels starting at the output column, i.e. the packed weights are
                           [2][cout_pad][kh * n*C] with k = ky*n*C + kx*C + c and zeros for kx >= kw.
                           n*C must be 32 or 64.  Filter row ky reads input row h + ky - pad (rows outside the tensor are zero);
                           out.H is taken from the output view (allows the asymmetric 2-above/1-below padding of the
                           7x7/s2/p3 ResNet stem expressed as 4 rows over the space-to-depth input). */
  int32_t halo;         /* activation loading of multi-tap layers.  0 = auto (cost model), -1 = classic: one TMA box per filter tap,
                           1 / 2 = force "halo" mode where the layer allows it: the 8x16-pixel tile is loaded once per K chunk including
                           the filter halo (2: one box per input map; 1: one box per horizontal tap offset) and every tap is a
                           row-shifted shared-memory descriptor view of it -- up to 6x less L2->SM fill traffic for 3x3 layers. */
  const CvbYoloDecode* yolo; /* required iff out_kind == CVB_OUT_YOLO */
  float residual_scale; /* out = act(...) + residual_scale * residual (0 is read as 1): the learnable shortcut weight `alpha` of the YOLOv6
                           BottleRep block (src/models/modules/yolo_modules.py:474-492) */
} CvbConvDesc;

typedef struct CvbConvPlan CvbConvPlan;

int cvb_conv_plan_create(const CvbConvDesc* desc, CvbConvPlan** plan);
int cvb_conv_plan_run(const CvbConvPlan* plan, void* stream);
void cvb_conv_plan_destroy(CvbConvPlan* plan);
/* Diagnostics: attach a device buffer of grid * 16 int64 cycle counters (NULL detaches); every CTA then records how long its TMA
 * producer, MMA issuer and epilogue spent in total and waiting on each pipeline barrier
 * ([0..2] producer: total, wait free A slot, wait free weight slot; [4..7] MMA: total, wait accumulator, wait activations, wait
 * weights; [8..9] epilogue: total, wait accumulator).  *grid receives the number of CTAs.  Used by tools/conv_pipeline_profile.py. */
int cvb_conv_plan_set_profile(CvbConvPlan* plan, long long* counters, int32_t* grid);
/* Run n plans back to back on one stream (one C call per forward instead of one per layer). */
int cvb_conv_plan_run_many(CvbConvPlan* const* plans, int32_t n, void* stream);

/*
 * Layout / precision conversion at the drop-in boundary (reference tensors are NCHW fp32).
 * replaces: nothing in the reference (it stays NCHW); this is the module-boundary adapter.
 */
int cvb_nchw_to_split(const float* src, int32_t B, int32_t C, int32_t H, int32_t W, const CvbView* dst,
                      void* stream);
int cvb_split_to_nchw(const CvbView* src, float* dst, void* stream);
/* fp32 NHWC view (e.g. head output) -> NCHW fp32 */
int cvb_f32nhwc_to_nchw(const CvbView* src, float* dst, void* stream);
/*
 * Stem input adapter: NCHW fp32 [B,3,H,W] -> space-to-depth split16 [B,H/2,W/2,16] (12 used,
 * channel = (dy*2+dx)*3 + c, 4 zero pad) so that the 6x6/s2/p2 stem conv
 * (src/models/backbones/det/yolov5_csp_darknet.py:36-45) becomes a 3x3/s1/p1 tensor-core conv.
 * dst may also be a zero-padded row-window layout [B,H/2,W/2+3,16] (data in columns pad_left..pad_left+W/2-1; the pad
 * columns must already be zero) consumed by a CvbConvDesc with w_window = 4: pad_left = 1 for the YOLOv5 6x6 stem,
 * pad_left = 2 for the ResNet 7x7/s2/p3 stem (== 4x4 taps over the space-to-depth input, torchvision resnet conv1).
 */
int cvb_stem_s2d(const float* src, int32_t B, int32_t H, int32_t W, const CvbView* dst, int32_t pad_left, void* stream);

/*
 * Same output from the camera-side format: uint8 HWC frames [B,H,W,3].  Fuses the reference's input
 * transforms into the stem loader: ToTensor (src/data/transforms/det_transforms.py:85-99: HWC->CHW, channel reversal BGR->RGB when
 * reverse_channels != 0, float32 / 255) and Normalize (:102-109 = torchvision F.normalize, (x - mean[c]) / std[c] in fp32, mean/std
 * indexed by the OUTPUT channel; conf/coco_yolov5_s.yml:59 mean [0.406,0.456,0.485] std [0.225,0.224,0.229]).  Same operation order
 * and IEEE division as the reference, so dst is bit-identical to cvb_stem_s2d on the tensor those transforms would have produced;
 * the fp32 NCHW image never exists and the host->device copy shrinks 4x.  mean/std: host pointers to 3 floats.
 */
int cvb_stem_s2d_u8(const uint8_t* src, int32_t B, int32_t H, int32_t W, const float* mean, const float* std, int32_t reverse_channels,
                    const CvbView* dst, int32_t pad_left, void* stream);

/*
 * ResNet stem max pool 3x3 / stride 2 / pad 1 on a split16 tensor.
 * replaces: nn.MaxPool2d(3, 2, 1) of torchvision.models.resnet (src/models/backbones/seg/resnet.py:80-83).
 */
int cvb_maxpool3x3s2(const CvbView* x, const CvbView* y, void* stream);

/* split16 -> fp32 NHWC copy (an FPN level that is both a conv input and an upsampled partial of the next level) */
int cvb_split_to_f32nhwc(const CvbView* x, const CvbView* y, void* stream);

/*
 * GroupNorm with 8 channels per group (+ optional ReLU), per-sample statistics, split16 in/out.
 * replaces: nn.GroupNorm(32, 256) + nn.ReLU of the FCOS head towers (src/models/heads/fcos_head.py:39-50).
 * workspace: cvb_groupnorm_workspace_bytes(B, groups) bytes (fp64 sums), caller-owned.
 */
size_t cvb_groupnorm_workspace_bytes(int32_t B, int32_t groups);
int cvb_groupnorm_relu(const CvbView* x, int32_t groups, const float* gamma, const float* beta, float eps, int32_t relu,
                       const CvbView* y, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Depthwise 3x3 convolution (stride 1, padding = dilation) + folded-BN bias + optional ReLU, split16 in/out.
 * weights: fp32 [9][C] tap-major (BN scale folded by the caller), bias fp32 [C].
 * replaces: the depthwise ConvModule of DepthwiseSeparableConvModule (src/models/bricks/depthwise_separable_conv_module.py:73-99)
 *           used by DepthwiseSeparableASPPModule / Deeplabv3PlusHead.fuse (src/models/heads/seg/deeplabv3plus_head.py:14-54).
 */
int cvb_dwconv3x3(const CvbView* x, const float* weights, const float* bias, int32_t dilation, int32_t relu, const CvbView* y,
                  void* stream);

/* nn.AdaptiveAvgPool2d(1) of the ASPP image-pool branch (src/models/heads/seg/deeplabv3_head.py:59-62): [B,H,W,C] -> [B,1,1,C] */
int cvb_global_avgpool(const CvbView* x, const CvbView* y, void* stream);

/* F.interpolate(mode='bilinear', align_corners=False) on split16 tensors, any size (deeplabv3plus_head.py:57,63) */
int cvb_bilinear_resize(const CvbView* x, const CvbView* y, void* stream);

/*
 * Fused bilinear upsample (align_corners=False) + argmax over classes: fp32 NHWC logits [B,h,w,>=nc] -> int64 labels [B,Ho,Wo].
 * replaces: F.interpolate(preds, size=targets.shape[-2:]) + torch.argmax(dim=1) (src/models/segmentors/encoder_decoder.py:132-133).
 */
int cvb_upsample_argmax(const CvbView* logits, int32_t nc, int64_t* labels, int32_t Ho, int32_t Wo, void* stream);

/*
 * SPPF pooling: y1=maxpool5(x), y2=maxpool5(y1), y3=maxpool5(y2) (stride 1, pad 2), written to
 * three channel slices.  replaces: SPPF.forward src/models/modules/yolo_modules.py:185-194 /
 * src/models/modules/yolo11_modules.py:282-288 (3 x nn.MaxPool2d + torch.cat).
 */
int cvb_sppf_pool(const CvbView* x, const CvbView* y1, const CvbView* y2, const CvbView* y3, void* stream);

/*
 * YOLOv5 decode of one level.  raw: fp32 NHWC [B,ny,nx,c_pitch>=na*no] conv output (+bias).
 *   z[b, z_off + (a*ny+y)*nx + x, :] = decode(sigmoid(raw))   (z row pitch = no floats, z_rows rows/img)
 *   xperm (optional) = raw permuted to [B,na,ny,nx,no]
 *   nms_workspace (optional): a workspace prepared with cvb_nms_workspace_reset(); the decode then also accumulates the
 *   NMS score histogram and the per-row best score (conf_thres / multi_label as later passed to cvb_yolo_nms with
 *   hist_ready = 1; z_rows must then be the A of that call, the workspace must have been sized for it), which saves one full
 *   pass over the prediction tensor and lets the NMS emit passes skip rows below their threshold.
 * replaces: YOLOv5Detect.forward src/models/detects/yolov5_detect.py:42-55, _make_grid :60-65.
 */
int cvb_yolo_decode(const CvbView* raw, int32_t na, int32_t no, const float* anchors_px /*[na*2]*/, float stride,
                    float* z, int64_t z_rows, int64_t z_off, float* xperm, void* nms_workspace, float conf_thres,
                    int32_t multi_label, void* stream);

/*
 * Batched YOLOv5 NMS.  prediction: fp32 [B, A, 5+nc].  Outputs (fixed capacity, device):
 *   det [B, max_det, 6] = (x1,y1,x2,y2,conf,cls), det_idx [B, max_det] = anchor*nc+cls of each kept
 *   row (the "NMS-surviving box index"), det_count [B].
 * replaces: non_max_suppression src/models/yolov5.py:62-153 (multi_label / best-class, class-offset
 *           4096, max_nms 30000, max_det 300; the 10 s wall-clock break :149-151 is NOT reproduced),
 *           xywh2xyxy :52-59, torchvision.ops.nms (third party) called at :137.
 * workspace: cvb_nms_workspace_bytes(B, A, nc) bytes, caller-owned, 256-byte aligned (histogram, counters, 64-bit candidate keys,
 *            per-row best scores).  The phase-B sort is one cooperative launch (grid = number of SMs).
 * status (device int32[4], optional): [0] overflow flag (candidate capacity exceeded), rest reserved.
 */
typedef struct CvbNmsParams {
  int32_t B, A, nc;
  float conf_thres;    /* compared in fp32: x > float(conf_thres) like the reference's tensor > scalar */
  double iou_thres;    /* compared in double: torchvision's CPU kernel tests float IoU > double threshold */
  int32_t multi_label; /* 1: one candidate per (box,class) pair; 0: best class only */
  int32_t max_nms;     /* 30000 */
  int32_t max_det;     /* 300   */
  float max_wh;        /* 4096 class offset; 0 = agnostic */
  int32_t hist_ready;  /* 1: histogram + per-row best scores were already produced by cvb_yolo_decode (all levels) */
} CvbNmsParams;

size_t cvb_nms_workspace_bytes(int32_t B, int32_t A, int32_t nc);
/* zero the per-image histogram / counters (only needed before cvb_yolo_decode(..., nms_workspace, ...)) */
int cvb_nms_workspace_reset(void* workspace, size_t workspace_bytes, int32_t B, void* stream);
int cvb_yolo_nms(const float* prediction, const CvbNmsParams* p, float* det, int32_t* det_idx, int32_t* det_count,
                 void* workspace, size_t workspace_bytes, int32_t* status, void* stream);

/*
 * FCOS decode of one pyramid level.  cls: fp32 NHWC [B,h,w,>=nc] logits; regcnt: fp32 NHWC [B,h,w,>=5] = (l,t,r,b raw
 * regression, centerness logit).  Writes at location offset loc_off (of n_total locations per image):
 *   scores[B,n_total] = sqrt(max_c sigmoid(cls) * sigmoid(cnt)), classes[B,n_total] = argmax + 1 (first maximum),
 *   boxes[B,n_total,4] = (cx - exp(l*scale), cy - exp(t*scale), cx + exp(r*scale), cy + exp(b*scale)),
 *   (cx, cy) = (x*stride + stride/2, y*stride + stride/2).
 * replaces: FCOSDetect.forward / _reshape_cat_out / _coords2boxes / coords_fmap2orig
 *           (src/models/detects/fcos_detect.py:42-62,14-31,155-186) and ScaleExp (src/models/heads/fcos_head.py:13-19).
 */
int cvb_fcos_decode(const CvbView* cls, const CvbView* regcnt, int32_t nc, float stride, float scale, float* scores,
                    int32_t* classes, float* boxes, int64_t n_total, int64_t loc_off, void* stream);

/*
 * FCOS post-processing, one image per CTA, no host sync: top-k by score (ties: lower location index), score >= thres,
 * class offset = cls * (max candidate coordinate + 1), greedy NMS with '+1' areas keeping boxes with iou <= thres (fp32).
 * Outputs are fixed capacity [B,topk] (+ out_count[B]); out_loc = kept location indices ("NMS-surviving box index").
 * replaces: torch.topk + FCOSDetect._post_process / batched_nms / box_nms (src/models/detects/fcos_detect.py:64-153; the
 *           reference's python while-loop with .item() per kept box, and its torch.stack that needs equal counts).
 */
int cvb_fcos_nms(const float* scores, const int32_t* classes, const float* boxes, int32_t B, int32_t N, float score_thres,
                 float iou_thres, int32_t topk, float* out_scores, int32_t* out_classes, float* out_boxes, int32_t* out_loc,
                 int32_t* out_count, int32_t* status, void* stream);

/*
 * YOLOX post-processing, replaces yolox_post_process (src/models/yolox.py:18-68) incl. the third-party
 * torchvision.ops.batched_nms call at :64.
 *
 * cvb_yolox_decode: one head level.  reg_obj / cls: fp32 NHWC views [B,h,w,>=5] (reg 4, obj 1) and [B,h,w,>=nc] (class logits),
 * i.e. the three predictors of heads/yolox_head.py:74-86 before their torch.cat.  Writes one 8-float record per location into
 * cand[B, A, 8] at rows off .. off + h*w:  (x1, y1, x2, y2, obj, class_conf, class_pred, obj * class_conf) with
 * xy = (p + grid) * stride, wh = exp(p) * stride (:33-35), sigmoids (:37-39), corner form (:46-51), best class (:57).
 *
 * cvb_yolox_nms: per image keeps locations with obj * class_conf >= conf_thres (:58), then batched_nms(boxes, scores, class, iou_thres):
 * stable descending score order; more than `vanilla_above` boxes (torchvision: numel > 4000 on CPU tensors, i.e. 1000 boxes;
 * 5000 on CUDA tensors) -> class-wise NMS on the raw boxes, else one NMS on boxes + class * (max_coordinate + 1); fp32 IoU
 * compared in double, areas without +1.  det[B, A, 7] receives the kept rows (x1, y1, x2, y2, obj, class_conf, class_pred) in
 * score order, det_count[B] their number (the reference returns None for 0).  A <= 16384.  workspace: cvb_yolox_workspace_bytes.
 */
size_t cvb_yolox_workspace_bytes(int32_t B, int32_t A);
int cvb_yolox_decode(const CvbView* reg_obj, const CvbView* cls, int32_t nc, float stride, float* cand, int64_t A, int64_t off, void* stream);
int cvb_yolox_nms(const float* cand, int32_t B, int32_t A, float conf_thres, double iou_thres, int32_t vanilla_above, float* det,
                  int32_t* det_count, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Output side of the path, on the device instead of a per-image `.cpu().numpy()` loop.
 *
 * cvb_rescale_clip_boxes: rows [B, M, row_stride] fp32 (x1, y1, x2, y2 first), in place for the first count[b] rows:
 *   x -= pad[1]; y -= pad[0]; x /= scale[1]; y /= scale[0]; clip to [0, width] / [0, height]
 *   (src/models/yolov5.py:274-281, src/models/yolox.py:171-178, src/models/fcos.py:150-161); pads / scales / wh = device [B,2] fp32
 *   (wh = width, height).  Same fp32 operations as the numpy lines: bit-identical results.
 * cvb_confusion_matrix: cm[num_classes, num_classes] (int64, accumulated) += bincount(num_classes * gt[mask] + pred[mask]) with
 *   mask = 0 <= gt < num_classes (src/evaluator/eval_segmentation.py:52-57); gt / pred = int64 label maps of n pixels; num_classes <= 64.
 */
int cvb_rescale_clip_boxes(float* rows, int32_t B, int32_t M, int32_t row_stride, const int32_t* count, const float* pads, const float* scales,
                           const float* wh, void* stream);
int cvb_confusion_matrix(const int64_t* gt, const int64_t* pred, int64_t n, int32_t num_classes, int64_t* cm, void* stream);

/*
 * Input side of the path: the reference's Resize(size=[640,640], keep_ratio=True, fill=[114,114,114]) transform
 * (src/data/transforms/det_transforms.py:162-198, conf/coco_yolov5_s.yml:56) on the device: cv2.resize(INTER_LINEAR) of each uint8 HWC
 * frame to (oh, ow) -- OpenCV's 11-bit fixed-point 8-bit bilinear kernel restated, bit-identical to cv2 4.x -- and the constant border of
 * cv2.copyMakeBorder, for a batch of frames of DIFFERENT sizes.  The output [B,out_h,out_w,3] uint8 is what cvb_stem_s2d_u8 consumes.
 *   src_ptrs: device array of B device pointers (frame b = uint8 [h_b, w_b, 3]); geom: device int32 [B,6] = (h, w, oh, ow, top, left) as
 *   det_transforms.py:177-189 computes them (host side: python round(), see cvpytorch_b200.ops.letterbox_geometry); fill: host int32[3].
 */
int cvb_letterbox_u8(const uint8_t* const* src_ptrs, const int32_t* geom, int32_t B, int32_t out_h, int32_t out_w, const int32_t* fill,
                     uint8_t* dst, void* stream);

/*
 * Output side, COCO records on the device: replaces CocoEvaluator.prepare_for_coco_detection + convert_to_xywh
 * (src/evaluator/eval_coco.py:87-111, 200-202) for the fixed-capacity detections of a batch.  rows [B, M, row_stride] fp32 =
 * (x1, y1, x2, y2, score, class, ...) after cvb_rescale_clip_boxes, count [B]; image_ids device int64 [B]; id2category device int32
 * [num_classes] or NULL (dataset.id2category).  Writes, compacted in image order, rec_ids [n,2] int64 = (image_id, category_id) and
 * rec_box [n,5] fp32 = (x, y, x2 - x1, y2 - y1, score) (the fp32 subtraction of the torch lines), and *total = n (device int32).
 */
int cvb_coco_pack(const float* rows, int32_t B, int32_t M, int32_t row_stride, const int32_t* count, const int64_t* image_ids,
                  const int32_t* id2category, int32_t num_classes, int64_t* rec_ids, float* rec_box, int32_t* total, void* stream);

/*
 * Training step of the YOLOX C3 block (BASELINE.json configs[3]): one  BaseConv = nn.Conv2d(bias=False) ->
 * nn.BatchNorm2d (batch statistics) -> SiLU  (src/models/modules/yolox_modules.py:35-55) forward and backward, i.e. what
 * trainer.py:177-207 gets from cuDNN / ATen autograd, as bf16 tcgen05 implicit GEMMs + element-wise passes (csrc/train_kernels.cu).
 * All activations / gradients: device bf16 NHWC, dense ([B,H,W,C], C a multiple of 64); convolutions k = 1 or 3, stride 1, pad k/2.
 *
 * cvb_train_pack_weights: fp32 master weights [cout,cin,k,k] -> w_fwd bf16 [cout][k*k][cin] and w_bwd bf16 [cin][k*k][cout] (taps rotated
 *   by 180 degrees: the operand of the backward-data convolution).
 * cvb_train_conv: out[B,H,W,cout] = conv(x[B,H,W,cin], w_packed [cout][k*k][cin]).  Forward: (x, w_fwd).  Backward-data: (dy, w_bwd) with
 *   cin/cout swapped.  y_prev / bn_stat_prev != NULL: SiLU' epilogue -- the result is multiplied by silu'(y_prev * scale + shift) of the
 *   layer that produced this conv's input (bn_stat layout below), giving dz of that layer directly.
 *   stride = 2 (k = 3, pad 1; the downsampling BaseConv in front of every CSPLayer, src/models/backbones/det/csp_darknet.py): the forward
 *   reads the input through four parity tensor maps, out is [B,(H-1)/2+1,(W-1)/2+1,cout].
 * cvb_train_conv_dgrad_s2: backward-data of that stride-2 convolution, dx[B,H,W,cin] from dy[B,Ho,Wo,cout] and w_bwd: four dense
 *   sub-convolutions, one per parity class of dx (1 + 2 + 2 + 4 taps), same optional SiLU' epilogue.
 * cvb_train_conv_wgrad: dw[cout][k*k][cin] (fp32, ZEROED BY THE CALLER, accumulated with atomics) += sum over pixels of dy x (shifted x);
 *   x is [B,H,W,cin], dy the conv output's gradient (stride 1: same H, W; stride 2: (H-1)/2+1 etc.); cout in {64, 128, multiples of 256}.
 * cvb_train_bn_stats: batch mean / biased variance of y over npix = B*H*W -> stat [4][C] fp32 = (mean, rstd, scale = gamma*rstd,
 *   shift = beta - mean*scale); running_mean / running_var (may be NULL) updated like nn.BatchNorm2d (momentum, unbiased variance).
 *   sums_scratch: [2][C] fp32.
 * cvb_train_bn_silu_fwd: out = silu(y * scale + shift).
 * cvb_train_bn_silu_bwd: g = d(loss)/d(out) (or, g_is_dz != 0, already multiplied by silu'): sums [2][C] = (dbeta, dgamma) and
 *   dy = gamma * rstd * (dz - dbeta / N - xhat * dgamma / N)   (the batch-norm backward of torch.autograd).
 */
int cvb_train_pack_weights(const float* w, int32_t cout, int32_t cin, int32_t k, void* w_fwd, void* w_bwd, void* stream);
int cvb_train_conv(const void* x, int32_t B, int32_t H, int32_t W, int32_t cin, const void* w_packed, int32_t cout, int32_t k, int32_t stride, void* out,
                   const void* y_prev, const float* bn_stat_prev, void* stream);
int cvb_train_conv_dgrad_s2(const void* dy, int32_t B, int32_t Ho, int32_t Wo, int32_t cout, const void* w_bwd, int32_t cin, int32_t H, int32_t W, void* dx,
                            const void* y_prev, const float* bn_stat_prev, void* stream);
int cvb_train_conv_wgrad(const void* x, const void* dy, int32_t B, int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t k, int32_t stride, float* dw,
                         void* stream);
int cvb_train_bn_stats(const void* y, int64_t npix, int32_t C, const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                       float* running_var, float* sums_scratch, float* stat, void* stream);
int cvb_train_bn_silu_fwd(const void* y, int64_t npix, int32_t C, const float* stat, void* out, void* stream);
int cvb_train_bn_silu_bwd(const void* g, int32_t g_is_dz, const void* y, int64_t npix, int32_t C, const float* stat, const float* gamma, float* sums,
                          void* dy, void* stream);

/* Library info / errors */
const char* cvb_last_error_string(void);
int cvb_version(void);
/* number of kernels this library has launched in this process (for bench.py's gpu_launches) */
int64_t cvb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CVB200_H_ */
