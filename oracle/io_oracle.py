"""CPU oracle (TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it) of the steps right
before and right after the detector path:

  letterbox(img, size, fill)   the reference's Resize(keep_ratio=True) transform (src/data/transforms/det_transforms.py:162-198,
                               conf/coco_yolov5_s.yml:56: size [640, 640], fill [114, 114, 114]): cv2.resize(INTER_LINEAR) of a uint8 HWC
                               frame + cv2.copyMakeBorder.  cv2 is a third-party dependency of the reference (not vendored); its 8-bit
                               bilinear kernel is restated here from OpenCV's published algorithm (imgproc/resize.cpp: 11-bit fixed-point
                               coefficients, the horizontal pass kept in int32, the vertical pass
                               ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2) and pinned against the installed cv2 and against the
                               reference's own Resize class by tools/make_golden_io.py (tests/golden/letterbox.npz).
  coco_records(...)            prepare_for_coco_detection + convert_to_xywh (src/evaluator/eval_coco.py:87-111, 200-202).
"""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def letterbox_geometry(h, w, size=(640, 640), scaleup=True):
    """(scale, oh, ow, top, bottom, left, right) exactly as det_transforms.py:177-189 computes them (Python round = half to even)."""
    scale = min(size[0] / h, size[1] / w)
    if not scaleup:
        scale = min(scale, 1.0)
    oh, ow = int(round(h * scale)), int(round(w * scale))
    padh, padw = (size[0] - oh) / 2, (size[1] - ow) / 2
    top, bottom = int(round(padh - 0.1)), int(round(padh + 0.1))
    left, right = int(round(padw - 0.1)), int(round(padw + 0.1))
    return scale, oh, ow, top, bottom, left, right


def _axis_tables(dst, src):
    """OpenCV resize.cpp: per destination index the source index and the two 11-bit coefficients (saturate_cast<short> = round to nearest even)."""
    scale = 1.0 / (dst / src)                      # double, like `scale_x = 1. / inv_scale_x`
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def resize_linear_u8(img, oh, ow):
    """cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR) for uint8 HWC images, bit for bit."""
    h, w, c = img.shape
    sx, fx = _axis_tables(ow, w)
    lo = sx < 0
    fx = np.where(lo, np.float32(0), fx)
    sx = np.where(lo, 0, sx)
    hi = sx >= w - 1
    fx = np.where(hi, np.float32(0), fx)
    sx = np.where(hi, w - 1, sx)
    a0 = np.rint((np.float32(1) - fx) * np.float32(COEF_SCALE)).astype(np.int64)
    a1 = np.rint(fx * np.float32(COEF_SCALE)).astype(np.int64)
    sx1 = np.minimum(sx + 1, w - 1)
    sy, fy = _axis_tables(oh, h)                    # rows: the coefficient is NOT zeroed at the borders, the row index is clipped
    b0 = np.rint((np.float32(1) - fy) * np.float32(COEF_SCALE)).astype(np.int64)
    b1 = np.rint(fy * np.float32(COEF_SCALE)).astype(np.int64)
    y0 = np.clip(sy, 0, h - 1)
    y1 = np.clip(sy + 1, 0, h - 1)
    src = img.astype(np.int64)
    rows = src[:, sx, :] * a0[None, :, None] + src[:, sx1, :] * a1[None, :, None]      # horizontal pass, int32 range
    r0, r1 = rows[y0], rows[y1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox(img, size=(640, 640), fill=(114, 114, 114), scaleup=True):
    """Returns (letterboxed uint8 [size0, size1, 3], pads (top, left), scales (scale, scale)) like Resize.__call__ (:176-198)."""
    h, w, _ = img.shape
    scale, oh, ow, top, bottom, left, right = letterbox_geometry(h, w, size, scaleup)
    r = resize_linear_u8(img, oh, ow) if (h != oh or w != ow) else img
    out = np.empty((oh + top + bottom, ow + left + right, 3), np.uint8)
    out[:] = np.asarray(fill, np.uint8)
    out[top:top + oh, left:left + ow] = r
    return out, (top, left), (scale, scale)


def coco_records(boxes, scores, labels, counts, image_ids, id2category=None):
    """Per kept detection of every image: (image_id, category_id, [x, y, w, h], score) with w = xmax - xmin, h = ymax - ymin in fp32
    (eval_coco.py:87-111 + convert_to_xywh :200-202).  boxes [B,M,4] xyxy f32, scores [B,M], labels [B,M] (class index), counts [B]."""
    ids, cats, xywh, sc = [], [], [], []
    for b in range(boxes.shape[0]):
        k = int(counts[b])
        bx = boxes[b, :k].astype(np.float32)
        xywh.append(np.stack([bx[:, 0], bx[:, 1], bx[:, 2] - bx[:, 0], bx[:, 3] - bx[:, 1]], 1))
        sc.append(scores[b, :k].astype(np.float32))
        lab = labels[b, :k].astype(np.int64)
        cats.append(np.asarray(id2category, np.int64)[lab] if id2category is not None else lab)
        ids.append(np.full(k, int(image_ids[b]), np.int64))
    return np.concatenate(ids), np.concatenate(cats), np.concatenate(xywh).astype(np.float32), np.concatenate(sc)
