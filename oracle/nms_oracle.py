"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of the reference's YOLOv5 post-processing: ``non_max_suppression``
(src/models/yolov5.py:62-153) including the third-party ``torchvision.ops.nms`` it calls at :137
(torchvision is not vendored in the reference; README.md:55 names 0.7.0; semantics of the CPU kernel
re-derived here and pinned against torchvision 0.26.0 in the build container, see tools/make_golden.py):

  * scores sorted descending with a STABLE sort (ties: lower index first)
  * greedy; box j is suppressed by kept box i when  inter / (area_i + area_j - inter) > iou_thres, all
    arithmetic in float32, the final comparison in DOUBLE (float IoU promoted, threshold is a double)
  * areas = (x2-x1)*(y2-y1), no +1

Deviations from the reference, on purpose:
  * the 10 s wall-clock break (yolov5.py:149-151) is not reproduced (non-deterministic);
  * when more than max_nms candidates exist the reference picks the top max_nms with an UNSTABLE
    ``argsort(descending=True)`` (:131-132) whose tie order is implementation defined; the oracle defines it as
    stable (ties: lower candidate index first).  Inputs without score ties at that point behave identically.
"""
import numpy as np


def xywh2xyxy(x):
    """src/models/yolov5.py:52-59 (float32 arithmetic)."""
    y = np.empty_like(x)
    y[:, 0] = x[:, 0] - x[:, 2] / np.float32(2)
    y[:, 1] = x[:, 1] - x[:, 3] / np.float32(2)
    y[:, 2] = x[:, 0] + x[:, 2] / np.float32(2)
    y[:, 3] = x[:, 1] + x[:, 3] / np.float32(2)
    return y


def greedy_nms(boxes, scores, iou_thres, max_keep=None):
    """torchvision.ops.nms CPU semantics (see module docstring).  Returns indices into boxes, score order."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    order = np.argsort(-scores.astype(np.float32), kind='stable')
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = float(iou_thres)
    for pos, i in enumerate(order):
        if suppressed[i]:
            continue
        keep.append(int(i))
        if max_keep is not None and len(keep) >= max_keep:
            break
        rest = order[pos + 1:]
        if rest.size == 0:
            break
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide='ignore', invalid='ignore'):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, multi_label=False, max_det=300, max_nms=30000,
                        max_wh=4096, agnostic=False):
    """src/models/yolov5.py:62-153.  prediction: float32 [B, A, 5+nc].
    Returns per image (det [n,6] float32 = xyxy, conf, cls ; idx [n] int64 = anchor*nc + cls)."""
    prediction = np.asarray(prediction, dtype=np.float32)
    nc = prediction.shape[2] - 5
    conf32 = np.float32(conf_thres)
    multi_label = multi_label and nc > 1
    out = []
    for x in prediction:
        sel = np.nonzero(x[:, 4] > conf32)[0]  # :71,:90
        x = x[sel].copy()
        if x.shape[0] == 0:
            out.append((np.zeros((0, 6), np.float32), np.zeros((0,), np.int64)))
            continue
        x[:, 5:] *= x[:, 4:5]  # :106
        box = xywh2xyxy(x[:, :4])  # :109
        if multi_label:
            i, j = np.nonzero(x[:, 5:] > conf32)  # row-major like tensor.nonzero() (:113)
            det = np.concatenate([box[i], x[i, j + 5][:, None], j[:, None].astype(np.float32)], 1)
            ids = sel[i].astype(np.int64) * nc + j
        else:
            j = x[:, 5:].argmax(1)  # first maximum, like torch.max
            conf = x[np.arange(x.shape[0]), j + 5]
            m = conf > conf32
            det = np.concatenate([box, conf[:, None], j[:, None].astype(np.float32)], 1)[m]
            ids = (sel.astype(np.int64) * nc + j)[m]
        n = det.shape[0]
        if n == 0:
            out.append((np.zeros((0, 6), np.float32), np.zeros((0,), np.int64)))
            continue
        if n > max_nms:  # :131-132 (stable here, see docstring)
            top = np.argsort(-det[:, 4], kind='stable')[:max_nms]
            det, ids = det[top], ids[top]
        c = det[:, 5:6] * np.float32(0 if agnostic else max_wh)  # :135
        boxes = det[:, :4] + c  # :136
        keep = greedy_nms(boxes, det[:, 4], iou_thres, max_keep=max_det)  # :137-139
        out.append((det[keep].astype(np.float32), ids[keep]))
    return out


def make_stress_prediction(B, A=25200, nc=80, regime='typical', seed=2, img=640):
    """NMS stress set of SURVEY.md §8(d): clustered jittered anchors with unique scores.
    regimes: 'few' (kept < max_det), 'sparse' (~300 candidates/img), 'typical' (~3k), 'capped' (>> 30000 -> hits max_nms)."""
    rng = np.random.default_rng(seed)
    pred = np.zeros((B, A, 5 + nc), dtype=np.float32)
    frac_obj = {'few': 0.0008, 'sparse': 0.004, 'typical': 0.04, 'capped': 0.5}[regime]
    bg_scale = {'few': 0.0005, 'sparse': 0.0008, 'typical': 0.002, 'capped': 0.02}[regime]
    extra_scale = {'few': 0.0, 'sparse': 0.3, 'typical': 0.5, 'capped': 0.5}[regime]
    for b in range(B):
        G = int(rng.integers(20, 200)) if regime != 'few' else int(rng.integers(3, 12))
        gxy = rng.uniform(40, img - 40, size=(G, 2))
        gwh = rng.uniform(16, 220, size=(G, 2))
        gcls = rng.integers(0, nc, size=G)
        owner = rng.integers(0, G, size=A)
        on = rng.random(A) < frac_obj
        pred[b, :, 0:2] = (gxy[owner] + rng.normal(0, 8, size=(A, 2))).astype(np.float32)
        pred[b, :, 2:4] = (gwh[owner] * (1 + rng.normal(0, 0.1, size=(A, 2)))).clip(2, None).astype(np.float32)
        obj = np.where(on, rng.beta(2, 2, size=A), rng.beta(1, 30, size=A) * bg_scale)
        pred[b, :, 4] = obj.astype(np.float32)
        cls = rng.beta(1, 40, size=(A, nc)) * 0.05
        hot = rng.beta(5, 1, size=A)
        cls[np.arange(A), gcls[owner]] = hot
        extra = rng.integers(0, nc, size=A)  # a second, weaker label so multi_label matters
        cls[np.arange(A), extra] = np.maximum(cls[np.arange(A), extra], rng.beta(2, 3, size=A) * extra_scale)
        pred[b, :, 5:] = cls.astype(np.float32)
    # make every obj*cls product unique-ish: perturb obj by k*2^-20 (ties are then vanishingly rare; checked by callers)
    k = (np.arange(B * A) % 997).astype(np.float32).reshape(B, A)
    pred[:, :, 4] = np.clip(pred[:, :, 4] + k * np.float32(2.0 ** -20), 0, 1)
    return pred


def same_up_to_score_ties(ref_scores, ref_ids, test_ids):
    """True iff `test_ids` is `ref_ids` up to permutations INSIDE groups of exactly equal scores (the reference sorts with an unstable
    argsort at its 30 000 cap, src/models/yolov5.py:131-132, so the order of exactly tied candidates is not defined by the reference)."""
    ref_scores, ref_ids, test_ids = np.asarray(ref_scores), np.asarray(ref_ids), np.asarray(test_ids)
    if ref_ids.shape != test_ids.shape:
        return False
    i, n = 0, ref_ids.shape[0]
    while i < n:
        j = i
        while j + 1 < n and ref_scores[j + 1] == ref_scores[i]:
            j += 1
        if sorted(ref_ids[i:j + 1].tolist()) != sorted(test_ids[i:j + 1].tolist()):
            return False
        i = j + 1
    return True
