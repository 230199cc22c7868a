"""TEST INFRASTRUCTURE -- CPU restatement (numpy, float32) of the reference's validation-loss path: the criterion
`Loss.__call__` (reference nets/yolo_training.py:371-430) with its task-aligned assigner (:75-225), CIoU (:227-265),
`bbox2dist` (:267-270) and `BboxLoss` (:272-303).  Nothing under dcfa-yolo_b200/ imports this module; only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may.

Pinned by tests/test_loss_cpu.py against tests/golden/loss_*.npz, which oracle/make_golden_loss.py produced by running
the REAL reference criterion (imported from /root/reference) on the same seeded inputs.

One place of the reference is unspecified: `torch.topk` (:190) does not define which of several EQUAL values it returns.
That matters only when a ground-truth box has fewer than `topk` anchors with a positive alignment metric and, at the same
time, anchors inside the box whose metric is exactly zero.  This restatement (and the CUDA path) takes the lowest anchor
index first; the goldens are generated on inputs where the choice cannot change the result (the generator checks it).
"""
import numpy as np

F32 = np.float32
EPS_IOU = F32(1e-7)      # bbox_iou's eps (:227)
EPS_TAL = F32(1e-9)      # TaskAlignedAssigner.eps (:77) and select_candidates_in_gts' eps (:12)
TOPK, ALPHA, BETA = 10, 0.5, 6.0   # Loss.__init__ (:333-337)
GAIN_BOX, GAIN_CLS, GAIN_DFL = F32(7.5), F32(0.5), F32(1.5)   # (:426-428)


def make_anchors(shapes, strides, offset=0.5):
    """utils/utils_bbox.py:16-28: cell centres (x, y) in grid units and the stride of each anchor, level after level."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = np.arange(w, dtype=F32) + F32(offset)
        sy = np.arange(h, dtype=F32) + F32(offset)
        yy, xx = np.meshgrid(sy, sx, indexing='ij')
        pts.append(np.stack((xx, yy), -1).reshape(-1, 2))
        st.append(np.full((h * w, 1), s, dtype=F32))
    return np.concatenate(pts), np.concatenate(st)


def ciou(b1, b2):
    """bbox_iou(box1, box2, xywh=False, CIoU=True) (:227-262); b1, b2 broadcastable (..., 4) float32 xyxy."""
    b1 = b1.astype(F32)
    b2 = b2.astype(F32)
    x11, y11, x12, y12 = (b1[..., i] for i in range(4))
    x21, y21, x22, y22 = (b2[..., i] for i in range(4))
    w1, h1 = x12 - x11, y12 - y11 + EPS_IOU
    w2, h2 = x22 - x21, y22 - y21 + EPS_IOU
    inter = np.maximum(np.minimum(x12, x22) - np.maximum(x11, x21), F32(0)) * \
        np.maximum(np.minimum(y12, y22) - np.maximum(y11, y21), F32(0))
    union = w1 * h1 + w2 * h2 - inter + EPS_IOU
    iou = inter / union
    cw = np.maximum(x12, x22) - np.minimum(x11, x21)
    ch = np.maximum(y12, y22) - np.minimum(y11, y21)
    c2 = cw * cw + ch * ch + EPS_IOU
    dx = x21 + x22 - x11 - x12
    dy = y21 + y22 - y11 - y12
    rho2 = (dx * dx + dy * dy) / F32(4)
    da = np.arctan(w2 / h2) - np.arctan(w1 / h1)
    v = F32(4 / np.pi ** 2) * (da * da)
    alpha = v / (v - iou + (F32(1) + EPS_IOU))
    return (iou - (rho2 / c2 + v * alpha)).astype(F32)


def preprocess(targets, batch_size, scale):
    """Loss.preprocess (:342-360): (n, 6) rows [image, class, cx, cy, w, h] (normalised) -> (B, G, 5) rows
    [class, x1, y1, x2, y2] in input pixels, zero padded, in order of appearance."""
    targets = np.asarray(targets, dtype=F32).reshape(-1, 6)
    if targets.shape[0] == 0:
        return np.zeros((batch_size, 0, 5), F32)
    img = targets[:, 0]
    counts = [int((img == v).sum()) for v in np.unique(img)]
    out = np.zeros((batch_size, max(counts), 5), F32)
    for j in range(batch_size):
        rows = targets[img == j, 1:]
        out[j, :len(rows)] = rows
    xywh = out[..., 1:5] * scale.astype(F32)
    half_w, half_h = xywh[..., 2] / F32(2), xywh[..., 3] / F32(2)
    out[..., 1:5] = np.stack((xywh[..., 0] - half_w, xywh[..., 1] - half_h, xywh[..., 0] + half_w, xywh[..., 1] + half_h), -1)
    return out


def dfl_expectation(dist):
    """Loss.bbox_decode (:362-369): softmax over the 16 bins of each side, dotted with 0..15.  dist (..., 4, 16)."""
    m = dist.max(-1, keepdims=True)
    e = np.exp((dist - m).astype(F32)).astype(F32)
    p = e / e.sum(-1, keepdims=True, dtype=F32)
    return (p * np.arange(dist.shape[-1], dtype=F32)).sum(-1, dtype=F32)


def log_sigmoid(x):
    return np.minimum(x, F32(0)) - np.log1p(np.exp(-np.abs(x)).astype(F32)).astype(F32)


def topk_lowest_index_first(values, k):
    """Indices of the k largest entries; among equal values the lowest index wins (see the module docstring)."""
    order = np.lexsort((np.arange(values.size), -values.astype(np.float64)))
    return order[:k]


def assign_image(scores, boxes_px, anc_px, labels, gts, valid):
    """TaskAlignedAssigner.forward for ONE image (:88-128).  scores (A, nc) sigmoid outputs, boxes_px (A, 4) predicted
    boxes in pixels, anc_px (A, 2), labels (G,), gts (G, 4) pixels, valid (G,) bool.  Returns fg (A,) bool,
    gt index (A,), norm_align_metric (A,) -- the value every foreground anchor's target score carries."""
    A, G = scores.shape[0], gts.shape[0]
    lab = labels.astype(np.int64)
    overlaps = np.maximum(ciou(gts[:, None, :], boxes_px[None, :, :]), F32(0))          # (G, A)  (:171)
    align = np.sqrt(scores[:, lab].T.astype(F32)) * np.power(overlaps, F32(BETA)).astype(F32)   # (:172)
    d = np.stack((anc_px[None, :, 0] - gts[:, None, 0], anc_px[None, :, 1] - gts[:, None, 1],
                  gts[:, None, 2] - anc_px[None, :, 0], gts[:, None, 3] - anc_px[None, :, 1]), -1)
    in_gts = d.min(-1) > EPS_TAL                                                          # (:34-38)
    mask_pos = np.zeros((G, A), bool)
    for g in range(G):
        if not valid[g]:
            continue          # topk indices forced to 0, counted `topk` times, zeroed by the `> 1` rule (:194-203)
        idx = topk_lowest_index_first(align[g] * in_gts[g].astype(F32), TOPK)
        mask_pos[g, idx] = True
        mask_pos[g] &= in_gts[g]
    fg_count = mask_pos.sum(0)
    multi = fg_count > 1                                                                  # (:56-69)
    if multi.any():
        best = overlaps.argmax(0)
        mask_pos[:, multi] = False
        mask_pos[best[multi], np.nonzero(multi)[0]] = True
    fg = mask_pos.any(0)
    gt_idx = mask_pos.argmax(0)
    al = align * mask_pos                                                                 # (:118-126)
    pos_al = al.max(1, keepdims=True)
    pos_ov = (overlaps * mask_pos).max(1, keepdims=True)
    norm = (al * pos_ov / (pos_al + EPS_TAL)).max(0)
    return fg, gt_idx, norm.astype(F32), mask_pos


def loss_forward(feats, targets, strides=(8., 16., 32.), nc=1, reg_max=16, details=None):
    """Loss.__call__ (:371-430).  feats: three (B, 4*reg_max + nc, Hi, Wi) float32 arrays (the `x` list that
    YoloBody.forward returns); targets (n, 6).  Returns float32 [box*7.5, cls*0.5, dfl*1.5, sum]."""
    feats = [np.asarray(f, dtype=F32) for f in feats]
    B = feats[0].shape[0]
    no = 4 * reg_max + nc
    pred = np.concatenate([f.reshape(B, no, -1) for f in feats], 2)
    distri = np.ascontiguousarray(pred[:, :4 * reg_max].transpose(0, 2, 1))     # (B, A, 64)
    logits = np.ascontiguousarray(pred[:, 4 * reg_max:].transpose(0, 2, 1))     # (B, A, nc)
    A = distri.shape[1]
    imgsz = np.array(feats[0].shape[2:], F32) * F32(strides[0])                 # (h, w)
    anc, st = make_anchors([f.shape[2:] for f in feats], strides)
    gt = preprocess(targets, B, imgsz[[1, 0, 1, 0]])
    G = gt.shape[1]
    ltrb = dfl_expectation(distri.reshape(B, A, 4, reg_max))
    pbox = np.concatenate((anc[None] - ltrb[..., :2], anc[None] + ltrb[..., 2:]), -1).astype(F32)   # grid units
    scores = (F32(1) / (F32(1) + np.exp(-logits).astype(F32))).astype(F32)
    anc_px = anc * st
    tscore = np.zeros((B, A, nc), F32)
    tbox = np.zeros((B, A, 4), F32)
    fg_all = np.zeros((B, A), bool)
    gi_all = np.zeros((B, A), np.int64)
    for b in range(B):
        if G == 0:
            break
        valid = gt[b, :, 1:].sum(-1, dtype=F32) > 0                              # mask_gt (:401)
        fg, gi, norm, _ = assign_image(scores[b], pbox[b] * st, anc_px, gt[b, :, 0], gt[b, :, 1:], valid)
        fg_all[b], gi_all[b] = fg, gi
        tbox[b] = gt[b, gi, 1:] / st
        rows = np.nonzero(fg)[0]
        tscore[b, rows, gt[b, gi[rows], 0].astype(np.int64)] = norm[rows]
    tss = max(F32(tscore.sum(dtype=np.float64)), F32(1))
    bce = (F32(1) - tscore) * logits - log_sigmoid(logits)
    loss_cls = F32(bce.sum(dtype=np.float64)) / tss
    loss_box = loss_dfl = F32(0)
    if fg_all.any():
        w = tscore.sum(-1, dtype=F32)[fg_all]
        iou = ciou(pbox[fg_all], tbox[fg_all])
        loss_box = F32(((F32(1) - iou) * w).sum(dtype=np.float64)) / tss
        a_full = np.broadcast_to(anc[None], (B, A, 2))[fg_all]
        tb = tbox[fg_all]
        t = np.clip(np.concatenate((a_full - tb[:, :2], tb[:, 2:] - a_full), -1), F32(0), F32(reg_max - 1 - 0.01)).astype(F32)
        tl = t.astype(np.int64)
        wl = (tl + 1).astype(F32) - t
        wr = F32(1) - wl
        d = distri[fg_all].reshape(-1, 4, reg_max)
        m = d.max(-1, keepdims=True)
        lse = (m + np.log(np.exp(d - m).sum(-1, keepdims=True, dtype=F32))).astype(F32)
        logp = d - lse
        ce_l = -np.take_along_axis(logp, tl[..., None], -1)[..., 0]
        ce_r = -np.take_along_axis(logp, (tl + 1)[..., None], -1)[..., 0]
        dfl = (ce_l * wl + ce_r * wr).mean(-1, dtype=F32)
        loss_dfl = F32((dfl * w).sum(dtype=np.float64)) / tss
    parts = np.array([loss_box * GAIN_BOX, loss_cls * GAIN_CLS, loss_dfl * GAIN_DFL], F32)
    if details is not None:
        details.update(fg=fg_all, gt_idx=gi_all, target_scores_sum=tss, n_max=G, target_scores=tscore)
    return np.append(parts, F32(parts.sum(dtype=F32)))


def synth_case(seed, B, nc, hw0, n_targets, tiny=0, crowd=0, empty_image=None, reg_max=16, spread=1.0):
    """Deterministic inputs for the loss tests (numpy MT19937): three head maps and a target list.  `tiny` extra boxes
    of a few pixels, `crowd` extra boxes packed around one point (anchors claimed by several boxes), one image may be
    left without targets."""
    rs = np.random.RandomState(seed)
    h0, w0 = hw0
    no = 4 * reg_max + nc
    feats = [(rs.standard_normal((B, no, h0 >> i, w0 >> i)) * spread).astype(F32) for i in range(3)]
    rows = []
    for b in range(B):
        if empty_image is not None and b == empty_image:
            continue
        for _ in range(n_targets):
            w, h = rs.uniform(0.08, 0.6, 2)
            cx, cy = rs.uniform(w / 2, 1 - w / 2), rs.uniform(h / 2, 1 - h / 2)
            rows.append([b, rs.randint(nc), cx, cy, w, h])
        for _ in range(tiny):
            cx, cy = rs.uniform(0.1, 0.9, 2)
            rows.append([b, rs.randint(nc), cx, cy, rs.uniform(0.004, 0.03), rs.uniform(0.004, 0.03)])
        if crowd:
            cx, cy = rs.uniform(0.3, 0.7, 2)
            for _ in range(crowd):
                rows.append([b, rs.randint(nc), cx + rs.uniform(-0.02, 0.02), cy + rs.uniform(-0.02, 0.02),
                             rs.uniform(0.2, 0.3), rs.uniform(0.2, 0.3)])
    targets = np.array(rows, F32).reshape(-1, 6)
    return feats, targets
