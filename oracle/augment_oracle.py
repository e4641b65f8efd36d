"""CPU restatement (numpy) of the reference's input transform chain for one batch — TEST INFRASTRUCTURE (the checker for
csrc/augment.cu; pinned against tests/golden/augment.npz, which the unmodified data/transforms.py produced).

Follows data/transforms.py: normalize (:127-139) -> scale (:28-35) -> random_crop_resize (:82-99) -> random_flip
(:101-114) -> random_exchange (:116-125; SCD :300-312) -> to_tensor (:141-154), with cv2.resize restated:
INTER_LINEAR on float32 (fx = (float)((dx + 0.5) * scale - 0.5), clamp, horizontal then vertical pass) and INTER_NEAREST
(sx = min(floor(dx * ifx), src - 1)).  `params` rows are input_pipeline.draw_params' rows."""
import numpy as np


def _lin(dst, src):
    scale = 1.0 / (float(dst) / float(src))
    d = np.arange(dst, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)
    lo = sx < 0
    sx[lo] = 0
    fx[lo] = 0
    hi = sx >= src - 1
    sx[hi] = src - 1
    fx[hi] = 0
    return sx, np.minimum(sx + 1, src - 1), fx


def resize_linear(img, H, W):
    """cv2.resize(img float32 (h, w, c), (W, H)) with INTER_LINEAR."""
    h, w = img.shape[:2]
    if (h, w) == (H, W):
        return img.copy()
    x0, x1, fx = _lin(W, w)
    y0, y1, fy = _lin(H, h)
    fx = fx[None, :, None]
    rows = img[:, x0] * (np.float32(1) - fx) + img[:, x1] * fx            # horizontal pass (float32)
    fy = fy[:, None, None]
    return (rows[y0] * (np.float32(1) - fy) + rows[y1] * fy).astype(np.float32)


def resize_nearest(lab, H, W):
    h, w = lab.shape[:2]
    if (h, w) == (H, W):
        return lab.copy()
    xs = np.minimum(np.floor(np.arange(W) * (1.0 / (W / w))).astype(np.int64), w - 1)
    ys = np.minimum(np.floor(np.arange(H) * (1.0 / (H / h))).astype(np.int64), h - 1)
    return lab[ys][:, xs]


def augment_batch(img_u8, label_u8, params, H, W, task, mean=0.5, std=0.5):
    """img_u8 (B, Hs, Ws, 6), label_u8 (B, Hs, Ws[, L]); returns (pre (B,3,H,W) f32, post, label (B,L,H,W))."""
    B = img_u8.shape[0]
    pre, post, labs = [], [], []
    for b in range(B):
        do_crop, x1, y1, f0, f1, ex, exl, _ = [int(v) for v in params[b]]
        im = ((img_u8[b].astype(np.float32) / np.float32(255.0)) - np.float32(mean)) / np.float32(std)
        lb = label_u8[b]
        if lb.ndim == 2:
            lb = lb[..., None]
        lb = np.ceil(lb / 255.0).astype(np.float32) if task == "bcd" else lb
        im, lb = resize_linear(im, H, W), resize_nearest(lb, H, W)
        if do_crop:
            im = resize_linear(im[y1:H - y1, x1:W - x1], H, W)
            lb = resize_nearest(lb[y1:H - y1, x1:W - x1], H, W)
        if f0:
            im, lb = im[::-1], lb[::-1]
        if f1:
            im, lb = im[:, ::-1], lb[:, ::-1]
        if ex:
            im = np.concatenate([im[:, :, 3:6], im[:, :, 0:3]], 2)
        if exl:
            lb = np.concatenate([lb[:, :, 1:2], lb[:, :, 0:1], lb[:, :, 2:]], 2)
        im = im.transpose(2, 0, 1)
        pre.append(im[0:3])
        post.append(im[3:6])
        labs.append(lb.transpose(2, 0, 1))
    lab = np.stack(labs)
    return np.stack(pre).astype(np.float32), np.stack(post).astype(np.float32), \
        (lab.astype(np.float32) if task == "bcd" else lab.astype(np.int64))
