"""Golden vectors for the GPU input pipeline (csrc/augment.cu, change3d_b200/input_pipeline.GpuAugment) — TEST
INFRASTRUCTURE, authoring container only (reads /root/reference and needs cv2).

Runs the reference's own, unmodified transform pipelines (data/transforms.py: BCDTransforms / SCDTransforms /
BDATransforms .get_transform_pipelines) sample by sample on seeded random uint8 images with `random.seed(SEED)`, and
stores inputs + outputs in tests/golden/augment.npz.  The GPU test re-draws the augmentation decisions with
`input_pipeline.draw_params` under the same seed, so it pins the draw ORDER as well as the kernel's arithmetic.

    python oracle/make_golden_augment.py
"""
import argparse
import importlib.util
import os
import random

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/data/transforms.py"
SEED = 16
CASES = [          # name, task, transforms class, (Hs, Ws) source, (H, W) output, samples, train
    ("bcd_train", "bcd", "BCDTransforms", (48, 40), (48, 40), 6, True),
    ("bcd_val", "bcd", "BCDTransforms", (48, 40), (48, 40), 2, False),
    ("bcd_scale_train", "bcd", "BCDTransforms", (37, 53), (64, 64), 4, True),
    ("scd_train", "scd", "SCDTransforms", (32, 32), (32, 32), 6, True),
    ("bda_train", "bda", "BDATransforms", (40, 32), (40, 32), 6, True),
    ("bda_scale_val", "bda", "BDATransforms", (20, 24), (40, 32), 2, False),
]


def main():
    spec = importlib.util.spec_from_file_location("ref_transforms", REF)
    T = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(T)
    out = {}
    g = np.random.default_rng(SEED)
    for name, task, cls_name, (Hs, Ws), (H, W), n, train in CASES:
        cls = getattr(T, cls_name)
        tr, va = cls.get_transform_pipelines(argparse.Namespace(in_height=H, in_width=W))
        fn = tr if train else va
        # smooth-ish images so that bilinear resampling is exercised on varying data; labels in blocks
        imgs = g.integers(0, 256, (n, Hs, Ws, 6), dtype=np.uint8)
        if task == "bcd":
            labels = (g.random((n, Hs, Ws)) < 0.3).astype(np.uint8) * 255
        elif task == "scd":
            labels = np.stack([g.integers(0, 7, (n, Hs, Ws)), g.integers(0, 7, (n, Hs, Ws)),
                               (g.random((n, Hs, Ws)) < 0.3).astype(np.int64)], -1).astype(np.uint8)
        else:
            labels = np.stack([(g.random((n, Hs, Ws)) < 0.4).astype(np.int64), g.integers(1, 5, (n, Hs, Ws))], -1).astype(np.uint8)
        random.seed(SEED)
        o_img, o_lab = [], []
        for i in range(n):
            im, lb = fn(imgs[i].copy(), labels[i].copy())
            o_img.append(im.numpy())
            o_lab.append(lb.numpy())
        out[name + "_img"] = imgs
        out[name + "_label"] = labels
        out[name + "_out_img"] = np.stack(o_img).astype(np.float32)       # (n, 6, H, W)
        out[name + "_out_label"] = np.stack(o_lab)                        # bcd (n,1,H,W) int64; scd/bda (n,L,H,W) uint8
    out["seed"] = np.int64(SEED)
    path = os.path.join(ROOT, "tests", "golden", "augment.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
