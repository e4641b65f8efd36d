"""GPU: csrc/augment.cu through input_pipeline.GpuAugment against the golden vectors the reference's own transform
pipelines produced (tests/golden/augment.npz) and against the numpy restatement at the training shape."""
import os
import random

import numpy as np
import pytest
import torch

from change3d_b200 import input_pipeline as IP
from oracle import augment_oracle as AO

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment.npz")
CASES = [("bcd_train", "bcd", True), ("bcd_val", "bcd", False), ("bcd_scale_train", "bcd", True),
         ("scd_train", "scd", True), ("bda_train", "bda", True), ("bda_scale_val", "bda", False)]


@pytest.mark.parametrize("name,task,train", CASES)
def test_matches_reference_pipeline_golden(name, task, train):
    z = np.load(GOLD)
    img, label, want_img, want_lab = z[name + "_img"], z[name + "_label"], z[name + "_out_img"], z[name + "_out_label"]
    H, W = want_img.shape[2], want_img.shape[3]
    random.seed(int(z["seed"]))
    params = IP.draw_params(img.shape[0], W, task, train)
    pre, post, lab = IP.GpuAugment(H, W, task)(torch.from_numpy(img).to(DEV), torch.from_numpy(label).to(DEV), params)
    got = torch.cat([pre, post], 1).cpu().numpy()
    assert np.abs(got - want_img).max() <= 2e-6                     # float32 bilinear, tolerance 2e-6 absolute
    assert lab.dtype == (torch.float32 if task == "bcd" else torch.int64)
    assert np.array_equal(lab.cpu().numpy().astype(np.int64), want_lab.astype(np.int64))


def test_training_shape_against_restatement_and_feeds_the_model():
    g = np.random.default_rng(1)
    B, H, W = 4, 256, 256
    img = g.integers(0, 256, (B, H, W, 6), dtype=np.uint8)
    label = (g.random((B, H, W)) < 0.05).astype(np.uint8) * 255
    random.seed(5)
    params = IP.draw_params(B, W, "bcd", True)
    pre, post, lab = IP.GpuAugment(H, W, "bcd")(torch.from_numpy(img).to(DEV), torch.from_numpy(label).to(DEV), params)
    wp, wq, wl = AO.augment_batch(img, label, params.numpy(), H, W, "bcd")
    assert np.abs(pre.cpu().numpy() - wp).max() <= 2e-6 and np.abs(post.cpu().numpy() - wq).max() <= 2e-6
    assert np.array_equal(lab.cpu().numpy(), wl)
    assert pre.shape == (B, 3, H, W) and pre.is_contiguous() and lab.shape == (B, 1, H, W)
    assert float(pre.min()) >= -1.0 and float(pre.max()) <= 1.0


def test_rejects_host_tensors():
    with pytest.raises(RuntimeError, match="no CPU path"):
        IP.GpuAugment(8, 8)(torch.zeros(1, 8, 8, 6, dtype=torch.uint8), None)
