"""CPU: the numpy restatement of the input transform chain (oracle/augment_oracle.py) and the decision draws of
input_pipeline.draw_params against tests/golden/augment.npz, which the reference's unmodified data/transforms.py
produced under random.seed(16) (oracle/make_golden_augment.py)."""
import os
import random

import numpy as np
import pytest

from change3d_b200 import input_pipeline as IP
from oracle import augment_oracle as AO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment.npz")
CASES = [("bcd_train", "bcd", True), ("bcd_val", "bcd", False), ("bcd_scale_train", "bcd", True),
         ("scd_train", "scd", True), ("bda_train", "bda", True), ("bda_scale_val", "bda", False)]


@pytest.mark.parametrize("name,task,train", CASES)
def test_restatement_and_draw_order_match_reference_pipeline(name, task, train):
    z = np.load(GOLD)
    img, label, want_img, want_lab = z[name + "_img"], z[name + "_label"], z[name + "_out_img"], z[name + "_out_label"]
    H, W = want_img.shape[2], want_img.shape[3]
    random.seed(int(z["seed"]))
    params = IP.draw_params(img.shape[0], W, task, train).numpy()
    if train:
        assert params[:, 0].any() and params[:, 3:6].any()          # the seed exercises crops, flips and exchanges
    pre, post, lab = AO.augment_batch(img, label, params, H, W, task)
    got = np.concatenate([pre, post], 1)
    assert np.abs(got - want_img).max() <= 2e-6                     # float32 bilinear, tolerance 2e-6 absolute
    assert np.array_equal(lab.astype(np.int64), want_lab.astype(np.int64))


def test_validation_params_are_identity():
    p = IP.draw_params(3, 256, "scd", train=False)
    assert p.shape == (3, 8) and int(p.abs().sum()) == 0
    with pytest.raises(ValueError):
        IP.draw_params(1, 256, "cc")
