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


def test_runner_takes_raw_uint8_batches(tmp_path):
    """A dataset that yields raw pairs (uint8 HWC + uint8 mask) trains through the BCD runner: the transform chain runs
    on the device (4x fewer host-to-device bytes, no CPU image arithmetic)."""
    from change3d_b200 import runner

    class RawPairs(torch.utils.data.Dataset):
        def __init__(self, n, seed):
            self.n, self.seed = n, seed

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            g = torch.Generator().manual_seed(self.seed * 1000 + i)
            img = torch.randint(0, 256, (64, 64, 6), generator=g, dtype=torch.uint8)
            lab = torch.zeros(64, 64, dtype=torch.uint8)
            lab[16:40, 8:32] = 255
            img[16:40, 8:32, 3:] = 255 - img[16:40, 8:32, :3]
            return img, lab

    args = runner.build_parser().parse_args(
        ["--in_height", "64", "--in_width", "64", "--batch_size", "2", "--num_workers", "0", "--max_steps", "6",
         "--pretrained", "/nonexistent/X3D_L.pyth", "--save_dir", str(tmp_path)])
    random.seed(3)
    scores = runner.train_validate(args, datasets=(RawPairs(6, 1), RawPairs(2, 2), RawPairs(2, 3)))
    assert args.max_epochs == 2 and 'F1' in scores and scores['OA'] == scores['OA']
