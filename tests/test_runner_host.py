"""CPU: host logic of change3d_b200.runner (the scripts/train_BCD.py mirror) — flag surface, learning-rate
schedule through FlatAdam-style param_groups, synthetic dataset layout, DistributedSampler sharding."""
import os
import re
from argparse import Namespace

import pytest
import torch

from change3d_b200 import runner
from change3d_b200.model.utils import adjust_learning_rate
from oracle import change3d_oracle as O

# scripts/train_BCD.py:363-484 (name -> default), written down from the reference
REFERENCE_FLAGS = {
    "dataset": "LEVIR-CD", "file_root": "path/to/LEVIR-CD", "in_height": 256, "in_width": 256,
    "num_perception_frame": 1, "num_class": 1, "max_steps": 80000, "batch_size": 16, "num_workers": 4, "lr": 2e-4,
    "lr_mode": "poly", "step_loss": 100, "pretrained": "model/X3D_L.pyth", "save_dir": "./exp", "resume": None,
    "log_file": "train_val_log.txt", "gpu_id": 0,
}


def test_flag_surface_matches_reference_script():
    args = runner.build_parser().parse_args([])
    for k, v in REFERENCE_FLAGS.items():
        assert getattr(args, k) == v, k
    assert set(vars(args)) - set(REFERENCE_FLAGS) == {"synthetic", "no_graph"}
    ref = "/root/reference/scripts/train_BCD.py"
    if os.path.isfile(ref):                     # authoring container: the list above is the script's own
        flags = set(re.findall(r"'--(\w+)'", open(ref).read()))
        assert flags == set(REFERENCE_FLAGS)


def test_poly_lr_with_warmup_matches_oracle():
    class Opt:                                   # FlatAdam exposes the same attribute
        param_groups = [{"lr": 0.0}]
    args = Namespace(lr=2e-4, lr_mode="poly", max_epochs=5, step_loss=100)
    for epoch, it in ((0, 0), (0, 199), (0, 200), (1, 450), (4, 999)):
        lr = adjust_learning_rate(args, Opt, epoch, it, 200)
        assert Opt.param_groups[0]["lr"] == lr
        assert abs(lr - O.poly_lr(2e-4, it, 200 * 5, epoch)) < 1e-12


def test_synthetic_dataset_layout_and_sharding():
    ds = runner.SyntheticBCD(10, 32, 24, seed=1)
    img, target = ds[3]
    assert img.shape == (6, 32, 24) and target.shape == (1, 32, 24) and img.dtype == torch.float32
    assert set(target.unique().tolist()) <= {0.0, 1.0} and target.sum() > 0
    assert torch.equal(ds[3][0], img)                                  # deterministic per index
    args = Namespace(synthetic=10, in_height=32, in_width=24, batch_size=2, num_workers=0)
    seen = []
    for rank in range(2):
        tl, _vl, te, max_batches = runner.make_loaders(args, world=2, rank=rank)
        assert max_batches == 3 and len(te.dataset) == 2
        tl.sampler.set_epoch(0)
        seen.append(list(iter(tl.sampler)))
    assert sorted(seen[0] + seen[1]) == list(range(10))                # the two ranks split the epoch
    with pytest.raises(RuntimeError, match="dataset"):
        runner.make_loaders(Namespace(synthetic=0), 1, 0)


def test_runner_refuses_to_run_without_cuda(tmp_path):
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    args = runner.build_parser().parse_args(["--synthetic", "4", "--save_dir", str(tmp_path)])
    with pytest.raises(RuntimeError, match="no CPU"):
        runner.train_validate(args)
