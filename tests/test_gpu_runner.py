"""GPU: the train_BCD.py mirror end to end on a tiny synthetic split — epoch structure, checkpoint format
(reference keys, reference state-dict schema), resume, log file."""
import os

import pytest
import torch

from change3d_b200 import runner
from oracle import change3d_oracle as O

pytestmark = pytest.mark.gpu


def _args(tmp_path, extra=()):
    return runner.build_parser().parse_args(
        ["--synthetic", "8", "--in_height", "64", "--in_width", "64", "--batch_size", "2", "--num_workers", "0",
         "--max_steps", "12", "--pretrained", "/nonexistent/X3D_L.pyth", "--save_dir", str(tmp_path)] + list(extra))


def test_train_validate_checkpoint_and_resume(tmp_path, capsys):
    args = _args(tmp_path)
    scores = runner.train_validate(args)
    assert args.max_epochs == 3                                            # ceil(12 / 4 batches)
    assert set(scores) == {'Kappa', 'IoU', 'F1', 'OA', 'recall', 'precision', 'Pre'}
    save = os.path.join(str(tmp_path), "LEVIR-CD_iter_12_lr_0.0002")
    ck = torch.load(os.path.join(save, "checkpoint.pth.tar"), map_location="cpu", weights_only=False)
    assert set(ck) == {'epoch', 'arch', 'state_dict', 'optimizer', 'loss_train', 'loss_val', 'F_train', 'F_val', 'lr'}
    assert ck['epoch'] == 3 and ck['optimizer']['state']['step'] == 12
    schema = dict(O.trainer_schema("bcd", 1, 64, 64, 1))                   # the reference's key schema / shapes
    assert set(ck['state_dict']) == set(schema)
    assert all(tuple(ck['state_dict'][k].shape) == tuple(schema[k]) for k in schema)
    best = torch.load(os.path.join(save, "best_model.pth"), map_location="cpu")
    assert set(best) == set(schema)
    log_text = open(os.path.join(save, "train_val_log.txt")).read()
    assert "Epoch\tKappa (val)" in log_text and "\nTest\t\t" in log_text and "\n1\t\t" in log_text and "\n2\t\t" in log_text
    out = capsys.readouterr().out
    assert "Epoch No. 2" in out and "Test:" in out
    # resume: picks up at epoch 3 of 5 (max_steps 20), model weights only, cur_iter = epoch * max_batches
    os.rename(save, os.path.join(str(tmp_path), "LEVIR-CD_iter_20_lr_0.0002"))
    args2 = _args(tmp_path, ["--resume", "yes", "--no_graph"])
    args2.max_steps = 20
    runner.train_validate(args2)
    out = capsys.readouterr().out
    assert "=> loaded checkpoint" in out and "Epoch No. 3" in out and "Epoch No. 4" in out and "Epoch No. 2" not in out
    ck2 = torch.load(os.path.join(str(tmp_path), "LEVIR-CD_iter_20_lr_0.0002", "checkpoint.pth.tar"),
                     map_location="cpu", weights_only=False)
    assert ck2['epoch'] == 5 and ck2['optimizer']['state']['step'] == 8    # optimizer state is not restored (reference)
    assert ck2['loss_train'] == ck2['loss_train'] and ck2['loss_train'] < 5.0


def test_ragged_last_batch_runs_eagerly(tmp_path):
    args = runner.build_parser().parse_args(
        ["--synthetic", "5", "--in_height", "64", "--in_width", "64", "--batch_size", "2", "--num_workers", "0",
         "--max_steps", "3", "--pretrained", "/nonexistent/X3D_L.pyth", "--save_dir", str(tmp_path)])
    scores = runner.train_validate(args)                                  # 3 batches: 2, 2, 1 -> one epoch
    assert args.max_epochs == 1 and 'F1' in scores
