"""GPU: the train_SCD.py / train_BDA.py mirrors end to end on tiny synthetic splits — epoch structure, checkpoint keys
of each script, log columns, CUDA-graph step with a ragged last batch, device-side metrics vs a host recomputation."""
import os

import numpy as np
import pytest
import torch

from change3d_b200 import runner_tasks as RT
from oracle import change3d_oracle as O

pytestmark = pytest.mark.gpu


def _args(task, tmp_path, n="7", extra=()):
    return RT.build_parser(task).parse_args(
        ["--synthetic", n, "--in_height", "64", "--in_width", "64", "--batch_size", "2", "--num_workers", "0",
         "--max_steps", "8", "--pretrained", "/nonexistent/X3D_L.pyth", "--save_dir", str(tmp_path)] + list(extra))


def test_scd_train_validate(tmp_path, capsys):
    args = _args("scd", tmp_path, extra=["--num_class", "7", "--dataset", "SECOND"])
    res = RT.train_validate(args, "scd")                 # 7 samples, batch 2: 4 batches (last ragged) x 2 epochs
    assert args.max_epochs == 2 and set(res) == {"Fscd", "IoU_mean", "Sek", "acc", "loss"}
    assert 0.0 <= res["acc"] <= 1.0 and res["loss"] == res["loss"]
    save = os.path.join(str(tmp_path), "SECOND_iter_8_lr_0.0002")
    ck = torch.load(os.path.join(save, "checkpoint.pth.tar"), map_location="cpu", weights_only=False)
    assert set(ck) == {'epoch', 'arch', 'state_dict', 'optimizer', 'loss_train', 'loss_val', 'acc_train', 'acc_val', 'lr'}
    schema = dict(O.trainer_schema("scd", 7, 64, 64, 3))
    assert set(ck['state_dict']) == set(schema) and ck['optimizer']['state']['step'] == 8
    log_text = open(os.path.join(save, "train_val_log.txt")).read()
    assert "epoch\ttrain_loss\ttrain_acc\tval_Fscd" in log_text and "\n1\t\t" in log_text and "\nTest\t\t" in log_text
    assert "Best rec: Train acc" in capsys.readouterr().out


def test_bda_train_validate(tmp_path, capsys):
    args = _args("bda", tmp_path)
    res = RT.train_validate(args, "bda")
    assert args.max_epochs == 2 and set(res) == {"loss", "loc_f1", "harmonic_mean_f1", "oa_f1", "damage_f1"}
    assert len(res["damage_f1"]) == 4
    save = os.path.join(str(tmp_path), "xBD_iter_8_lr_0.0002")
    ck = torch.load(os.path.join(save, "checkpoint.pth.tar"), map_location="cpu", weights_only=False)
    assert set(ck) == {'epoch', 'arch', 'state_dict', 'optimizer', 'loss_train', 'loss_val', 'loc_f1_score',
                       'harmonic_mean_f1', 'lr'}
    assert set(ck['state_dict']) == set(dict(O.trainer_schema("bda", 5, 64, 64, 2)))
    assert "epoch\tloss_val\tloc_f1_score" in open(os.path.join(save, "train_val_log.txt")).read()
    assert "oaf1 =" in capsys.readouterr().out


def test_device_meters_match_host_recomputation():
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(3)
    B, C, H, W = 3, 7, 20, 24
    pre_mask, post_mask = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    change = torch.rand(B, 1, H, W, generator=g)
    la, lb = torch.randint(0, C, (B, H, W), generator=g), torch.randint(0, C, (B, H, W), generator=g)
    m = RT._ScdMeter(C, dev)
    m.update(pre_mask.to(dev), post_mask.to(dev), change.to(dev), la.to(dev), lb.to(dev), with_hist=True)
    chg = (change > 0.5).squeeze(1).long()
    pa, pb = (pre_mask.argmax(1) * chg).numpy(), (post_mask.argmax(1) * chg).numpy()
    hist = np.zeros((C, C), dtype=np.int64)
    accs = []
    for i in range(B):                                   # train_SCD.py:160-170 + model/utils.py:313-328
        for p, l in ((pa[i], la[i].numpy()), (pb[i], lb[i].numpy())):
            hist += np.bincount(C * p.reshape(-1) + l.reshape(-1), minlength=C * C).reshape(C, C)
        accs.append(0.5 * ((pa[i] == la[i].numpy()).mean() + (pb[i] == lb[i].numpy()).mean()))
    assert np.array_equal(m.hist.cpu().numpy(), hist)
    assert abs(m.average() - float(np.mean(accs))) < 1e-9
