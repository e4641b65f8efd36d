"""CPU: host logic of change3d_b200.runner_tasks (the scripts/train_SCD.py / train_BDA.py mirrors) — flag surfaces,
synthetic dataset layouts, and the histogram-based metric formulas against the reference's own functions
(model/utils.py:313-430: accuracy / SCDD_eval_all / Evaluator) where the reference tree is mounted, and against
hand-computed values everywhere."""
import os
import re

import numpy as np
import pytest
import torch

from change3d_b200 import runner_tasks as RT
from oracle import reference_loader as R

# scripts/train_SCD.py:443-552 and scripts/train_BDA.py:373-482 (name -> default), written down from the reference
COMMON = {"in_height": 256, "in_width": 256, "num_workers": 4, "lr": 2e-4, "lr_mode": "poly", "step_loss": 100,
          "pretrained": "model/X3D_L.pyth", "save_dir": "./exp", "resume": None, "log_file": "train_val_log.txt", "gpu_id": 0}
FLAGS = {
    "scd": dict(COMMON, dataset="HRSCD", file_root="path/to/HRSCD", num_perception_frame=3, num_class=6, max_steps=80000,
                batch_size=8),
    "bda": dict(COMMON, dataset="xBD", file_root="path/to/xBD", num_perception_frame=2, num_class=5, max_steps=200000,
                batch_size=12),
}


@pytest.mark.parametrize("task,script", [("scd", "train_SCD.py"), ("bda", "train_BDA.py")])
def test_flag_surface_matches_reference_script(task, script):
    args = RT.build_parser(task).parse_args([])
    for k, v in FLAGS[task].items():
        assert getattr(args, k) == v, k
    assert set(vars(args)) - set(FLAGS[task]) == {"synthetic", "no_graph"}
    ref = os.path.join("/root/reference/scripts", script)
    if os.path.isfile(ref):                     # authoring container: the list above is the script's own
        assert set(re.findall(r"'--(\w+)'", open(ref).read())) == set(FLAGS[task])


def test_synthetic_layouts():
    img, lab = RT.SyntheticSCD(4, 32, 24, 7, seed=1)[2]
    assert img.shape == (6, 32, 24) and lab.shape == (3, 32, 24) and lab.dtype == torch.int64
    assert set(lab[2].unique().tolist()) == {0, 1} and 1 <= int(lab[0].min()) and int(lab[0].max()) <= 6
    img, lab = RT.SyntheticBDA(4, 32, 24, 5, seed=1)[2]
    assert img.shape == (6, 32, 24) and lab.shape == (2, 32, 24)
    cls = torch.prod(lab, dim=0)                                    # train_BDA.py:118
    assert set(cls.unique().tolist()) <= {0, 1, 2, 3, 4} and int((cls > 0).sum()) == int(lab[0].sum())


def _rand_maps(n_class, seed, n=6, hw=(17, 13)):
    g = np.random.default_rng(seed)
    preds = [g.integers(0, n_class, hw) for _ in range(n)]
    labels = [np.where(g.random(hw) < 0.6, p, g.integers(0, n_class, hw)) for p in preds]
    return preds, labels


def test_scd_scores_from_hist_hand_case():
    # 2 classes + background; hist[pred][label]
    hist = np.array([[50, 2, 3], [4, 20, 1], [5, 2, 13]], dtype=np.float64)
    fscd, miou, sek = RT.scd_scores_from_hist(hist)
    c2 = np.array([[50, 5], [9, 36]], dtype=np.float64)
    iu = np.diag(c2) / (c2.sum(1) + c2.sum(0) - np.diag(c2))
    assert abs(miou - iu.mean()) < 1e-12
    prec, rec = 33 / (100 - 55), 33 / (100 - 59)
    assert abs(fscd - 2 / (1 / prec + 1 / rec)) < 1e-12
    h0 = hist.copy(); h0[0, 0] = 0
    po = np.diag(h0).sum() / h0.sum(); pe = (h0.sum(1) * h0.sum(0)).sum() / h0.sum() ** 2
    assert abs(sek - (po - pe) / (1 - pe) * np.exp(iu[1]) / np.e) < 1e-12


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted")
def test_metrics_match_reference_functions():
    R.load()
    import model.utils as ref_utils            # the reference's file, unchanged (path set up by reference_loader)
    # SCD: SCDD_eval_all over per-image maps == formulas on the accumulated hist[pred][label]
    preds, labels = _rand_maps(7, 3)
    want = ref_utils.SCDD_eval_all(preds, labels, 7)
    hist = np.zeros((7, 7))
    for p, l in zip(preds, labels):
        hist += np.bincount(7 * p.reshape(-1) + l.reshape(-1), minlength=49).reshape(7, 7)
    got = RT.scd_scores_from_hist(hist)
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    # BDA: the two Evaluators of train_BDA.py:120-147
    g = np.random.default_rng(5)
    ev_loc, ev_cls = ref_utils.Evaluator(2), ref_utils.Evaluator(5)
    cm_loc, cm_cls = np.zeros((2, 2)), np.zeros((5, 5))
    for _ in range(4):
        label_loc = (g.random((3, 9, 11)) < 0.4).astype(np.float32)
        label_cls = (label_loc * g.integers(1, 5, label_loc.shape)).astype(np.int64)
        pred_loc = g.random(label_loc.shape) > 0.5
        pred_cls = g.integers(0, 5, label_loc.shape)
        ev_loc.add_batch(label_loc, pred_loc)
        ev_cls.add_batch(label_cls[label_loc > 0], pred_cls[label_loc > 0])
        cm_loc += np.bincount(2 * label_loc.astype(int).reshape(-1) + pred_loc.astype(int).reshape(-1), minlength=4).reshape(2, 2)
        m = label_loc > 0
        cm_cls += np.bincount(5 * label_cls[m] + pred_cls[m], minlength=25).reshape(5, 5)
    loc_f1, harm, oaf1, dmg = RT.bda_scores_from_hists(cm_loc, cm_cls)
    d_ref = ev_cls.Damage_F1_socore()
    assert abs(loc_f1 - ev_loc.Pixel_F1_score()) < 1e-12 and np.allclose(dmg, d_ref, rtol=1e-12)
    assert abs(harm - len(d_ref) / np.sum(1.0 / d_ref)) < 1e-12 and abs(oaf1 - (0.3 * loc_f1 + 0.7 * harm)) < 1e-12


def test_runner_refuses_to_run_without_cuda(tmp_path):
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    args = RT.build_parser("scd").parse_args(["--synthetic", "4", "--save_dir", str(tmp_path)])
    with pytest.raises(RuntimeError, match="no CPU"):
        RT.train_validate(args, "scd")
    with pytest.raises(ValueError, match="unknown task"):
        RT.train_validate(args, "cc")
