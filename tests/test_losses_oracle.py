"""CPU: pins the oracle's loss / metric restatements (oracle/change3d_oracle.py) against tests/golden/losses.npz,
which holds outputs and autograd gradients of the REFERENCE's own model/utils.py and utils/metric_tool.py
(oracle/make_golden_losses.py); checks the C ABI's argument validation and that the product losses refuse CPU."""
import os

import numpy as np
import pytest
import torch

from oracle import change3d_oracle as O


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "losses.npz"))


def test_bce_dice_oracle_matches_reference(g):
    p = torch.tensor(g["bce_pred"], requires_grad=True)
    t = torch.tensor(g["bce_target"])
    loss = O.bce_dice_loss(p, t)
    loss.backward()
    assert abs(loss.item() - float(g["bce_loss"])) < 1e-6 * abs(float(g["bce_loss"]))
    np.testing.assert_allclose(p.grad.numpy(), g["bce_grad"], rtol=1e-5, atol=1e-9)


def test_confusion_matrix_oracle_matches_reference(g):
    mask = (g["bce_pred"] > 0.5).astype(np.int64)
    cm = O.confusion_matrix(2, g["bce_target"], mask)
    assert np.array_equal(cm, g["bce_cm"]) and cm.sum() == g["bce_pred"].size
    sc = O.cm_scores(cm)
    got = np.array([sc[k] for k in ('Kappa', 'IoU', 'F1', 'OA', 'recall', 'precision', 'Pre')])
    np.testing.assert_allclose(got, g["bce_scores"], rtol=1e-12)
    assert np.array_equal(O.confusion_matrix(7, g["scd_pre_label"], g["scd_argmax"]), g["scd_cm"])


def test_ce_and_similarity_oracle_match_reference(g):
    pre = torch.tensor(g["scd_pre"], requires_grad=True)
    post = torch.tensor(g["scd_post"], requires_grad=True)
    lc = torch.tensor(g["scd_label_change"])
    seg = O.cross_entropy_2d(pre, torch.tensor(g["scd_pre_label"]), ignore_index=0)
    sim = O.change_similarity(pre[:, 1:], post[:, 1:], lc.unsqueeze(1))
    (seg * 0.5 + sim).backward()
    assert abs(seg.item() - float(g["scd_seg_loss"])) < 1e-5
    assert abs(sim.item() - float(g["scd_sim_loss"])) < 1e-5
    np.testing.assert_allclose(pre.grad.numpy(), g["scd_pre_grad"], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(post.grad.numpy(), g["scd_post_grad"], rtol=1e-4, atol=1e-8)
    x = torch.tensor(g["bda_x"], requires_grad=True)
    t = torch.tensor(g["bda_t"])
    for ign, key in ((0, "ign0"), (-1, "ignm1")):
        x.grad = None
        loss = O.cross_entropy_2d(x, t, ignore_index=ign)
        loss.backward()
        assert abs(loss.item() - float(g["bda_loss_" + key])) < 1e-5
        np.testing.assert_allclose(x.grad.numpy(), g["bda_grad_" + key], rtol=1e-4, atol=1e-8)


def test_loss_abi_rejects_bad_arguments_without_device():
    from change3d_b200 import _lib
    lib = _lib.load()
    assert lib.c3d_bce_dice_fwd(None, None, 0, None, None, None, None) == 1
    assert lib.c3d_bce_dice_bwd(None, None, None, None, 1.0, None, 0, None) == 1
    assert lib.c3d_ce2d_fwd(1, 1, 1, 17, 4, 0, 0, 1, 1, None, None, None) == 1            # > 16 classes
    assert lib.c3d_ce2d_bwd(None, None, 1, 5, 4, 0, 0, None, None, 1.0, None, None) == 1
    assert lib.c3d_change_similarity_fwd(None, None, None, 1, 6, 4, 0, 0, None, None, None) == 1
    assert lib.c3d_change_similarity_bwd(1, 1, 1, 0, 6, 4, 0, 0, None, 1.0, 1, 1, None) == 1   # B = 0
    assert lib.c3d_confusion_matrix(None, 0, None, 0, 2, None, None) == 1


def test_product_losses_refuse_cpu_tensors():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from change3d_b200 import losses, metrics
    from change3d_b200.model.utils import BCEDiceLoss, ChangeSimilarity, CrossEntropyLoss2d
    with pytest.raises(RuntimeError, match="no CPU"):
        BCEDiceLoss(torch.rand(1, 1, 4, 4), torch.zeros(1, 1, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU"):
        CrossEntropyLoss2d(ignore_index=0)(torch.rand(1, 5, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU"):
        ChangeSimilarity()(torch.rand(1, 6, 4, 4), torch.rand(1, 6, 4, 4), torch.zeros(1, 1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU"):
        losses.confusion_matrix(torch.zeros(4), torch.zeros(4, dtype=torch.long), 2)
    with pytest.raises(NotImplementedError):
        CrossEntropyLoss2d(weight=torch.ones(5))
    sc = metrics.cm2score(np.array([[90, 2], [3, 5]]))
    ref = O.cm_scores(np.array([[90, 2], [3, 5]]))
    assert all(abs(sc[k] - ref[k]) < 1e-15 for k in ref)


def test_task_loss_matches_reference_loss_functions():
    """O.task_loss (the per-script loss formulas) against the reference's own loss functions on random heads
    (authoring container only: needs /root/reference)."""
    import importlib
    from oracle import reference_loader as R
    if not R.available():
        pytest.skip("reference tree not present")
    R.load()
    mu = importlib.import_module("model.utils")
    gen = torch.Generator().manual_seed(9)
    B, H, W = 2, 12, 10
    # scripts/train_SCD.py:215-229
    pre_m, post_m = torch.randn(B, 7, H, W, generator=gen), torch.randn(B, 7, H, W, generator=gen)
    chg = torch.sigmoid(torch.randn(B, 1, H, W, generator=gen))
    labels = O.synth_labels("scd", B, H, W, 7, 9)
    pl, ql, lc = labels[0] * labels[2], labels[1] * labels[2], labels[2]
    seg = mu.CrossEntropyLoss2d(ignore_index=0)
    want = (seg(pre_m, pl) + seg(post_m, ql)) * 0.5 + mu.BCEDiceLoss(chg, lc.unsqueeze(1).float()) \
        + mu.ChangeSimilarity()(pre_m[:, 1:], post_m[:, 1:], lc.unsqueeze(1))
    got = O.task_loss("scd", (pre_m, post_m, chg), labels)
    assert abs(got.item() - want.item()) < 1e-5
    # scripts/train_BDA.py:181-199
    cls, loc = torch.randn(B, 5, H, W, generator=gen), torch.sigmoid(torch.randn(B, 1, H, W, generator=gen))
    l_loc, l_cls = O.synth_labels("bda", B, H, W, 5, 9)
    want = seg(cls, l_cls) + mu.BCEDiceLoss(loc, l_loc.unsqueeze(1))
    assert abs(O.task_loss("bda", (cls, loc), (l_loc, l_cls)).item() - want.item()) < 1e-5
