"""Generates tests/golden/losses.npz by running the REFERENCE's own loss / metric functions (model/utils.py
BCEDiceLoss, CrossEntropyLoss2d, ChangeSimilarity; utils/metric_tool.py get_confuse_matrix, cm2score) on seeded
inputs, with their autograd gradients.  TEST INFRASTRUCTURE — authoring container only (needs /root/reference):

    python -m oracle.make_golden_losses
"""
import importlib
import os

import numpy as np
import torch

from oracle import reference_loader as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main() -> None:
    R.load()
    mu = importlib.import_module("model.utils")            # the reference's file, unchanged
    mt = importlib.import_module("utils.metric_tool")
    g = torch.Generator().manual_seed(16)
    out = {}

    # BCD: sigmoid prediction (B,1,H,W) against a sparse binary mask (scripts/train_BCD.py:200-225)
    B, H, W = 3, 24, 20
    logit = torch.randn(B, 1, H, W, generator=g) * 3
    pred = torch.sigmoid(logit)
    pred[0, 0, 0, 0], pred[0, 0, 0, 1], pred[0, 0, 0, 2] = 0.0, 1.0, 0.5     # clamp / threshold edge cases
    pred.requires_grad_(True)
    target = (torch.rand(B, 1, H, W, generator=g) < 0.2).float()
    loss = mu.BCEDiceLoss(pred, target)
    loss.backward()
    mask = torch.where(pred > 0.5, torch.ones_like(pred), torch.zeros_like(pred)).long()
    cm = mt.get_confuse_matrix(num_classes=2, label_gts=target.numpy(), label_preds=mask.numpy())
    out.update(bce_pred=pred.detach().numpy(), bce_target=target.numpy(), bce_loss=np.float64(loss.item()),
               bce_grad=pred.grad.numpy(), bce_cm=cm.astype(np.int64))
    sc = mt.cm2score(cm)
    out["bce_scores"] = np.array([sc[k] for k in ('Kappa', 'IoU', 'F1', 'OA', 'recall', 'precision', 'Pre')])
    out["bce_f1"] = np.float64(mt.cm2F1(cm))

    # SCD: 7-class heads, labels masked by the change map (scripts/train_SCD.py:212-229)
    C = 7
    pre_mask = (torch.randn(B, C, H, W, generator=g) * 2).requires_grad_(True)
    post_mask = (torch.randn(B, C, H, W, generator=g) * 2).requires_grad_(True)
    label_change = (torch.rand(B, H, W, generator=g) < 0.3).long()
    pre_label = torch.randint(1, C, (B, H, W), generator=g) * label_change
    post_label = torch.randint(1, C, (B, H, W), generator=g) * label_change
    seg = mu.CrossEntropyLoss2d(ignore_index=0)
    sim = mu.ChangeSimilarity()
    l_seg = seg(pre_mask, pre_label)
    l_sim = sim(pre_mask[:, 1:], post_mask[:, 1:], label_change.unsqueeze(1))
    (l_seg * 0.5 + l_sim).backward()
    out.update(scd_pre=pre_mask.detach().numpy(), scd_post=post_mask.detach().numpy(),
               scd_label_change=label_change.numpy(), scd_pre_label=pre_label.numpy(),
               scd_seg_loss=np.float64(l_seg.item()), scd_sim_loss=np.float64(l_sim.item()),
               scd_pre_grad=pre_mask.grad.numpy(), scd_post_grad=post_mask.grad.numpy(),
               scd_argmax=torch.argmax(pre_mask.detach(), dim=1).numpy())
    out["scd_cm"] = mt.get_confuse_matrix(num_classes=C, label_gts=pre_label.numpy(),
                                          label_preds=out["scd_argmax"]).astype(np.int64)

    # BDA: 5 classes, ignore_index 0, default ignore (-1) as a second case
    C = 5
    x = (torch.randn(2, C, 9, 7, generator=g)).requires_grad_(True)
    t = torch.randint(0, C, (2, 9, 7), generator=g)
    l0 = mu.CrossEntropyLoss2d(ignore_index=0)(x, t)
    l0.backward()
    out.update(bda_x=x.detach().numpy(), bda_t=t.numpy(), bda_loss_ign0=np.float64(l0.item()),
               bda_grad_ign0=x.grad.numpy().copy())
    x.grad = None
    l1 = mu.CrossEntropyLoss2d()(x, t)
    l1.backward()
    out.update(bda_loss_ignm1=np.float64(l1.item()), bda_grad_ignm1=x.grad.numpy().copy())

    path = os.path.join(ROOT, "tests", "golden", "losses.npz")
    np.savez_compressed(path, **out)
    print(path, f"{os.path.getsize(path) / 1024:.0f} KiB", sorted(out))


if __name__ == "__main__":
    main()
