"""Host-side helpers with the reference's names and semantics (model/utils.py): weight_init,
adjust_learning_rate, and the losses BCEDiceLoss / CrossEntropyLoss2d / ChangeSimilarity, which run as fused
sm_100a kernels (change3d_b200/losses.py, csrc/loss.cu) — CUDA tensors only."""
import torch
import torch.nn as nn

from ..losses import ChangeSimilarity, CrossEntropyLoss2d, bce_dice_loss  # noqa: F401  (reference names)


def weight_init(module):
    """Same traversal and initialisers as model/utils.py:20-82: kaiming-normal for nn.Conv2d / nn.Linear
    (bias zero), ones/zeros for BatchNorm2d / GroupNorm, recursing through containers.
    (nn.ConvTranspose2d is not an nn.Conv2d, so it keeps its default init — as in the reference.)"""
    skip = (nn.AdaptiveAvgPool2d, nn.AdaptiveMaxPool2d, nn.ModuleList, nn.BCELoss)

    def init_leaf(m) -> bool:
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='relu')
            if m.bias is not None:
                nn.init.zeros_(m.bias)
            return True
        if isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
            nn.init.ones_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
            return True
        return False

    for _, child in module.named_children():
        if isinstance(child, skip) or init_leaf(child):
            continue
        if isinstance(child, nn.Sequential):
            for _, sub in child.named_children():
                if not init_leaf(sub):
                    weight_init(sub)
        elif len(list(child.children())) > 0:
            weight_init(child)


def adjust_learning_rate(args, optimizer, epoch=None, iter=None, max_batches=None, lr_factor=1.0,
                         shrink_factor=None, verbose=True):
    """Poly / step schedule with the 200-iteration warm-up of model/utils.py:84-150."""
    if shrink_factor is not None:
        if not 0 < shrink_factor < 1:
            raise ValueError(f"Shrink factor must be between 0 and 1, got {shrink_factor}")
        for g in optimizer.param_groups:
            g['lr'] = g['lr'] * shrink_factor
        return optimizer.param_groups[0]['lr']
    if args.lr_mode == 'step':
        if epoch is None:
            raise ValueError("Epoch must be provided for step lr_mode")
        lr = args.lr * (0.1 ** (epoch // args.step_loss))
    elif args.lr_mode == 'poly':
        if any(v is None for v in (epoch, iter, max_batches)):
            raise ValueError("Epoch, iter, and max_batches must be provided for poly lr_mode")
        lr = args.lr * (1 - iter * 1.0 / (max_batches * args.max_epochs)) ** 0.9
    else:
        raise ValueError(f'Unknown lr mode {args.lr_mode}')
    if epoch == 0 and iter is not None and iter < 200:
        lr = args.lr * 0.9 * (iter + 1) / 200 + 0.1 * args.lr
    lr *= lr_factor
    for g in optimizer.param_groups:
        g['lr'] = lr
    return lr


def BCEDiceLoss(inputs, targets):
    """model/utils.py:154-169: F.binary_cross_entropy(inputs, targets) + 1 - dice, one fused kernel each way."""
    return bce_dice_loss(inputs, targets)
