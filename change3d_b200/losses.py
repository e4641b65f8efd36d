"""On-device losses of the train step over the C ABI (csrc/loss.cu).

Same arithmetic as the reference's `model/utils.py` (BCEDiceLoss :154-169, CrossEntropyLoss2d :171-178,
ChangeSimilarity :180-203), but each forward is ONE streaming kernel (loss + backward coefficients written on
the device by the last CTA, no host synchronisation) and each backward is one kernel.  The BCE+Dice forward can
also accumulate the 2x2 confusion matrix of `(target, pred > 0.5)` in the same pass — what the reference does on
the CPU with `pred.cpu().numpy()` + `np.bincount` every step (scripts/train_BCD.py:203-225).

No CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from . import ops

_WS = {}


def _ws(device: torch.device) -> torch.Tensor:
    """Per-device reduction workspace: zeroed once here, left zeroed by every kernel (self-cleaning)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _WS:
        _WS[key] = torch.zeros(L.LOSS_WS_BYTES // 8, dtype=torch.float64, device=device)
    return _WS[key]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("change3d_b200 losses run on CUDA tensors only (there is no CPU fallback)")


def _gout(g: torch.Tensor) -> torch.Tensor:
    g = g.detach()
    if g.dtype != torch.float32:
        g = g.float()
    return g.reshape(1) if g.numel() == 1 else g


def _chw_view(x: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) tensor whose (C,H,W) part is dense (batch stride free), else a contiguous copy."""
    B, C, H, W = x.shape
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(3) == 1 and x.stride(2) == W and x.stride(1) == H * W:
        return x
    return x.contiguous()


# ----------------------------------------------------------------------------------------------- BCE + Dice
class _BCEDiceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs: torch.Tensor, targets: torch.Tensor, cm: Optional[torch.Tensor], holder):
        _need_cuda(inputs, targets, cm)
        p = inputs.detach().float().contiguous()
        t = targets.detach().float().contiguous()
        if p.numel() != t.numel():
            raise ValueError(f"BCEDiceLoss: input {tuple(inputs.shape)} and target {tuple(targets.shape)} differ in size")
        if cm is not None and (cm.dtype != torch.int64 or cm.numel() != 4 or not cm.is_contiguous()):
            raise ValueError("BCEDiceLoss: cm must be a contiguous int64 tensor of 4 elements")
        out = torch.empty(L.LOSS_OUT_FLOATS, dtype=torch.float32, device=p.device)
        with ops._Timed("loss_fwd", 8 * p.numel()):
            L.check(L.load().c3d_bce_dice_fwd(p.data_ptr(), t.data_ptr(), p.numel(), _ws(p.device).data_ptr(),
                                              out.data_ptr(), None if cm is None else cm.data_ptr(), _stream()),
                    "c3d_bce_dice_fwd")
        ctx.save_for_backward(p, t, out)
        ctx.in_shape = inputs.shape
        if holder is not None:
            holder.append(out)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        p, t, out = ctx.saved_tensors
        g = _gout(g)
        dp = torch.empty_like(p)
        with ops._Timed("loss_bwd", 12 * p.numel()):
            L.check(L.load().c3d_bce_dice_bwd(p.data_ptr(), t.data_ptr(), out.data_ptr(), g.data_ptr(), 1.0,
                                              dp.data_ptr(), p.numel(), _stream()), "c3d_bce_dice_bwd")
        return dp.view(ctx.in_shape), None, None, None


def bce_dice_loss(inputs: torch.Tensor, targets: torch.Tensor, cm: Optional[torch.Tensor] = None,
                  return_parts: bool = False):
    """BCEDiceLoss(inputs, targets) (model/utils.py:154-169).  `cm` (int64[2,2] on the device, optional) receives
    += the confusion matrix hist[gt][pred > 0.5] of this batch.  With return_parts also returns the device
    vector (loss, 1/n, 2/S, dice/S, bce, dice, -, -)."""
    holder = [] if return_parts else None
    loss = _BCEDiceFn.apply(inputs, targets, cm, holder)
    return (loss, holder[0]) if return_parts else loss


# -------------------------------------------------------------------------------------------- cross entropy
class _CE2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, targets, ignore_index: int, argmax_out, cm):
        _need_cuda(inputs, targets, argmax_out, cm)
        if inputs.dim() != 4:
            raise ValueError("CrossEntropyLoss2d: inputs must be (B, C, H, W)")
        B, C, H, W = inputs.shape
        if C > 16:
            raise ValueError("CrossEntropyLoss2d: at most 16 classes")
        x = _chw_view(inputs.detach())
        t = targets.detach()
        if t.dtype != torch.int64:
            raise ValueError("CrossEntropyLoss2d: targets must be int64 (like nn.NLLLoss)")
        t = t.contiguous()
        if t.numel() != B * H * W:
            raise ValueError(f"CrossEntropyLoss2d: target {tuple(targets.shape)} does not match {tuple(inputs.shape)}")
        for name, a, n in (("argmax_out", argmax_out, B * H * W), ("cm", cm, C * C)):
            if a is not None and (a.dtype != torch.int64 or a.numel() != n or not a.is_contiguous()):
                raise ValueError(f"CrossEntropyLoss2d: {name} must be contiguous int64 with {n} elements")
        out = torch.empty(L.LOSS_OUT_FLOATS, dtype=torch.float32, device=x.device)
        L.check(L.load().c3d_ce2d_fwd(x.data_ptr(), t.data_ptr(), B, C, H * W, x.stride(0), ignore_index,
                                      _ws(x.device).data_ptr(), out.data_ptr(),
                                      None if argmax_out is None else argmax_out.data_ptr(),
                                      None if cm is None else cm.data_ptr(), _stream()), "c3d_ce2d_fwd")
        ctx.save_for_backward(x, t, out)
        ctx.ignore_index = ignore_index
        return out[0]

    @staticmethod
    def backward(ctx, g):
        x, t, out = ctx.saved_tensors
        B, C, H, W = x.shape
        g = _gout(g)
        dx = torch.empty(B, C, H, W, dtype=torch.float32, device=x.device)
        L.check(L.load().c3d_ce2d_bwd(x.data_ptr(), t.data_ptr(), B, C, H * W, x.stride(0), ctx.ignore_index,
                                      out.data_ptr(), g.data_ptr(), 1.0, dx.data_ptr(), _stream()), "c3d_ce2d_bwd")
        return dx, None, None, None, None


def cross_entropy_2d(inputs: torch.Tensor, targets: torch.Tensor, ignore_index: int = -1,
                     argmax_out: Optional[torch.Tensor] = None, cm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NLLLoss(log_softmax(inputs, 1), targets, ignore_index, 'mean') (model/utils.py:171-178).  Optionally the same
    pass writes torch.argmax(inputs, 1) into `argmax_out` (int64 (B,H,W)) and adds hist[target][argmax] to `cm`."""
    return _CE2dFn.apply(inputs, targets, int(ignore_index), argmax_out, cm)


class CrossEntropyLoss2d(torch.nn.Module):
    """model/utils.py:171-178.  Class weights are not used by the reference scripts and are not implemented."""

    def __init__(self, weight=None, ignore_index: int = -1):
        super().__init__()
        if weight is not None:
            raise NotImplementedError("CrossEntropyLoss2d(weight=...) is not implemented (unused by the reference)")
        self.ignore_index = ignore_index

    def forward(self, inputs, targets):
        return cross_entropy_2d(inputs, targets, self.ignore_index)


# ------------------------------------------------------------------------------------------ change similarity
class _SimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, label_change):
        _need_cuda(x1, x2, label_change)
        if x1.shape != x2.shape or x1.dim() != 4:
            raise ValueError("ChangeSimilarity: x1 and x2 must be (B, C, H, W) of the same shape")
        B, C, H, W = x1.shape
        if C > 16:
            raise ValueError("ChangeSimilarity: at most 16 classes")
        a, b = _chw_view(x1.detach()), _chw_view(x2.detach())
        lc = label_change.detach()
        if lc.dtype != torch.int64:
            lc = lc.long()
        lc = lc.contiguous()
        if lc.numel() != B * H * W:
            raise ValueError("ChangeSimilarity: label_change must have B*H*W elements")
        out = torch.empty(L.LOSS_OUT_FLOATS, dtype=torch.float32, device=a.device)
        L.check(L.load().c3d_change_similarity_fwd(a.data_ptr(), b.data_ptr(), lc.data_ptr(), B, C, H * W,
                                                   a.stride(0), b.stride(0), _ws(a.device).data_ptr(),
                                                   out.data_ptr(), _stream()), "c3d_change_similarity_fwd")
        ctx.save_for_backward(a, b, lc)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        a, b, lc = ctx.saved_tensors
        B, C, H, W = a.shape
        g = _gout(g)
        d1 = torch.empty(B, C, H, W, dtype=torch.float32, device=a.device)
        d2 = torch.empty(B, C, H, W, dtype=torch.float32, device=a.device)
        L.check(L.load().c3d_change_similarity_bwd(a.data_ptr(), b.data_ptr(), lc.data_ptr(), B, C, H * W,
                                                   a.stride(0), b.stride(0), g.data_ptr(), 1.0, d1.data_ptr(),
                                                   d2.data_ptr(), _stream()), "c3d_change_similarity_bwd")
        return d1, d2, None


class ChangeSimilarity(torch.nn.Module):
    """model/utils.py:180-203 (reduction 'mean' only, the reference's default and only use)."""

    def __init__(self, reduction: str = 'mean'):
        super().__init__()
        if reduction != 'mean':
            raise NotImplementedError("ChangeSimilarity: only reduction='mean' is implemented")

    def forward(self, x1, x2, label_change):
        return _SimFn.apply(x1, x2, label_change)


# ------------------------------------------------------------------------------------------- confusion matrix
def confusion_matrix(gt: torch.Tensor, pred: torch.Tensor, num_classes: int,
                     cm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cm[gt][pred] += 1 over entries with 0 <= gt < num_classes (utils/metric_tool.py:111-128); gt float32 or
    int64, pred int64.  Returns the (num_classes, num_classes) int64 device tensor (allocated zeroed if None)."""
    _need_cuda(gt, pred, cm)
    if gt.dtype not in (torch.float32, torch.int64):
        gt = gt.float() if gt.is_floating_point() else gt.long()
    if pred.dtype != torch.int64:
        pred = pred.long()
    gt, pred = gt.contiguous(), pred.contiguous()
    if gt.numel() != pred.numel():
        raise ValueError("confusion_matrix: gt and pred differ in size")
    if cm is None:
        cm = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=gt.device)
    elif cm.dtype != torch.int64 or cm.numel() != num_classes * num_classes or not cm.is_contiguous():
        raise ValueError("confusion_matrix: cm must be contiguous int64 (num_classes, num_classes)")
    L.check(L.load().c3d_confusion_matrix(gt.data_ptr(), 1 if gt.dtype == torch.float32 else 0, pred.data_ptr(),
                                          gt.numel(), num_classes, cm.data_ptr(), _stream()), "c3d_confusion_matrix")
    return cm
