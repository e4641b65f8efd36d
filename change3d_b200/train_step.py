"""One training iteration of the reference loop (scripts/train_BCD.py:179-216) on the B200 engine.

    output = model.update_bcd(pre, post); loss = BCEDiceLoss(output, target)
    optimizer.zero_grad(); loss.backward(); optimizer.step()

with the reference's optimizer (torch.optim.Adam(lr, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-4),
scripts/train_BCD.py:284-290) restated as ONE fused kernel over a flat parameter buffer:

  * every parameter that the task trains becomes a view of one flat fp32 buffer; the wgrad kernels
    accumulate straight into the matching views of a flat gradient buffer (one memset per step);
  * parameters that never receive a gradient (blocks.4 / blocks.5 in BCD/SCD/BDA — torch's Adam skips
    `grad is None`, scripts/train_BCD.py:284) are left out of the flat buffers, hence untouched;
  * data parallel: one process per GPU, the flat gradient buffer is all-reduced once (NCCL, sum) and the
    mean is folded into the Adam kernel (grad_scale = 1 / world_size); BatchNorm statistics stay per rank
    (the reference uses plain nn.BatchNorm3d);
  * the whole iteration can be captured in a CUDA graph (static shapes) to remove launch overhead.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import ops
from .losses import ChangeSimilarity, bce_dice_loss, cross_entropy_2d


def _trained_parameters(model) -> List[torch.nn.Parameter]:
    """Parameters that receive gradients on the change-decoder tasks: everything except x3d.blocks[4], blocks[5]
    (never executed, model/trainer.py:127-139)."""
    out = []
    for name, p in model.named_parameters():
        if ".x3d.blocks.4." in name or ".x3d.blocks.5." in name:
            continue
        if p.requires_grad:
            out.append(p)
    return out


class FlatAdam:
    """Flat-buffer Adam with the reference's hyper-parameters; gradients are written in place by the kernels."""

    def __init__(self, model, lr: float = 2e-4, betas=(0.9, 0.99), eps: float = 1e-8, weight_decay: float = 1e-4):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.params = _trained_parameters(model)
        dev = self.params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]
        total = sum(sizes)
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.v = torch.zeros(total, device=dev, dtype=torch.float32)
        self.step_count = 0
        # `adjust_learning_rate` (model/utils.py:84-150) writes param_groups[i]['lr']; step() reads it back
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]
        off = 0
        gview = {}
        for p, n in zip(self.params, sizes):
            pv = self.flat_p[off:off + p.numel()].view(p.shape)
            pv.copy_(p.data)
            p.data = pv                                   # the module's parameter now lives in the flat buffer
            gview[id(p)] = self.flat_g[off:off + p.numel()].view(p.shape)
            off += n
        self.numel = total
        self._install(gview)

    def _install(self, gview) -> None:
        """Tell the engine where each module's parameter gradients live (see engine.GradArena)."""
        m = self.model
        enc = m.encoder
        stem = enc.x3d.blocks[0]
        object.__setattr__(stem, "_c3d_grad_views", [gview[id(p)] for p in (stem.conv.conv_t.weight,
                                                                           stem.conv.conv_xy.weight,
                                                                           stem.norm.weight, stem.norm.bias)])
        object.__setattr__(stem, "_c3d_perc_grad_view", gview[id(enc.perception_frames)])
        for i in range(1, 4):
            stage = enc.x3d.blocks[i]
            object.__setattr__(stage, "_c3d_grad_views", [gview[id(p)] for p in stage.param_list()])
        for fc in enc.fc:
            fc[0].weight._c3d_grad_view = gview[id(fc[0].weight)]
        for name in ("decoder", "decoder_pre", "decoder_post", "decoder_change", "decoder_cls", "decoder_loc"):
            dec = getattr(m, name, None)
            if dec is not None and hasattr(dec, "param_list"):
                object.__setattr__(dec, "_c3d_grad_views", [gview[id(p)] for p in dec.param_list()])

    def zero_grad(self) -> None:
        self.flat_g.zero_()

    def all_reduce(self) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)

    def step(self, lr: Optional[float] = None) -> None:
        self.step_count += 1
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, self.param_groups[0]["lr"] if lr is None else lr, self.betas[0],
                      self.betas[1], self.eps, self.weight_decay, self.step_count, 1.0 / world)


    def state_dict(self) -> dict:
        """Optimizer state for the reference's checkpoint dict (scripts/train_BCD.py:333-343 saves it, nothing in the
        reference restores it): flat moment buffers + step + hyper-parameters."""
        return {"state": {"step": self.step_count, "exp_avg": self.m.clone(), "exp_avg_sq": self.v.clone()},
                "param_groups": [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd: dict) -> None:
        self.step_count = int(sd["state"]["step"])
        self.m.copy_(sd["state"]["exp_avg"])
        self.v.copy_(sd["state"]["exp_avg_sq"])
        self.param_groups[0].update(sd["param_groups"][0])


class TrainStep:
    """model.update_<task> -> the script's loss -> backward -> (all-reduce) -> fused Adam, optionally as a CUDA graph.

    task 'bcd' (scripts/train_BCD.py:179-216): labels = (target float (B,1,H,W),); BCEDiceLoss.
    task 'scd' (scripts/train_SCD.py:205-233): labels = (pre_label, post_label, label_change) int64 (B,H,W); the class
         labels are masked by label_change, loss = 0.5*(CE0(pre)+CE0(post)) + BCEDice(change) + ChangeSimilarity.
    task 'bda' (scripts/train_BDA.py:174-204): labels = (label_loc float (B,H,W), label_cls int64 (B,H,W));
         loss = CE0(cls) + BCEDice(loc).           CE0 = CrossEntropyLoss2d(ignore_index=0).

    The BCEDice kernel also accumulates the confusion matrix hist[target][output > 0.5] of the binary head into
    `self.cm` (int64 (2,2), device) — the per-step `eval_meter.update_cm(pred.cpu().numpy(), target.cpu().numpy())`
    of scripts/train_BCD.py:203-225 without leaving the GPU; read it with `scores()` at the end of an epoch.
    `self.parts` holds the device scalars of the last iteration's loss terms (seg / binary / sim), as the scripts log."""

    def __init__(self, model, lr: float = 2e-4, use_graph: bool = False, task: str = "bcd"):
        if task not in ("bcd", "scd", "bda"):
            raise ValueError(f"TrainStep: unknown task {task!r}")
        self.task = task
        self.model = model.train()
        self.opt = FlatAdam(model, lr=lr)
        self.cm = torch.zeros(2, 2, dtype=torch.int64, device=self.opt.flat_p.device)
        self.sim = ChangeSimilarity()
        self.parts = {}
        self.use_graph = use_graph
        self.graph = None
        self.static = None
        self.loss = None

    def _loss(self, pre, post, labels) -> torch.Tensor:
        if self.task == "bcd":
            return bce_dice_loss(self.model.update_bcd(pre, post), labels[0], cm=self.cm)
        if self.task == "scd":
            pre_label, post_label, label_change = labels
            pre_label, post_label = pre_label * label_change, post_label * label_change
            pre_mask, post_mask, change_mask = self.model.update_scd(pre, post)
            seg = cross_entropy_2d(pre_mask, pre_label, 0) + cross_entropy_2d(post_mask, post_label, 0)
            binary = bce_dice_loss(change_mask, label_change.unsqueeze(1).float(), cm=self.cm)
            sim = self.sim(pre_mask[:, 1:], post_mask[:, 1:], label_change.unsqueeze(1))
            self.parts = {"seg": seg.detach(), "binary": binary.detach(), "sim": sim.detach()}
            return seg * 0.5 + binary + sim
        label_loc, label_cls = labels
        pred_cls, pred_loc = self.model.update_bda(pre, post)
        seg = cross_entropy_2d(pred_cls, label_cls, 0)
        binary = bce_dice_loss(pred_loc, label_loc.unsqueeze(1), cm=self.cm)
        self.parts = {"seg": seg.detach(), "binary": binary.detach()}
        return seg + binary

    def _iteration(self, pre, post, *labels) -> torch.Tensor:
        self.opt.zero_grad()
        loss = self._loss(pre, post, labels)
        loss.backward()
        return loss.detach()

    def eager(self, pre: torch.Tensor, post: torch.Tensor, *labels: torch.Tensor, lr: Optional[float] = None):
        """One iteration without the graph (any batch size — e.g. the ragged last batch of an epoch)."""
        loss = self._iteration(pre, post, *labels)
        self.opt.all_reduce()
        self.opt.step(lr)
        return loss

    def __call__(self, pre: torch.Tensor, post: torch.Tensor, *labels: torch.Tensor, lr: Optional[float] = None):
        """pre/post (B,3,H,W) and the task's labels on the GPU.  Returns the (device) loss of this iteration."""
        if not self.use_graph:
            return self.eager(pre, post, *labels, lr=lr)
        if self.graph is None:
            # warm-up on a side stream (also sets every kernel's shared-memory attribute), then capture
            self.static = tuple(t.clone() for t in (pre, post) + labels)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._iteration(*self.static)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            from . import _lib
            n0 = _lib.LAUNCHES[0]
            self.graph = torch.cuda.CUDAGraph()
            # thread_local: a DataLoader pin-memory thread may call cudaHostAlloc while this thread captures
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.loss = self._iteration(*self.static)
            self.captured_launches = _lib.LAUNCHES[0] - n0 + 1      # + the Adam launch outside the graph
            self.cm.zero_()                                         # drop the warm-up iteration's counts
        for dst, src in zip(self.static, (pre, post) + labels):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.opt.all_reduce()
        self.opt.step(lr)                 # Adam outside the graph: step count / lr change every iteration
        return self.loss

    def scores(self) -> dict:
        """cm2score of the accumulated training confusion matrix (one small device->host copy)."""
        from .metrics import cm2score
        return cm2score(self.cm.cpu().numpy())


class BCDTrainStep(TrainStep):
    """TrainStep(task='bcd') — the BASELINE.json headline configuration."""

    def __init__(self, model, lr: float = 2e-4, use_graph: bool = False):
        super().__init__(model, lr=lr, use_graph=use_graph, task="bcd")
