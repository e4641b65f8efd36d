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

`CCTrainStep` is the change-captioning iteration (scripts/train_CC.py:105-146): encoder feature path + CaptionDecoder
+ packed cross-entropy, gradient clamp and the script's TWO Adam optimizers (scripts/train_CC.py:440-463), each a
flat buffer with the clamp fused into its Adam launch.
"""
from __future__ import annotations

import os

from typing import List, Optional

import torch
import torch.distributed as dist

from . import ops
from .losses import ChangeSimilarity, bce_dice_loss, cross_entropy_2d

BN_BUFFERS = ("running_mean", "running_var", "num_batches_tracked")


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
    """Flat-buffer Adam with the reference's hyper-parameters; gradients are written in place by the kernels.
    `params`: the parameters this optimizer owns (default: the change-decoder tasks' trained set); `grad_clip` > 0
    clamps every gradient element first (clip_gradient, model/utils.py:481-491), fused into the Adam launch.
    Parameters whose gradient torch's autograd produces (the captioning head) get `.grad` pointed at their slice of
    the flat gradient buffer, so autograd accumulates in place and no per-parameter copies are needed."""

    def __init__(self, model, lr: float = 2e-4, betas=(0.9, 0.99), eps: float = 1e-8, weight_decay: float = 1e-4,
                 params: Optional[List[torch.nn.Parameter]] = None, grad_clip: float = 0.0, autograd_params: bool = False):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.grad_clip = grad_clip
        self.params = _trained_parameters(model) if params is None else list(params)
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
        self.offsets = {}
        for p, n in zip(self.params, sizes):
            pv = self.flat_p[off:off + p.numel()].view(p.shape)
            pv.copy_(p.data)
            p.data = pv                                   # the module's parameter now lives in the flat buffer
            gview[id(p)] = self.flat_g[off:off + p.numel()].view(p.shape)
            self.offsets[id(p)] = off
            off += n
        self.numel = total
        if autograd_params:
            for p in self.params:
                p.grad = gview[id(p)]
        else:
            self._install(gview)

    def grad_of(self, p: torch.nn.Parameter) -> torch.Tensor:
        """View of the flat gradient buffer that belongs to parameter `p`."""
        o = self.offsets[id(p)]
        return self.flat_g[o:o + p.numel()].view(p.shape)

    def _install(self, gview) -> None:
        """Tell the engine where each module's parameter gradients live (see engine.GradArena).  Only modules whose
        parameters this optimizer owns are wired (the captioning task trains blocks[4] but not the enhance convs)."""
        m = self.model
        enc = m.encoder
        stem = enc.x3d.blocks[0]
        if id(stem.conv.conv_t.weight) in gview:
            object.__setattr__(stem, "_c3d_grad_views", [gview[id(p)] for p in (stem.conv.conv_t.weight,
                                                                               stem.conv.conv_xy.weight,
                                                                               stem.norm.weight, stem.norm.bias)])
            object.__setattr__(stem, "_c3d_perc_grad_view", gview[id(enc.perception_frames)])
        for i in range(1, 5):
            stage = enc.x3d.blocks[i]
            plist = stage.param_list()
            if plist and id(plist[0]) in gview:
                object.__setattr__(stage, "_c3d_grad_views", [gview[id(p)] for p in plist])
        for fc in enc.fc:
            if id(fc[0].weight) in gview:
                fc[0].weight._c3d_grad_view = gview[id(fc[0].weight)]
        for name in ("decoder", "decoder_pre", "decoder_post", "decoder_change", "decoder_cls", "decoder_loc"):
            dec = getattr(m, name, None)
            if dec is not None and hasattr(dec, "param_list") and id(dec.param_list()[0]) in gview:
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
                      self.betas[1], self.eps, self.weight_decay, self.step_count, 1.0 / world, self.grad_clip)


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
    `self.parts` holds the device scalars of the last iteration's loss terms (seg / binary / sim), as the scripts log;
    `self.outs` the heads' outputs of the last iteration (SCD / BDA: the scripts compute training accuracy from them).
    Under the CUDA graph both live in the graph's memory pool and are overwritten by every replay."""

    def __init__(self, model, lr: float = 2e-4, use_graph: bool = False, task: str = "bcd"):
        if task not in ("bcd", "scd", "bda"):
            raise ValueError(f"TrainStep: unknown task {task!r}")
        self.task = task
        self.model = model.train()
        self.opt = FlatAdam(model, lr=lr)
        self.cm = torch.zeros(2, 2, dtype=torch.int64, device=self.opt.flat_p.device)
        self.sim = ChangeSimilarity()
        self.parts = {}
        self.outs = ()
        self.use_graph = use_graph
        self.graph = None
        self.static = None
        self.loss = None
        self.adam_launches = 1

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
            self.outs = (pre_mask.detach(), post_mask.detach(), change_mask.detach())
            return seg * 0.5 + binary + sim
        label_loc, label_cls = labels
        pred_cls, pred_loc = self.model.update_bda(pre, post)
        seg = cross_entropy_2d(pred_cls, label_cls, 0)
        binary = bce_dice_loss(pred_loc, label_loc.unsqueeze(1), cm=self.cm)
        self.parts = {"seg": seg.detach(), "binary": binary.detach()}
        self.outs = (pred_cls.detach(), pred_loc.detach())
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
            self._capture((pre, post) + labels)
        for dst, src in zip(self.static, (pre, post) + labels):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.opt.all_reduce()
        self.opt.step(lr)                 # Adam outside the graph: step count / lr change every iteration
        return self.loss

    def _capture(self, tensors) -> None:
        """Warm-up on a side stream (also sets every kernel's shared-memory attribute), then capture.  The warm-up is
        a real train-mode iteration on the first batch, which the first replay repeats: the BatchNorm buffers
        (running_mean / running_var / num_batches_tracked) are snapshotted and restored around it so they advance once
        per batch exactly like the eager path and the reference."""
        self.static = tuple(t.clone() for t in tensors)
        buffers = [b for n, b in self.model.named_buffers() if n.rsplit(".", 1)[-1] in BN_BUFFERS]
        snap = [b.clone() for b in buffers]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._iteration(*self.static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for b, v in zip(buffers, snap):
            b.copy_(v)
        from . import _lib
        n0 = _lib.LAUNCHES[0]
        self.graph = torch.cuda.CUDAGraph()
        # The capture stream is a HIGH-priority stream: the kernel nodes of the dgrad chain inherit it, the weight-gradient
        # kernels forked onto engine._side_stream() (default priority) do not, so whenever both have thread blocks pending
        # the block scheduler places the critical path first (C3D_GRAPH_PRIORITY=0: default-priority capture stream).
        prio = -1 if os.environ.get("C3D_GRAPH_PRIORITY", "1") == "1" else 0
        cap_stream = torch.cuda.Stream(priority=prio)
        # thread_local: a DataLoader pin-memory thread may call cudaHostAlloc while this thread captures
        with torch.cuda.graph(self.graph, stream=cap_stream, capture_error_mode="thread_local"):
            self.loss = self._iteration(*self.static)
        self.captured_launches = _lib.LAUNCHES[0] - n0 + self.adam_launches      # + the Adam launch(es) outside the graph
        if getattr(self, "cm", None) is not None:
            self.cm.zero_()                                     # drop the warm-up iteration's counts

    def scores(self) -> dict:
        """cm2score of the accumulated training confusion matrix (one small device->host copy)."""
        from .metrics import cm2score
        return cm2score(self.cm.cpu().numpy())


class BCDTrainStep(TrainStep):
    """TrainStep(task='bcd') — the BASELINE.json headline configuration."""

    def __init__(self, model, lr: float = 2e-4, use_graph: bool = False):
        super().__init__(model, lr=lr, use_graph=use_graph, task="bcd")


def packed_caption_loss(scores: torch.Tensor, caps_sorted: torch.Tensor, decode_lengths: torch.Tensor) -> torch.Tensor:
    """The captioning loss of scripts/train_CC.py:124-133 without leaving the device:

        targets = caps_sorted[:, 1:]
        scores  = pack_padded_sequence(scores,  decode_lengths, batch_first=True).data
        targets = pack_padded_sequence(targets, decode_lengths, batch_first=True).data
        loss = nn.CrossEntropyLoss(ignore_index=0)(scores, targets)

    Packing keeps step t of sequence b iff t < decode_lengths[b]; the mean runs over the kept tokens whose target is
    not 0.  Same set, same mean: mark the dropped steps with the ignored target instead of gathering them (no
    `.tolist()`, so the iteration can be captured in a CUDA graph)."""
    B, L, V = scores.shape
    keep = torch.arange(L - 1, device=scores.device).unsqueeze(0) < decode_lengths.to(scores.device).unsqueeze(1)
    tgt = torch.where(keep, caps_sorted[:, 1:], torch.zeros_like(caps_sorted[:, 1:]))
    return torch.nn.functional.cross_entropy(scores[:, :L - 1].reshape(-1, V), tgt.reshape(-1), ignore_index=0)


class CCTrainStep(TrainStep):
    """One iteration of the change-captioning loop (scripts/train_CC.py:105-146) on the B200 engine:

        feat = encoder(A, B, output_final=True); memory = rearrange(feat, 'b c h w -> (h w) b c')
        scores, caps_sorted, decode_lengths, _ = decoder(memory, caps, caplens)
        loss = CrossEntropyLoss(ignore_index=0)(packed scores, packed targets)
        zero_grad; loss.backward(); clip_gradient(+-grad_clip) on both optimizers; encoder_optimizer.step(); decoder_optimizer.step()

    The encoder (stem .. res5, 55 residual blocks) runs on the sm_100a kernels with its gradients written straight into
    the encoder optimizer's flat buffer; the captioning head is torch ops whose autograd
    accumulates into the decoder optimizer's flat buffer.  Both optimizers are torch.optim.Adam(lr, weight_decay=1e-5)
    with default betas (scripts/train_CC.py:440-458), restated as one fused clamp+Adam launch each.  Parameters that
    never receive a gradient (enhance convs and blocks[5] of the encoder; the unused attention / feed-forward /
    fc_alpha modules of the decoder layers) are left out, as torch's Adam skips `grad is None`."""

    def __init__(self, model, encoder_lr: float = 1e-4, decoder_lr: float = 1e-4, grad_clip: float = 5.0,
                 use_graph: bool = False):
        self.task = "cc"
        self.model = model.train()
        enc_params = [p for n, p in model.encoder.named_parameters()
                      if p.requires_grad and not n.startswith("fc.") and ".blocks.5." not in n]
        self.opt = FlatAdam(model, lr=encoder_lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, params=enc_params,
                            grad_clip=grad_clip)
        self.dec_opt = FlatAdam(model, lr=decoder_lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5,
                                params=model.decoder.live_parameters(), grad_clip=grad_clip, autograd_params=True)
        self.cm = None
        self.parts = {}
        self.outs = ()
        self.use_graph = use_graph
        self.graph = None
        self.static = None
        self.loss = None
        self.adam_launches = 2

    def _iteration(self, pre, post, caps, caplens) -> torch.Tensor:
        self.opt.zero_grad()
        self.dec_opt.zero_grad()
        feat = self.model.update_cc(pre, post)                                     # (B, 192, H/16, W/16)
        B, C, H, W = feat.shape
        memory = feat.permute(2, 3, 0, 1).reshape(H * W, B, C)                     # 'b c h w -> (h w) b c'
        scores, caps_sorted, decode_lengths, _ = self.model.decoder.forward_device(memory, caps, caplens)
        loss = packed_caption_loss(scores, caps_sorted, decode_lengths)
        self.outs = (scores.detach(), caps_sorted, decode_lengths)      # what scripts/train_CC.py:148-151 scores
        loss.backward()
        return loss.detach()

    def eager(self, pre, post, caps, caplens, lr: Optional[float] = None):
        loss = self._iteration(pre, post, caps, caplens)
        self._step()
        return loss

    def _step(self) -> None:
        self.opt.all_reduce()
        self.dec_opt.all_reduce()
        self.opt.step()
        self.dec_opt.step()

    def __call__(self, pre, post, caps, caplens, lr: Optional[float] = None):
        if not self.use_graph:
            return self.eager(pre, post, caps, caplens)
        if self.graph is None:
            self._capture((pre, post, caps, caplens))
        for dst, src in zip(self.static, (pre, post, caps, caplens)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self._step()
        return self.loss

    def scores(self) -> dict:
        raise NotImplementedError("the captioning loop scores captions (BLEU/CIDEr), not a confusion matrix")
