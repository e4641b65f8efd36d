"""CPU oracle for the Change3D hot path — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product package `change3d_b200` never does.

This is a plain-PyTorch (CPU, fp32 or fp64) *functional* restatement of the reference
algorithm, written against a state dict that uses the reference's key schema
(SURVEY.md §9.3).  Every function cites the reference lines it follows.  Third-party
arithmetic (pytorchvideo 0.1.5 / fvcore 0.1.5.post20221221, not vendored in the
reference) is restated from its published behaviour.

Pinning: the reference ships no tests or golden vectors for this path, so the oracle is
pinned against the reference's own `model/*.py` executed in the authoring container
(through `oracle/pv_shim`, see `oracle/reference_loader.py`); the resulting vectors are
committed under `tests/golden/` by `oracle/make_golden.py`.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

BN_EPS = 1e-5        # model/x3d.py:33 (norm_eps)
BN_MOMENTUM = 0.1    # model/x3d.py:34 (norm_momentum)
_momentum_override: List[float] = []   # calibrate_running_stats() pushes 1.0 here

# X3D-L instance built by create_x3d(input_clip_length=3, depth_factor=5.0)
# (model/trainer.py:40, model/x3d.py:653-685): (dim_in, dim_inner, dim_out, depth)
STAGES = ((24, 54, 24, 5), (24, 108, 48, 10), (48, 216, 96, 25), (96, 432, 192, 15))
EMBED_DIMS = (24, 24, 48, 96)  # model/trainer.py:186


def se_reduced(dim_inner: int, ratio: float = 0.0625) -> int:
    """pytorchvideo round_width(dim_inner, se_ratio) (call site model/x3d.py:197)."""
    w = dim_inner * ratio
    out = max(8, int(w + 4) // 8 * 8)
    if out < 0.9 * w:
        out += 8
    return int(out)


# --------------------------------------------------------------------------------------
# schema
# --------------------------------------------------------------------------------------
def _bn_keys(prefix: str, c: int) -> List[Tuple[str, Tuple[int, ...]]]:
    return [(prefix + ".weight", (c,)), (prefix + ".bias", (c,)),
            (prefix + ".running_mean", (c,)), (prefix + ".running_var", (c,)),
            (prefix + ".num_batches_tracked", ())]


def x3d_schema() -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered (key, shape) list of create_x3d()'s state dict — SURVEY.md §9.3 / model/x3d.py."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    out.append(("blocks.0.conv.conv_t.weight", (24, 3, 1, 3, 3)))      # x3d.py:70-77 (spatial)
    out.append(("blocks.0.conv.conv_xy.weight", (24, 1, 5, 1, 1)))     # x3d.py:78-86 (temporal dw)
    out += _bn_keys("blocks.0.norm", 24)
    for s, (cin, ci, cout, depth) in enumerate(STAGES, start=1):
        for b in range(depth):
            p = f"blocks.{s}.res_blocks.{b}."
            bin_ = cin if b == 0 else cout
            if b == 0:                                                   # x3d.py:301-311
                out.append((p + "branch1_conv.weight", (cout, bin_, 1, 1, 1)))
                if bin_ != cout:                                         # x3d.py:296-298,312
                    out += _bn_keys(p + "branch1_norm", cout)
            out.append((p + "branch2.conv_a.weight", (ci, bin_, 1, 1, 1)))
            out += _bn_keys(p + "branch2.norm_a", ci)
            out.append((p + "branch2.conv_b.weight", (ci, 1, 3, 3, 3)))
            out += _bn_keys(p + "branch2.norm_b.0", ci)
            if (b + 1) % 2:                                              # x3d.py:406
                r = se_reduced(ci)
                out.append((p + "branch2.norm_b.1.block.0.weight", (r, ci, 1, 1, 1)))
                out.append((p + "branch2.norm_b.1.block.0.bias", (r,)))
                out.append((p + "branch2.norm_b.1.block.2.weight", (ci, r, 1, 1, 1)))
                out.append((p + "branch2.norm_b.1.block.2.bias", (ci,)))
            out.append((p + "branch2.conv_c.weight", (cout, ci, 1, 1, 1)))
            out += _bn_keys(p + "branch2.norm_c", cout)
    out.append(("blocks.5.pool.pre_conv.weight", (432, 192, 1, 1, 1)))   # x3d.py:469-471
    out += _bn_keys("blocks.5.pool.pre_norm", 432)
    out.append(("blocks.5.pool.post_conv.weight", (2048, 432, 1, 1, 1)))
    out.append(("blocks.5.proj.weight", (400, 2048)))
    out.append(("blocks.5.proj.bias", (400,)))
    return out


def decoder_schema(prefix: str, num_class: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """model/change_decoder.py:30-55 with in_dim=[24,24,48,96]."""
    c1, c2, c3, c4 = EMBED_DIMS
    p = prefix
    return [
        (p + "up_c4.0.weight", (c3, c4, 1, 1)), (p + "up_c4.1.weight", (c3, c3, 4, 4)), (p + "up_c4.1.bias", (c3,)),
        (p + "up_c3.0.weight", (c2, c3, 1, 1)), (p + "up_c3.1.weight", (c2, c2, 4, 4)), (p + "up_c3.1.bias", (c2,)),
        (p + "up_c2.0.weight", (c1, c2, 1, 1)), (p + "up_c2.1.weight", (c1, c1, 4, 4)), (p + "up_c2.1.bias", (c1,)),
        (p + "up_c1.0.weight", (num_class, c1, 3, 3)),
    ]


def trainer_heads(task: str) -> List[Tuple[str, bool, int]]:
    """(attribute name, has_sigmoid, perception-frame index fed to it) per task, in module
    registration AND return order — model/trainer.py:192-213 (construction), :236-239 (bcd),
    :258-266 (scd returns pre, post, change; pre<-p0, change<-p1, post<-p2), :283-290 (bda)."""
    return {"bcd": [("decoder", True, 0)],
            "scd": [("decoder_pre", False, 0), ("decoder_post", False, 2), ("decoder_change", True, 1)],
            "bda": [("decoder_cls", False, 0), ("decoder_loc", True, 1)]}[task]


def trainer_schema(task: str, P: int, H: int, W: int, num_class: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered state dict of model.trainer.Trainer(args) for a change-decoder task
    (registration order: encoder.x3d, encoder.perception_frames?, ...).  Order of nn.Module
    registration in Encoder.__init__: x3d (module), perception_frames (parameter), fc (module);
    state_dict lists a module's own parameters before its sub-modules (model/trainer.py:40-69)."""
    out = [("encoder.perception_frames", (1, 3, P, H, W))]
    out += [("encoder.x3d." + k, s) for k, s in x3d_schema()]
    out += [(f"encoder.fc.{i}.0.weight", (c, c, 1, 1)) for i, c in enumerate(EMBED_DIMS)]
    for name, sig, _ in trainer_heads(task):
        out += decoder_schema(name + ".", 1 if sig else num_class)
    return out


# --------------------------------------------------------------------------------------
# deterministic synthesis (weights and inputs) — shared by golden generation and GPU tests
# --------------------------------------------------------------------------------------
def synth_tensor(key: str, shape: Tuple[int, ...], seed: int, dtype=torch.float32) -> Tensor:
    """Value of one state-dict entry, a function of (seed, key, shape) only.

    BN affine/running stats and SE biases are randomised (default init hides
    folding/ordering bugs); conv weights are fan-in scaled so activations stay O(1)."""
    h = 0
    for ch in key:
        h = (h * 131 + ord(ch)) % 2147483647
    g = torch.Generator().manual_seed((seed * 1000003 + h) % 2147483647)
    if key.endswith("num_batches_tracked"):
        return torch.zeros((), dtype=torch.int64)
    if key.endswith("running_mean"):
        return (0.2 * torch.randn(shape, generator=g)).to(dtype)
    if key.endswith("running_var"):
        return (0.5 + torch.rand(shape, generator=g)).to(dtype)
    is_norm = ".norm" in key or "branch1_norm" in key or "pre_norm" in key
    if is_norm and key.endswith(".weight"):
        return (0.75 + 0.5 * torch.rand(shape, generator=g)).to(dtype)
    if is_norm and key.endswith(".bias"):
        return (0.2 * torch.randn(shape, generator=g)).to(dtype)
    if key.endswith(".bias"):
        return (0.3 * torch.randn(shape, generator=g)).to(dtype)
    if key.endswith("perception_frames"):
        return torch.randn(shape, generator=g).to(dtype)
    fan_in = 1
    for d in shape[1:]:
        fan_in *= d
    if ".1.weight" in key and len(shape) == 4 and shape[-1] == 4:
        fan_in = shape[0] * 4        # ConvTranspose2d k4 s2: 4 taps reach each output pixel
    gain = 2.0
    if "up_c1.0.weight" in key:
        gain = 2.0 * 0.02 ** 2       # keep the head's logits O(1): saturated sigmoids make BCE gradients meaningless
    return (torch.randn(shape, generator=g) * math.sqrt(gain / max(fan_in, 1))).to(dtype)


def synth_state_dict(schema: Sequence[Tuple[str, Tuple[int, ...]]], seed: int, dtype=torch.float32) -> StateDict:
    return {k: synth_tensor(k, tuple(s), seed, dtype) for k, s in schema}


def synth_inputs(B: int, H: int, W: int, seed: int = 16, dtype=torch.float32):
    """pre, post ~ N(0,1) (B,3,H,W); BCD target ~ Bernoulli(0.05)."""
    g = torch.Generator().manual_seed(seed)
    pre = torch.randn(B, 3, H, W, generator=g).to(dtype)
    post = torch.randn(B, 3, H, W, generator=g).to(dtype)
    target = (torch.rand(B, 1, H, W, generator=g) < 0.05).to(dtype)
    return pre, post, target


# --------------------------------------------------------------------------------------
# the algorithm
# --------------------------------------------------------------------------------------
def _bn(sd: StateDict, p: str, x: Tensor, training: bool) -> Tensor:
    """nn.BatchNorm3d(eps=1e-5, momentum=0.1): batch statistics + running-stat update in
    training, running statistics in eval (model/x3d.py:94-98,176-180,203-208,217-221,296-298)."""
    if training and (p + ".num_batches_tracked") in sd:
        sd[p + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training, _momentum_override[-1] if _momentum_override else BN_MOMENTUM,
                        BN_EPS)


def stem(sd: StateDict, x: Tensor, training: bool, p: str = "blocks.0.") -> Tensor:
    """model/x3d.py:23-106 + pytorchvideo ResNetBasicStem/Conv2plus1d: the module stored as
    conv_t (spatial 1x3x3, 3->24) runs first, then conv_xy (depthwise 5x1x1), BN, ReLU; stride 1
    (model/x3d.py:563-564)."""
    x = F.conv3d(x, sd[p + "conv.conv_t.weight"], None, stride=1, padding=(0, 1, 1))
    x = F.conv3d(x, sd[p + "conv.conv_xy.weight"], None, stride=1, padding=(2, 0, 0), groups=24)
    return F.relu(_bn(sd, p + "norm", x, training))


def squeeze_excite(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """fvcore SqueezeExcitation(is_3d=True): x * sigmoid(W2 relu(W1 mean_{T,H,W}(x) + b1) + b2)
    (call site model/x3d.py:194-202); the pool spans T as well as H, W."""
    s = x.mean(dim=[2, 3, 4], keepdim=True)
    s = F.relu(F.conv3d(s, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"]))
    s = torch.sigmoid(F.conv3d(s, sd[p + ".block.2.weight"], sd[p + ".block.2.bias"]))
    return x * s


def bottleneck(sd: StateDict, p: str, x: Tensor, stride: int, has_se: bool, training: bool) -> Tensor:
    """model/x3d.py:109-232 + pytorchvideo BottleneckBlock:
    conv_a 1x1x1 -> BN -> ReLU -> depthwise 3x3x3 (stride (1,s,s), pad 1) -> BN -> [SE] -> Swish
    -> conv_c 1x1x1 -> BN."""
    ci = sd[p + "conv_a.weight"].shape[0]
    x = F.relu(_bn(sd, p + "norm_a", F.conv3d(x, sd[p + "conv_a.weight"]), training))
    x = F.conv3d(x, sd[p + "conv_b.weight"], None, stride=(1, stride, stride), padding=1, groups=ci)
    x = _bn(sd, p + "norm_b.0", x, training)
    if has_se:
        x = squeeze_excite(sd, p + "norm_b.1", x)
    x = x * torch.sigmoid(x)                                      # pytorchvideo Swish
    return _bn(sd, p + "norm_c", F.conv3d(x, sd[p + "conv_c.weight"]), training)


def res_block(sd: StateDict, p: str, x: Tensor, first: bool, has_se: bool, training: bool) -> Tensor:
    """model/x3d.py:235-328 + pytorchvideo ResBlock: relu(shortcut(x) + branch2(x)); the first
    block of a stage has a 1x1x1 stride-(1,2,2) shortcut conv, with BN only where C_in != C_out."""
    if first:
        sc = F.conv3d(x, sd[p + "branch1_conv.weight"], None, stride=(1, 2, 2))
        if (p + "branch1_norm.weight") in sd:
            sc = _bn(sd, p + "branch1_norm", sc, training)
    else:
        sc = x
    return F.relu(sc + bottleneck(sd, p + "branch2.", x, 2 if first else 1, has_se, training))


def res_stage(sd: StateDict, s: int, x: Tensor, training: bool, prefix: str = "", depth: Optional[int] = None) -> Tensor:
    """model/x3d.py:331-412: `depth` blocks, stride on block 0 only, SE on even block indices.
    (`depth` override: tests run a truncated stage to keep fp32 round-off from compounding.)"""
    depth = STAGES[s - 1][3] if depth is None else depth
    for b in range(depth):
        x = res_block(sd, f"{prefix}blocks.{s}.res_blocks.{b}.", x, b == 0, (b + 1) % 2 == 1, training)
    return x


def x3d_block(sd: StateDict, i: int, x: Tensor, training: bool, prefix: str = "") -> Tensor:
    """create_x3d(...).blocks[i](x) for i in 0..4 (model/x3d.py:543-744)."""
    return stem(sd, x, training, prefix + "blocks.0.") if i == 0 else res_stage(sd, i, x, training, prefix)


def enhance(x: Tensor, w: Tensor, P: int) -> Tensor:
    """Encoder.enhance (model/trainer.py:71-108): frame T//2 += relu(conv1x1(|x[:,:,0]-x[:,:,P+1]|))."""
    mid = x.shape[2] // 2
    e = F.relu(F.conv2d(torch.abs(x[:, :, 0] - x[:, :, P + 1]), w))
    out = x.clone()
    out[:, :, mid] = x[:, :, mid] + e
    return out


def assemble_frames(pre: Tensor, post: Tensor, perception: Tensor) -> Tensor:
    """Encoder.forward (model/trainer.py:143-162): [pre, P learnable frames, post] on dim 2."""
    B = pre.shape[0]
    return torch.cat([pre.unsqueeze(2), perception.expand(B, -1, -1, -1, -1), post.unsqueeze(2)], dim=2)


def encoder_forward(sd: StateDict, pre: Tensor, post: Tensor, P: int, training: bool,
                    output_final: bool = False, prefix: str = "encoder."):
    """Encoder.base_forward (model/trainer.py:110-141)."""
    x = assemble_frames(pre, post, sd[prefix + "perception_frames"])
    if output_final:
        for i in range(5):
            x = x3d_block(sd, i, x, training, prefix + "x3d.")
        return x[:, :, P]
    feats = []
    for i in range(4):
        x = x3d_block(sd, i, x, training, prefix + "x3d.")
        x = enhance(x, sd[f"{prefix}fc.{i}.0.weight"], P)
        feats.append([x[:, :, k + 1] for k in range(P)])
    return feats


def change_decoder(sd: StateDict, p: str, f: Sequence[Tensor], has_sigmoid: bool) -> Tensor:
    """ChangeDecoder.forward (model/change_decoder.py:57-81)."""
    c1, c2, c3, c4 = f

    def up(name, t):
        t = F.conv2d(t, sd[p + name + ".0.weight"])
        return F.conv_transpose2d(t, sd[p + name + ".1.weight"], sd[p + name + ".1.bias"], stride=2, padding=1)

    c3f = c3 + up("up_c4", c4)
    c2f = c2 + up("up_c3", c3f)
    c1f = c1 + up("up_c2", c2f)
    pred = F.conv2d(c1f, sd[p + "up_c1.0.weight"], None, padding=1)
    return torch.sigmoid(pred) if has_sigmoid else pred


def trainer_forward(sd: StateDict, task: str, pre: Tensor, post: Tensor, training: bool):
    """Trainer.update_bcd / update_scd / update_bda (model/trainer.py:221-290)."""
    heads = trainer_heads(task)
    P = len(heads)
    feats = encoder_forward(sd, pre, post, P, training)
    outs = [change_decoder(sd, name + ".", [lvl[k] for lvl in feats], sig) for name, sig, k in heads]
    return outs[0] if task == "bcd" else tuple(outs)


def calibrate_running_stats(sd: StateDict, task: str, pre: Tensor, post: Tensor) -> StateDict:
    """Test helper (not reference behaviour): replace the synthetic running statistics by the
    batch statistics of one training-mode pass (momentum 1.0), so that eval-mode activations stay
    O(1) through 40 residual blocks the way they do with real pretrained statistics.  Applied
    identically when the golden vectors are generated (there on the reference modules, by setting
    every BatchNorm's `momentum` attribute to 1.0) and when they are checked."""
    out = clone_sd(sd)
    _momentum_override.append(1.0)
    try:
        with torch.no_grad():
            trainer_forward(out, task, pre, post, training=True)
    finally:
        _momentum_override.pop()
    for k in out:
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.int64)
    return out


def bce_dice_loss(inputs: Tensor, targets: Tensor) -> Tensor:
    """BCEDiceLoss (model/utils.py:154-169)."""
    bce = F.binary_cross_entropy(inputs, targets)
    inter = (inputs * targets).sum()
    eps = 1e-5
    return bce + 1 - (2 * inter + eps) / (inputs.sum() + targets.sum() + eps)


def adam_reference_step(params: Sequence[Tensor], grads: Sequence[Optional[Tensor]], state: dict, lr: float,
                        betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-4) -> None:
    """torch.optim.Adam as configured at scripts/train_BCD.py:284-290 (L2 added to the gradient,
    not decoupled; parameters whose grad is None are skipped)."""
    b1, b2 = betas
    for i, (p, g) in enumerate(zip(params, grads)):
        if g is None:
            continue
        st = state.setdefault(i, {"step": 0, "m": torch.zeros_like(p), "v": torch.zeros_like(p)})
        st["step"] += 1
        g = g + weight_decay * p
        st["m"].mul_(b1).add_(g, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** st["step"]
        bc2 = 1 - b2 ** st["step"]
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(st["m"], denom, value=-lr / bc1)


def poly_lr(base_lr: float, it: int, max_it: int, epoch: int) -> float:
    """adjust_learning_rate, poly mode + 200-iteration warm-up (model/utils.py:130-146)."""
    lr = base_lr * (1 - it * 1.0 / max_it) ** 0.9
    if epoch == 0 and it < 200:
        lr = base_lr * 0.9 * (it + 1) / 200 + 0.1 * base_lr
    return lr


def clone_sd(sd: StateDict, dtype=None, requires_grad: bool = False) -> StateDict:
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        if requires_grad and t.is_floating_point() and "running_" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


# ------------------------------------------------------------------------------------------------------------
# Losses and metrics of the train step — restated; pinned to the reference's own functions by
# tests/golden/losses.npz (oracle/make_golden_losses.py).
def cross_entropy_2d(inputs: Tensor, targets: Tensor, ignore_index: int = -1) -> Tensor:
    """CrossEntropyLoss2d (model/utils.py:171-178): mean over non-ignored pixels of -log_softmax(inputs, 1)[target]."""
    logp = inputs - torch.logsumexp(inputs, dim=1, keepdim=True)
    valid = targets != ignore_index
    picked = logp.gather(1, targets.clamp(min=0).unsqueeze(1)).squeeze(1)
    return -(picked * valid).sum() / valid.sum()


def change_similarity(x1: Tensor, x2: Tensor, label_change: Tensor) -> Tensor:
    """ChangeSimilarity (model/utils.py:180-203): CosineEmbeddingLoss(margin 0, mean) between the per-pixel class
    distributions; target +1 on unchanged pixels (1 - cos), -1 on changed pixels (max(0, cos))."""
    b, c, h, w = x1.shape
    p1 = torch.softmax(x1, dim=1).permute(0, 2, 3, 1).reshape(b * h * w, c)
    p2 = torch.softmax(x2, dim=1).permute(0, 2, 3, 1).reshape(b * h * w, c)
    changed = label_change.reshape(b * h * w).bool()
    eps = 1e-12                                          # aten cosine_embedding_loss EPSILON
    cos = (p1 * p2).sum(1) / torch.sqrt(((p1 * p1).sum(1) + eps) * ((p2 * p2).sum(1) + eps))
    per = torch.where(changed, cos.clamp(min=0), 1 - cos)
    return per.mean()


def confusion_matrix(num_classes: int, label_gt, label_pred) -> np.ndarray:
    """get_confuse_matrix for one batch (utils/metric_tool.py:111-128): hist[gt][pred] over 0 <= gt < num_classes."""
    gt = np.asarray(label_gt).reshape(-1)
    pr = np.asarray(label_pred).reshape(-1)
    mask = (gt >= 0) & (gt < num_classes)
    return np.bincount(num_classes * gt[mask].astype(int) + pr[mask], minlength=num_classes ** 2) \
        .reshape(num_classes, num_classes).astype(np.int64)


def cm_scores(cm) -> dict:
    """cm2score (utils/metric_tool.py:84-108)."""
    hist = np.asarray(cm, dtype=np.float64)
    e = np.finfo(np.float32).eps
    tp, fn, fp, tn = hist[1, 1], hist[1, 0], hist[0, 1], hist[0, 0]
    oa = (tp + tn) / (tp + fn + fp + tn + e)
    recall = tp / (tp + fn + e)
    precision = tp / (tp + fp + e)
    f1 = 2 * recall * precision / (recall + precision + e)
    iou = tp / (tp + fp + fn + e)
    pre = ((tp + fn) * (tp + fp) + (tn + fp) * (tn + fn)) / (tp + fp + tn + fn) ** 2
    kappa = (oa - pre) / (1 - pre)
    return {'Kappa': kappa, 'IoU': iou, 'F1': f1, 'OA': oa, 'recall': recall, 'precision': precision, 'Pre': pre}


def task_loss(task: str, outputs, labels) -> Tensor:
    """The training loss of each reference script on the model outputs:
    bcd (scripts/train_BCD.py:200-201)  BCEDice(output, target)
    scd (scripts/train_SCD.py:215-229)  labels (pre_label, post_label, label_change) int64 (B,H,W), the class labels
                                        masked by the change map; 0.5*(CE0(pre)+CE0(post)) + BCEDice(change) + Sim
    bda (scripts/train_BDA.py:181-199)  labels (label_loc float (B,H,W), label_cls int64 (B,H,W)); CE0(cls) + BCEDice(loc)
    CE0 = CrossEntropyLoss2d(ignore_index=0)."""
    if task == "bcd":
        return bce_dice_loss(outputs, labels[0])
    if task == "scd":
        pre_mask, post_mask, change_mask = outputs
        pre_label, post_label, label_change = labels
        pre_label, post_label = pre_label * label_change, post_label * label_change
        seg = cross_entropy_2d(pre_mask, pre_label, 0) + cross_entropy_2d(post_mask, post_label, 0)
        binary = bce_dice_loss(change_mask, label_change.unsqueeze(1).to(change_mask.dtype))
        sim = change_similarity(pre_mask[:, 1:], post_mask[:, 1:], label_change.unsqueeze(1))
        return seg * 0.5 + binary + sim
    if task == "bda":
        pred_cls, pred_loc = outputs
        label_loc, label_cls = labels
        return cross_entropy_2d(pred_cls, label_cls, 0) + bce_dice_loss(pred_loc, label_loc.unsqueeze(1).to(pred_loc.dtype))
    raise ValueError(task)


def synth_labels(task: str, B: int, H: int, W: int, num_class: int, seed: int = 16):
    """Synthetic labels with the statistics of SURVEY.md §8(d) items 2-4."""
    g = torch.Generator().manual_seed(seed + 1000)
    if task == "bcd":
        return ((torch.rand(B, 1, H, W, generator=g) < 0.05).float(),)
    if task == "scd":
        change = (torch.rand(B, H, W, generator=g) < 0.2).long()
        return (torch.randint(1, num_class, (B, H, W), generator=g), torch.randint(1, num_class, (B, H, W), generator=g),
                change)
    if task == "bda":
        loc = (torch.rand(B, H, W, generator=g) < 0.1).float()
        return (loc, (loc * torch.randint(1, num_class, (B, H, W), generator=g)).long())
    raise ValueError(task)
