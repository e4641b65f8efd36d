"""Data-parallel runners with the surface of the reference's `scripts/train_SCD.py` and `scripts/train_BDA.py`
(`runner.py` is the `train_BCD.py` counterpart and supplies the shared plumbing).

Same command-line flags and defaults (train_SCD.py:443-552, train_BDA.py:373-482), same epoch structure
(`trainValidate`: train every epoch, validate on the TEST split from epoch 1, poly learning rate per iteration,
`checkpoint.pth.tar` with each script's keys + `best_model.pth`, final test with the best model, the log columns of
model/utils.py:235-262), same loss (`train_step.TrainStep(task=...)`: train_SCD.py:226-229, train_BDA.py:192-194).

What differs is where the work happens (as in runner.py): one process per GPU under torchrun with a DistributedSampler
and one NCCL all-reduce of the flat gradient buffer per step, batches staged by `DevicePrefetcher`, the step as a CUDA
graph, and the metrics accumulated ON THE DEVICE:
  * SCD (train_SCD.py:139-170, 236-257): the scripts copy both class maps and the labels to the host every step and
    loop over samples in numpy (`accuracy`, `SCDD_eval_all`).  Here the per-image accuracies are one reduction per
    batch and the 7x7 histogram hist[pred][label] is accumulated by `c3d_confusion_matrix`; Fscd / mIoU / Sek are
    evaluated from that histogram once per epoch with the reference's formulas (model/utils.py:345-377).
  * BDA (train_BDA.py:120-147): `Evaluator(2)` for the localisation mask and `Evaluator(num_class)` over the pixels
    with `label_loc > 0` become two device histograms; F1 scores from them as model/utils.py:379-430.
"""
from __future__ import annotations

import math
import os
import time
from argparse import ArgumentParser
from os.path import join as osp

import numpy as np
import torch
import torch.distributed as dist

from . import losses
from .input_pipeline import DevicePrefetcher
from .losses import ChangeSimilarity, bce_dice_loss, cross_entropy_2d
from .model.trainer import Trainer
from .model.utils import adjust_learning_rate
from .runner import _state_dict_copy, load_checkpoint
from .train_step import TrainStep


# ---------------------------------------------------------------------------------------------
# command line (train_SCD.py:443-552, train_BDA.py:373-482): same names / types / defaults
# ---------------------------------------------------------------------------------------------
_DEFAULTS = {
    "scd": dict(dataset="HRSCD", file_root="path/to/HRSCD", num_perception_frame=3, num_class=6, max_steps=80000,
                batch_size=8),
    "bda": dict(dataset="xBD", file_root="path/to/xBD", num_perception_frame=2, num_class=5, max_steps=200000,
                batch_size=12),
}


def build_parser(task: str) -> ArgumentParser:
    d = _DEFAULTS[task]
    p = ArgumentParser()
    p.add_argument('--dataset', default=d["dataset"], help='Dataset selection')
    p.add_argument('--file_root', default=d["file_root"], help='path to the dataset directory')
    p.add_argument('--in_height', type=int, default=256, help='Height of RGB image')
    p.add_argument('--in_width', type=int, default=256, help='Width of RGB image')
    p.add_argument('--num_perception_frame', type=int, default=d["num_perception_frame"], help='Number of perception frames')
    p.add_argument('--num_class', type=int, default=d["num_class"], help='Number of classes')
    p.add_argument('--max_steps', type=int, default=d["max_steps"], help='Max number of iterations')
    p.add_argument('--batch_size', type=int, default=d["batch_size"], help='Batch size (per process, as in the reference)')
    p.add_argument('--num_workers', type=int, default=4, help='Number of parallel threads')
    p.add_argument('--lr', type=float, default=2e-4, help='Initial learning rate')
    p.add_argument('--lr_mode', default='poly', help='Learning rate policy: step or poly')
    p.add_argument('--step_loss', type=int, default=100, help='Decrease learning rate after how many epochs')
    p.add_argument('--pretrained', default='model/X3D_L.pyth', type=str, help='Path to pretrained weight')
    p.add_argument('--save_dir', default='./exp', help='Directory to save the experiment results')
    p.add_argument('--resume', default=None, help='Checkpoint to resume training')
    p.add_argument('--log_file', default='train_val_log.txt', help='File that stores the training and validation logs')
    p.add_argument('--gpu_id', default=0, type=int, help='GPU ID number (ignored under torchrun: LOCAL_RANK wins)')
    # additions
    p.add_argument('--synthetic', type=int, default=0, help='use N seeded random samples per split instead of --file_root')
    p.add_argument('--no_graph', action='store_true', help='run the step eagerly instead of as a CUDA graph')
    return p


# ---------------------------------------------------------------------------------------------
# synthetic stand-ins with the item layout of data/dataset.py (SCDDataset / BDADataset)
# ---------------------------------------------------------------------------------------------
class SyntheticSCD(torch.utils.data.Dataset):
    """img (6,H,W) float, label (3,H,W) int64 = [pre class 1..C-1, post class 1..C-1, change 0/1] (train_SCD.py:205-217
    masks the class maps by the change map)."""

    def __init__(self, n: int, H: int, W: int, num_class: int, seed: int):
        self.n, self.H, self.W, self.C, self.seed = n, H, W, num_class, seed

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        pre = torch.rand(3, self.H, self.W, generator=g) * 2 - 1
        post = pre + 0.05 * torch.randn(3, self.H, self.W, generator=g)
        change = torch.zeros(self.H, self.W, dtype=torch.int64)
        s = max(4, self.H // 4)
        y = int(torch.randint(0, self.H - s + 1, (1,), generator=g))
        x = int(torch.randint(0, self.W - s + 1, (1,), generator=g))
        change[y:y + s, x:x + s] = 1
        ca = int(torch.randint(1, self.C, (1,), generator=g))
        cb = int(torch.randint(1, self.C, (1,), generator=g))
        post[:, y:y + s, x:x + s] = (cb / self.C) - pre[:, y:y + s, x:x + s]
        pre[0, y:y + s, x:x + s] = ca / self.C
        la = torch.full((self.H, self.W), ca, dtype=torch.int64)
        lb = torch.full((self.H, self.W), cb, dtype=torch.int64)
        return torch.cat([pre, post], 0), torch.stack([la, lb, change], 0)


class SyntheticBDA(torch.utils.data.Dataset):
    """img (6,H,W) float, label (2,H,W) int64 = [building 0/1, damage class 1..C-1]; the scripts use
    label_loc = label[:, 0] and label_cls = prod(label, dim=1) (train_BDA.py:117-118)."""

    def __init__(self, n: int, H: int, W: int, num_class: int, seed: int):
        self.n, self.H, self.W, self.C, self.seed = n, H, W, num_class, seed

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        pre = torch.rand(3, self.H, self.W, generator=g) * 2 - 1
        post = pre + 0.05 * torch.randn(3, self.H, self.W, generator=g)
        loc = torch.zeros(self.H, self.W, dtype=torch.int64)
        s = max(4, self.H // 4)
        y = int(torch.randint(0, self.H - s + 1, (1,), generator=g))
        x = int(torch.randint(0, self.W - s + 1, (1,), generator=g))
        loc[y:y + s, x:x + s] = 1
        dmg = int(torch.randint(1, self.C, (1,), generator=g))
        pre[:, y:y + s, x:x + s] = 0.8
        post[:, y:y + s, x:x + s] = 0.8 - 0.4 * dmg / self.C
        cls = torch.full((self.H, self.W), dmg, dtype=torch.int64)
        return torch.cat([pre, post], 0), torch.stack([loc, cls], 0)


def make_loaders(args, task: str, world: int, rank: int, datasets=None):
    """create_data_loaders (train_SCD.py:35-94 / train_BDA.py:30-91) with a DistributedSampler on the training split."""
    if datasets is None:
        if args.synthetic <= 0:
            raise RuntimeError("the reference's dataset readers (data/dataset.py) are outside this package: pass "
                               "datasets=(train, val, test), or use --synthetic N")
        cls = SyntheticSCD if task == "scd" else SyntheticBDA
        n = args.synthetic
        datasets = tuple(cls(m, args.in_height, args.in_width, args.num_class, sd)
                         for m, sd in ((n, 1), (max(1, n // 4), 2), (max(1, n // 4), 3)))
    train, val_, test = datasets
    sampler = None
    if world > 1:
        sampler = torch.utils.data.distributed.DistributedSampler(train, num_replicas=world, rank=rank, shuffle=True,
                                                                  seed=16, drop_last=False)
    kw = dict(batch_size=args.batch_size, num_workers=args.num_workers, pin_memory=True)
    train_loader = torch.utils.data.DataLoader(train, shuffle=sampler is None, sampler=sampler, drop_last=False, **kw)
    val_loader = torch.utils.data.DataLoader(val_, shuffle=False, **kw)
    test_loader = torch.utils.data.DataLoader(test, shuffle=False, **kw)
    return train_loader, val_loader, test_loader, len(train_loader)


# ---------------------------------------------------------------------------------------------
# metrics from device histograms (formulas: model/utils.py:313-430)
# ---------------------------------------------------------------------------------------------
def _cal_kappa(hist: np.ndarray) -> float:
    """model/utils.py:330-342."""
    if hist.sum() == 0:
        return 0.0
    po = np.diag(hist).sum() / hist.sum()
    pe = np.matmul(hist.sum(1), hist.sum(0).T) / hist.sum() ** 2
    return 0.0 if pe == 1 else float((po - pe) / (1 - pe))


def scd_scores_from_hist(hist: np.ndarray):
    """`SCDD_eval_all` (model/utils.py:345-377) given its accumulated hist[pred][label]: (Fscd, IoU_mean, Sek)."""
    hist = np.asarray(hist, dtype=np.float64)
    c2 = np.zeros((2, 2))
    c2[0][0] = hist[0][0]
    c2[0][1] = hist.sum(1)[0] - hist[0][0]
    c2[1][0] = hist.sum(0)[0] - hist[0][0]
    c2[1][1] = hist[1:, 1:].sum()
    hist_n0 = hist.copy()
    hist_n0[0][0] = 0
    kappa_n0 = _cal_kappa(hist_n0)
    with np.errstate(divide="ignore", invalid="ignore"):
        iu = np.diag(c2) / (c2.sum(1) + c2.sum(0) - np.diag(c2))
        pixel_sum = hist.sum()
        change_pred_sum = pixel_sum - hist.sum(1)[0].sum()
        change_label_sum = pixel_sum - hist.sum(0)[0].sum()
        sc_tp = np.diag(hist[1:, 1:]).sum()
        prec, rec = sc_tp / change_pred_sum, sc_tp / change_label_sum
    sek = (kappa_n0 * math.exp(iu[1])) / math.e
    fscd = 0.0 if not (prec > 0 and rec > 0) else 2.0 / (1.0 / prec + 1.0 / rec)      # scipy.stats.hmean of two values
    return float(fscd), float((iu[0] + iu[1]) / 2), float(sek)


def bda_scores_from_hists(cm_loc: np.ndarray, cm_cls: np.ndarray):
    """train_BDA.py:142-147 from the two `Evaluator` matrices (hist[gt][pred]): (loc F1, harmonic mean of the damage
    F1s, overall F1, per-class damage F1s)."""
    cm_loc = np.asarray(cm_loc, dtype=np.float64)
    cm_cls = np.asarray(cm_cls, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        pre = cm_loc[1, 1] / (cm_loc[0, 1] + cm_loc[1, 1])
        rec = cm_loc[1, 1] / (cm_loc[1, 0] + cm_loc[1, 1])
        loc_f1 = 2 * rec * pre / (rec + pre)
        tps = np.diag(cm_cls)[1:]
        fns = cm_cls.sum(1)[1:] - tps
        fps = cm_cls.sum(0)[1:] - tps
        precisions = tps / (tps + fps + 1e-7)
        recalls = tps / (tps + fns + 1e-7)
        dmg = 2 * (precisions * recalls) / (precisions + recalls + 1e-7)
        harm = len(dmg) / np.sum(1.0 / dmg)
    return float(loc_f1), float(harm), float(0.3 * loc_f1 + 0.7 * harm), dmg


class _ScdMeter:
    """Per-image accuracy average (train_SCD.py:160-170, `accuracy` counts every pixel: labels are >= 0) and the
    `SCDD_eval_all` histogram, both on the device."""

    def __init__(self, num_class: int, dev):
        self.hist = torch.zeros(num_class, num_class, dtype=torch.int64, device=dev)     # [pred][label]
        self.acc_sum = torch.zeros((), dtype=torch.float64, device=dev)
        self.n_img = 0
        self.num_class = num_class

    def update(self, pre_mask, post_mask, change_mask, pre_label, post_label, with_hist: bool):
        chg = (change_mask > 0.5).squeeze(1).long()
        pa = torch.argmax(pre_mask, dim=1) * chg
        pb = torch.argmax(post_mask, dim=1) * chg
        px = float(pa[0].numel())
        acc = ((pa == pre_label).flatten(1).sum(1).double() / (px + 1e-10) +
               (pb == post_label).flatten(1).sum(1).double() / (px + 1e-10)) * 0.5
        self.acc_sum += acc.sum()
        self.n_img += pa.shape[0]
        if with_hist:
            # c3d_confusion_matrix accumulates hist[first][second]; SCDD_eval_all indexes [pred][label]
            losses.confusion_matrix(pa, pre_label, self.num_class, self.hist)
            losses.confusion_matrix(pb, post_label, self.num_class, self.hist)
        return acc.mean()

    def average(self) -> float:
        return float(self.acc_sum.item()) / max(1, self.n_img)


def _scd_batch(batch):
    imgs, labels = batch
    pre = imgs[:, 0:3].float().contiguous()
    post = imgs[:, 3:6].float().contiguous()
    return pre, post, labels[:, 0].long().contiguous(), labels[:, 1].long().contiguous(), labels[:, 2].long().contiguous()


def _bda_batch(batch):
    img, label = batch
    pre = img[:, 0:3].float().contiguous()
    post = img[:, 3:6].float().contiguous()
    return pre, post, label[:, 0].float().contiguous(), torch.prod(label, dim=1).long().contiguous()


# ---------------------------------------------------------------------------------------------
# SCD: train_SCD.py:96-277
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def val_scd(args, val_loader, model, dev, verbose: bool = True):
    """train_SCD.py:96-178.  Returns (Fscd, IoU_mean, Sek, accuracy average, loss average)."""
    model.eval()
    sim = ChangeSimilarity()
    meter = _ScdMeter(args.num_class, dev)
    loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
    n = 0
    t0 = time.time()
    for batch in DevicePrefetcher(val_loader, dev):
        pre, post, pre_label, post_label, label_change = _scd_batch(batch)
        pre_label, post_label = pre_label * label_change, post_label * label_change
        pre_mask, post_mask, change_mask = model.update_scd(pre, post)
        binary = bce_dice_loss(change_mask, label_change.unsqueeze(1).float())
        seg = cross_entropy_2d(pre_mask, pre_label, 0) + cross_entropy_2d(post_mask, post_label, 0)
        loss_sum += binary + seg * 0.5 + sim(pre_mask[:, 1:], post_mask[:, 1:], label_change.unsqueeze(1))
        meter.update(pre_mask, post_mask, change_mask, pre_label, post_label, with_hist=True)
        n += 1
    fscd, iou_mean, sek = scd_scores_from_hist(meter.hist.cpu().numpy())
    loss_val = float(loss_sum.item()) / max(1, n)
    acc = meter.average()
    if verbose:
        print(f"{time.time() - t0:.1f}s Val loss: {loss_val:.2f} Fscd: {fscd * 100:.2f} IoU: {iou_mean * 100:.2f} "
              f"Sek: {sek * 100:.2f} Accuracy: {acc * 100:.2f}")
    return fscd, iou_mean, sek, acc, loss_val


def train_scd(args, train_loader, step: TrainStep, epoch: int, max_batches: int, cur_iter: int, dev, verbose: bool = True):
    """train_SCD.py:181-277.  Returns (loss average, accuracy average, lr).  The per-step training accuracy comes from
    the step's own head outputs (`step.outs`: static tensors of the captured graph), reduced on the device."""
    step.model.train()
    meter = _ScdMeter(args.num_class, dev)
    loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
    full, lr, n = args.batch_size, args.lr, 0
    t_epoch = time.time()
    for iter_idx, batch in enumerate(DevicePrefetcher(train_loader, dev)):
        pre, post, pre_label, post_label, label_change = _scd_batch(batch)
        lr = adjust_learning_rate(args, step.opt, epoch, iter_idx + cur_iter, max_batches)
        run = step.eager if (pre.shape[0] != full and step.use_graph) else step
        loss = run(pre, post, pre_label, post_label, label_change)
        loss_sum += loss
        n += 1
        pm, qm, cm_ = step.outs
        acc = meter.update(pm, qm, cm_, pre_label * label_change, post_label * label_change, with_hist=False)
        if verbose and (iter_idx + 1) % 5 == 0:
            done = iter_idx + 1
            res_time = (max_batches * args.max_epochs - iter_idx - cur_iter) * (time.time() - t_epoch) / done / 3600
            parts = {k: float(v.item()) for k, v in step.parts.items()}
            print(f"[epoch {epoch}] [iter {done}/{len(train_loader)} {res_time:.2f}h] "
                  f"[lr {step.opt.param_groups[0]['lr']:.6f}] "
                  f"[train seg_loss {parts.get('seg', 0.0):.4f} sim_loss {parts.get('sim', 0.0):.4f} "
                  f"bn_loss {parts.get('binary', 0.0):.4f} sum_loss {float(loss.item()):.4f} "
                  f"acc {float(acc.item()) * 100:.2f}]")
    return float(loss_sum.item()) / max(1, n), meter.average(), lr


# ---------------------------------------------------------------------------------------------
# BDA: train_BDA.py:94-222
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def val_bda(args, val_loader, model, dev, verbose: bool = True):
    """train_BDA.py:94-152.  Returns (loss of the last batch — what the script returns —, loc F1, harmonic mean F1,
    overall F1, per-class damage F1s)."""
    model.eval()
    cm_loc = torch.zeros(2, 2, dtype=torch.int64, device=dev)
    cm_cls = torch.zeros(args.num_class, args.num_class, dtype=torch.int64, device=dev)
    loss = torch.zeros((), device=dev)
    for batch in DevicePrefetcher(val_loader, dev):
        pre, post, label_loc, label_cls = _bda_batch(batch)
        pred_cls, pred_loc = model.update_bda(pre, post)
        loss = cross_entropy_2d(pred_cls, label_cls, 0) + bce_dice_loss(pred_loc, label_loc.unsqueeze(1), cm=cm_loc)
        # Evaluator(num_class).add_batch(label_cls[label_loc > 0], argmax(pred_cls)[label_loc > 0]): pixels outside
        # buildings are dropped -- mark them with an out-of-range label, which the histogram kernel skips
        gt = torch.where(label_loc > 0, label_cls, torch.full_like(label_cls, -1))
        losses.confusion_matrix(gt, torch.argmax(pred_cls, dim=1), args.num_class, cm_cls)
    loc_f1, harm, oaf1, dmg = bda_scores_from_hists(cm_loc.cpu().numpy(), cm_cls.cpu().numpy())
    if verbose:
        print(f"lofF1 is {loc_f1}, clfF1 is {harm}, oaF1 is {oaf1}, sub class F1 score is {dmg} ")
    return float(loss.item()), loc_f1, harm, oaf1, dmg


def train_bda(args, train_loader, step: TrainStep, epoch: int, max_batches: int, cur_iter: int, dev, verbose: bool = True):
    """train_BDA.py:155-222.  Returns (loss average, lr)."""
    step.model.train()
    loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
    full, lr, n = args.batch_size, args.lr, 0
    t_epoch = time.time()
    for iter_idx, batch in enumerate(DevicePrefetcher(train_loader, dev)):
        pre, post, label_loc, label_cls = _bda_batch(batch)
        lr = adjust_learning_rate(args, step.opt, epoch, iter_idx + cur_iter, max_batches)
        run = step.eager if (pre.shape[0] != full and step.use_graph) else step
        loss = run(pre, post, label_loc, label_cls)
        loss_sum += loss
        n += 1
        if verbose and (iter_idx + 1) % 5 == 0:
            done = iter_idx + 1
            res_time = (max_batches * args.max_epochs - iter_idx - cur_iter) * (time.time() - t_epoch) / done / 3600
            parts = {k: float(v.item()) for k, v in step.parts.items()}
            print(f"[epoch {epoch}] [iter {done}/{len(train_loader)} {res_time:.2f}h] "
                  f"[lr {step.opt.param_groups[0]['lr']:.6f}] [seg_loss {parts.get('seg', 0.0):.4f} "
                  f"bn_loss {parts.get('binary', 0.0):.4f} sum_loss {float(loss.item()):.4f}] ")
    return float(loss_sum.item()) / max(1, n), lr


# ---------------------------------------------------------------------------------------------
# trainValidate (train_SCD.py:279-441, train_BDA.py:224-371)
# ---------------------------------------------------------------------------------------------
def _setup_logger(args, save_path: str, task: str):
    """model/utils.py:235-277 (the SCD / BDA header rows)."""
    logger = open(osp(save_path, args.log_file), 'a+')
    logger.write("Model Configurations:\n")
    for arg, value in vars(args).items():
        logger.write(f"{arg}: {value}\n")
    logger.write('\n' + '-' * 60)
    if task == "scd":
        logger.write("\n%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s" % ('epoch', 'train_loss', 'train_acc', 'val_Fscd', 'val_IoU_mean',
                                                         'val_Sek', 'val_loss', 'val_acc'))
    else:
        logger.write("\n%s\t%s\t%s\t%s\t%s\t%s" % ('epoch', 'loss_val', 'loc_f1_score', 'harmonic_mean_f1', 'oa_f1',
                                                 'damage_f1_scores'))
    logger.flush()
    return logger


def train_validate(args, task: str, datasets=None) -> dict:
    """Returns the final test metrics on rank 0 ({} on other ranks)."""
    if task not in ("scd", "bda"):
        raise ValueError(f"runner_tasks.train_validate: unknown task {task!r} (bcd lives in change3d_b200.runner)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(args.gpu_id)))
    if not torch.cuda.is_available():
        raise RuntimeError("change3d_b200.runner_tasks needs a CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(16)
    torch.cuda.manual_seed(16)
    model = Trainer(args).to(dev).float()
    save_path = osp(args.save_dir, f"{args.dataset}_iter_{args.max_steps}_lr_{args.lr}")
    if rank == 0:
        os.makedirs(save_path, exist_ok=True)
    train_loader, _val_loader, test_loader, max_batches = make_loaders(args, task, world, rank, datasets)
    args.max_epochs = int(np.ceil(args.max_steps / max_batches))
    start_epoch, cur_iter = load_checkpoint(args, model, save_path, max_batches, dev)
    logger = _setup_logger(args, save_path, task) if rank == 0 else None
    step = TrainStep(model, lr=args.lr, use_graph=not args.no_graph, task=task)
    best_file = osp(save_path, 'best_model.pth')
    best_acc_t, best_miou, best_acc_v, best_loss, best_oa = 0.0, 0.0, 0.0, 1.0, 0.0
    result: dict = {}
    for epoch in range(start_epoch, args.max_epochs):
        if hasattr(train_loader.sampler, "set_epoch"):
            train_loader.sampler.set_epoch(epoch)
        if task == "scd":
            loss_train, acc_train, lr = train_scd(args, train_loader, step, epoch, max_batches, cur_iter, dev, rank == 0)
        else:
            loss_train, lr = train_bda(args, train_loader, step, epoch, max_batches, cur_iter, dev, rank == 0)
        cur_iter += len(train_loader)
        if epoch == 0:
            continue
        if rank == 0 and task == "scd":
            fscd, iou_mean, sek, acc_val, loss_val = val_scd(args, test_loader, model, dev)
            logger.write("\n%d\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f" % (
                epoch, loss_train, acc_train, fscd, iou_mean, sek, loss_val, acc_val))
            logger.flush()
            torch.save({'epoch': epoch + 1, 'arch': str(model), 'state_dict': _state_dict_copy(model),
                        'optimizer': step.opt.state_dict(), 'loss_train': loss_train, 'loss_val': loss_val,
                        'acc_train': acc_train, 'acc_val': acc_val, 'lr': lr}, osp(save_path, 'checkpoint.pth.tar'))
            best_acc_t = max(best_acc_t, acc_train)
            if iou_mean > best_miou or not os.path.isfile(best_file):
                best_miou, best_acc_v, best_loss = max(best_miou, iou_mean), acc_val, loss_val
                torch.save(_state_dict_copy(model), best_file)
            print(f"Epoch {epoch}: Details\nBest rec: Train acc {best_acc_t * 100:.2f}, Val mIoU {best_miou * 100:.2f} "
                  f"acc {best_acc_v * 100:.2f} loss {best_loss:.4f}")
        elif rank == 0:
            loss_val, loc_f1, harm, oaf1, dmg = val_bda(args, test_loader, model, dev)
            logger.write(("\n%d\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f\t\t" + "\t\t".join(["%.4f"] * len(dmg))) % (
                epoch, loss_val, loc_f1, harm, oaf1, *dmg))
            logger.flush()
            torch.save({'epoch': epoch + 1, 'arch': str(model), 'state_dict': _state_dict_copy(model),
                        'optimizer': step.opt.state_dict(), 'loss_train': loss_train, 'loss_val': loss_val,
                        'loc_f1_score': loc_f1, 'harmonic_mean_f1': harm, 'lr': lr}, osp(save_path, 'checkpoint.pth.tar'))
            if oaf1 > best_oa or not os.path.isfile(best_file):
                best_oa = max(best_oa, oaf1) if oaf1 == oaf1 else best_oa          # NaN scores (no damage pixels yet) never win
                torch.save(_state_dict_copy(model), best_file)
            print(f"\nEpoch No. {epoch}:\tTrain Loss = {loss_train:.4f}\tVal Loss = {loss_val:.4f}\tloc_f1_score = "
                  f"{loc_f1:.4f}\tharmonic_mean_f1 = {harm:.4f}\toaf1 = {oaf1:.4f}\tdamage_f1_score = {dmg}")
        if world > 1:
            dist.barrier()
    if rank == 0:
        if os.path.isfile(best_file):
            model.load_state_dict(torch.load(best_file, map_location=dev))
        if task == "scd":
            fscd, iou_mean, sek, acc_val, loss_val = val_scd(args, test_loader, model, dev)
            logger.write("\n%s\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f" % ('Test', fscd, iou_mean, sek, loss_val, acc_val))
            result = {"Fscd": fscd, "IoU_mean": iou_mean, "Sek": sek, "acc": acc_val, "loss": loss_val}
        else:
            loss_val, loc_f1, harm, oaf1, dmg = val_bda(args, test_loader, model, dev)
            logger.write(("\n%s\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f\t\t" + "\t\t".join(["%.4f"] * len(dmg))) % (
                'Test', loss_val, loc_f1, harm, oaf1, *dmg))
            result = {"loss": loss_val, "loc_f1": loc_f1, "harmonic_mean_f1": harm, "oa_f1": oaf1,
                      "damage_f1": [float(x) for x in dmg]}
        logger.flush()
        logger.close()
    if world > 1:
        dist.barrier()
    return result


def main(argv=None) -> None:
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("scd", "bda"):
        raise SystemExit("usage: python -m change3d_b200.runner_tasks {scd|bda} [flags of scripts/train_SCD.py / train_BDA.py]")
    task = argv.pop(0)
    args = build_parser(task).parse_args(argv)
    train_validate(args, task)
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
