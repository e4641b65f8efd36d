"""Device-agnostic / data-parallel runner with the surface of the reference's `scripts/train_BCD.py`
(SURVEY.md §8 f2): same command-line flags and defaults (:363-484), same epoch structure (`trainValidate`
:239-360: train every epoch, validate on the test split from epoch 1, poly learning rate with warm-up per
iteration, `checkpoint.pth.tar` with the reference's keys + `best_model.pth` state_dict, final test with the
best model, tab-separated log file), same `--resume` semantics (model/utils.py:205-232: model weights only,
`cur_iter = epoch * max_batches`).

What differs is where the work happens:
  * one process per GPU (`torchrun --nproc-per-node N -m change3d_b200.runner ...`), `DistributedSampler` shards the
    training set, one NCCL all-reduce of the flat gradient buffer per step (train_step.FlatAdam); rank 0 logs,
    validates and writes checkpoints; no hard-coded `.cuda()` / `CUDA_VISIBLE_DEVICES` (`--gpu_id` is honoured
    only when not launched by torchrun);
  * batches are staged by `input_pipeline.DevicePrefetcher` (pinned memory, copy stream, one batch ahead) instead of
    synchronous `.cuda()` calls on the compute stream;
  * the step is the CUDA-graph-captured `BCDTrainStep`; loss and the 2x2 confusion matrix stay on the device
    (losses.bce_dice_loss(cm=...)) and are read back when something is printed (every 5 iterations, like the
    reference's print cadence) and at the end of the epoch — not `loss.item()` + a 16.8 MB mask copy per step.

Datasets are outside the hot path: anything yielding `(img (6,H,W) float, target (1,H,W) float)`
like the reference's `BCDDataset` plugs in through `make_loaders`; `--synthetic N` uses N seeded random pairs
per split so the runner can be exercised without a dataset tree.
"""
from __future__ import annotations

import os
import time
from argparse import ArgumentParser
from os.path import join as osp

import numpy as np
import torch
import torch.distributed as dist

from .input_pipeline import DevicePrefetcher, augment_raw_batch, is_raw_batch
from .losses import bce_dice_loss
from .metrics import cm2score
from .model.trainer import Trainer
from .model.utils import adjust_learning_rate
from .train_step import BCDTrainStep


def build_parser() -> ArgumentParser:
    """The flags of scripts/train_BCD.py:363-484, same names / types / defaults, plus --synthetic and --no_graph."""
    p = ArgumentParser()
    p.add_argument('--dataset', default="LEVIR-CD", help='Dataset selection | LEVIR-CD | WHU-CD | CLCD')
    p.add_argument('--file_root', default="path/to/LEVIR-CD", help='path to the dataset directory')
    p.add_argument('--in_height', type=int, default=256, help='Height of RGB image')
    p.add_argument('--in_width', type=int, default=256, help='Width of RGB image')
    p.add_argument('--num_perception_frame', type=int, default=1, help='Number of perception frames')
    p.add_argument('--num_class', type=int, default=1, help='Number of classes')
    p.add_argument('--max_steps', type=int, default=80000, help='Max number of iterations')
    p.add_argument('--batch_size', type=int, default=16, help='Batch size (per process, as in the reference)')
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


class SyntheticBCD(torch.utils.data.Dataset):
    """Seeded stand-in with BCDDataset's item layout (data/dataset.py: img (6,H,W) = [pre, post] normalised to
    roughly [-1, 1], target (1,H,W) in {0,1}); the 'change' is a bright square so that there is something to learn."""

    def __init__(self, n: int, H: int, W: int, seed: int):
        self.n, self.H, self.W, self.seed = n, H, W, seed

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        pre = torch.rand(3, self.H, self.W, generator=g) * 2 - 1
        post = pre + 0.05 * torch.randn(3, self.H, self.W, generator=g)
        target = torch.zeros(1, self.H, self.W)
        s = max(4, self.H // 4)
        y = int(torch.randint(0, self.H - s + 1, (1,), generator=g))
        x = int(torch.randint(0, self.W - s + 1, (1,), generator=g))
        post[:, y:y + s, x:x + s] = 1.0 - pre[:, y:y + s, x:x + s]
        target[:, y:y + s, x:x + s] = 1.0
        return torch.cat([pre, post], 0), target


def make_loaders(args, world: int, rank: int, datasets=None):
    """create_data_loaders (scripts/train_BCD.py:28-88) with a DistributedSampler on the training split.
    `datasets` = (train, val, test) overrides --synthetic.  Returns (train_loader, val_loader, test_loader, max_batches)."""
    if datasets is None:
        if args.synthetic <= 0:
            raise RuntimeError("the reference's dataset readers (data/dataset.py) are outside this package: pass "
                               "datasets=(train, val, test) to make_loaders / train_validate, or use --synthetic N")
        datasets = (SyntheticBCD(args.synthetic, args.in_height, args.in_width, 1),
                    SyntheticBCD(max(1, args.synthetic // 4), args.in_height, args.in_width, 2),
                    SyntheticBCD(max(1, args.synthetic // 4), args.in_height, args.in_width, 3))
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


def _split(batch, args=None, train: bool = False):
    """(img (B,6,H,W), target) already on the device -> pre, post, target as the scripts slice them
    (scripts/train_BCD.py:182-185), contiguous fp32.  Raw uint8 batches (img (B,Hs,Ws,6), label (B,Hs,Ws)) go through the
    GPU transform chain instead (input_pipeline.GpuAugment: data/transforms.py on the device; 4x fewer H2D bytes)."""
    if args is not None and is_raw_batch(batch):
        return augment_raw_batch(batch, "bcd", args.in_height, args.in_width, train)
    img, target = batch
    return img[:, 0:3].float().contiguous(), img[:, 3:6].float().contiguous(), target.float()


@torch.no_grad()
def val(args, val_loader, model, epoch, dev):
    """scripts/train_BCD.py:91-154: eval-mode forward, BCEDiceLoss, confusion matrix of `output > 0.5` — loss sum and
    matrix accumulated on the device, one read-back at the end.  Returns (average loss, cm2score dict)."""
    model.eval()
    cm = torch.zeros(2, 2, dtype=torch.int64, device=dev)
    loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
    n = 0
    for batch in DevicePrefetcher(val_loader, dev):          # pinned batches staged one ahead on a copy stream
        pre, post, target = _split(batch, args, train=False)
        output = model.update_bcd(pre, post)
        loss_sum += bce_dice_loss(output, target, cm=cm)
        n += 1
    return float(loss_sum.item()) / max(1, n), cm2score(cm.cpu().numpy())


def train(args, train_loader, step: BCDTrainStep, epoch: int, max_batches: int, cur_iter: int = 0,
          lr_factor: float = 1.0, dev=None, verbose: bool = True):
    """scripts/train_BCD.py:157-236.  Returns (average loss, cm2score of the epoch's training predictions, lr)."""
    step.model.train()
    step.cm.zero_()
    loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
    full = args.batch_size
    lr = args.lr
    n = 0
    t_epoch = time.time()
    for iter_idx, batch in enumerate(DevicePrefetcher(train_loader, dev)):
        pre, post, target = _split(batch, args, train=True)
        lr = adjust_learning_rate(args, step.opt, epoch, iter_idx + cur_iter, max_batches, lr_factor=lr_factor)
        if pre.shape[0] != full and step.use_graph:
            # ragged last batch (drop_last=False in the reference): static-shape graph does not apply
            loss = step.eager(pre, post, target)
        else:
            loss = step(pre, post, target)
        loss_sum += loss
        n += 1
        if verbose and (iter_idx + 1) % 5 == 0:
            done = iter_idx + 1
            res_time = (max_batches * args.max_epochs - iter_idx - cur_iter) * (time.time() - t_epoch) / done / 3600
            print(f"[epoch {epoch}] [iter {done}/{len(train_loader)} {res_time:.2f}h] "
                  f"[lr {step.opt.param_groups[0]['lr']:.6f}] [bn_loss {loss.item():.4f}] ")
    # data parallel: every rank saw 1/world of the epoch -- reduce the loss sum, the step count and the training confusion
    # matrix so that rank 0 logs / checkpoints whole-epoch numbers (the reference is single-process)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        tot = torch.stack([loss_sum.double(), torch.tensor(float(n), dtype=torch.float64, device=dev)])
        dist.all_reduce(tot)
        dist.all_reduce(step.cm)
        return float(tot[0].item()) / max(1.0, float(tot[1].item())), step.scores(), lr
    return float(loss_sum.item()) / max(1, n), step.scores(), lr


def load_checkpoint(args, model, save_path: str, max_batches: int, dev):
    """model/utils.py:205-232: `--resume <anything>` loads save_path/checkpoint.pth.tar (model weights only)."""
    start_epoch, cur_iter = 0, 0
    if args.resume is not None:
        path = osp(save_path, 'checkpoint.pth.tar')
        if os.path.isfile(path):
            print(f"=> loading checkpoint '{path}'")
            ck = torch.load(path, map_location=dev, weights_only=False)
            start_epoch = ck['epoch']
            cur_iter = start_epoch * max_batches
            model.load_state_dict(ck['state_dict'])
            print(f"=> loaded checkpoint '{path}' (epoch {ck['epoch']})")
        else:
            print(f"=> no checkpoint found at '{path}'")
    return start_epoch, cur_iter


def setup_logger(args, save_path: str):
    """model/utils.py:235-262 (BCD columns)."""
    logger = open(osp(save_path, args.log_file), 'a+')
    logger.write("Model Configurations:\n")
    for arg, value in vars(args).items():
        logger.write(f"{arg}: {value}\n")
    logger.write('\n' + '-' * 60)
    logger.write("\n%s\t%s\t%s\t%s\t%s\t%s" % ('Epoch', 'Kappa (val)', 'IoU (val)', 'F1 (val)', 'R (val)', 'P (val)'))
    return logger


def _state_dict_copy(model) -> dict:
    # parameters are views of the flat Adam buffer: save independent tensors, reference key schema
    # (BatchNorm running statistics are this rank's: per-rank statistics are the reference's semantics for plain
    # nn.BatchNorm3d, and rank 0's buffers are what DDP's broadcast_buffers would keep.)  The 'optimizer' entry of a
    # checkpoint is FlatAdam.state_dict() -- flat exp_avg / exp_avg_sq buffers in parameter order and one step count, not
    # torch.optim.Adam's per-parameter dict; the reference's --resume never restores it (model/utils.py:205-232).
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def train_validate(args, datasets=None) -> dict:
    """trainValidate (scripts/train_BCD.py:239-360).  Returns the final test scores (rank 0) / {} (other ranks)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(args.gpu_id)))
    if not torch.cuda.is_available():
        raise RuntimeError("change3d_b200.runner needs a CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(16)                       # identical initial weights on every rank (train_BCD.py:253-254)
    torch.cuda.manual_seed(16)
    model = Trainer(args).to(dev).float()
    save_path = osp(args.save_dir, f"{args.dataset}_iter_{args.max_steps}_lr_{args.lr}")
    if rank == 0:
        os.makedirs(save_path, exist_ok=True)
    train_loader, _val_loader, test_loader, max_batches = make_loaders(args, world, rank, datasets)
    args.max_epochs = int(np.ceil(args.max_steps / max_batches))
    start_epoch, cur_iter = load_checkpoint(args, model, save_path, max_batches, dev)
    logger = setup_logger(args, save_path) if rank == 0 else None
    step = BCDTrainStep(model, lr=args.lr, use_graph=not args.no_graph)
    max_f1, best_file, score_test = 0.0, osp(save_path, 'best_model.pth'), {}
    for epoch in range(start_epoch, args.max_epochs):
        if hasattr(train_loader.sampler, "set_epoch"):
            train_loader.sampler.set_epoch(epoch)
        loss_train, score_tr, lr = train(args, train_loader, step, epoch, max_batches, cur_iter, dev=dev,
                                         verbose=rank == 0)
        cur_iter += len(train_loader)
        if epoch == 0:                          # the reference skips validation after the first epoch (:300-302)
            continue
        if rank == 0:
            loss_val, score_val = val(args, test_loader, model, epoch, dev)
            logger.write("\n%d\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f" % (
                epoch, score_val['Kappa'], score_val['IoU'], score_val['F1'], score_val['recall'],
                score_val['precision']))
            logger.flush()
            torch.save({'epoch': epoch + 1, 'arch': str(model), 'state_dict': _state_dict_copy(model),
                        'optimizer': step.opt.state_dict(), 'loss_train': loss_train, 'loss_val': loss_val,
                        'F_train': score_tr['F1'], 'F_val': score_val['F1'], 'lr': lr},
                       osp(save_path, 'checkpoint.pth.tar'))
            if max_f1 <= score_val['F1']:
                max_f1 = score_val['F1']
                torch.save(_state_dict_copy(model), best_file)
            print(f"\nEpoch No. {epoch}:\tTrain Loss = {loss_train:.4f}\tVal Loss = {loss_val:.4f}\t"
                  f"F1(tr) = {score_tr['F1']:.4f}\tF1(val) = {score_val['F1']:.4f}")
        if world > 1:
            dist.barrier()
    if rank == 0:
        if os.path.isfile(best_file):
            model.load_state_dict(torch.load(best_file, map_location=dev))
        _loss_test, score_test = val(args, test_loader, model, 0, dev)
        print(f"\nTest:\t Kappa (te) = {score_test['Kappa']:.4f}\t IoU (te) = {score_test['IoU']:.4f}\t"
              f"F1 (te) = {score_test['F1']:.4f}\t R (te) = {score_test['recall']:.4f}\t"
              f"P (te) = {score_test['precision']:.4f}")
        logger.write("\n%s\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f\t\t%.4f" % (
            'Test', score_test['Kappa'], score_test['IoU'], score_test['F1'], score_test['recall'],
            score_test['precision']))
        logger.flush()
        logger.close()
    if world > 1:
        dist.barrier()
    return score_test


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    train_validate(args)
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
