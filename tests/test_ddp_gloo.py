"""world_size-2 data-parallel logic on CPU (gloo): the flat gradient buffer is summed across ranks and the
mean is folded into the Adam update, so every rank ends with identical parameters equal to a single-process
step on the averaged gradient (BatchNorm statistics stay per rank, as in the reference's plain BatchNorm3d)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import change3d_oracle as O


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import argparse
        import contextlib
        import io
        from change3d_b200.model.trainer import Trainer
        from change3d_b200.train_step import FlatAdam
        torch.manual_seed(16)                                   # identical init on every rank
        a = argparse.Namespace(num_perception_frame=1, num_class=1, in_height=32, in_width=32, dataset="LEVIR-CD",
                               pretrained="/nonexistent")
        with contextlib.redirect_stdout(io.StringIO()):
            m = Trainer(a)
        opt = FlatAdam(m)
        g = torch.Generator().manual_seed(100 + rank)           # different data shard -> different gradient
        local = torch.randn(opt.numel, generator=g)
        opt.flat_g.copy_(local)
        opt.all_reduce()                                        # NCCL on the GPU box, gloo here: same call
        summed = opt.flat_g.clone()
        # the fused kernel's arithmetic restated on CPU: Adam on grad_scale * summed gradient
        p = opt.flat_p.clone()
        O.adam_reference_step([p], [summed / world], {}, lr=2e-4)
        # numpy arrays travel by value: torch tensors are shared through file descriptors that die with this process
        q.put((rank, local.numpy(), summed.numpy(), p.detach().numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_update():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, s0, p0), (_, l1, s1, p1) = [(r, *(torch.from_numpy(x) for x in xs)) for r, *xs in res]
    assert torch.allclose(s0, l0 + l1) and torch.equal(s0, s1)
    assert torch.equal(p0, p1)                                   # ranks stay bit-identical
    assert not torch.equal(l0, l1)
