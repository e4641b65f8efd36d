"""GPU: DevicePrefetcher hands out exactly the host batches, in order, also when the consumer's stream is busy
(a buffer set must not be overwritten before the consumer's queued work has read it) and for a ragged last batch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prefetcher_order_overlap_and_ragged_batch():
    from change3d_b200.input_pipeline import DevicePrefetcher
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(1)
    sizes = [8, 8, 8, 8, 8, 8, 3]
    host = [(torch.randn(n, 3, 64, 64, generator=g).pin_memory(), torch.randint(0, 5, (n, 64, 64), generator=g).pin_memory())
            for n in sizes]
    big = torch.randn(4096, 4096, device=dev)
    pf = DevicePrefetcher(host, dev)
    results = []
    for x, y in pf:
        for _ in range(3):
            big = torch.tanh(big @ big * 1e-3)          # keep the consumer's stream busy before the batch is read
        results.append((x.double().sum(), y.sum(), x.shape[0]))   # reads are queued behind the busy work
    torch.cuda.synchronize()
    assert len(results) == len(host)
    for (sx, sy, n), (hx, hy) in zip(results, host):
        assert n == hx.shape[0]
        assert abs(sx.item() - hx.double().sum().item()) < 1e-6 * hx.numel()
        assert sy.item() == hy.sum().item()
    assert pf.bytes_staged == sum(hx.numel() * 4 + hy.numel() * 8 for hx, hy in host)
    with pytest.raises(RuntimeError, match="CUDA"):
        DevicePrefetcher(host, "cpu")
