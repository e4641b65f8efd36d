"""Helpers for the `-m gpu` parity tests: build the product model from the oracle's synthetic
state dict, compare tensors, and log error margins to gpurun_out/gpu_test_report.txt."""
import argparse
import contextlib
import io
import os

import numpy as np
import torch

from oracle import change3d_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "gpu_test_report.txt")

TASK_P = {"bcd": 1, "bda": 2, "scd": 3}
TASK_DATASET = {"bcd": "LEVIR-CD", "bda": "xBD", "scd": "SECOND"}


def log(msg: str) -> None:
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(msg + "\n")


def rel_err(got, ref) -> float:
    got = got.detach().double().cpu() if torch.is_tensor(got) else torch.as_tensor(np.asarray(got)).double()
    ref = ref.detach().double().cpu() if torch.is_tensor(ref) else torch.as_tensor(np.asarray(ref)).double()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    denom = ref.abs().max().item() + 1e-30
    return (got - ref).abs().max().item() / denom


def check(name: str, got, ref, tol: float) -> None:
    e = rel_err(got, ref)
    log(f"{name}: max-abs-err/max-abs-ref = {e:.3e} (tol {tol:.1e})")
    assert e < tol, f"{name}: {e:.3e} >= {tol:.1e}"


def build_trainer(task: str, H: int, W: int, num_class: int, sd=None, device="cuda"):
    from change3d_b200.model.trainer import Trainer
    args = argparse.Namespace(num_perception_frame=TASK_P[task], num_class=num_class, in_height=H, in_width=W,
                              dataset=TASK_DATASET[task], pretrained="/nonexistent/X3D_L.pyth")
    with contextlib.redirect_stdout(io.StringIO()):
        m = Trainer(args)
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    return m.to(device).float()


def ndhwc(x: torch.Tensor) -> torch.Tensor:
    """(B,C,T,H,W) -> dense (B,T,H,W,C)"""
    return x.permute(0, 2, 3, 4, 1).contiguous()


def pad_c(x: torch.Tensor, cs: int) -> torch.Tensor:
    """zero-pad the last (channel) dim to cs"""
    if x.shape[-1] == cs:
        return x.contiguous()
    out = torch.zeros(*x.shape[:-1], cs, dtype=x.dtype, device=x.device)
    out[..., :x.shape[-1]] = x
    return out


def check_vs_noise(name: str, got, ref64, ref32, factor: float = 4.0, floor: float = 2e-5) -> float:
    """Noise-floor criterion for long fp32 chains: `got` must be as close to the fp64 truth as torch's own
    fp32 path is, within `factor` (plus a small absolute floor, relative to max |ref64|)."""
    e_mine = rel_err(got, ref64)
    e_ref = rel_err(ref32, ref64)
    log(f"{name}: |mine-fp64| = {e_mine:.3e}, |torch_fp32-fp64| = {e_ref:.3e}")
    assert e_mine <= max(factor * e_ref, floor), f"{name}: {e_mine:.3e} vs fp32 noise {e_ref:.3e}"
    return e_mine
