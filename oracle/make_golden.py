"""Generate tests/golden/*.npz by running the UNMODIFIED reference model files (through
oracle/pv_shim) on deterministic synthetic weights and inputs — TEST INFRASTRUCTURE.

Run in the authoring container:  python oracle/make_golden.py
The weights/inputs are functions of (seed, key) only (oracle.change3d_oracle.synth_*), so the
fixtures hold only reference OUTPUTS: predictions, per-level perception features, the loss,
selected gradients and updated BN running statistics.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import change3d_oracle as O   # noqa: E402
from oracle import reference_loader as R  # noqa: E402

CASES = [
    # name, task, B, H, W, num_class, seed
    ("bcd_b2_64", "bcd", 2, 64, 64, 1, 16),
    ("bda_b1_32", "bda", 1, 32, 32, 5, 17),
    ("scd_b1_32", "scd", 1, 32, 32, 7, 18),
]

GRAD_KEYS = [
    "encoder.perception_frames",
    "encoder.x3d.blocks.0.conv.conv_t.weight", "encoder.x3d.blocks.0.conv.conv_xy.weight",
    "encoder.x3d.blocks.0.norm.weight", "encoder.x3d.blocks.0.norm.bias",
    "encoder.x3d.blocks.1.res_blocks.0.branch1_conv.weight",
    "encoder.x3d.blocks.1.res_blocks.0.branch2.conv_a.weight",
    "encoder.x3d.blocks.1.res_blocks.0.branch2.conv_b.weight",
    "encoder.x3d.blocks.1.res_blocks.0.branch2.norm_b.1.block.0.weight",
    "encoder.x3d.blocks.1.res_blocks.0.branch2.norm_b.1.block.2.bias",
    "encoder.x3d.blocks.2.res_blocks.0.branch1_norm.weight",
    "encoder.x3d.blocks.2.res_blocks.3.branch2.conv_c.weight",
    "encoder.x3d.blocks.3.res_blocks.24.branch2.norm_c.bias",
    "encoder.x3d.blocks.3.res_blocks.7.branch2.conv_b.weight",
    "encoder.fc.0.0.weight", "encoder.fc.3.0.weight",
    "decoder.up_c4.0.weight", "decoder.up_c4.1.weight", "decoder.up_c4.1.bias",
    "decoder.up_c2.1.weight", "decoder.up_c1.0.weight",
]
STAT_KEYS = [
    "encoder.x3d.blocks.0.norm.running_mean", "encoder.x3d.blocks.0.norm.running_var",
    "encoder.x3d.blocks.1.res_blocks.0.branch2.norm_b.0.running_var",
    "encoder.x3d.blocks.3.res_blocks.24.branch2.norm_c.running_mean",
    "encoder.x3d.blocks.3.res_blocks.24.branch2.norm_c.running_var",
]


def run_case(name, task, B, H, W, num_class, seed):
    P = {"bcd": 1, "bda": 2, "scd": 3}[task]
    schema = O.trainer_schema(task, P, H, W, num_class)
    sd = O.synth_state_dict(schema, seed)
    pre, post, target = O.synth_inputs(B, H, W, seed)
    out = {}
    model = R.build_trainer(task, H, W, num_class, sd)
    assert [k for k, _ in schema] == list(model.state_dict().keys()), "schema order differs from the reference"

    # calibrate running statistics on the reference modules themselves (momentum 1.0, one pass);
    # oracle.calibrate_running_stats mirrors this
    bns = [m for m in model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    for m in bns:
        m.momentum = 1.0
    model.train()
    with torch.no_grad():
        {"bcd": model.update_bcd, "bda": model.update_bda, "scd": model.update_scd}[task](pre, post)
    for m in bns:
        m.momentum = 0.1
        m.num_batches_tracked.zero_()
    out["calib:encoder.x3d.blocks.3.res_blocks.24.branch2.norm_c.running_var"] = \
        model.state_dict()["encoder.x3d.blocks.3.res_blocks.24.branch2.norm_c.running_var"].numpy().copy()

    # eval forward (BN running statistics)
    model.eval()
    with torch.no_grad():
        feats = model.encoder(pre, post)
        pred = {"bcd": model.update_bcd, "bda": model.update_bda, "scd": model.update_scd}[task](pre, post)
    preds = [pred] if task == "bcd" else list(pred)
    for i, p_ in enumerate(preds):
        out[f"eval_pred{i}"] = p_.numpy()
    for lvl, fl in enumerate(feats):
        for k, f in enumerate(fl):
            out[f"eval_feat_l{lvl}_p{k}"] = f.numpy()[:, :, ::4, ::4].copy()   # subsample: keep fixtures small

    if task == "bcd":
        # train-mode forward/backward (batch-statistics BN), loss as scripts/train_BCD.py:200-201
        model.train()
        model.zero_grad()
        pred = model.update_bcd(pre, post)
        from model.utils import BCEDiceLoss  # the reference's loss, unchanged
        loss = BCEDiceLoss(pred, target)
        loss.backward()
        out["train_pred0"] = pred.detach().numpy()
        out["train_loss"] = np.array(loss.item(), dtype=np.float64)
        named = dict(model.named_parameters())
        for k in GRAD_KEYS:
            g = named[k].grad
            out["grad:" + k] = g.numpy() if g.numel() < 20000 else g.numpy().reshape(-1)[::97].copy()
        out["grad_none_count"] = np.array(sum(p.grad is None for p in model.parameters()))
        new_sd = model.state_dict()
        for k in STAT_KEYS:
            out["stat:" + k] = new_sd[k].numpy()
        # one Adam step exactly as scripts/train_BCD.py:284-290, report a few updated parameters
        opt = torch.optim.Adam(model.parameters(), 2e-4, (0.9, 0.99), eps=1e-8, weight_decay=1e-4)
        opt.step()
        for k in GRAD_KEYS[:6]:
            v = named[k].detach().numpy()
            out["adam:" + k] = v if v.size < 20000 else v.reshape(-1)[::97].copy()
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, f"{os.path.getsize(path) / 1024:.0f} KiB", "keys", len(out))


if __name__ == "__main__":
    torch.manual_seed(16)
    torch.set_num_threads(8)
    for case in CASES:
        run_case(*case)
