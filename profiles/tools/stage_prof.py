"""Run fwd+bwd of a truncated ResStage at bench shapes (for ncu captures and quick per-kernel timing).

    python profiles/tools/stage_prof.py --stage 3 --depth 2 --H 64 --time          # per-launch CUDA-event times (us)
    C3D_TC_DBG=1 python profiles/tools/stage_prof.py --stage 1 --depth 2 --H 256   # per-warp wait/work cycle counters
    ncu --set full --import-source on -k regex:pw_gemm_tc ... python profiles/tools/stage_prof.py ...
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import change3d_oracle as O
from change3d_b200.model.x3d import create_x3d
from change3d_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--stage", type=int, default=3)
ap.add_argument("--depth", type=int, default=2)
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--H", type=int, default=64)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--time", action="store_true")
a = ap.parse_args()
sd = O.synth_state_dict(O.x3d_schema(), 21)
cin, _, cout, full_depth = O.STAGES[a.stage - 1]
net = create_x3d(input_clip_length=3, depth_factor=5.0)
net.load_state_dict(sd, strict=True)
net.blocks[a.stage].res_blocks = net.blocks[a.stage].res_blocks[:a.depth]
net = net.cuda().train()
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.relu(torch.randn(a.B, 3, a.H, a.H, cin, device="cuda", generator=g)).permute(0, 4, 1, 2, 3).requires_grad_(True)
wgt = torch.randn(a.B, 3, a.H // 2, a.H // 2, cout, device="cuda", generator=g).permute(0, 4, 1, 2, 3)
for it in range(a.iters):
    if a.time and it == a.iters - 1:
        torch.cuda.synchronize(); ops.PROF = {}
    x.grad = None
    (net.blocks[a.stage](x) * wgt).sum().backward()
torch.cuda.synchronize()
if a.time:
    prof, ops.PROF = ops.PROF, None
    for fam, recs in prof.items():
        print(fam, " ".join(f"{r[0].elapsed_time(r[1])*1e3:.0f}" for r in recs))
