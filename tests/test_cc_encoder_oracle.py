"""CPU: pins the oracle's change-captioning feature path (Encoder.forward(output_final=True): x3d.blocks[0..4],
frame P — model/trainer.py:120-124,143-167) against tests/golden/cc_enc_b2_32.npz, produced
by the reference's own Encoder (oracle/make_golden_cc.py)."""
import os

import numpy as np

from oracle import change3d_oracle as O
from oracle.make_golden_cc import B, GRAD_KEYS, H, SEED, W, weights


def test_cc_feature_path_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "cc_enc_b2_32.npz"))
    sd = O.clone_sd(O.synth_state_dict(O.trainer_schema("bcd", 1, H, W, 1), SEED), requires_grad=True)
    pre, post, _ = O.synth_inputs(B, H, W, SEED)
    out = O.encoder_forward(sd, pre, post, 1, True, output_final=True)
    assert tuple(out.shape) == (B, 192, H // 16, W // 16)
    (out * weights(SEED)).sum().backward()
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=1e-4, atol=1e-5)
    k = "x3d.blocks.4.res_blocks.14.branch2.norm_c.running_mean"
    np.testing.assert_allclose(sd["encoder." + k].detach().numpy(), g["stat:" + k], rtol=1e-3, atol=1e-5)
    for k in GRAD_KEYS:
        got, ref = sd["encoder." + k].grad.numpy(), g["grad:" + k]
        err = np.abs(got - ref).max() / (np.abs(ref).max() + 1e-30)
        assert err < 1e-4, (k, err)          # same fp32 torch ops in the same order: ~1e-6 here
