"""CPU: host logic of change3d_b200.runner_cc (the scripts/train_CC.py mirror) — flag surface, synthetic dataset layout,
corpus BLEU against hand-computed values."""
import math
import os
import re

import pytest
import torch

from change3d_b200 import runner_cc as RC

# scripts/train_CC.py:533-684 (name -> default), written down from the reference
FLAGS = {"file_root": "path/to/LEVIR-CC", "dataset": "LEVIR_CC_5_cap_per_img_5_min_word_freq", "n_head": 8, "n_layer": 3,
         "decoder_n_layers": 1, "embed_dim": 192, "dropout": 0.1, "num_perception_frame": 1, "in_height": 256,
         "in_width": 256, "epochs": 200, "batch_size": 32, "print_freq": 100, "workers": 1, "encoder_lr": 1e-4,
         "decoder_lr": 1e-4, "grad_clip": 5., "fine_tune_encoder": True, "checkpoint": None,
         "pretrained": "model/X3D_L.pyth", "gpu_id": 0, "Split": "TEST", "beam_size": 1, "save_dir": "./exp"}


def test_flag_surface_matches_reference_script():
    args = RC.build_parser().parse_args([])
    for k, v in FLAGS.items():
        assert getattr(args, k) == v, k
    assert set(vars(args)) - set(FLAGS) == {"synthetic", "vocab_size", "no_graph"}
    ref = "/root/reference/scripts/train_CC.py"
    if os.path.isfile(ref):
        assert set(re.findall(r"'--(\w+)'", open(ref).read())) == set(FLAGS)


def test_synthetic_dataset_layout():
    ds = RC.SyntheticCC(3, 32, 32, 20, "TRAIN", seed=1)
    assert len(ds) == 15
    pair, cap, ln = ds[7]
    ids = RC.special_ids(20)
    assert pair.shape == (2, 3, 32, 32) and cap.shape == (52,) and cap.dtype == torch.int64 and ln.shape == (1,)
    assert int(cap[0]) == ids["<start>"] and int(cap[int(ln) - 1]) == ids["<end>"] and int(cap[int(ln):].sum()) == 0
    pair2, cap2, ln2, allcaps = RC.SyntheticCC(3, 32, 32, 20, "TEST", seed=1)[7]
    assert allcaps.shape == (5, 52) and torch.equal(allcaps[7 % 5], cap2) and torch.equal(pair, pair2)


def test_bleu_hand_cases():
    refs = [[[1, 2, 3, 4, 5]], [[7, 8, 9]]]
    assert all(abs(b - 1.0) < 1e-6 for b in RC.bleu_scores(refs, [[1, 2, 3, 4, 5], [7, 8, 9]], 3))
    # one substitution in a 5-word caption, single pair: p1 = 4/5, p2 = 2/4, p3 = 0/3, no brevity penalty
    b = RC.bleu_scores([[[1, 2, 3, 4, 5]]], [[1, 2, 9, 4, 5]], 2)
    assert abs(b[0] - 0.8) < 1e-6 and abs(b[1] - math.sqrt(0.8 * 0.5)) < 1e-6
    # brevity penalty with the closest reference length: hypothesis of 3 words, references of 4 and 9 words
    b = RC.bleu_scores([[[1, 2, 3, 4], [1, 2, 3, 4, 5, 6, 7, 8, 9]]], [[1, 2, 3]], 1)
    assert abs(b[0] - math.exp(1 - 4 / 3)) < 1e-6
    # clipping: repeated word counted at most as often as in a reference
    b = RC.bleu_scores([[[1, 2]]], [[1, 1, 1, 1]], 1)
    assert abs(b[0] - 0.25) < 1e-6


def test_runner_refuses_to_run_without_cuda(tmp_path):
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    args = RC.build_parser().parse_args(["--synthetic", "4", "--vocab_size", "20", "--save_dir", str(tmp_path)])
    with pytest.raises(RuntimeError, match="no CPU"):
        RC.train_validate(args)
