"""GPU: the train_CC.py mirror end to end on a tiny synthetic split — training epochs through CCTrainStep (CUDA graph +
ragged last batch), batched cached caption search in `evaluate`, BLEU bookkeeping, checkpoint keys of the script."""
import os

import pytest
import torch

from change3d_b200 import runner_cc as RC

pytestmark = pytest.mark.gpu


def test_cc_train_validate(tmp_path, capsys):
    args = RC.build_parser().parse_args(
        ["--synthetic", "3", "--vocab_size", "20", "--in_height", "64", "--in_width", "64", "--batch_size", "4",
         "--workers", "0", "--epochs", "2", "--n_layer", "1", "--print_freq", "1", "--dropout", "0.0",
         "--encoder_lr", "1e-3", "--decoder_lr", "1e-3", "--pretrained", "/nonexistent/X3D_L.pyth", "--save_dir", str(tmp_path)])
    m = RC.train_validate(args)            # 3 pairs x 5 captions = 15 items: batches 4, 4, 4, 3 (last one eager)
    assert set(m) == {"Bleu_1", "Bleu_2", "Bleu_3", "Bleu_4", "n_captions", "train_loss", "train_top1"}
    assert m["train_loss"] == m["train_loss"] and 0.0 <= m["train_top1"] <= 100.0 and 0.0 <= m["Bleu_4"] <= 1.0
    save = os.path.join(str(tmp_path), f"{args.dataset}_iter_2_lr_0.001")
    ck = torch.load(os.path.join(save, f"checkpoint_{args.dataset}.pth.tar"), map_location="cpu", weights_only=False)
    assert set(ck) == {'epoch', 'bleu-4', 'encoder_image', 'decoder', 'encoder_image_optimizer', 'decoder_optimizer'}
    assert ck['epoch'] == 1 and os.path.isfile(os.path.join(save, f"checkpoint_{args.dataset}_epoch_0.pth.tar"))
    assert "wdc.weight" in ck['decoder'] and any(k.startswith("x3d.blocks.4.") for k in ck['encoder_image'])
    out = capsys.readouterr().out
    assert "Epoch: 1/2 step: 3/4" in out and "evaluated" in out
