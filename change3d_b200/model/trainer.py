"""B200-native Encoder / Trainer — drop-in for the reference's `model/trainer.py`.

Same constructors (`Trainer(args)`, `Encoder(args, embed_dims)`), attributes (`.encoder`, `.decoder*`),
methods (`update_bcd/scd/bda/cc`) and state-dict keys.  Differences are all underneath:
  * the [pre, P perception frames, post] clip is never materialised — the stem kernel reads the three
    sources in place (model/trainer.py:154-162);
  * `enhance` (model/trainer.py:71-108) is one in-place kernel on frame T//2 fused into the same
    autograd node as the stage that produced the tensor (the reference clones the whole 5-D tensor);
  * features stay NDHWC; the per-frame (B,C,H,W) views handed to the decoders are strided views.
"""
from typing import Any, List

import torch
import torch.nn as nn

from .. import engine
from .caption_decoder import CaptionDecoder
from .change_decoder import ChangeDecoder
from .utils import weight_init
from .x3d import create_x3d, _as_ncdhw_view, _as_ndhwc


def _merge_slice_grads(g, gs, like: torch.Tensor, P: int) -> torch.Tensor:
    """Gradient w.r.t. a published (B,C,T,H,W) feature tensor from its two consumers: `g` from the next stage (dense NDHWC
    once permuted, or None) and `gs[k]` from the decoder head that read perception frame k+1 (or None).  The frame
    gradients are added in place into the owned full-size gradient — one strided add per head instead of autograd's
    zero-filled full-size tensor + slice copy + full-size add per slice (2.4 GB of traffic at the stem's resolution)."""
    if g is None:
        g = torch.zeros(like.shape, device=like.device, dtype=torch.float32)       # (B,T,H,W,C): last stage only
    else:
        g = engine.owned_ndhwc(g)
    for k, gk in enumerate(gs):
        if gk is not None:
            g[:, k + 1].add_(gk.permute(0, 2, 3, 1))
    return g


class _StemEnhFn(torch.autograd.Function):
    """blocks[0] on [pre, perception, post] + enhance, one autograd node.  Outputs: the (B,C,T,H,W) feature tensor and,
    when `n_slices` > 0, its perception-frame slices f[:, :, k+1] as separate outputs (views of the same memory) so
    that their gradients arrive separately (see _merge_slice_grads)."""

    @staticmethod
    def forward(ctx, pre, post, perc, stem, fc_w, P, n_slices, grad_on, w_xy, w_t, gamma, beta):
        B, _, H, W = pre.shape
        pre = pre.contiguous()
        post = post.contiguous()
        percc = perc.contiguous()
        hw = H * W
        frames = [(pre, 3 * hw, hw)] + [(percc[0, :, f], 0, P * hw) for f in range(P)] + [(post, 3 * hw, hw)]
        need = grad_on and any(ctx.needs_input_grad)
        if need and not stem.training:
            raise RuntimeError("change3d_b200: backward through eval-mode BatchNorm is not implemented")
        out, saved = engine.stem_forward(stem, frames, B, H, W, stem.training, need)
        mid_pre = None
        if fc_w is not None:
            # the pre-enhance mid frame is only kept to restore the published tensor for a backward that reads it
            mid_pre = engine.enhance_forward(out, fc_w, P, need and not engine.stem_masked())
        ctx.stem, ctx.saved, ctx.frames, ctx.mid_pre, ctx.fc_w, ctx.P = stem, saved, frames, mid_pre, fc_w, P
        ctx.keep = (pre, post, percc)
        ctx.set_materialize_grads(False)
        f = _as_ncdhw_view(out)
        return (f,) + tuple(f[:, :, k + 1] for k in range(n_slices))

    @staticmethod
    def backward(ctx, g, *gs):
        y, bnp, out = ctx.saved
        g = _merge_slice_grads(g, gs, out, ctx.P)
        dfc = None
        if ctx.fc_w is not None:
            dfc = engine.enhance_backward(out, ctx.mid_pre, ctx.fc_w, ctx.P, g, restore=not engine.stem_masked())
        dperc, dwxy, dwt, dgamma, dbeta = engine.stem_backward(ctx.stem, ctx.frames, y, bnp, out, g, ctx.P)
        ctx.saved = None
        return None, None, dperc, None, dfc, None, None, None, dwxy, dwt, dgamma, dbeta


class _StageEnhFn(torch.autograd.Function):
    """blocks[i] (ResStage) + enhance, one autograd node; outputs as in _StemEnhFn."""

    @staticmethod
    def forward(ctx, x, stage, fc_w, P, grad_on, *params):
        need = grad_on and any(ctx.needs_input_grad)
        if need and not stage.training:
            raise RuntimeError("change3d_b200: backward through eval-mode BatchNorm is not implemented")
        out, saved = engine.res_stage_forward(stage, _as_ndhwc(x), stage.training, need)
        mid_pre = engine.enhance_forward(out, fc_w, P, need)
        ctx.stage, ctx.saved, ctx.mid_pre, ctx.fc_w, ctx.P, ctx.out = stage, saved, mid_pre, fc_w, P, out
        ctx.set_materialize_grads(False)
        f = _as_ncdhw_view(out)
        return (f,) + tuple(f[:, :, k + 1] for k in range(P))

    @staticmethod
    def backward(ctx, g, *gs):
        g = _merge_slice_grads(g, gs, ctx.out, ctx.P)
        dfc = engine.enhance_backward(ctx.out, ctx.mid_pre, ctx.fc_w, ctx.P, g)
        dx, grads = engine.res_stage_backward(ctx.stage, ctx.saved, g)
        ctx.saved = None
        return (_as_ncdhw_view(dx), None, dfc, None, None) + tuple(grads)


class Encoder(nn.Module):
    def __init__(self, args: Any, embed_dims: List[int]) -> None:
        super().__init__()
        self.args = args
        self.x3d = create_x3d(input_clip_length=3, depth_factor=5.0)       # model/trainer.py:40
        try:                                                               # model/trainer.py:43-48
            state_dict = torch.load(args.pretrained, map_location='cpu')['model_state']
            msg = self.x3d.load_state_dict(state_dict, strict=True)
            print(f'Load pretrained weight: {args.pretrained}, {msg}.')
        except Exception as e:
            print(f"Failed to load pretrained weights: {e}")
        self.perception_frames = nn.Parameter(
            torch.randn(1, 3, args.num_perception_frame, args.in_height, args.in_width), requires_grad=True)
        self.fc = nn.ModuleList([
            nn.Sequential(nn.Conv2d(dim, dim, kernel_size=1, stride=1, padding=0, bias=False), nn.ReLU())
            for dim in embed_dims])

    def enhance(self, x: torch.Tensor, fc: nn.Module) -> torch.Tensor:
        """Stand-alone (out-of-place, inference-only) enhance kept for API parity; training uses the
        fused path in forward()."""
        out = _as_ndhwc(x.detach()).clone()
        engine.enhance_forward(out, fc[0].weight, self.args.num_perception_frame, False)
        return _as_ncdhw_view(out)

    def forward(self, x: torch.Tensor, y: torch.Tensor, output_final: bool = False):
        if not x.is_cuda:
            raise RuntimeError("change3d_b200.Encoder: CUDA tensors required (no CPU/eager fallback)")
        P = self.args.num_perception_frame
        blocks = self.x3d.blocks
        stem = blocks[0]
        grad_on = torch.is_grad_enabled()
        stem_args = (grad_on, stem.conv.conv_t.weight, stem.conv.conv_xy.weight, stem.norm.weight, stem.norm.bias)
        if output_final:                                                   # model/trainer.py:120-124
            f = _StemEnhFn.apply(x.float(), y.float(), self.perception_frames, stem, None, P, 0, *stem_args)[0]
            for i in range(1, 5):
                f = blocks[i](f)
            return f[:, :, P]
        out = []
        res = _StemEnhFn.apply(x.float(), y.float(), self.perception_frames, stem, self.fc[0][0].weight, P, P, *stem_args)
        out.append(list(res[1:]))
        for i in range(1, 4):                                              # model/trainer.py:127-139
            res = _StageEnhFn.apply(res[0], blocks[i], self.fc[i][0].weight, P, grad_on, *blocks[i].param_list())
            out.append(list(res[1:]))
        return out


class Trainer(nn.Module):
    def __init__(self, args: Any) -> None:
        super().__init__()
        self.args = args
        self.embed_dims = [24, 24, 48, 96]
        self.encoder = Encoder(args, self.embed_dims)
        if args.num_perception_frame == 1 and 'CD' in args.dataset:        # model/trainer.py:192-195
            self.decoder = ChangeDecoder(args, in_dim=self.embed_dims, has_sigmoid=True)
            weight_init(self.decoder)
        elif args.num_perception_frame == 3:                               # model/trainer.py:198-205
            self.decoder_pre = ChangeDecoder(args, in_dim=self.embed_dims)
            self.decoder_post = ChangeDecoder(args, in_dim=self.embed_dims)
            self.decoder_change = ChangeDecoder(args, in_dim=self.embed_dims, has_sigmoid=True)
            weight_init(self.decoder_pre)
            weight_init(self.decoder_post)
            weight_init(self.decoder_change)
        elif args.num_perception_frame == 2:                               # model/trainer.py:208-213
            self.decoder_cls = ChangeDecoder(args, in_dim=self.embed_dims)
            self.decoder_loc = ChangeDecoder(args, in_dim=self.embed_dims, has_sigmoid=True)
            weight_init(self.decoder_cls)
            weight_init(self.decoder_loc)
        elif args.num_perception_frame == 1 and 'CC' in args.dataset:      # model/trainer.py:216-217
            self.decoder = CaptionDecoder(args)
        else:
            assert False

    def update_bcd(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        features = self.encoder(x, y)
        return self.decoder([lvl[0] for lvl in features])

    def update_scd(self, x: torch.Tensor, y: torch.Tensor):
        features = self.encoder(x, y)
        pre_mask = self.decoder_pre([lvl[0] for lvl in features])
        post_mask = self.decoder_post([lvl[2] for lvl in features])
        change_mask = self.decoder_change([lvl[1] for lvl in features])
        return pre_mask, post_mask, change_mask

    def update_bda(self, x: torch.Tensor, y: torch.Tensor):
        features = self.encoder(x, y)
        pred_cls = self.decoder_cls([lvl[0] for lvl in features])
        pred_loc = self.decoder_loc([lvl[1] for lvl in features])
        return pred_cls, pred_loc

    def update_cc(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return self.encoder(x, y, output_final=True)
