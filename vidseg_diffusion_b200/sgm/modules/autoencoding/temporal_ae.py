"""SVD's video decoder on B200 (reference: sgm/modules/autoencoding/temporal_ae.py, svd.yaml:119-133).

``VideoDecoder`` (:293-349, ``time_mode="conv-only"``) is the image ``Decoder`` whose ResnetBlocks are followed by a
(3,1,1) ResBlock over the frames of the clip, mixed in with ``sigmoid(mix_factor)`` (``VideoResBlock`` :18-83), and
whose output convolution is followed by a (3,1,1) convolution over frames (``AE3DConv`` :86-110).  As in the UNet the
activations stay in the frame-major channels-last layout: the temporal convolutions shift the frame coordinate of the
TMA box instead of rearranging '(b t) c h w -> b c t h w', GroupNorm statistics run over whole clips, and the mix with
the spatial branch is the epilogue of the second temporal convolution.
"""
import torch
import torch.nn as nn

from .... import _lib
from .... import kernels as K
from ..diffusionmodules.model import Decoder, ResnetBlock, _unsupported, make_attn
from ..diffusionmodules.openaimodel import ResBlock


class VideoResBlock(ResnetBlock):
    """reference :18-83."""

    def __init__(self, out_channels, *args, dropout=0.0, video_kernel_size=3, alpha=0.0, merge_strategy="learned", **kwargs):
        super().__init__(out_channels=out_channels, dropout=dropout, *args, **kwargs)
        if video_kernel_size is None:
            video_kernel_size = [3, 1, 1]
        self.time_stack = ResBlock(channels=out_channels, emb_channels=0, dropout=dropout, dims=3, use_scale_shift_norm=False,
                                   use_conv=False, up=False, down=False, kernel_size=video_kernel_size, use_checkpoint=False,
                                   skip_t_emb=True)
        self.merge_strategy = merge_strategy
        if merge_strategy == "fixed":
            self.register_buffer("mix_factor", torch.Tensor([alpha]))
        elif merge_strategy == "learned":
            self.register_parameter("mix_factor", nn.Parameter(torch.Tensor([alpha])))
        else:
            raise ValueError(f"unknown merge strategy {self.merge_strategy}")

    def get_alpha(self, bs):
        if self.merge_strategy == "fixed":
            return self.mix_factor
        return torch.sigmoid(self.mix_factor)

    def forward(self, x, temb=None, skip_video=False, timesteps=None):
        if timesteps is None:
            timesteps = self.timesteps
        x = super().forward(x, temb)
        if skip_video:
            return x
        bt = x.shape[0]
        videos = bt // timesteps
        if videos * timesteps != bt:
            raise _lib.VidsegError(f"VideoResBlock: batch {bt} is not a multiple of {timesteps} frames")
        ts = self.time_stack
        # x = alpha * time_stack(x) + (1 - alpha) * x: the kernel's blend weights its `blend` input, so it gets 1 - alpha
        keep = (1.0 - self.get_alpha(videos).float()).reshape(1).expand(bt).contiguous()
        hs, _, _ = K.group_norm_split(x, ts.in_layers[0], silu=True, samples=videos)
        h = K.conv_temporal(hs, ts.in_layers[2], videos, timesteps)
        hs2, _, _ = K.group_norm_split(h, ts.out_layers[0], silu=True, samples=videos)
        return K.conv_temporal(hs2, ts.out_layers[3], videos, timesteps, residual=x, blend=x, blend_alpha=keep)


class AE3DConv(nn.Conv2d):
    """reference :86-110."""

    def __init__(self, in_channels, out_channels, video_kernel_size=3, *args, **kwargs):
        super().__init__(in_channels, out_channels, *args, **kwargs)
        if isinstance(video_kernel_size, (list, tuple)):
            padding = [int(k // 2) for k in video_kernel_size]
        else:
            padding = int(video_kernel_size // 2)
        self.time_mix_conv = nn.Conv3d(in_channels=out_channels, out_channels=out_channels, kernel_size=video_kernel_size,
                                       padding=padding)

    def forward_split(self, hs, timesteps, skip_video=False):
        """hs: Split [(b t), H, W, Cin] (the normalised activation).  Returns fp32 [(b t), Cout, H, W]."""
        x = K.conv2d(hs, self)
        if skip_video:
            return x
        bt = x.shape[0]
        return K.conv_temporal(K.image_split(x), self.time_mix_conv, bt // timesteps, timesteps)


class VideoDecoder(Decoder):
    """reference :293-349."""
    available_time_modes = ["all", "conv-only", "attn-only"]

    def __init__(self, *args, video_kernel_size=3, alpha=0.0, merge_strategy="learned", time_mode="conv-only", **kwargs):
        self.video_kernel_size = video_kernel_size
        self.alpha = alpha
        self.merge_strategy = merge_strategy
        self.time_mode = time_mode
        assert self.time_mode in self.available_time_modes, f"time_mode parameter has to be in {self.available_time_modes}"
        if time_mode != "conv-only":
            _unsupported(f"VideoDecoder(time_mode={time_mode!r}) (svd.yaml uses the default 'conv-only')")
        super().__init__(*args, **kwargs)

    def get_last_layer(self, skip_time_mix=False, **kwargs):
        return self.conv_out.weight if skip_time_mix else self.conv_out.time_mix_conv.weight

    def _make_attn(self):
        return make_attn

    def _make_conv(self):
        return lambda *a, **k: AE3DConv(*a, video_kernel_size=self.video_kernel_size, **k)

    def _make_resblock(self):
        return lambda *a, **k: VideoResBlock(*a, video_kernel_size=self.video_kernel_size, alpha=self.alpha,
                                             merge_strategy=self.merge_strategy, **k)

    def _conv_out(self, hs, timesteps=None, skip_video=False, **kwargs):
        return self.conv_out.forward_split(hs, timesteps, skip_video=skip_video)
