"""First-stage (VAE) encoder / decoder on B200 (reference: sgm/modules/diffusionmodules/model.py).

``Encoder`` (:487-601) and ``Decoder`` (:604-748) keep the reference's constructor kwargs (the ``ddconfig`` of
sd_2_1.yaml:49-59 / svd.yaml:105-118), module tree and state-dict keys (``down.{l}.block.{b}.conv1``,
``mid.attn_1.q``, ``up.{l}.upsample.conv`` ...).  Everything runs on the kernels of the UNet: GroupNorm + swish fused
with the operand conversion, 3x3 / 1x1 convolutions as implicit-GEMM tcgen05 kernels on channels-last activations, the
nearest x2 upsampling fused into the operand of the following convolution.  Two things the UNet does not have:

  * ``Downsample`` pads (0, 1, 0, 1) and convolves with stride 2 and no padding (:77-94): the same rank-5 TMA view as
    the UNet's stride-2 convolution with the tap table shifted by one pixel (the zero fill beyond the last row / column
    is the TMA out-of-bounds fill);
  * ``AttnBlock`` (:160-204) is single-head attention over all H*W sites with head dimension C (512 at full width): two
    tensor-core GEMMs per image (q k^T, then p v) around a row-softmax kernel that writes the operand of the second.
"""
import torch
import torch.nn as nn

from .... import _lib
from .... import kernels as K
from ....linear import Split


def Normalize(in_channels, num_groups=32):
    """reference :54-57."""
    return nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


def _unsupported(what):
    raise NotImplementedError(f"{what}: not on the B200 hot path (SURVEY.md section 8f)")


class Upsample(nn.Module):
    """reference :60-74."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        if not with_conv:
            _unsupported("Upsample(with_conv=False)")
        self.with_conv = with_conv
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        return K.conv2d(K.upsample_nearest2x_split(x), self.conv)


class Downsample(nn.Module):
    """reference :77-94."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        if not with_conv:
            _unsupported("Downsample(with_conv=False)")
        self.with_conv = with_conv
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def forward(self, x):
        return K.conv2d_pad_after(K.split(K.nhwc(x)), self.conv)


class ResnetBlock(nn.Module):
    """reference :97-149 (the autoencoder builds it with temb_channels=0)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        if conv_shortcut or temb_channels > 0:
            _unsupported("ResnetBlock(conv_shortcut / temb_channels > 0)")
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, x, temb=None):
        if temb is not None:
            _unsupported("ResnetBlock(temb)")
        same = self.in_channels == self.out_channels
        hs, raw, _ = K.group_norm_split(x, self.norm1, silu=True, want_raw=not same)    # x * sigmoid(x) == SiLU
        h = K.conv2d(hs, self.conv1)
        skip = x if same else K.conv2d(raw, self.nin_shortcut)
        hs2, _, _ = K.group_norm_split(h, self.norm2, silu=True)
        return K.conv2d(hs2, self.conv2, residual=skip)


class AttnBlock(nn.Module):
    """reference :160-204."""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, x, **kwargs):
        b, c, h, w = x.shape
        n = h * w
        hs, _, x_nhwc = K.group_norm_split(x, self.norm, silu=False)
        tok = hs.reshape(b, n, c)
        lin = lambda conv: K.linear(tok, conv.weight.view(c, c), conv.bias, want_f32=True)[0]   # 1x1 conv == Linear per site
        q, k, v = lin(self.q), lin(self.k), lin(self.v)
        out = torch.empty((b, n, c), dtype=torch.float32, device=x.device)
        for i in range(b):   # one head of dimension C over the n sites of an image: q k^T and p v as tensor-core GEMMs
            out[i] = K.single_head_attention(q[i], k[i], v[i], float(c) ** -0.5)
        o, _ = K.linear(K.split(out), self.proj_out.weight.view(c, c), self.proj_out.bias,
                        residual=x_nhwc.reshape(b, n, c), want_f32=True)
        return K.as_nchw(o.reshape(b, h, w, c))


def make_attn(in_channels, attn_type="vanilla", attn_kwargs=None):
    """reference :277-306: every attention type computes the same function; ``vanilla`` / ``vanilla-xformers`` map here."""
    if attn_type == "none":
        return nn.Identity()
    if attn_type not in ("vanilla", "vanilla-xformers"):
        _unsupported(f"make_attn({attn_type!r})")
    return AttnBlock(in_channels)


class Encoder(nn.Module):
    """reference :487-601."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, use_linear_attn=False,
                 attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        self.ch = ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.conv_in = nn.Conv2d(in_channels, ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.in_ch_mult = in_ch_mult
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = curr_res // 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        h = K.conv2d(K.image_split(x.float()), self.conv_in)
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block](h)
                if len(self.down[i_level].attn) > 0:
                    h = self.down[i_level].attn[i_block](h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample(h)
        h = self.mid.block_1(h)
        h = self.mid.attn_1(h)
        h = self.mid.block_2(h)
        hs, _, _ = K.group_norm_split(h, self.norm_out, silu=True)
        return K.conv2d(hs, self.conv_out)


class Decoder(nn.Module):
    """reference :604-748.  ``_make_attn`` / ``_make_resblock`` / ``_make_conv`` are the factory hooks VideoDecoder
    overrides (:637-639, temporal_ae.py:324-349)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if give_pre_end or tanh_out:
            _unsupported("Decoder(give_pre_end / tanh_out)")
        self.ch = ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.out_ch = out_ch
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        make_attn_cls, make_resblock_cls, make_conv_cls = self._make_attn(), self._make_resblock(), self._make_conv()
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = make_resblock_cls(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.mid.attn_1 = make_attn_cls(block_in, attn_type=attn_type)
        self.mid.block_2 = make_resblock_cls(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(make_resblock_cls(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn_cls(block_in, attn_type=attn_type))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = make_conv_cls(block_in, out_ch, kernel_size=3, stride=1, padding=1)

    def _make_attn(self):
        return make_attn

    def _make_resblock(self):
        return ResnetBlock

    def _make_conv(self):
        return nn.Conv2d

    def get_last_layer(self, **kwargs):
        return self.conv_out.weight

    def _conv_out(self, hs, **kwargs):
        return K.conv2d(hs, self.conv_out)

    def forward(self, z, **kwargs):
        self.last_z_shape = z.shape
        h = K.conv2d(K.image_split(z.float()), self.conv_in)
        h = self.mid.block_1(h, None, **kwargs)
        h = self.mid.attn_1(h, **kwargs)
        h = self.mid.block_2(h, None, **kwargs)
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block](h, None, **kwargs)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block](h, **kwargs)
            if i_level != 0:
                h = self.up[i_level].upsample(h)
        hs, _, _ = K.group_norm_split(h, self.norm_out, silu=True)
        return self._conv_out(hs, **kwargs)
