"""Oracle for the next row: the SD-2.1 first stage (AutoencoderKL) either side of the UNet.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain fp32 torch-CPU, functional, driven by a state dict with the
reference's key names.  No product code uses this row yet: the oracle and its goldens are laid down first (parity is the
first gate), the CUDA path follows.

Follows, in the reference tree:
  sgm/modules/diffusionmodules/model.py
    Normalize :54-57 (GroupNorm 32, eps 1e-6), nonlinearity :49-51 (x * sigmoid(x))
    ResnetBlock :97-149     GN+swish+conv3x3, GN+swish+conv3x3, 1x1 nin_shortcut when the width changes, no temb
    AttnBlock :160-204      GN, 1x1 q / k / v, single-head softmax(q k^T / sqrt(C)) v over the H*W sites, 1x1 proj_out, + x
    Downsample :77-94       zero pad (0,1,0,1) then conv3x3 stride 2;  Upsample :60-74  nearest x2 then conv3x3
    Encoder :487-601, Decoder :604-748     block plan and forward order
  sgm/models/autoencoder.py :440-506      quant_conv after the encoder, post_quant_conv before the decoder
  sgm/modules/distributions/distributions.py :24-41   mean | logvar split, logvar clamped to [-30, 20], sample
  sgm/models/diffusion.py :117-151        scale_factor around encode / decode
  sgm/modules/autoencoding/temporal_ae.py (SVD's decoder, svd.yaml: VideoDecoder, time_mode "conv-only")
    VideoResBlock :18-83    the 2-D ResnetBlock, then a 3-D ResBlock over (t, h, w) with kernel (3,1,1)
                            (openaimodel.py ResBlock: GroupNorm32 eps 1e-5 + SiLU + conv3d, twice, + x; no embedding),
                            merged as sigmoid(mix_factor) * temporal + (1 - sigmoid(mix_factor)) * spatial
    AE3DConv :86-110        conv_out: the 2-D conv followed by a Conv3d (3,1,1) over the frames
    VideoDecoder :293-349   Decoder with those factories; the mid attention stays the 2-D AttnBlock
Pinned by tests/golden/make_vae_goldens.py against the reference Encoder / Decoder / VideoDecoder modules.
"""
import torch
import torch.nn.functional as F

SD_VAE_CONFIG = dict(  # configs/inference/sd_2_1.yaml:49-59
    double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2,
    attn_resolutions=(), dropout=0.0)
TINY_VAE_CONFIG = dict(SD_VAE_CONFIG, ch=64, resolution=32)   # same topology at toy width (64: the stride-2 TMA view needs Cin % 64 == 0)


def _gn_swish(sd, p, x):
    h = F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)
    return h * torch.sigmoid(h)


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def resnet_block(sd, p, x):
    h = _conv(sd, p + ".conv1", _gn_swish(sd, p + ".norm1", x))
    h = _conv(sd, p + ".conv2", _gn_swish(sd, p + ".norm2", h))
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(sd, p + ".nin_shortcut", x, padding=0)
    return x + h


def attn_block(sd, p, x):
    b, c, hh, ww = x.shape
    h = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-6)
    q, k, v = (_conv(sd, f"{p}.{n}", h, padding=0).reshape(b, c, hh * ww).transpose(1, 2) for n in "qkv")
    att = torch.softmax(q @ k.transpose(1, 2) * (c ** -0.5), dim=-1) @ v
    return x + _conv(sd, p + ".proj_out", att.transpose(1, 2).reshape(b, c, hh, ww), padding=0)


def encoder_forward(sd, cfg, x, prefix="encoder"):
    """Encoder.forward (model.py:577-601): moments [B, 2*z, H/8, W/8] before quant_conv."""
    n_res, nrb = len(cfg["ch_mult"]), cfg["num_res_blocks"]
    res = cfg["resolution"]
    h = _conv(sd, f"{prefix}.conv_in", x)
    for lvl in range(n_res):
        for blk in range(nrb):
            h = resnet_block(sd, f"{prefix}.down.{lvl}.block.{blk}", h)
            if res in cfg["attn_resolutions"]:
                h = attn_block(sd, f"{prefix}.down.{lvl}.attn.{blk}", h)
        if lvl != n_res - 1:
            h = _conv(sd, f"{prefix}.down.{lvl}.downsample.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
            res //= 2
    h = resnet_block(sd, f"{prefix}.mid.block_1", h)
    h = attn_block(sd, f"{prefix}.mid.attn_1", h)
    h = resnet_block(sd, f"{prefix}.mid.block_2", h)
    return _conv(sd, f"{prefix}.conv_out", _gn_swish(sd, f"{prefix}.norm_out", h))


def _frames_first(x, t):
    """"(b t) c h w -> b c t h w" """
    bt, c, hh, ww = x.shape
    return x.reshape(bt // t, t, c, hh, ww).permute(0, 2, 1, 3, 4)


def _frames_last(x):
    b, c, t, hh, ww = x.shape
    return x.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh, ww)


def video_resnet_block(sd, p, x, timesteps):
    """VideoResBlock.forward (temporal_ae.py:62-83)."""
    x = resnet_block(sd, p, x)
    v = _frames_first(x, timesteps)
    ts = p + ".time_stack"
    h = F.silu(F.group_norm(v, 32, sd[ts + ".in_layers.0.weight"], sd[ts + ".in_layers.0.bias"], eps=1e-5))
    h = F.conv3d(h, sd[ts + ".in_layers.2.weight"], sd[ts + ".in_layers.2.bias"], padding=(1, 0, 0))
    h = F.silu(F.group_norm(h, 32, sd[ts + ".out_layers.0.weight"], sd[ts + ".out_layers.0.bias"], eps=1e-5))
    h = F.conv3d(h, sd[ts + ".out_layers.3.weight"], sd[ts + ".out_layers.3.bias"], padding=(1, 0, 0))
    alpha = torch.sigmoid(sd[p + ".mix_factor"])
    return _frames_last(alpha * (v + h) + (1.0 - alpha) * v)


def decoder_forward(sd, cfg, z, prefix="decoder", timesteps=None):
    """Decoder.forward (model.py:716-748): image [B, out_ch, 8H, 8W] from the post_quant_conv output.
    ``timesteps`` (frames per clip) selects the VideoDecoder of SVD (temporal_ae.py:293-349, "conv-only")."""
    n_res, nrb = len(cfg["ch_mult"]), cfg["num_res_blocks"]
    res = cfg["resolution"] // 2 ** (n_res - 1)
    resnet = resnet_block if timesteps is None else (lambda sd_, p_, x_: video_resnet_block(sd_, p_, x_, timesteps))
    h = _conv(sd, f"{prefix}.conv_in", z)
    h = resnet(sd, f"{prefix}.mid.block_1", h)
    h = attn_block(sd, f"{prefix}.mid.attn_1", h)
    h = resnet(sd, f"{prefix}.mid.block_2", h)
    for lvl in reversed(range(n_res)):
        for blk in range(nrb + 1):
            h = resnet(sd, f"{prefix}.up.{lvl}.block.{blk}", h)
            if res in cfg["attn_resolutions"]:
                h = attn_block(sd, f"{prefix}.up.{lvl}.attn.{blk}", h)
        if lvl != 0:
            h = _conv(sd, f"{prefix}.up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
            res *= 2
    out = _conv(sd, f"{prefix}.conv_out", _gn_swish(sd, f"{prefix}.norm_out", h))
    if timesteps is not None:   # AE3DConv (temporal_ae.py:102-110)
        tm = f"{prefix}.conv_out.time_mix_conv"
        out = _frames_last(F.conv3d(_frames_first(out, timesteps), sd[tm + ".weight"], sd[tm + ".bias"], padding=(1, 0, 0)))
    return out


def encode_first_stage(sd, cfg, x, scale_factor, noise=None):
    """AutoencoderKL.encode (autoencoder.py:468-487) + DiagonalGaussianDistribution (:24-41) + scale_factor
    (diffusion.py:136-151).  ``noise``: the standard-normal draw of ``sample()``; None -> the mode."""
    moments = F.conv2d(encoder_forward(sd, cfg, x), sd["quant_conv.weight"], sd["quant_conv.bias"])
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    z = mean if noise is None else mean + torch.exp(0.5 * logvar) * noise
    return scale_factor * z


def decode_first_stage(sd, cfg, z, scale_factor, timesteps=None):
    """diffusion.py:117-134 + AutoencoderKL.decode (autoencoder.py:489-506).  SVD's autoencoder
    (AutoencodingEngine with a VideoDecoder, svd.yaml) has no post_quant_conv: it is applied only when present."""
    z = 1.0 / scale_factor * z
    if "post_quant_conv.weight" in sd:
        z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    return decoder_forward(sd, cfg, z, timesteps=timesteps)


def video_decoder_param_shapes(cfg):
    """State-dict key -> shape of VideoDecoder(**cfg, video_kernel_size=[3,1,1]) (prefix ``decoder.``)."""
    base = {k: v for k, v in param_shapes(cfg).items() if k.startswith("decoder.")}
    out = dict(base)
    for k, shape in base.items():
        if k.endswith(".norm1.weight"):     # one entry per ResnetBlock: add its time stack
            p = k[: -len(".norm1.weight")]
            c = base[p + ".conv2.weight"][0]
            for n in ("in_layers.0", "out_layers.0"):
                out[f"{p}.time_stack.{n}.weight"], out[f"{p}.time_stack.{n}.bias"] = (c,), (c,)
            for n in ("in_layers.2", "out_layers.3"):
                out[f"{p}.time_stack.{n}.weight"], out[f"{p}.time_stack.{n}.bias"] = (c, c, 3, 1, 1), (c,)
            out[p + ".mix_factor"] = (1,)
    oc = cfg["out_ch"]
    out["decoder.conv_out.time_mix_conv.weight"], out["decoder.conv_out.time_mix_conv.bias"] = (oc, oc, 3, 1, 1), (oc,)
    return out


def param_shapes(cfg, embed_dim=4):
    """State-dict key -> shape of AutoencoderKL(ddconfig=cfg, embed_dim) restricted to encoder / decoder / quant convs."""
    out = {}
    ch, mult, nrb, zc = cfg["ch"], tuple(cfg["ch_mult"]), cfg["num_res_blocks"], cfg["z_channels"]

    def conv(p, cin, cout, k):
        out[p + ".weight"], out[p + ".bias"] = (cout, cin, k, k), (cout,)

    def norm(p, c):
        out[p + ".weight"], out[p + ".bias"] = (c,), (c,)

    def res(p, cin, cout):
        norm(p + ".norm1", cin); conv(p + ".conv1", cin, cout, 3); norm(p + ".norm2", cout); conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cin, cout, 1)

    def attn(p, c):
        norm(p + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(f"{p}.{n}", c, c, 1)

    in_mult = (1,) + mult
    conv("encoder.conv_in", cfg["in_channels"], ch, 3)
    r = cfg["resolution"]
    for lvl in range(len(mult)):
        bi, bo = ch * in_mult[lvl], ch * mult[lvl]
        for blk in range(nrb):
            res(f"encoder.down.{lvl}.block.{blk}", bi, bo)
            bi = bo
            if r in cfg["attn_resolutions"]:
                attn(f"encoder.down.{lvl}.attn.{blk}", bi)
        if lvl != len(mult) - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", bi, bi, 3)
            r //= 2
    res("encoder.mid.block_1", bi, bi); attn("encoder.mid.attn_1", bi); res("encoder.mid.block_2", bi, bi)
    norm("encoder.norm_out", bi); conv("encoder.conv_out", bi, 2 * zc if cfg["double_z"] else zc, 3)
    bi = ch * mult[-1]
    conv("decoder.conv_in", zc, bi, 3)
    res("decoder.mid.block_1", bi, bi); attn("decoder.mid.attn_1", bi); res("decoder.mid.block_2", bi, bi)
    for lvl in reversed(range(len(mult))):
        bo = ch * mult[lvl]
        for blk in range(nrb + 1):
            res(f"decoder.up.{lvl}.block.{blk}", bi, bo)
            bi = bo
            if r in cfg["attn_resolutions"]:
                attn(f"decoder.up.{lvl}.attn.{blk}", bi)
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", bi, bi, 3)
            r *= 2
    norm("decoder.norm_out", bi); conv("decoder.conv_out", bi, cfg["out_ch"], 3)
    conv("quant_conv", (1 + cfg["double_z"]) * zc, (1 + cfg["double_z"]) * embed_dim, 1)
    conv("post_quant_conv", embed_dim, zc, 1)
    return out
