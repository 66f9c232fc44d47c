"""Oracle for A2/A5/A6/A8: the SVD ``VideoUNet`` single-step forward with the spatial AND temporal Q/K stash.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain fp32 torch-CPU, functional, driven by a
state dict with the reference's key names -- no nn.Module tree, no shared code with the product.
Spatial layers reuse the restatement in oracle/unet.py (they are the same reference code).

Follows, in the reference tree:
  sgm/modules/diffusionmodules/video_model.py
    VideoUNet.__init__ :92-449     same block plan as UNetModel with VideoResBlock / SpatialVideoTransformer,
                                   label_emb ("sequential": Linear-SiLU-Linear on y, :186-194)
    VideoUNet.forward  :451-566    emb = time_embed(t_emb) + label_emb(y); blocks; out
    VideoResBlock.forward :66-89   2-D ResBlock -> '(b t) c h w -> b c t h w' -> 3-D ResBlock with (3,1,1) kernels
                                   and per-(b,t) embedding (exchange_temb_dims, openaimodel.py:364-366)
                                   -> AlphaBlender 'b t -> b 1 t 1 1'
  sgm/modules/video_attention.py
    SpatialVideoTransformer.forward :378-489  time_context = context[::T] repeated h*w times (:395-404);
                                   frame-index sinusoidal embedding -> time_pos_embed MLP (:417-427); per depth:
                                   spatial block, x_mix = x + emb, temporal block, AlphaBlender (:472-476)
    VideoTransformerBlock._forward :145-285   '(b t) s c -> (b s) t c'; ff_in(norm_in) + skip; attn1(norm1) + x;
                                   attn2(norm2, time_context) + x; ff(norm3) + skip; back
  sgm/modules/diffusionmodules/util.py
    AlphaBlender :314-380          learned_with_images: alpha = where(indicator, 1, sigmoid(mix_factor));
                                   x = alpha * x_spatial + (1 - alpha) * x_temporal
  sgm/modules/attention.py         CrossAttention stash :330-331 (xformers class additionally stashes v, :446-448;
                                   the importable 'softmax' class is numerically the same path)
"""
import torch
import torch.nn.functional as F

from . import unet as ou

SVD_CONFIG = dict(  # configs/inference/svd.yaml:15-34
    in_channels=8, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1), num_res_blocks=2,
    channel_mult=(1, 2, 4, 4), num_head_channels=64, context_dim=1024, adm_in_channels=768)

TINY_VIDEO_CONFIG = dict(
    in_channels=8, out_channels=4, model_channels=64, attention_resolutions=(4, 2, 1), num_res_blocks=2,
    channel_mult=(1, 2, 4, 4), num_head_channels=64, context_dim=96, adm_in_channels=48)


def _video_layer_shapes(prefix, layer, cfg, out):
    kind, cin, cout = layer
    temb = 4 * cfg["model_channels"]
    ou._layer_shapes(prefix, layer, cfg, out)
    if kind == "res":
        ts = prefix + ".time_stack"
        out[ts + ".in_layers.0.weight"] = (cout,)
        out[ts + ".in_layers.0.bias"] = (cout,)
        out[ts + ".in_layers.2.weight"] = (cout, cout, 3, 1, 1)
        out[ts + ".in_layers.2.bias"] = (cout,)
        out[ts + ".emb_layers.1.weight"] = (cout, temb)
        out[ts + ".emb_layers.1.bias"] = (cout,)
        out[ts + ".out_layers.0.weight"] = (cout,)
        out[ts + ".out_layers.0.bias"] = (cout,)
        out[ts + ".out_layers.3.weight"] = (cout, cout, 3, 1, 1)
        out[ts + ".out_layers.3.bias"] = (cout,)
        out[prefix + ".time_mixer.mix_factor"] = (1,)
    elif kind == "attn":
        ch, ctx = cin, cfg["context_dim"]
        tb = prefix + ".time_stack.0"
        out[tb + ".norm_in.weight"] = (ch,)
        out[tb + ".norm_in.bias"] = (ch,)
        out[tb + ".ff_in.net.0.proj.weight"] = (8 * ch, ch)
        out[tb + ".ff_in.net.0.proj.bias"] = (8 * ch,)
        out[tb + ".ff_in.net.2.weight"] = (ch, 4 * ch)
        out[tb + ".ff_in.net.2.bias"] = (ch,)
        for a, kdim in (("attn1", ch), ("attn2", ctx)):
            out[f"{tb}.{a}.to_q.weight"] = (ch, ch)
            out[f"{tb}.{a}.to_k.weight"] = (ch, kdim)
            out[f"{tb}.{a}.to_v.weight"] = (ch, kdim)
            out[f"{tb}.{a}.to_out.0.weight"] = (ch, ch)
            out[f"{tb}.{a}.to_out.0.bias"] = (ch,)
        out[f"{tb}.ff.net.0.proj.weight"] = (8 * ch, ch)
        out[f"{tb}.ff.net.0.proj.bias"] = (8 * ch,)
        out[f"{tb}.ff.net.2.weight"] = (ch, 4 * ch)
        out[f"{tb}.ff.net.2.bias"] = (ch,)
        for n in ("norm1", "norm2", "norm3"):
            out[f"{tb}.{n}.weight"] = (ch,)
            out[f"{tb}.{n}.bias"] = (ch,)
        out[prefix + ".time_pos_embed.0.weight"] = (4 * ch, ch)
        out[prefix + ".time_pos_embed.0.bias"] = (4 * ch,)
        out[prefix + ".time_pos_embed.2.weight"] = (ch, 4 * ch)
        out[prefix + ".time_pos_embed.2.bias"] = (ch,)
        out[prefix + ".time_mixer.mix_factor"] = (1,)


def param_shapes(cfg):
    """{state-dict key: shape} of the reference VideoUNet built with ``cfg`` (svd.yaml options)."""
    mc = cfg["model_channels"]
    out = {"time_embed.0.weight": (4 * mc, mc), "time_embed.0.bias": (4 * mc,),
           "time_embed.2.weight": (4 * mc, 4 * mc), "time_embed.2.bias": (4 * mc,),
           "label_emb.0.0.weight": (4 * mc, cfg["adm_in_channels"]), "label_emb.0.0.bias": (4 * mc,),
           "label_emb.0.2.weight": (4 * mc, 4 * mc), "label_emb.0.2.bias": (4 * mc,)}
    inputs, middle, outputs = ou.block_plan(cfg)
    for i, layers in enumerate(inputs):
        for j, layer in enumerate(layers):
            _video_layer_shapes(f"input_blocks.{i}.{j}", layer, cfg, out)
    for j, layer in enumerate(middle):
        _video_layer_shapes(f"middle_block.{j}", layer, cfg, out)
    for i, layers in enumerate(outputs):
        for j, layer in enumerate(layers):
            _video_layer_shapes(f"output_blocks.{i}.{j}", layer, cfg, out)
    out["out.0.weight"] = (mc,)
    out["out.0.bias"] = (mc,)
    out["out.2.weight"] = (cfg["out_channels"], mc, 3, 3)
    out["out.2.bias"] = (cfg["out_channels"],)
    return out


def _alpha(sd, key, indicator):
    """AlphaBlender.get_alpha, 'learned_with_images' (util.py:357-366): [b, t]."""
    a = torch.sigmoid(sd[key])
    return torch.where(indicator.bool(), torch.ones(1, 1), a[..., None])


def _video_resblock(sd, p, x, emb, T, indicator):
    """VideoResBlock.forward (video_model.py:66-89)."""
    x = ou._resblock(sd, p, x, emb)
    bt, c, h, w = x.shape
    b = bt // T
    xs = x.reshape(b, T, c, h, w).permute(0, 2, 1, 3, 4)  # b c t h w
    ts = p + ".time_stack"
    hh = F.silu(F.group_norm(xs, 32, sd[ts + ".in_layers.0.weight"], sd[ts + ".in_layers.0.bias"], eps=1e-5))
    hh = F.conv3d(hh, sd[ts + ".in_layers.2.weight"], sd[ts + ".in_layers.2.bias"], padding=(1, 0, 0))
    e = F.linear(F.silu(emb.reshape(b, T, -1)), sd[ts + ".emb_layers.1.weight"], sd[ts + ".emb_layers.1.bias"])  # b t c
    hh = hh + e.permute(0, 2, 1)[:, :, :, None, None]
    hh = F.silu(F.group_norm(hh, 32, sd[ts + ".out_layers.0.weight"], sd[ts + ".out_layers.0.bias"], eps=1e-5))
    hh = F.conv3d(hh, sd[ts + ".out_layers.3.weight"], sd[ts + ".out_layers.3.bias"], padding=(1, 0, 0))
    xt = xs + hh
    alpha = _alpha(sd, p + ".time_mixer.mix_factor", indicator)[:, None, :, None, None]  # b 1 t 1 1
    out = alpha * xs + (1.0 - alpha) * xt
    return out.permute(0, 2, 1, 3, 4).reshape(bt, c, h, w)


def _temporal_attention(sd, p, x, context, heads, stash, tag, kind, injected=None):
    """CrossAttention.forward on the '(b s) t c' layout; stash under the reference's dump names
    ``temporal_{self|cross}_attn_{q|k}`` (svd_single_video_inference.py:121-125).  ``injected``: {key: tensor}, a key
    containing ``temporal_{kind}_attn_q`` / ``_k`` replaces the projection (video_attention.py:167-176; the SDPA class
    never injects v)."""
    ctx = x if context is None else context
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(ctx, sd[p + ".to_k.weight"])
    for key, val in (injected or {}).items():
        if f"temporal_{kind}_attn_q" in key:
            q = val
        elif f"temporal_{kind}_attn_k" in key:
            k = val
    v = F.linear(ctx, sd[p + ".to_v.weight"])
    if stash is not None:
        stash[(tag, f"temporal_{kind}_attn_q")] = q
        stash[(tag, f"temporal_{kind}_attn_k")] = k
    b, n, c = q.shape
    d = c // heads
    qh, kh, vh = (t.reshape(b, -1, heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
    w = torch.softmax((qh @ kh.transpose(-1, -2)) * (d ** -0.5), dim=-1)
    o = (w @ vh).permute(0, 2, 1, 3).reshape(b, n, c)
    return F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _geglu_ff(sd, p, x):
    hgate = F.linear(x, sd[p + ".net.0.proj.weight"], sd[p + ".net.0.proj.bias"])
    val, gate = hgate.chunk(2, dim=-1)
    return F.linear(val * F.gelu(gate), sd[p + ".net.2.weight"], sd[p + ".net.2.bias"])


def _temporal_modulation(mod, bs, T):
    """video_attention.py:197-216: out[half_hw:, i] += lambda_i * mask_i[:, None] on the '(b s) t c' layout -> [bs, T, 1]."""
    if mod is None:
        return None
    half = bs // 2
    add = torch.zeros(bs, T, 1)
    for i, mask in enumerate(mod["feature_masks"]):
        if i in mod["block_frames"] and i in mod["layer_frames"] and i in mod["timestep_frames"]:
            lam = mod["lambda_start"] if mod["schedule"] == "constant" else \
                mod["lambda_start"] + (mod["lambda_end"] - mod["lambda_start"]) * i / mod["num_frames"]
            m = torch.as_tensor(mask, dtype=torch.float32).reshape(-1, 1)
            add[half:, i] += lam * m
            if mod["uc"]:
                add[:half, i] += lam * m
    return add


def _video_transformer(sd, p, x, context, heads, T, indicator, stash, tag, mod_spatial=None, mod_temporal=None, injected=None):
    """SpatialVideoTransformer.forward (video_attention.py:378-489), depth 1, use_linear, use_spatial_context."""
    bt, c, h, w = x.shape
    b, s = bt // T, h * w
    x_in = x
    time_context = context[::T].repeat_interleave(s, dim=0)  # 'b ... -> (b n) ...', n = h*w
    x = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-6)
    x = x.permute(0, 2, 3, 1).reshape(bt, s, c)
    x = F.linear(x, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    frames = torch.arange(T).repeat(b)
    emb = F.linear(ou.timestep_embedding(frames, c), sd[p + ".time_pos_embed.0.weight"], sd[p + ".time_pos_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd[p + ".time_pos_embed.2.weight"], sd[p + ".time_pos_embed.2.bias"])[:, None, :]
    # spatial block (attention.py:609-759)
    tb = p + ".transformer_blocks.0"
    ln = lambda t, pre, n: F.layer_norm(t, (c,), sd[f"{pre}.{n}.weight"], sd[f"{pre}.{n}.bias"], eps=1e-5)
    add_s = ou._modulation(mod_spatial, bt, s)
    site_s = lambda kind: add_s if (add_s is not None and kind in mod_spatial["attn_types"]) else 0.0
    x = x + (ou._attention(sd, tb + ".attn1", ln(x, tb, "norm1"), None, heads, stash, tag, "self", injected) + site_s("self_attn"))
    x = x + (ou._attention(sd, tb + ".attn2", ln(x, tb, "norm2"), context, heads, stash, tag, "cross", injected) + site_s("cross_attn"))
    x = x + (_geglu_ff(sd, tb + ".ff", ln(x, tb, "norm3")) + site_s("ff_out"))
    # temporal block (video_attention.py:145-285)
    ts = p + ".time_stack.0"
    xm = (x + emb).reshape(b, T, s, c).permute(0, 2, 1, 3).reshape(b * s, T, c)  # '(b t) s c -> (b s) t c'
    xm = _geglu_ff(sd, ts + ".ff_in", ln(xm, ts, "norm_in")) + xm
    add_t = _temporal_modulation(mod_temporal, b * s, T)
    site_t = lambda kind: add_t if (add_t is not None and kind in mod_temporal["attn_types"]) else 0.0
    xm = (_temporal_attention(sd, ts + ".attn1", ln(xm, ts, "norm1"), None, heads, stash, tag, "self", injected)
          + site_t("self_attn")) + xm
    xm = (_temporal_attention(sd, ts + ".attn2", ln(xm, ts, "norm2"), time_context, heads, stash, tag, "cross")
          + site_t("cross_attn")) + xm
    xm = xm + (_geglu_ff(sd, ts + ".ff", ln(xm, ts, "norm3")) + site_t("ff_out"))
    xm = xm.reshape(b, s, T, c).permute(0, 2, 1, 3).reshape(bt, s, c)
    alpha = _alpha(sd, p + ".time_mixer.mix_factor", indicator).reshape(bt, 1, 1)  # 'b t -> (b t) 1 1'
    x = alpha * x + (1.0 - alpha) * xm
    x = F.linear(x, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x.reshape(bt, h, w, c).permute(0, 3, 1, 2) + x_in


def _run_block(sd, prefix, layers, h, emb, context, T, indicator, stash, tag, mods=(None, None), injected=None):
    for j, (kind, cin, cout) in enumerate(layers):
        p = f"{prefix}.{j}"
        if kind == "res":
            h = _video_resblock(sd, p, h, emb, T, indicator)
        elif kind == "attn":
            h = _video_transformer(sd, p, h, context, cout, T, indicator, stash, tag, mods[0], mods[1], injected)
        else:
            h = _plain_layer(sd, p, kind, h)
    return h


def _plain_layer(sd, p, kind, h):
    if kind == "conv":
        return F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
    if kind == "down":
        return F.conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
    if kind == "up":
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        return F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
    raise ValueError(kind)


@torch.no_grad()
def video_unet_forward(sd, cfg, x, timesteps, context, y, num_video_frames, image_only_indicator=None, stash=None,
                       modulate_params=None, injection=None):
    """VideoUNet.forward (video_model.py:451-566), inference path without modulation / injection.

    x [B=(b t), Cin, H, W]; timesteps [B]; context [B, L, Cctx]; y [B, adm]; image_only_indicator [b, t] (zeros).
    ``stash`` receives {(block_tag, "{spatial|temporal}_{self|cross}_attn_{q|k}"): tensor}; temporal tensors are in
    the reference's '(b s) t c' layout."""
    mc, T = cfg["model_channels"], num_video_frames
    b = x.shape[0] // T
    ind = torch.zeros(b, T) if image_only_indicator is None else image_only_indicator.float()
    inputs, middle, outputs = ou.block_plan(cfg)
    emb = F.linear(ou.timestep_embedding(timesteps, mc), sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    lab = F.linear(y.float(), sd["label_emb.0.0.weight"], sd["label_emb.0.0.bias"])
    emb = emb + F.linear(F.silu(lab), sd["label_emb.0.2.weight"], sd["label_emb.0.2.bias"])
    def inj(kind, i, layers):
        """video_model.py:480-497, 532-550 (same convention as oracle/unet.py)."""
        if injection is None or kind not in injection["block_types"] or i not in injection[f"{kind}_block_indices"] \
                or len(layers) < 2 or layers[1][0] != "attn":
            return None
        keys = [f"{kind}_block_{i}_{ft}_time_{injection['timestep']}" for ft in injection["feature_types"]]
        return {k: injection["features"][k] for k in keys if k in injection["features"]}

    def mods(i, layers):
        """video_model.py:523-530 (block frames) + video_attention.py:437-463 (layer frames per layer type)."""
        mp = modulate_params
        if mp is None or i not in mp["modulate_block_idx"] or len(layers) < 2 or layers[1][0] != "attn":
            return (None, None)
        every = list(range(mp["num_frames"]))
        out = []
        for kind in ("spatial", "temporal"):
            if kind not in mp["modulate_layer_type"]:
                out.append(None)
                continue
            out.append(dict(feature_masks=mp["feature_masks"], block_frames=mp["modulate_block_frames"].get(i, every),
                            layer_frames=mp["modulate_layer_frames"].get(kind, every),
                            timestep_frames=mp["modulate_timestep_frames_group"], lambda_start=mp["modulate_lambda_start"],
                            lambda_end=mp["modulate_lambda_end"], schedule=mp["modulate_schedule"],
                            num_frames=mp["num_frames"], uc=mp["modulate_uc"], attn_types=mp["modulate_attn_type"]))
        return tuple(out)

    hs = []
    h = x.float()
    for i, layers in enumerate(inputs):
        h = _run_block(sd, f"input_blocks.{i}", layers, h, emb, context, T, ind, stash, f"input_block_{i}",
                       injected=inj("input", i, layers))
        hs.append(h)
    h = _run_block(sd, "middle_block", middle, h, emb, context, T, ind, stash, "middle_block")
    for i, layers in enumerate(outputs):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, f"output_blocks.{i}", layers, h, emb, context, T, ind, stash, f"output_block_{i}",
                       mods(i, layers), inj("output", i, layers))
    h = F.silu(F.group_norm(h, 32, sd["out.0.weight"], sd["out.0.bias"], eps=1e-5))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)
