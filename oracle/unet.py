"""Oracle for A1/A3/A4/A7: the SD-2.1 UNet single-step forward with the Q/K stash.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain fp32 torch-CPU, functional, driven by a
state dict with the reference's key names -- no nn.Module tree, no shared code with the product.

Follows, in the reference tree:
  sgm/modules/diffusionmodules/openaimodel.py
    UNetModel.__init__ :517-829   block plan (which level gets ResBlock / SpatialTransformer /
                                  Downsample / Upsample, channel bookkeeping for the skips)
    UNetModel.forward  :831-954   time_embed, input blocks (push skips), middle, output blocks
                                  (cat with popped skip), out = GroupNorm32 + SiLU + conv
    ResBlock._forward  :341-369   GN32+SiLU+conv3x3, + emb_layers(SiLU+Linear)[:, :, None, None],
                                  GN32+SiLU+conv3x3, + skip (identity or 1x1 conv)
    Downsample :161-217 (3x3 conv stride 2 pad 1), Upsample :117-158 (nearest x2 + 3x3 conv)
  sgm/modules/diffusionmodules/util.py
    timestep_embedding :209-233 (cos | sin, max_period 1e4), GroupNorm32 :276-278 (eps 1e-5)
  sgm/modules/attention.py
    SpatialTransformer.forward :889-927   GroupNorm(32, eps=1e-6) -> b (hw) c -> proj_in Linear ->
                                          blocks -> proj_out Linear -> b c h w -> + x_in
    BasicTransformerBlock._forward :609-759  x += attn1(LN1 x); x += attn2(LN2 x, ctx); x += ff(LN3 x)
    CrossAttention.forward :286-364       q/k/v bias-free Linears, stash self.q / self.k (:330-331,
                                          pre-head-split [B, N, heads*64]), softmax(q k^T / sqrt(64)) v
                                          per head, to_out Linear + bias
    GEGLU :89-96, FeedForward :99-115     proj -> chunk(2) -> x * gelu(gate) (erf form) -> Linear
"""
import math

import torch
import torch.nn.functional as F

SD21_CONFIG = dict(  # configs/inference/sd_2_1.yaml:18-30
    in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1), num_res_blocks=2,
    channel_mult=(1, 2, 4, 4), num_head_channels=64, context_dim=1024)

# same topology (4 levels, attention at the first three, head dim 64) at toy width for fast tests
TINY_CONFIG = dict(
    in_channels=4, out_channels=4, model_channels=64, attention_resolutions=(4, 2, 1), num_res_blocks=2,
    channel_mult=(1, 2, 4, 4), num_head_channels=64, context_dim=96)


def block_plan(cfg):
    """Layer list of every input / middle / output block (openaimodel.py:606-823)."""
    mc, mult, nres = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    attn_at, hd = set(cfg["attention_resolutions"]), cfg["num_head_channels"]
    inputs = [[("conv", cfg["in_channels"], mc)]]
    skip_ch = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nres):
            layers = [("res", ch, m * mc)]
            ch = m * mc
            if ds in attn_at:
                layers.append(("attn", ch, ch // hd))
            inputs.append(layers)
            skip_ch.append(ch)
        if level != len(mult) - 1:
            inputs.append([("down", ch, ch)])
            skip_ch.append(ch)
            ds *= 2
    middle = [("res", ch, ch), ("attn", ch, ch // hd), ("res", ch, ch)]
    outputs = []
    for level, m in reversed(list(enumerate(mult))):
        for i in range(nres + 1):
            layers = [("res", ch + skip_ch.pop(), m * mc)]
            ch = m * mc
            if ds in attn_at:
                layers.append(("attn", ch, ch // hd))
            if level and i == nres:
                layers.append(("up", ch, ch))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs


def _layer_shapes(prefix, layer, cfg, out):
    kind, cin, cout = layer
    temb = 4 * cfg["model_channels"]
    if kind == "conv":
        out[prefix + ".weight"] = (cout, cin, 3, 3)
        out[prefix + ".bias"] = (cout,)
    elif kind == "down":
        out[prefix + ".op.weight"] = (cout, cin, 3, 3)
        out[prefix + ".op.bias"] = (cout,)
    elif kind == "up":
        out[prefix + ".conv.weight"] = (cout, cin, 3, 3)
        out[prefix + ".conv.bias"] = (cout,)
    elif kind == "res":
        out[prefix + ".in_layers.0.weight"] = (cin,)
        out[prefix + ".in_layers.0.bias"] = (cin,)
        out[prefix + ".in_layers.2.weight"] = (cout, cin, 3, 3)
        out[prefix + ".in_layers.2.bias"] = (cout,)
        out[prefix + ".emb_layers.1.weight"] = (cout, temb)
        out[prefix + ".emb_layers.1.bias"] = (cout,)
        out[prefix + ".out_layers.0.weight"] = (cout,)
        out[prefix + ".out_layers.0.bias"] = (cout,)
        out[prefix + ".out_layers.3.weight"] = (cout, cout, 3, 3)
        out[prefix + ".out_layers.3.bias"] = (cout,)
        if cin != cout:
            out[prefix + ".skip_connection.weight"] = (cout, cin, 1, 1)
            out[prefix + ".skip_connection.bias"] = (cout,)
    elif kind == "attn":
        ch, ctx = cin, cfg["context_dim"]
        out[prefix + ".norm.weight"] = (ch,)
        out[prefix + ".norm.bias"] = (ch,)
        out[prefix + ".proj_in.weight"] = (ch, ch)
        out[prefix + ".proj_in.bias"] = (ch,)
        tb = prefix + ".transformer_blocks.0"
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
        out[prefix + ".proj_out.weight"] = (ch, ch)
        out[prefix + ".proj_out.bias"] = (ch,)
    else:
        raise ValueError(kind)


def param_shapes(cfg):
    """{state-dict key: shape} of the reference UNetModel built with ``cfg``."""
    mc = cfg["model_channels"]
    out = {"time_embed.0.weight": (4 * mc, mc), "time_embed.0.bias": (4 * mc,),
           "time_embed.2.weight": (4 * mc, 4 * mc), "time_embed.2.bias": (4 * mc,)}
    inputs, middle, outputs = block_plan(cfg)
    for i, layers in enumerate(inputs):
        for j, layer in enumerate(layers):
            _layer_shapes(f"input_blocks.{i}.{j}", layer, cfg, out)
    for j, layer in enumerate(middle):
        _layer_shapes(f"middle_block.{j}", layer, cfg, out)
    for i, layers in enumerate(outputs):
        for j, layer in enumerate(layers):
            _layer_shapes(f"output_blocks.{i}.{j}", layer, cfg, out)
    out["out.0.weight"] = (mc,)
    out["out.0.bias"] = (mc,)
    out["out.2.weight"] = (cfg["out_channels"], mc, 3, 3)
    out["out.2.bias"] = (cfg["out_channels"],)
    return out


def timestep_embedding(timesteps, dim, max_period=10000):
    """util.py:209-233."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _attention(sd, p, x, context, heads, stash, tag, kind, injected=None):
    """attention.py:286-364.  x [B, N, C]; context [B, L, Cctx] or None (self-attention).  ``injected``: {key: tensor};
    a key containing ``spatial_{kind}_attn_q`` / ``_k`` replaces the projection (:305-315; v is never injected here)."""
    ctx = x if context is None else context
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(ctx, sd[p + ".to_k.weight"])
    for key, val in (injected or {}).items():
        if f"spatial_{kind}_attn_q" in key:
            q = val
        elif f"spatial_{kind}_attn_k" in key:
            k = val
    v = F.linear(ctx, sd[p + ".to_v.weight"])
    if stash is not None:
        stash[(tag, f"spatial_{kind}_attn_q")] = q
        stash[(tag, f"spatial_{kind}_attn_k")] = k
    b, n, c = q.shape
    d = c // heads
    qh, kh, vh = (t.reshape(b, -1, heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
    w = torch.softmax((qh @ kh.transpose(-1, -2)) * (d ** -0.5), dim=-1)
    o = (w @ vh).permute(0, 2, 1, 3).reshape(b, n, c)
    return F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _modulation(mod, batch, tokens):
    """attention.py:646-663 / :697-719 / :730-752: per-(sample, token) additive term [B, N, 1] or None."""
    if mod is None:
        return None
    masks = mod["feature_masks"]
    nm = len(masks)
    add = torch.zeros(batch, tokens, 1)
    for i, mask in enumerate(masks):
        if i in mod["block_frames"] and i in mod["layer_frames"] and i in mod["timestep_frames"]:
            lam = mod["lambda_start"] if mod["schedule"] == "constant" else \
                mod["lambda_start"] + (mod["lambda_end"] - mod["lambda_start"]) * i / mod["num_frames"]
            add[i + nm] += lam * torch.as_tensor(mask, dtype=torch.float32).reshape(-1, 1)
            if mod["uc"]:
                add[i] += lam * torch.as_tensor(mask, dtype=torch.float32).reshape(-1, 1)
    return add


def _transformer(sd, p, x, context, heads, stash, tag, mod=None, injected=None):
    """SpatialTransformer.forward (attention.py:889-927) with one BasicTransformerBlock (:609-759).  ``mod``: the mask
    modulation of this block (dict with feature_masks, block_frames, layer_frames, timestep_frames, lambda_start,
    lambda_end, schedule, num_frames, uc, attn_types) or None."""
    b, c, h, w = x.shape
    x_in = x
    x = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-6)
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    x = F.linear(x, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    tb = p + ".transformer_blocks.0"
    ln = lambda t, n: F.layer_norm(t, (c,), sd[f"{tb}.{n}.weight"], sd[f"{tb}.{n}.bias"], eps=1e-5)
    add = _modulation(mod, x.shape[0], x.shape[1])
    site = lambda kind: add if (add is not None and kind in mod["attn_types"]) else 0.0
    x = x + (_attention(sd, tb + ".attn1", ln(x, "norm1"), None, heads, stash, tag, "self", injected) + site("self_attn"))
    x = x + (_attention(sd, tb + ".attn2", ln(x, "norm2"), context, heads, stash, tag, "cross", injected) + site("cross_attn"))
    hgate = F.linear(ln(x, "norm3"), sd[tb + ".ff.net.0.proj.weight"], sd[tb + ".ff.net.0.proj.bias"])
    val, gate = hgate.chunk(2, dim=-1)
    x = x + (F.linear(val * F.gelu(gate), sd[tb + ".ff.net.2.weight"], sd[tb + ".ff.net.2.bias"]) + site("ff_out"))
    x = F.linear(x, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x.reshape(b, h, w, c).permute(0, 3, 1, 2) + x_in


def _resblock(sd, p, x, emb):
    """ResBlock._forward (openaimodel.py:341-369), no up/down, no scale-shift."""
    h = F.silu(F.group_norm(x, 32, sd[p + ".in_layers.0.weight"], sd[p + ".in_layers.0.bias"], eps=1e-5))
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.silu(F.group_norm(h, 32, sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"], eps=1e-5))
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def _run_block(sd, prefix, layers, h, emb, context, stash, tag, mod=None, injected=None):
    for j, (kind, cin, cout) in enumerate(layers):
        p = f"{prefix}.{j}"
        if kind == "conv":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "res":
            h = _resblock(sd, p, h, emb)
        elif kind == "attn":
            h = _transformer(sd, p, h, context, cout, stash, tag, mod, injected)
        elif kind == "down":
            h = F.conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
        elif kind == "up":
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
    return h


@torch.no_grad()
def unet_forward(sd, cfg, x, timesteps, context, stash=None, modulate_params=None, injection=None):
    """UNetModel.forward (openaimodel.py:831-954), inference path without modulation/injection.

    sd: {key: fp32 CPU tensor}; x [B, Cin, H, W]; timesteps [B]; context [B, L, Cctx].
    ``stash`` (dict) receives {(block_tag, "spatial_{self|cross}_attn_{q|k}"): tensor} with block_tag
    in the reference's dump naming (``input_block_i`` / ``middle_block`` / ``output_block_i``,
    svd_single_video_inference.py:117-125)."""
    mc = cfg["model_channels"]
    inputs, middle, outputs = block_plan(cfg)
    emb = F.linear(timestep_embedding(timesteps, mc), sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    hs = []
    h = x.float()
    def inj(kind, i, layers):
        """openaimodel.py:880-893, 918-935: {key: tensor} of the features stashed for this block, or None.
        ``injection``: dict(block_types, input_block_indices, output_block_indices, feature_types, timestep, features)."""
        if injection is None or kind not in injection["block_types"] or i not in injection[f"{kind}_block_indices"] \
                or len(layers) < 2 or layers[1][0] != "attn":
            return None
        keys = [f"{kind}_block_{i}_{ft}_time_{injection['timestep']}" for ft in injection["feature_types"]]
        return {k: injection["features"][k] for k in keys if k in injection["features"]}

    for i, layers in enumerate(inputs):
        h = _run_block(sd, f"input_blocks.{i}", layers, h, emb, context, stash, f"input_block_{i}", None, inj("input", i, layers))
        hs.append(h)
    h = _run_block(sd, "middle_block", middle, h, emb, context, stash, "middle_block")
    for i, layers in enumerate(outputs):
        h = torch.cat([h, hs.pop()], dim=1)
        mod = None
        mp = modulate_params
        if mp is not None and i in mp["modulate_block_idx"] and len(layers) > 1 and layers[1][0] == "attn" \
                and "spatial" in mp["modulate_layer_type"]:
            # openaimodel.py:907-916 (block frames) and attention.py:906-913 (layer frames)
            every = list(range(mp["num_frames"]))
            mod = dict(feature_masks=mp["feature_masks"], block_frames=mp["modulate_block_frames"].get(i, every),
                       layer_frames=mp["modulate_layer_frames"].get("spatial", every),
                       timestep_frames=mp["modulate_timestep_frames_group"], lambda_start=mp["modulate_lambda_start"],
                       lambda_end=mp["modulate_lambda_end"], schedule=mp["modulate_schedule"], num_frames=mp["num_frames"],
                       uc=mp["modulate_uc"], attn_types=mp["modulate_attn_type"])
        h = _run_block(sd, f"output_blocks.{i}", layers, h, emb, context, stash, f"output_block_{i}", mod,
                       inj("output", i, layers))
    h = F.silu(F.group_norm(h, 32, sd["out.0.weight"], sd["out.0.bias"], eps=1e-5))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)
