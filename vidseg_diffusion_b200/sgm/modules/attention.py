"""Attention operators of the UNet transformer blocks on B200 (reference: sgm/modules/attention.py).

``CrossAttention`` (:257-364) keeps the reference's constructor, forward signature and the Q/K
"hook": after every forward ``self.q`` / ``self.k`` hold the projected, pre-head-split activations
[B, N, heads*64] in fp32 (:330-331) -- written by the epilogue of the projection GEMM, so harvesting
them costs no extra pass.  ``BasicTransformerBlock`` (:504-759), ``SpatialTransformer`` (:806-927),
``FeedForward`` / ``GEGLU`` (:89-115) keep their attribute names so that state dicts and
``output_blocks[i][1].transformer_blocks[0].attn1.q`` resolve exactly as in the reference.

Every Linear runs as a split-fp16 3-MMA tcgen05 GEMM and the softmax attention as the tcgen05 flash
kernel (csrc/gemm_tc.cu, csrc/attention_tc.cu); normalisations and the GEGLU gate are fused
streaming kernels (csrc/elementwise.cu).  There is no eager-PyTorch fallback.
"""
import torch
import torch.nn as nn

from ... import _lib
from ... import kernels as K
from ...linear import Split, split as _split


def get_modulate_lambda(modulate_lambda_start, modulate_lambda_end, modulate_schedule, total_steps, current_step):
    """reference sgm/modules/diffusionmodules/util.py:382-391."""
    assert modulate_schedule in ["constant", "linear"]
    if modulate_schedule == "constant":
        return modulate_lambda_start
    return modulate_lambda_start + (modulate_lambda_end - modulate_lambda_start) * current_step / total_steps


def modulation_rows(modulate_params, batch, tokens, device):
    """The mask modulation of one attention / feed-forward output (reference attention.py:646-663, 697-719, 730-752):
    ``out[i + num_masks] += lambda_i * mask_i[:, None]`` (and ``out[i]`` with modulate_uc) for every frame i selected by
    the block / layer / timestep frame groups.  Returned as one value per output row, fp32 [batch * tokens]; it is
    added in the epilogue of the GEMM that produces the output."""
    masks = modulate_params["feature_masks"]
    num_masks = len(masks)
    rows = torch.zeros(batch, tokens, dtype=torch.float32, device=device)
    for i, mask in enumerate(masks):
        if i in modulate_params["modulate_block_frames_group"] and i in modulate_params["modulate_layer_frames_group"] \
                and i in modulate_params["modulate_timestep_frames_group"]:
            lam = get_modulate_lambda(modulate_params["modulate_lambda_start"], modulate_params["modulate_lambda_end"],
                                      modulate_params["modulate_schedule"], total_steps=modulate_params["num_frames"],
                                      current_step=i)
            m = torch.as_tensor(mask).to(device=device, dtype=torch.float32).reshape(-1)
            if m.numel() != tokens:
                raise _lib.VidsegError(f"modulation mask {i} has {m.numel()} cells, the layer has {tokens} tokens")
            rows[i + num_masks] += lam * m
            if modulate_params["modulate_uc"]:
                rows[i] += lam * m
    return rows.reshape(-1)


def _unsupported(name):
    raise NotImplementedError(f"{name}: not on the B200 hot path (SURVEY.md section 8f, next rows)")


class GEGLU(nn.Module):
    """reference :89-96.  proj -> chunk(2) -> value * gelu(gate), computed in the GEMM's consumer kernel."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward_split(self, xs):
        if (self.proj.out_features // 2) % 32 == 0:
            return K.linear_geglu(xs, self.proj)      # gate applied in the GEMM epilogue, no fp32 [.., 2D] intermediate
        h, _ = K.linear(xs, self.proj.weight, self.proj.bias, want_f32=True)
        return K.geglu_split(h)

    def forward(self, x):
        return self.forward_split(K.split(x)).float()


class FeedForward(nn.Module):
    """reference :99-115 (only the gated form is on the path: BasicTransformerBlock uses glu=True)."""

    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.0):
        super().__init__()
        if not glu:
            _unsupported("FeedForward(glu=False)")
        inner_dim = int(dim * mult)
        dim_out = dim if dim_out is None else dim_out
        self.net = nn.Sequential(GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))

    def forward_split(self, xs, residual=None, want_split=False, want_f32=True, **epilogue):
        """``epilogue``: row_bias / blend terms of the output GEMM (linear.gemm_split).  Returns fp32, or
        (fp32 or None, Split) with ``want_split``."""
        hs = self.net[0].forward_split(xs)
        out, outs = K.linear(hs, self.net[2].weight, self.net[2].bias, residual=residual, want_f32=want_f32,
                             want_split=want_split, **epilogue)
        return (out, outs) if want_split else out

    def forward(self, x):
        return self.forward_split(K.split(x))


class CrossAttention(nn.Module):
    """reference :257-364.  Self-attention when ``context`` is None."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0, backend=None):
        super().__init__()
        if dim_head != 64:
            raise _lib.VidsegError(f"CrossAttention: the tcgen05 attention kernel is built for head dim 64, got {dim_head}")
        inner_dim = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))
        self.backend = backend
        # the reference's SDPA class ignores injected_v (attention.py:316), its xformers class uses it (:435-444);
        # BasicTransformerBlock sets this from attn_mode
        self.inject_v = False
        self.stash_f32 = True   # see forward_split
        self._stash = {"q": None, "k": None, "v": None}

    # the Q/K "hook" (reference :330-331): plain attributes there; here the temporal layers keep their activations in
    # the frame-major layout and hand out the reference's '(b s) t c' view on first access
    def _get_stash(self, name):
        v = self._stash[name]
        if callable(v):
            v = self._stash[name] = v()
        return v

    q = property(lambda self: self._get_stash("q"), lambda self, v: self._stash.__setitem__("q", v))
    k = property(lambda self: self._get_stash("k"), lambda self, v: self._stash.__setitem__("k", v))
    # the reference's xformers class also keeps ``self.v`` (:446-448).  Nothing on the path reads it, so it is decoded
    # from the attention kernel's fp16-pair operand (22 significant bits) on first access instead of costing an fp32 write
    v = property(lambda self: self._get_stash("v"), lambda self, v: self._stash.__setitem__("v", v))

    def forward_single_token(self, xs, context):
        """Cross-attention to a context of ONE token (SVD: the CLIP image embedding, svd.yaml): softmax over a single
        key is exactly 1, so every query of a sample receives to_out(to_v(context)).  Returns that vector per sample,
        fp32 [B, C]; the caller adds it as a row bias.  q is still projected because the pipelines dump it."""
        q, _ = K.linear(xs, self.to_q.weight, want_f32=True)
        k, _ = K.linear(context, self.to_k.weight, want_f32=True)
        _, vs = K.linear(context, self.to_v.weight, want_f32=False, want_split=True)
        self.q = q
        self.k = k
        if self.inject_v:
            self.v = vs.float
        lin = self.to_out[0]
        out, _ = K.linear(vs, lin.weight, lin.bias, want_f32=True)
        return out.reshape(out.shape[0], out.shape[-1])

    def forward_split(self, xs, context=None, residual=None, injected_q=None, injected_k=None, injected_v=None,
                      row_scalar=None):
        """xs: Split [B, N, C] (already normalised); context: Split [B, L, Cctx] or None.
        Returns to_out(attention) (+ residual) as fp32 [B, N, C]."""
        cs = xs if context is None else context
        # q / k / v feed the attention kernel, which takes fp16-pair operands whatever the GEMM operand policy is.
        # ``stash_f32`` (default): q and k are also written as fp32 by the projection's epilogue -- the reference's
        # ``self.q`` / ``self.k`` attributes.  A caller that reads only some layers' stash (pipeline.ClipSegmenter) clears
        # the flag on the others: their q / k are then decoded from the fp16 pair (22 significant bits) on first access.
        f32 = self.stash_f32
        no_inj = injected_q is None and injected_k is None and not (injected_v is not None and self.inject_v)
        fused = None
        if context is None and no_inj:   # self-attention: ONE GEMM reads the activation once for all three projections
            fused = K.linear_stacked(xs, (self.to_q.weight, self.to_k.weight, self.to_v.weight), (f32, f32, False),
                                     (True, True, True))
        if fused is not None:
            (q, qs), (k, ks), (_, vs) = fused
        else:
            if injected_q is not None:
                q, qs = injected_q, K.split(injected_q.float().contiguous(), pair16=True)
            else:
                q, qs = K.linear(xs, self.to_q.weight, want_f32=f32, want_split=True, split_pair16=True)
            if injected_k is not None:
                k, ks = injected_k, K.split(injected_k.float().contiguous(), pair16=True)
            else:
                k, ks = K.linear(cs, self.to_k.weight, want_f32=f32, want_split=True, split_pair16=True)
            if injected_v is not None and self.inject_v:
                vs = K.split(injected_v.float().contiguous(), pair16=True)
            else:
                _, vs = K.linear(cs, self.to_v.weight, want_f32=False, want_split=True, split_pair16=True)
        self.q = q if q is not None else qs.float
        self.k = k if k is not None else ks.float
        if self.inject_v:   # "softmax-xformers" layers only, as in the reference
            self.v = vs.float
        _, os_ = K.attention(qs, ks, vs, self.heads, self.scale)
        lin = self.to_out[0]
        out, _ = K.linear(os_, lin.weight, lin.bias, residual=residual, want_f32=True, row_scalar=row_scalar)
        return out

    def forward(self, x, context=None, mask=None, additional_tokens=None, n_times_crossframe_attn_in_self=0,
                injected_q=None, injected_k=None, injected_v=None):
        if mask is not None or additional_tokens is not None or n_times_crossframe_attn_in_self:
            _unsupported("CrossAttention(mask / additional_tokens / n_times_crossframe_attn_in_self)")
        xs = x if isinstance(x, Split) else K.split(x)
        if context is not None and not isinstance(context, Split):
            context = K.split(context)
        return self.forward_split(xs, context, None, injected_q, injected_k, injected_v)


class BasicTransformerBlock(nn.Module):
    """reference :504-759 (inference path; modulation :646-663,:697-719 is a next row)."""
    ATTENTION_MODES = {"softmax": CrossAttention, "softmax-xformers": CrossAttention}  # one kernel serves both

    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True, checkpoint=True,
                 disable_self_attn=False, attn_mode="softmax", sdp_backend=None):
        super().__init__()
        if attn_mode not in self.ATTENTION_MODES:
            raise AssertionError(attn_mode)
        attn_cls = self.ATTENTION_MODES[attn_mode]
        self.disable_self_attn = disable_self_attn
        self.attn1 = attn_cls(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout,
                              context_dim=context_dim if disable_self_attn else None, backend=sdp_backend)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = attn_cls(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                              dropout=dropout, backend=sdp_backend)
        self.attn1.inject_v = self.attn2.inject_v = (attn_mode == "softmax-xformers")
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.checkpoint = checkpoint  # inference only: torch.utils.checkpoint is a no-op without grad

    def forward(self, x, context=None, additional_tokens=None, n_times_crossframe_attn_in_self=0,
                is_modulate_step=False, is_injected_step=False, modulate_params=None, out_split=False):
        return self._forward(x, context, additional_tokens, n_times_crossframe_attn_in_self, is_modulate_step,
                             is_injected_step, modulate_params, out_split)

    def _forward(self, x, context=None, additional_tokens=None, n_times_crossframe_attn_in_self=0,
                 is_modulate_step=False, is_injected_step=False, modulate_params=None, out_split=False):
        """``out_split`` (not in the reference's signature): return the block's output as the tensor-core operand of the
        next GEMM (SpatialTransformer.proj_out) instead of fp32 -- the feed-forward GEMM writes it directly."""
        if additional_tokens is not None or n_times_crossframe_attn_in_self:
            _unsupported("BasicTransformerBlock(additional_tokens / n_times_crossframe_attn_in_self)")
        inj = {"self": [None] * 3, "cross": [None] * 3}
        if is_injected_step:  # reference :616-623, :674-681: lookup by substring of the dict key
            for key, val in modulate_params["injected_features_group"].items():
                for kind in ("self", "cross"):
                    for i, name in enumerate("qkv"):
                        if f"spatial_{kind}_attn_{name}" in key:
                            inj[kind][i] = val
        if context is not None and not isinstance(context, Split):
            context = K.split(context.float().contiguous())
        x = x.float().contiguous()
        # mask modulation (reference :646-663, :697-719, :730-752): one added value per (sample, token), identical for the
        # three sites, applied in the epilogue of the GEMM that produces attn1_out / attn2_out / ff_out
        mod = {"self_attn": None, "cross_attn": None, "ff_out": None}
        if is_modulate_step:
            rows = modulation_rows(modulate_params, x.shape[0], x.shape[1], x.device)
            for kind in mod:
                if kind in modulate_params["modulate_attn_type"]:
                    mod[kind] = rows
        ctx1 = context if self.disable_self_attn else None
        x = self.attn1.forward_split(K.layer_norm_split(x, self.norm1), ctx1, x, *inj["self"], row_scalar=mod["self_attn"])
        self.attn1_out = None  # the reference keeps attn1_out / attn2_out / ff_out for modulation only
        if context is not None and context.hi.shape[1] == 1 and not any(v is not None for v in inj["cross"]) \
                and not is_modulate_step:
            # one context token: attn2's output is one vector per sample, carried as a row bias of norm3 / ff
            n = x.shape[1]
            a = self.attn2.forward_single_token(K.layer_norm_split(x, self.norm2), context)
            y = self.ff.forward_split(K.layer_norm_split(x, self.norm3, row_bias=a, rows_per_bias=n), x,
                                      want_split=out_split, want_f32=not out_split, row_bias=a, rows_per_bias=n)
            return y[1] if out_split else y
        x = self.attn2.forward_split(K.layer_norm_split(x, self.norm2), context, x, *inj["cross"], row_scalar=mod["cross_attn"])
        ff_kw = {"row_scalar": mod["ff_out"]} if mod["ff_out"] is not None else {}
        y = self.ff.forward_split(K.layer_norm_split(x, self.norm3), x, want_split=out_split, want_f32=not out_split, **ff_kw)
        return y[1] if out_split else y


class SpatialTransformer(nn.Module):
    """reference :806-927 (``use_linear=True`` form used by sd_2_1.yaml / svd.yaml)."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None, disable_self_attn=False,
                 use_linear=False, attn_type="softmax", use_checkpoint=True, sdp_backend=None):
        super().__init__()
        if not use_linear:
            _unsupported("SpatialTransformer(use_linear=False)")
        if context_dim is not None and not isinstance(context_dim, (list, tuple)):
            context_dim = [context_dim]
        if context_dim is None:
            context_dim = [None] * depth
        elif len(context_dim) != depth:
            context_dim = depth * [context_dim[0]]
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.norm = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner_dim)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim[d],
                                  disable_self_attn=disable_self_attn, attn_mode=attn_type, checkpoint=use_checkpoint,
                                  sdp_backend=sdp_backend) for d in range(depth)])
        self.proj_out = nn.Linear(inner_dim, in_channels)
        self.use_linear = use_linear

    def forward(self, x, context=None, is_modulate_step=False, is_injected_step=False, modulate_params=None):
        if not isinstance(context, (list, tuple)):
            context = [context]
        context = [c if (c is None or isinstance(c, Split)) else K.split(c.float().contiguous()) for c in context]
        b, c, h, w = x.shape
        # GroupNorm (eps 1e-6) + operand split in one kernel; activations are channels-last, so the reference's
        # 'b c h w -> b (h w) c' is a view and the tokens of the transformer are the pixels of the convolutions
        xs, _, x_nhwc = K.group_norm_split(x.float(), self.norm, silu=False)
        x_tok = x_nhwc.reshape(b, h * w, c)
        t, _ = K.linear(xs.reshape(b, h * w, c), self.proj_in.weight, self.proj_in.bias, want_f32=True)
        for i, block in enumerate(self.transformer_blocks):
            ctx = context[0] if (i > 0 and len(context) == 1) else context[i]
            mod_spatial = False
            if is_modulate_step and "spatial" in modulate_params["modulate_layer_type"]:   # reference :906-913
                mod_spatial = True
                if "spatial" in modulate_params["modulate_layer_frames"].keys():
                    modulate_params["modulate_layer_frames_group"] = modulate_params["modulate_layer_frames"]["spatial"]
                else:
                    modulate_params["modulate_layer_frames_group"] = list(range(modulate_params["num_frames"]))
            t = block(t, context=ctx, is_modulate_step=mod_spatial, is_injected_step=is_injected_step,
                      modulate_params=modulate_params, out_split=(i == len(self.transformer_blocks) - 1))
        # proj_out + residual x_in, still in token layout; the result is handed on as b c h w (channels_last view)
        out, _ = K.linear(t if isinstance(t, Split) else K.split(t), self.proj_out.weight, self.proj_out.bias, residual=x_tok,
                          want_f32=True)
        return K.as_nchw(out.reshape(b, h, w, c))
