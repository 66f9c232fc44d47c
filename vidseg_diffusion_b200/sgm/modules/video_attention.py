"""Temporal transformer layers of the SVD UNet on B200 (reference: sgm/modules/video_attention.py).

``VideoTransformerBlock`` (:18-292) and ``SpatialVideoTransformer`` (:295-489) keep the reference's constructor
kwargs, attribute names (``time_stack``, ``time_pos_embed``, ``time_mixer``, ``norm_in``, ``ff_in`` ...), state-dict
keys and the Q/K hook (``time_stack[0].attn1.q`` etc. in the reference's '(b s) t c' shape).

B200 design: activations never leave the frame-major token layout '(b t) s c' of the spatial layers.  Every Linear /
LayerNorm / GEGLU of the temporal block is per token, so it runs on that layout unchanged; the only op that mixes
frames -- self-attention over T -- indexes frames with a stride inside its kernel (csrc/temporal_attn.cu), so the
reference's two rearranges per block are never executed.  The broadcast adds (frame-position embedding, single-token
cross-attention output) ride as row biases of the following LayerNorm / GEMM epilogue, and the AlphaBlender mix is the
epilogue of the block's last GEMM.
"""
import torch
import torch.nn as nn

from ... import _lib
from ... import kernels as K
from ...linear import Split
from .attention import CrossAttention, FeedForward, SpatialTransformer, _unsupported, modulation_rows
from .diffusionmodules.util import AlphaBlender, timestep_embedding


def _to_frame_major(t, videos, frames):
    """the reference's '(b s) t c' tensor (an injected feature dump) -> frame-major [(b t), s, c]."""
    bs, T, c = t.shape
    s = bs // videos
    return t.reshape(videos, s, frames, c).permute(0, 2, 1, 3).reshape(videos * frames, s, c).float().contiguous()


def _to_site_major(t, videos, frames):
    """frame-major [(b t), s, c] -> the reference's '(b s) t c' tensor (what its hooks dump)."""
    bt, s, c = t.shape
    return t.view(videos, frames, s, c).permute(0, 2, 1, 3).reshape(videos * s, frames, c)


class VideoTransformerBlock(nn.Module):
    """reference :18-292 (inference path; modulation :197-216, :233-254 is a next row)."""
    ATTENTION_MODES = {"softmax": CrossAttention, "softmax-xformers": CrossAttention}

    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True, checkpoint=True,
                 timesteps=None, ff_in=False, inner_dim=None, attn_mode="softmax", disable_self_attn=False,
                 disable_temporal_crossattention=False, switch_temporal_ca_to_sa=False):
        super().__init__()
        if disable_self_attn or switch_temporal_ca_to_sa:
            _unsupported("VideoTransformerBlock(disable_self_attn / switch_temporal_ca_to_sa)")
        attn_cls = self.ATTENTION_MODES[attn_mode]
        self.ff_in = ff_in or inner_dim is not None
        if inner_dim is None:
            inner_dim = dim
        assert int(n_heads * d_head) == inner_dim
        self.is_res = inner_dim == dim
        if not self.is_res:
            _unsupported("VideoTransformerBlock(inner_dim != dim)")
        if self.ff_in:
            self.norm_in = nn.LayerNorm(dim)
            self.ff_in = FeedForward(dim, dim_out=inner_dim, dropout=dropout, glu=gated_ff)
        self.timesteps = timesteps
        self.disable_self_attn = disable_self_attn
        self.attn1 = attn_cls(query_dim=inner_dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.attn1.inject_v = (attn_mode == "softmax-xformers")   # the SDPA class ignores injected_v (attention.py:316)
        self.ff = FeedForward(inner_dim, dim_out=dim, dropout=dropout, glu=gated_ff)
        if disable_temporal_crossattention:
            self.attn2 = None
        else:
            self.norm2 = nn.LayerNorm(inner_dim)
            self.attn2 = attn_cls(query_dim=inner_dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                  dropout=dropout)
        self.norm1 = nn.LayerNorm(inner_dim)
        self.norm3 = nn.LayerNorm(inner_dim)
        self.switch_temporal_ca_to_sa = switch_temporal_ca_to_sa
        self.checkpoint = checkpoint
        self.attn2_out = None

    def forward(self, x, context=None, timesteps=None, is_modulate_step=False, is_injected_step=False,
                modulate_params=None):
        """Reference-shaped call: x [(b t), s, c] fp32, context = the per-site time context [(b s), L, D] (or its
        un-repeated form [b, L, D]).  Returns [(b t), s, c]."""
        assert self.timesteps or timesteps
        assert not (self.timesteps and timesteps) or self.timesteps == timesteps
        timesteps = self.timesteps or timesteps
        videos = x.shape[0] // timesteps
        if context is not None and not isinstance(context, Split):
            if context.shape[0] == videos * x.shape[1]:   # repeated 'b ... -> (b n) ...': every site has the same row
                context = context[:: x.shape[1]]
            context = K.split(context.float().contiguous())
        return self.forward_frames(x.float().contiguous(), context, timesteps, is_modulate_step=is_modulate_step,
                                   is_injected_step=is_injected_step, modulate_params=modulate_params)

    def forward_frames(self, x, context, timesteps, frame_bias=None, blend=None, blend_alpha=None, want_split=False,
                       is_modulate_step=False, is_injected_step=False, modulate_params=None):
        """x: fp32 [(b t), s, c]; frame_bias [(b t), c] is added to x first (x_mix = x + emb, reference :452-453);
        context: Split [b, 1, D], the time context of every clip; blend / blend_alpha [(b t)]: the AlphaBlender mix
        with the spatial branch applied to the result.  Returns fp32 [(b t), s, c] (and its Split with want_split)."""
        bt, s, c = x.shape
        T = timesteps
        videos = bt // T
        if bt != videos * T:
            raise _lib.VidsegError(f"VideoTransformerBlock: batch {bt} is not a multiple of {T} frames")
        fb = dict(row_bias=frame_bias, rows_per_bias=s) if frame_bias is not None else {}
        if self.ff_in:
            # x_skip = x (+ emb); x = ff_in(norm_in(x_skip)) + x_skip
            h = self.ff_in.forward_split(K.layer_norm_split(x, self.norm_in, **fb), x, **fb)
        elif frame_bias is not None:
            h = x + frame_bias[:, None, :]
        else:
            h = x
        # mask modulation (reference :197-216, :233-254, :260-278): 'out[half_hw:, i] += lambda_i * mask_i[:, None]' on the
        # '(b s) t c' layout is row (T + i) * S + s of the frame-major layout -- the same per-row term as in the spatial
        # blocks, added in the epilogue of the GEMM that produces attn1_out / attn2_out / ff_out
        mod = {"self_attn": None, "cross_attn": None, "ff_out": None}
        if is_modulate_step:
            rows = modulation_rows(modulate_params, bt, s, x.device)
            for kind in mod:
                if kind in modulate_params["modulate_attn_type"]:
                    mod[kind] = rows
        inj = {}
        if is_injected_step:   # reference :167-176: lookup by substring of the dict key, tensors in '(b s) t c'
            for key, val in modulate_params["injected_features_group"].items():
                for name in "qkv":
                    if f"temporal_self_attn_{name}" in key:
                        inj[name] = _to_frame_major(val, videos, T)
        # attn1: self-attention over the frames of every site
        hs = K.layer_norm_split(h, self.norm1)
        a1 = self.attn1
        q = inj["q"] if "q" in inj else K.linear(hs, a1.to_q.weight, want_f32=True)[0]
        k = inj["k"] if "k" in inj else K.linear(hs, a1.to_k.weight, want_f32=True)[0]
        v = inj["v"] if ("v" in inj and a1.inject_v) else K.linear(hs, a1.to_v.weight, want_f32=True)[0]
        a1.q = lambda: _to_site_major(q, videos, T)
        a1.k = lambda: _to_site_major(k, videos, T)
        if a1.inject_v:   # the xformers class also keeps v (attention.py:446-448)
            a1.v = lambda: _to_site_major(v, videos, T)
        o = K.temporal_attention(q, k, v, videos, T, a1.heads, a1.scale)
        lin = a1.to_out[0]
        h, _ = K.linear(o, lin.weight, lin.bias, residual=h, want_f32=True, row_scalar=mod["self_attn"])
        # attn2: cross-attention to the clip's time context
        cb = {}
        if self.attn2 is not None:
            if context is None or context.hi.shape[1] != 1 or context.hi.shape[0] != videos:
                _unsupported("VideoTransformerBlock: temporal cross-attention to a context of more than one token")
            a2 = self.attn2
            q2, _ = K.linear(K.layer_norm_split(h, self.norm2), a2.to_q.weight, want_f32=True)
            k2, _ = K.linear(context, a2.to_k.weight, want_f32=True)
            _, v2 = K.linear(context, a2.to_v.weight, want_f32=False, want_split=True)
            a2.q = lambda: _to_site_major(q2, videos, T)
            a2.k = lambda: k2.repeat_interleave(s, dim=0)     # [(b s), 1, c]: the repeated time context's keys
            if a1.inject_v:
                a2.v = lambda: v2.float().repeat_interleave(s, dim=0)
            lin2 = a2.to_out[0]
            av, _ = K.linear(v2, lin2.weight, lin2.bias, want_f32=True)   # one vector per clip (softmax over 1 key = 1)
            if mod["cross_attn"] is not None:
                # modulated attn2_out is no longer one vector per clip: materialise x = x + attn2_out (rare path)
                h = (h.view(videos, T * s, c) + av.reshape(videos, 1, c)).view(bt, s, c) + mod["cross_attn"].view(bt, s, 1)
                h = h.contiguous()
            else:
                cb = dict(row_bias=av.reshape(videos, c), rows_per_bias=T * s)
        ep = dict(blend=blend, blend_alpha=blend_alpha, rows_per_alpha=s) if blend is not None else {}
        if mod["ff_out"] is not None:
            ep["row_scalar"] = mod["ff_out"]
        # x = x + attn2_out;  x = x + ff(norm3(x))   [is_res]
        return self.ff.forward_split(K.layer_norm_split(h, self.norm3, **cb), h, want_split=want_split, **cb, **ep)


class SpatialVideoTransformer(SpatialTransformer):
    """reference :295-489."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, use_linear=False, context_dim=None,
                 use_spatial_context=False, timesteps=None, merge_strategy="fixed", merge_factor=0.5,
                 time_context_dim=None, ff_in=False, checkpoint=False, time_depth=1, attn_mode="softmax",
                 disable_self_attn=False, disable_temporal_crossattention=False, max_time_embed_period=10000):
        super().__init__(in_channels, n_heads, d_head, depth=depth, dropout=dropout, attn_type=attn_mode,
                         use_checkpoint=checkpoint, context_dim=context_dim, use_linear=use_linear,
                         disable_self_attn=disable_self_attn)
        self.time_depth = time_depth
        self.depth = depth
        self.max_time_embed_period = max_time_embed_period
        inner_dim = n_heads * d_head
        if use_spatial_context:
            time_context_dim = context_dim
        self.time_stack = nn.ModuleList([
            VideoTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=time_context_dim,
                                  timesteps=timesteps, checkpoint=checkpoint, ff_in=ff_in, inner_dim=inner_dim,
                                  attn_mode=attn_mode, disable_self_attn=disable_self_attn,
                                  disable_temporal_crossattention=disable_temporal_crossattention)
            for _ in range(self.depth)])
        assert len(self.time_stack) == len(self.transformer_blocks)
        self.use_spatial_context = use_spatial_context
        self.in_channels = in_channels
        time_embed_dim = self.in_channels * 4
        self.time_pos_embed = nn.Sequential(nn.Linear(self.in_channels, time_embed_dim), nn.SiLU(),
                                            nn.Linear(time_embed_dim, self.in_channels))
        self.time_mixer = AlphaBlender(alpha=merge_factor, merge_strategy=merge_strategy)
        self.features_after_temporal = None

    def forward(self, x, context=None, time_context=None, timesteps=None, image_only_indicator=None,
                is_modulate_step=False, is_injected_step=False, modulate_params=None):
        bt, c, h, w = x.shape
        T = timesteps
        videos = bt // T
        s = h * w
        spatial_context = context
        if spatial_context is not None and not isinstance(spatial_context, Split):
            spatial_context = K.split(spatial_context.float().contiguous())
        if self.use_spatial_context:
            assert context is not None and len(spatial_context.hi.shape) == 3, "n dims of spatial context should be 3"
            time_ctx = spatial_context[::T]          # first frame of every clip (:401); the h*w repeat (:402-404) stays implicit
        elif time_context is not None:
            tc = time_context if time_context.dim() == 3 else time_context[:, None, :]
            time_ctx = K.split(tc.float().contiguous())
        else:
            time_ctx = None
        xs, _, x_nhwc = K.group_norm_split(x.float(), self.norm, silu=False)
        x_tok = x_nhwc.reshape(bt, s, c)
        t, _ = K.linear(xs.reshape(bt, s, c), self.proj_in.weight, self.proj_in.bias, want_f32=True)
        # frame-position embedding (:417-427): sinusoidal(frame index) -> MLP, one row per (clip, frame)
        frames = torch.arange(T, device=x.device).repeat(videos)
        t_emb = timestep_embedding(frames, self.in_channels, max_period=self.max_time_embed_period)
        emb = K.dense(K.dense(t_emb, self.time_pos_embed[0]), self.time_pos_embed[2], act_silu_in=True)
        alpha = self.time_mixer.frame_alpha(image_only_indicator, videos, T)
        ts = None
        n = len(self.transformer_blocks)
        def layer_mod(kind):   # reference :437-445 / :455-463: which frames this layer type modulates
            if not (is_modulate_step and kind in modulate_params["modulate_layer_type"]):
                return False
            if kind in modulate_params["modulate_layer_frames"].keys():
                modulate_params["modulate_layer_frames_group"] = modulate_params["modulate_layer_frames"][kind]
            else:
                modulate_params["modulate_layer_frames_group"] = list(range(modulate_params["num_frames"]))
            return True

        for i, (block, mix_block) in enumerate(zip(self.transformer_blocks, self.time_stack)):
            t = block(t, context=spatial_context, is_modulate_step=layer_mod("spatial"), is_injected_step=is_injected_step,
                      modulate_params=modulate_params)
            # x_mix = x + emb -> temporal block -> alpha * x + (1 - alpha) * x_mix, the last two fused in its output GEMM
            res = mix_block.forward_frames(t, time_ctx, T, frame_bias=emb, blend=t, blend_alpha=alpha,
                                           want_split=(i == n - 1), is_modulate_step=layer_mod("temporal"),
                                           is_injected_step=is_injected_step, modulate_params=modulate_params)
            t, ts = res if i == n - 1 else (res, None)
        out, _ = K.linear(ts, self.proj_out.weight, self.proj_out.bias, residual=x_tok, want_f32=True)
        out = K.as_nchw(out.reshape(bt, h, w, c))
        self.features_after_temporal = None   # reference keeps x here for ad-hoc inspection only (:487)
        return out
