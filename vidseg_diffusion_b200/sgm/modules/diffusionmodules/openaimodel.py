"""SD-2.1 UNet single-step forward on B200 (reference: sgm/modules/diffusionmodules/openaimodel.py).

``UNetModel`` keeps the reference's constructor kwargs (sd_2_1.yaml:18-30), forward signature
(:831-841), module tree and state-dict keys (``input_blocks.i.j...``, ``middle_block.j...``,
``output_blocks.i.j...``, ``time_embed``, ``out``), and the class-name convention the pipelines rely on
to find the transformer layers (``"SpatialTransformer" in str(type(module[1]))``, :878/:914/:923).
Inference only: dropout and gradient checkpointing are identities here.
"""
import torch
import torch.nn as nn

from .... import _lib
from .... import kernels as K
from ..attention import SpatialTransformer, _unsupported
from ...util import load_target_features
from .util import normalization, timestep_embedding, zero_module


class TimestepBlock(nn.Module):
    """Marker base class: forward(x, emb)."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """reference :67-114: routes emb / context to the children that take them."""

    def forward(self, x, emb, context=None, image_only_indicator=None, time_context=None, num_video_frames=None,
                is_modulate_step=False, is_injected_step=False, modulate_params=None):
        from ..video_attention import SpatialVideoTransformer
        from .video_model import VideoResBlock
        for layer in self:
            if isinstance(layer, VideoResBlock):
                x = layer(x, emb, num_video_frames, image_only_indicator)
            elif isinstance(layer, SpatialVideoTransformer):
                x = layer(x, context, time_context, num_video_frames, image_only_indicator,
                          is_modulate_step=is_modulate_step, is_injected_step=is_injected_step,
                          modulate_params=modulate_params)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context, is_modulate_step=is_modulate_step, is_injected_step=is_injected_step,
                          modulate_params=modulate_params)
            elif isinstance(layer, TimestepBlock):
                x = layer(x, emb)
            elif isinstance(layer, nn.Conv2d):  # input_blocks[0]: 4 (SD) / 8 (SVD) latent channels, padded to 8 for the TMA rows
                x = K.conv2d(K.image_split(x), layer)
            else:
                x = layer(x)
        return x


class Upsample(nn.Module):
    """reference :117-158: nearest x2 then 3x3 conv."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1, third_up=False, kernel_size=3,
                 scale_factor=2):
        super().__init__()
        if dims != 2 or kernel_size != 3 or scale_factor != 2 or padding != 1:
            _unsupported("Upsample(dims != 2 / kernel != 3 / scale != 2)")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        if use_conv:
            self.conv = nn.Conv2d(self.channels, self.out_channels, 3, padding=1)

    def forward(self, x):
        assert x.shape[1] == self.channels
        if not self.use_conv:
            _unsupported("Upsample(use_conv=False)")
        return K.conv2d(K.upsample_nearest2x_split(x), self.conv)


class Downsample(nn.Module):
    """reference :161-217: 3x3 conv, stride 2, padding 1."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1, third_down=False):
        super().__init__()
        if dims != 2 or not use_conv or padding != 1:
            _unsupported("Downsample(dims != 2 / avg-pool form)")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        self.op = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=1)

    def forward(self, x):
        assert x.shape[1] == self.channels
        return K.conv2d(K.image_split(x), self.op)


class ResBlock(TimestepBlock):
    """reference :220-369 (plain form used by both configs: no up/down, no scale-shift norm)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False, kernel_size=3, exchange_temb_dims=False,
                 skip_t_emb=False):
        super().__init__()
        video = dims == 3
        if up or down or use_scale_shift_norm or dims not in (2, 3) or (skip_t_emb and not video):
            _unsupported("ResBlock(up / down / scale-shift / dims not in (2, 3) / skip_t_emb outside a time stack)")
        # the (3,1,1) time stacks: the UNet's carries the per-frame embedding (exchange_temb_dims), the first-stage
        # VideoDecoder's has none (skip_t_emb, temporal_ae.py:32-44)
        if video and not (list(kernel_size) == [3, 1, 1] and (exchange_temb_dims or skip_t_emb) and not use_conv):
            _unsupported("ResBlock(dims=3) other than the (3,1,1) time_stack of VideoResBlock")
        if not video and (kernel_size != 3 or exchange_temb_dims):
            _unsupported("ResBlock(dims=2, kernel != 3 / exchange_temb_dims)")
        self.dims = dims
        self.exchange_temb_dims = exchange_temb_dims
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint
        conv = (lambda ci, co: nn.Conv3d(ci, co, (3, 1, 1), padding=(1, 0, 0))) if video else \
            (lambda ci, co: nn.Conv2d(ci, co, 3, padding=1))
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(), conv(channels, self.out_channels))
        self.skip_t_emb = skip_t_emb
        if not skip_t_emb:
            self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(conv(self.out_channels, self.out_channels)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif video:
            _unsupported("ResBlock(dims=3) with a channel change")
        elif use_conv:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)

    def forward(self, x, emb):
        return self._forward(x, emb)

    def forward_video(self, x, emb, videos, frames, blend_alpha):
        """The dims=3 form as VideoResBlock uses it (video_model.py:75-85): x is the '(b t) c h w' output of the 2-D
        block (the 'b c t h w' rearrange is index arithmetic in the kernels), GroupNorm statistics run over whole
        clips, the convolutions are (3,1,1) over frames, the embedding is per frame (exchange_temb_dims), and the
        AlphaBlender mix with the block input is the epilogue of the second convolution."""
        emb_out = K.dense(emb, self.emb_layers[1], act_silu_in=True)  # [(b t), Cout]
        hs, _, _ = K.group_norm_split(x, self.in_layers[0], silu=True, samples=videos)
        h = K.conv_temporal(hs, self.in_layers[2], videos, frames, frame_bias=emb_out)
        hs2, _, _ = K.group_norm_split(h, self.out_layers[0], silu=True, samples=videos)
        return K.conv_temporal(hs2, self.out_layers[3], videos, frames, residual=x, blend=x, blend_alpha=blend_alpha)

    def _forward(self, x, emb):
        """x: image-shaped fp32 tensor or a K.ChannelCat (the skip concatenation of the output blocks)."""
        if self.dims != 2:
            _unsupported("ResBlock(dims=3).forward outside VideoResBlock")
        emb_out = K.dense(emb, self.emb_layers[1], act_silu_in=True)  # [B, Cout], added per (sample, channel)
        identity_skip = isinstance(self.skip_connection, nn.Identity)
        if identity_skip and isinstance(x, K.ChannelCat):
            x = x.materialize()
        hs, raw, _ = K.group_norm_split(x, self.in_layers[0], silu=True, want_raw=not identity_skip)
        h = K.conv2d(hs, self.in_layers[2], chan_bias=emb_out)
        skip = x if identity_skip else K.conv2d(raw, self.skip_connection)
        hs2, _, _ = K.group_norm_split(h, self.out_layers[0], silu=True)
        return K.conv2d(hs2, self.out_layers[3], residual=skip)


class UNetModel(nn.Module):
    """reference :487-954."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0.0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
                 num_heads=-1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, transformer_depth=1, context_dim=None, disable_self_attentions=None,
                 num_attention_blocks=None, disable_middle_self_attn=False, disable_middle_transformer=False,
                 use_linear_in_transformer=False, spatial_transformer_attn_type="softmax", adm_in_channels=None):
        super().__init__()
        if num_classes is not None or resblock_updown or use_scale_shift_norm or dims != 2 or not conv_resample:
            _unsupported("UNetModel(num_classes / resblock_updown / scale-shift / dims != 2 / conv_resample=False)")
        if num_heads == -1 and num_head_channels == -1:
            raise AssertionError("Either num_heads or num_head_channels has to be set")
        if isinstance(transformer_depth, int):
            transformer_depth = len(channel_mult) * [transformer_depth]
        transformer_depth_middle = transformer_depth[-1]
        if isinstance(num_res_blocks, int):
            num_res_blocks = len(channel_mult) * [num_res_blocks]
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.channel_mult = channel_mult
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels

        def transformer(ch, depth, disabled_sa=False):
            heads = num_heads if num_head_channels == -1 else ch // num_head_channels
            dim_head = ch // num_heads if num_head_channels == -1 else num_head_channels
            return SpatialTransformer(ch, heads, dim_head, depth=depth, context_dim=context_dim,
                                      disable_self_attn=disabled_sa, use_linear=use_linear_in_transformer,
                                      attn_type=spatial_transformer_attn_type, use_checkpoint=use_checkpoint)

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, time_embed_dim), nn.SiLU(),
                                        nn.Linear(time_embed_dim, time_embed_dim))
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        skip_chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for nr in range(num_res_blocks[level]):
                layers = [ResBlock(ch, time_embed_dim, dropout, out_channels=mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions and (num_attention_blocks is None or nr < num_attention_blocks[level]):
                    dsa = disable_self_attentions[level] if (context_dim is not None and disable_self_attentions) else False
                    layers.append(transformer(ch, transformer_depth[level], dsa))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, out_channels=ch)))
                skip_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, time_embed_dim, dropout, out_channels=ch),
            transformer(ch, transformer_depth_middle, disable_middle_self_attn) if not disable_middle_transformer
            else nn.Identity(),
            ResBlock(ch, time_embed_dim, dropout))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks[level] + 1):
                layers = [ResBlock(ch + skip_chans.pop(), time_embed_dim, dropout, out_channels=model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions and (num_attention_blocks is None or i < num_attention_blocks[level]):
                    dsa = disable_self_attentions[level] if disable_self_attentions else False
                    layers.append(transformer(ch, transformer_depth[level], dsa))
                if level and i == num_res_blocks[level]:
                    layers.append(Upsample(ch, conv_resample, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(nn.Conv2d(model_channels, out_channels, 3, padding=1)))

    @torch.no_grad()
    def forward(self, x, timesteps=None, context=None, y=None, is_modulate_step=False, is_injected_step=False,
                modulate_params=None, **kwargs):
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        if is_modulate_step:
            assert modulate_params is not None
            modulate_block_idx = modulate_params["modulate_block_idx"]
        if not x.is_cuda:
            raise _lib.VidsegError("UNetModel.forward: expected CUDA tensors (the hot path has no CPU fallback)")
        in_dtype = x.dtype
        x = x.float()
        emb = timestep_embedding(timesteps, self.model_channels)
        emb = K.dense(K.dense(emb, self.time_embed[0]), self.time_embed[2], act_silu_in=True)
        if context is not None:
            context = K.split(context.float().contiguous())  # split once, shared by the 16 cross-attention layers
        def injected_features(kind, i, module):
            """reference :880-893 / :918-935: the q / k tensors stashed by another pass for block i, or None."""
            if not (is_injected_step and kind in modulate_params["injected_block_types"] and len(module) > 1
                    and "SpatialTransformer" in str(type(module[1])) and i in modulate_params[f"{kind}_block_indices"]):
                return False
            modulate_params["injected_features_group"] = load_target_features(
                modulate_params.get("feature_folder"), modulate_params.get("exp_name"), modulate_params["timestep"], kind,
                modulate_params["injected_feature_types"], i, x.device, features=modulate_params.get("features"))
            return len(modulate_params["injected_features_group"]) > 0

        hs = []
        h = x
        for i, module in enumerate(self.input_blocks):
            h = module(h, emb, context=context, is_injected_step=injected_features("input", i, module),
                       modulate_params=modulate_params)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        for i, module in enumerate(self.output_blocks):
            h = K.concat_channels(h, hs.pop())
            # mask modulation happens in the transformer layers of the selected output blocks only (reference :907-916)
            mod_block = False
            if is_modulate_step and i in modulate_block_idx and len(module) > 1 and "SpatialTransformer" in str(type(module[1])):
                mod_block = True
                if i in modulate_params["modulate_block_frames"].keys():
                    modulate_params["modulate_block_frames_group"] = modulate_params["modulate_block_frames"][i]
                else:
                    modulate_params["modulate_block_frames_group"] = list(range(modulate_params["num_frames"]))
            h = module(h, emb, context=context, is_modulate_step=mod_block,
                       is_injected_step=injected_features("output", i, module), modulate_params=modulate_params)
        hs_out, _, _ = K.group_norm_split(h, self.out[0], silu=True)
        h = K.conv2d(hs_out, self.out[2])
        return h.contiguous().to(in_dtype)
