"""SVD ``VideoUNet`` single-step forward on B200 (reference: sgm/modules/diffusionmodules/video_model.py).

Keeps the reference's constructor kwargs (svd.yaml:15-34), forward signature (:451-463), module tree and state-dict
keys (``...time_stack...``, ``...time_mixer.mix_factor``, ``label_emb.0.{0,2}``) and the class-name convention the
pipelines use to find transformer layers (``"SpatialVideoTransformer" in str(type(module[1]))``, :485/:523/:532).
Inference only.
"""
import torch
import torch.nn as nn

from .... import _lib
from .... import kernels as K
from ..attention import _unsupported
from ..video_attention import SpatialVideoTransformer
from .openaimodel import Downsample, ResBlock, TimestepEmbedSequential, Upsample
from ...util import load_target_features
from .util import AlphaBlender, normalization, timestep_embedding, zero_module


class VideoResBlock(ResBlock):
    """reference :15-89: 2-D ResBlock, then a (3,1,1) ResBlock over frames, mixed by an AlphaBlender."""

    def __init__(self, channels, emb_channels, dropout, video_kernel_size=3, merge_strategy="fixed", merge_factor=0.5,
                 out_channels=None, use_conv=False, use_scale_shift_norm=False, dims=2, use_checkpoint=False, up=False,
                 down=False):
        super().__init__(channels, emb_channels, dropout, out_channels=out_channels, use_conv=use_conv,
                         use_scale_shift_norm=use_scale_shift_norm, dims=dims, use_checkpoint=use_checkpoint, up=up,
                         down=down)
        ch = out_channels if out_channels is not None else channels
        self.time_stack = ResBlock(ch, emb_channels, dropout=dropout, dims=3, out_channels=ch,
                                   use_scale_shift_norm=False, use_conv=False, up=False, down=False,
                                   kernel_size=video_kernel_size, use_checkpoint=use_checkpoint, exchange_temb_dims=True)
        self.time_mixer = AlphaBlender(alpha=merge_factor, merge_strategy=merge_strategy,
                                       rearrange_pattern="b t -> b 1 t 1 1")
        self.video_features = None

    def forward(self, x, emb, num_video_frames, image_only_indicator=None):
        x = super().forward(x, emb)
        bt = x.shape[0]
        videos = bt // num_video_frames
        if videos * num_video_frames != bt:
            raise _lib.VidsegError(f"VideoResBlock: batch {bt} is not a multiple of {num_video_frames} frames")
        alpha = self.time_mixer.frame_alpha(image_only_indicator, videos, num_video_frames)
        return self.time_stack.forward_video(x, emb, videos, num_video_frames, alpha)


class VideoUNet(nn.Module):
    """reference :92-566."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0.0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
                 num_heads=-1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, transformer_depth=1, transformer_depth_middle=None, context_dim=None,
                 time_downup=False, time_context_dim=None, extra_ff_mix_layer=False, use_spatial_context=False,
                 merge_strategy="fixed", merge_factor=0.5, spatial_transformer_attn_type="softmax", video_kernel_size=3,
                 use_linear_in_transformer=False, adm_in_channels=None, disable_temporal_crossattention=False,
                 max_ddpm_temb_period=10000):
        super().__init__()
        assert context_dim is not None
        if resblock_updown or use_scale_shift_norm or dims != 2 or not conv_resample or time_downup:
            _unsupported("VideoUNet(resblock_updown / scale-shift / dims != 2 / conv_resample=False / time_downup)")
        if num_heads == -1:
            assert num_head_channels != -1
        if num_head_channels == -1:
            assert num_heads != -1
        if isinstance(transformer_depth, int):
            transformer_depth = len(channel_mult) * [transformer_depth]
        if transformer_depth_middle is None:
            transformer_depth_middle = transformer_depth[-1]
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, time_embed_dim), nn.SiLU(),
                                        nn.Linear(time_embed_dim, time_embed_dim))
        if num_classes is not None:
            if num_classes != "sequential":
                _unsupported(f"VideoUNet(num_classes={num_classes!r}) (svd.yaml uses 'sequential')")
            assert adm_in_channels is not None
            self.label_emb = nn.Sequential(nn.Sequential(nn.Linear(adm_in_channels, time_embed_dim), nn.SiLU(),
                                                         nn.Linear(time_embed_dim, time_embed_dim)))

        def attention(ch, depth):
            heads = num_heads if num_head_channels == -1 else ch // num_head_channels
            dim_head = ch // num_heads if num_head_channels == -1 else num_head_channels
            return SpatialVideoTransformer(
                ch, heads, dim_head, depth=depth, context_dim=context_dim, time_context_dim=time_context_dim,
                dropout=dropout, ff_in=extra_ff_mix_layer, use_spatial_context=use_spatial_context,
                merge_strategy=merge_strategy, merge_factor=merge_factor, checkpoint=use_checkpoint,
                use_linear=use_linear_in_transformer, attn_mode=spatial_transformer_attn_type, disable_self_attn=False,
                disable_temporal_crossattention=disable_temporal_crossattention, max_time_embed_period=max_ddpm_temb_period)

        def resblock(ch, out_ch):
            return VideoResBlock(merge_factor=merge_factor, merge_strategy=merge_strategy,
                                 video_kernel_size=video_kernel_size, channels=ch, emb_channels=time_embed_dim,
                                 dropout=dropout, out_channels=out_ch, dims=dims, use_checkpoint=use_checkpoint,
                                 use_scale_shift_norm=use_scale_shift_norm)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        skip_chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [resblock(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(attention(ch, transformer_depth[level]))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                ds *= 2
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                skip_chans.append(ch)
        self.middle_block = TimestepEmbedSequential(resblock(ch, None), attention(ch, transformer_depth_middle),
                                                    resblock(ch, None))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [resblock(ch + skip_chans.pop(), model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(attention(ch, transformer_depth[level]))
                if level and i == num_res_blocks:
                    ds //= 2
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(nn.Conv2d(model_channels, out_channels, 3, padding=1)))

    @torch.no_grad()
    def forward(self, x, timesteps, context=None, y=None, time_context=None, num_video_frames=None,
                image_only_indicator=None, is_modulate_step=False, is_injected_step=False, modulate_params=None):
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional -> no, relax this TODO"
        if is_modulate_step:
            assert modulate_params is not None
            modulate_block_idx = modulate_params["modulate_block_idx"]
        if not x.is_cuda:
            raise _lib.VidsegError("VideoUNet.forward: expected CUDA tensors (the hot path has no CPU fallback)")
        if num_video_frames is None or x.shape[0] % num_video_frames:
            raise _lib.VidsegError("VideoUNet.forward: num_video_frames must divide the batch")
        in_dtype = x.dtype
        x = x.float()
        emb = timestep_embedding(timesteps, self.model_channels)
        emb = K.dense(K.dense(emb, self.time_embed[0]), self.time_embed[2], act_silu_in=True)
        if self.num_classes is not None:
            assert y.shape[0] == x.shape[0]
            lab = self.label_emb[0]
            emb = emb + K.dense(K.dense(y.float().contiguous(), lab[0]), lab[2], act_silu_in=True)
        if image_only_indicator is None:
            image_only_indicator = torch.zeros(x.shape[0] // num_video_frames, num_video_frames, device=x.device)
        context = K.split(context.float().contiguous())  # split once, shared by all cross-attention layers
        kw = dict(context=context, image_only_indicator=image_only_indicator, time_context=time_context,
                  num_video_frames=num_video_frames)
        def injected_features(kind, i, module):
            """reference :480-497 / :532-550: the q / k (/ v) tensors stashed by another pass for block i, or False."""
            if not (is_injected_step and kind in modulate_params["injected_block_types"] and len(module) > 1
                    and "SpatialVideoTransformer" in str(type(module[1])) and i in modulate_params[f"{kind}_block_indices"]):
                return False
            modulate_params["injected_features_group"] = load_target_features(
                modulate_params.get("feature_folder"), modulate_params.get("exp_name"), modulate_params["timestep"], kind,
                modulate_params["injected_feature_types"], i, x.device, features=modulate_params.get("features"))
            return len(modulate_params["injected_features_group"]) > 0

        hs = []
        h = x
        for i, module in enumerate(self.input_blocks):
            h = module(h, emb, is_injected_step=injected_features("input", i, module), modulate_params=modulate_params, **kw)
            hs.append(h)
        h = self.middle_block(h, emb, **kw)
        for i, module in enumerate(self.output_blocks):
            h = K.concat_channels(h, hs.pop())
            mod_block = False   # reference :523-530
            if is_modulate_step and i in modulate_block_idx and len(module) > 1 and "SpatialVideoTransformer" in str(type(module[1])):
                mod_block = True
                if i in modulate_params["modulate_block_frames"].keys():
                    modulate_params["modulate_block_frames_group"] = modulate_params["modulate_block_frames"][i]
                else:
                    modulate_params["modulate_block_frames_group"] = list(range(modulate_params["num_frames"]))
            h = module(h, emb, is_modulate_step=mod_block, is_injected_step=injected_features("output", i, module),
                       modulate_params=modulate_params, **kw)
        hs_out, _, _ = K.group_norm_split(h, self.out[0], silu=True)
        h = K.conv2d(hs_out, self.out[2])
        return h.contiguous().to(in_dtype)
