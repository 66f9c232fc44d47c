"""UNet hyper-parameters of the two inference configs the pipelines use (reference: configs/inference/)."""

# configs/inference/sd_2_1.yaml:18-30 (network_config.params of sgm...openaimodel.UNetModel)
SD21_UNET = dict(
    use_checkpoint=True, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
    num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_head_channels=64, use_linear_in_transformer=True,
    transformer_depth=1, context_dim=1024)

# same topology at toy width (head dim stays 64): plumbing checks and fast tests
TINY_UNET = dict(
    use_checkpoint=True, in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[4, 2, 1],
    num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_head_channels=64, use_linear_in_transformer=True,
    transformer_depth=1, context_dim=96)

# configs/inference/svd.yaml:15-34 (network_config.params of sgm...video_model.VideoUNet)
SVD_UNET = dict(
    adm_in_channels=768, num_classes="sequential", use_checkpoint=True, in_channels=8, out_channels=4,
    model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4],
    num_head_channels=64, use_linear_in_transformer=True, transformer_depth=1, context_dim=1024,
    spatial_transformer_attn_type="softmax-xformers", extra_ff_mix_layer=True, use_spatial_context=True,
    merge_strategy="learned_with_images", video_kernel_size=[3, 1, 1])

TINY_VIDEO_UNET = dict(SVD_UNET, model_channels=64, context_dim=96, adm_in_channels=48)
