"""Pin oracle/video_unet.py to the UNMODIFIED reference VideoUNet and write SVD UNet goldens.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_video_unet_goldens.py
The reference ``sgm.modules.diffusionmodules.video_model.VideoUNet`` is built with the svd.yaml options
(``spatial_transformer_attn_type="softmax"`` instead of ``softmax-xformers``: xformers is not installed and the
two classes compute the same function), loaded with the seeded synthetic state dict
(tests/synth.py) and run in fp32 on the CPU; the oracle restatement must agree to 2e-5 of the tensor scale.
Stored (from the REFERENCE run): the stashed spatial ``attn1.q`` of output blocks 6/7/8, the temporal
``time_stack[0].attn1.q`` / ``attn2.k`` of output block 7 (svd_single_video_inference.py:117-125), the UNet
output and the state-dict key/shape table.  Large tensors are stored strided to keep the fixtures small.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import video_unet as ov  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402
from synth import synthetic_modulate_params, synthetic_unet_weights, synthetic_video_unet_inputs  # noqa: E402

# (name, cfg, seed, F, latent_hw, (token stride, channel stride) for the stored q)
CASES = [
    ("video_tiny", ov.TINY_VIDEO_CONFIG, 4, 3, 16, (1, 1)),
    ("svd_c1", ov.SVD_CONFIG, 1, 4, 32, (4, 8)),   # SVD at the size of BASELINE.json configs[0]
]


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max())


def build_reference(cfg):
    vm = import_reference("sgm.modules.diffusionmodules.video_model")
    return vm.VideoUNet(
        num_classes="sequential", use_checkpoint=False, use_linear_in_transformer=True, transformer_depth=1,
        spatial_transformer_attn_type="softmax", extra_ff_mix_layer=True, use_spatial_context=True,
        merge_strategy="learned_with_images", video_kernel_size=[3, 1, 1], **cfg).eval()


def main():
    only = sys.argv[1:]
    for name, cfg, seed, F, hw, stride in CASES:
        if only and name not in only:
            continue
        model = build_reference(cfg)
        ref_shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        shapes = ov.param_shapes(cfg)
        assert ref_shapes == shapes, sorted(set(ref_shapes) ^ set(shapes))[:10]
        sd = {k: torch.from_numpy(v) for k, v in synthetic_unet_weights(shapes, seed).items()}
        model.load_state_dict(sd, strict=True)
        x, t, ctx, y = (torch.from_numpy(a) for a in
                        synthetic_video_unet_inputs(seed, F, hw, cfg["in_channels"], cfg["context_dim"], cfg["adm_in_channels"]))
        ind = torch.zeros(2, F)
        with torch.no_grad():
            out_ref = model(x, timesteps=t, context=ctx, y=y, num_video_frames=F, image_only_indicator=ind)
        blk = lambda i: model.output_blocks[i][1]
        q_ref = {i: blk(i).transformer_blocks[0].attn1.q for i in (6, 7, 8)}
        tq_ref = blk(7).time_stack[0].attn1.q
        tk2_ref = blk(7).time_stack[0].attn2.k
        stash = {}
        out_or = ov.video_unet_forward(sd, cfg, x, t, ctx, y, F, ind, stash)
        errs = {"out": relerr(out_or, out_ref)}
        for i in (6, 7, 8):
            errs[f"q{i}"] = relerr(stash[(f"output_block_{i}", "spatial_self_attn_q")], q_ref[i])
        errs["tq7"] = relerr(stash[("output_block_7", "temporal_self_attn_q")], tq_ref)
        errs["tk2_7"] = relerr(stash[("output_block_7", "temporal_cross_attn_k")], tk2_ref)
        assert tuple(tq_ref.shape) == tuple(stash[("output_block_7", "temporal_self_attn_q")].shape)
        print(name, "oracle vs reference:", {k: f"{v:.2e}" for k, v in errs.items()},
              "| out absmax", float(out_ref.abs().max()), "tq7 shape", tuple(tq_ref.shape))
        assert max(errs.values()) < 2e-5, errs
        extra = {}
        if name == "video_tiny":
            # mask modulation of the spatial AND temporal layers (video_attention.py:197-278, 437-463; video_model.py:523-530)
            mp = dict(synthetic_modulate_params(seed, F, (hw // 2) ** 2), modulate_layer_type=["spatial", "temporal"],
                      modulate_attn_type=["self_attn", "cross_attn", "ff_out"], modulate_layer_frames={"temporal": [0, 2]})
            mp_t = dict(mp, feature_masks=[torch.from_numpy(m) for m in mp["feature_masks"]])
            with torch.no_grad():
                out_mod = model(x, timesteps=t, context=ctx, y=y, num_video_frames=F, image_only_indicator=ind,
                                is_modulate_step=True, modulate_params=mp_t)
            out_mod_or = ov.video_unet_forward(sd, cfg, x, t, ctx, y, F, ind, None, modulate_params=mp)
            e1 = relerr(out_mod_or, out_mod)
            print("video_tiny + modulation: oracle vs reference", f"{e1:.2e}", "| changed output by", f"{relerr(out_mod, out_ref):.2e}")
            assert e1 < 2e-5 and relerr(out_mod, out_ref) > 1e-2
            # feature injection into the spatial and temporal self-attention of input block 5 / output block 7 through the
            # reference's .pt files (video_model.py:480-497, 532-550; sgm/util.py:277-296), second pass on another latent
            import tempfile
            with tempfile.TemporaryDirectory() as root:
                fm = os.path.join(root, "src", "feature_maps")
                os.makedirs(fm)
                with torch.no_grad():
                    model(x, timesteps=t, context=ctx, y=y, num_video_frames=F, image_only_indicator=ind)
                types = ["spatial_self_attn_q", "spatial_self_attn_k", "temporal_self_attn_q", "temporal_self_attn_k"]
                for kind, i in (("input", 5), ("output", 7)):
                    layer = getattr(model, f"{kind}_blocks")[i][1]
                    for ft in types:
                        blk = layer.transformer_blocks[0] if ft.startswith("spatial") else layer.time_stack[0]
                        torch.save(getattr(blk.attn1, ft[-1]).clone(), os.path.join(fm, f"{kind}_block_{i}_{ft}_time_24.pt"))
                x2 = torch.from_numpy(synthetic_video_unet_inputs(seed + 50, F, hw, cfg["in_channels"], cfg["context_dim"],
                                                                  cfg["adm_in_channels"])[0])
                inj = dict(injected_block_types=["input", "output"], input_block_indices=[5], output_block_indices=[7],
                           feature_folder=root, exp_name="src", timestep=24, injected_feature_types=types)
                with torch.no_grad():
                    out_inj = model(x2, timesteps=t, context=ctx, y=y, num_video_frames=F, image_only_indicator=ind,
                                    is_injected_step=True, modulate_params=dict(inj))
                    out_plain2 = model(x2, timesteps=t, context=ctx, y=y, num_video_frames=F, image_only_indicator=ind)
                st1 = {}
                ov.video_unet_forward(sd, cfg, x, t, ctx, y, F, ind, st1)
                feats = {f"{kind}_block_{i}_{ft}_time_24": st1[(f"{kind}_block_{i}", ft)]
                         for kind, i in (("input", 5), ("output", 7)) for ft in types}
                out_inj_or = ov.video_unet_forward(sd, cfg, x2, t, ctx, y, F, ind, None, injection=dict(
                    block_types=["input", "output"], input_block_indices=[5], output_block_indices=[7], feature_types=types,
                    timestep=24, features=feats))
                e2 = relerr(out_inj_or, out_inj)
                print("video_tiny + injection: oracle vs reference", f"{e2:.2e}", "| changed output by", f"{relerr(out_inj, out_plain2):.2e}")
                assert e2 < 2e-5 and relerr(out_inj, out_plain2) > 1e-3
            extra = dict(out_mod=out_mod.numpy(), out_inj=out_inj.numpy())
        ts, cs = stride
        keys = np.array(sorted(shapes))
        np.savez_compressed(
            os.path.join(HERE, f"unet_{name}.npz"),
            out=out_ref.numpy(), q6=q_ref[6][:, ::ts, ::cs].numpy(), q7=q_ref[7][:, ::ts, ::cs].numpy(),
            q8=q_ref[8][:, ::ts, ::cs].numpy(), tq7=tq_ref[::ts, :, ::cs].numpy(), tk2_7=tk2_ref[::ts, :, ::cs].numpy(),
            q_stride=np.array(stride), keys=keys, shapes=np.array([",".join(map(str, shapes[k])) for k in keys]),
            meta=np.array([seed, F, hw]), **extra)


if __name__ == "__main__":
    main()
