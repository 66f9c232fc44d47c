"""Pin oracle/unet.py to the UNMODIFIED reference UNetModel and write UNet goldens.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_unet_goldens.py
For each case the reference ``sgm.modules.diffusionmodules.openaimodel.UNetModel`` is built with
the case's config, loaded with the seeded synthetic state dict (tests/synth.py), and run in fp32
on the CPU; the oracle restatement must agree to 2e-5 of the tensor scale.  What is stored comes
from the REFERENCE run: the stashed ``attn1.q`` of output blocks 6/7/8
(svd_single_video_inference.py:117-125), the UNet output, and the state-dict key/shape table.
Large tensors are stored strided (``q_stride``) to keep the fixtures small.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import unet as ounet  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402
from synth import synthetic_modulate_params, synthetic_unet_inputs, synthetic_unet_weights  # noqa: E402

# (name, cfg, seed, F, latent_hw, context_len, (token stride, channel stride) for the stored q)
CASES = [
    ("tiny", ounet.TINY_CONFIG, 3, 2, 16, 7, (1, 1)),
    ("sd21_c1", ounet.SD21_CONFIG, 1, 4, 32, 77, (4, 8)),   # BASELINE.json configs[0]: 4 frames, 256x256
]


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max())


def main():
    om = import_reference("sgm.modules.diffusionmodules.openaimodel")
    only = sys.argv[1:]
    for name, cfg, seed, F, hw, L, stride in CASES:
        if only and name not in only:
            continue
        model = om.UNetModel(use_checkpoint=False, use_linear_in_transformer=True, transformer_depth=1, **cfg).eval()
        ref_shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        shapes = ounet.param_shapes(cfg)
        assert ref_shapes == shapes, set(ref_shapes) ^ set(shapes)
        w = synthetic_unet_weights(shapes, seed)
        sd = {k: torch.from_numpy(v) for k, v in w.items()}
        model.load_state_dict(sd, strict=True)
        x, t, ctx = synthetic_unet_inputs(seed, F, hw, cfg["in_channels"], L, cfg["context_dim"])
        x, t, ctx = torch.from_numpy(x), torch.from_numpy(t), torch.from_numpy(ctx)
        with torch.no_grad():
            out_ref = model(x, timesteps=t, context=ctx)
        q_ref = {i: model.output_blocks[i][1].transformer_blocks[0].attn1.q for i in (6, 7, 8)}
        k2_ref = model.output_blocks[7][1].transformer_blocks[0].attn2.k
        stash = {}
        out_or = ounet.unet_forward(sd, cfg, x, t, ctx, stash)
        errs = {"out": relerr(out_or, out_ref)}
        for i in (6, 7, 8):
            errs[f"q{i}"] = relerr(stash[(f"output_block_{i}", "spatial_self_attn_q")], q_ref[i])
        errs["k2"] = relerr(stash[("output_block_7", "spatial_cross_attn_k")], k2_ref)
        print(name, "oracle vs reference:", {k: f"{v:.2e}" for k, v in errs.items()},
              "| out absmax", float(out_ref.abs().max()), "q7 absmax", float(q_ref[7].abs().max()))
        assert max(errs.values()) < 2e-5, errs
        extra = {}
        if name == "tiny":
            # the same step with mask modulation switched on (is_modulate_step=True, attention.py:646-752,
            # openaimodel.py:907-916): reference output + stashed q of block 8, and the oracle must agree
            mp = synthetic_modulate_params(seed, F, (hw // 2) ** 2)
            mp_t = dict(mp, feature_masks=[torch.from_numpy(m) for m in mp["feature_masks"]])
            with torch.no_grad():
                out_mod = model(x, timesteps=t, context=ctx, is_modulate_step=True, modulate_params=mp_t)
            q8_mod = model.output_blocks[8][1].transformer_blocks[0].attn1.q
            st2 = {}
            out_mod_or = ounet.unet_forward(sd, cfg, x, t, ctx, st2, modulate_params=mp)
            e1, e2 = relerr(out_mod_or, out_mod), relerr(st2[("output_block_8", "spatial_self_attn_q")], q8_mod)
            print("tiny + modulation: oracle vs reference", f"{e1:.2e} {e2:.2e}", "| changed output by",
                  f"{relerr(out_mod, out_ref):.2e}")
            assert max(e1, e2) < 2e-5 and relerr(out_mod, out_ref) > 1e-2
            extra = dict(out_mod=out_mod.numpy(), q8_mod=q8_mod.numpy())
            # feature injection (is_injected_step=True, openaimodel.py:880-893 / 918-935, sgm/util.py:277-296): the q / k
            # stashed by the pass above are written as .pt files in the reference's feature_maps layout
            # (svd_single_video_inference.py:113-130) and injected into a SECOND pass on a different latent
            import tempfile
            with tempfile.TemporaryDirectory() as root:
                fm = os.path.join(root, "src", "feature_maps")
                os.makedirs(fm)
                with torch.no_grad():
                    model(x, timesteps=t, context=ctx)
                for kind, blocks in (("input", (4, 5)), ("output", (7, 8))):
                    for i in blocks:
                        tb = getattr(model, f"{kind}_blocks")[i][1].transformer_blocks[0]
                        torch.save(tb.attn1.q.clone(), os.path.join(fm, f"{kind}_block_{i}_spatial_self_attn_q_time_24.pt"))
                        torch.save(tb.attn1.k.clone(), os.path.join(fm, f"{kind}_block_{i}_spatial_self_attn_k_time_24.pt"))
                x2 = torch.from_numpy(synthetic_unet_inputs(seed + 50, F, hw, cfg["in_channels"], L, cfg["context_dim"])[0])
                inj = dict(injected_block_types=["input", "output"], input_block_indices=[4, 5], output_block_indices=[7, 8],
                           feature_folder=root, exp_name="src", timestep=24,
                           injected_feature_types=["spatial_self_attn_q", "spatial_self_attn_k"])
                with torch.no_grad():
                    out_inj = model(x2, timesteps=t, context=ctx, is_injected_step=True, modulate_params=dict(inj))
                    out_plain2 = model(x2, timesteps=t, context=ctx)
                st1 = {}
                ounet.unet_forward(sd, cfg, x, t, ctx, st1)
                feats = {}
                for kind, blocks in (("input", (4, 5)), ("output", (7, 8))):
                    for i in blocks:
                        for n in ("q", "k"):
                            feats[f"{kind}_block_{i}_spatial_self_attn_{n}_time_24"] = st1[(f"{kind}_block_{i}", f"spatial_self_attn_{n}")]
                out_inj_or = ounet.unet_forward(sd, cfg, x2, t, ctx, None, injection=dict(
                    block_types=["input", "output"], input_block_indices=[4, 5], output_block_indices=[7, 8],
                    feature_types=["spatial_self_attn_q", "spatial_self_attn_k"], timestep=24, features=feats))
                e3 = relerr(out_inj_or, out_inj)
                print("tiny + injection: oracle vs reference", f"{e3:.2e}", "| changed output by", f"{relerr(out_inj, out_plain2):.2e}")
                assert e3 < 2e-5 and relerr(out_inj, out_plain2) > 1e-3
                extra.update(out_inj=out_inj.numpy())
        ts, cs = stride
        keys = np.array(sorted(shapes))
        np.savez_compressed(
            os.path.join(HERE, f"unet_{name}.npz"),
            out=out_ref.numpy(), q6=q_ref[6][:, ::ts, ::cs].numpy(), q7=q_ref[7][:, ::ts, ::cs].numpy(),
            q8=q_ref[8][:, ::ts, ::cs].numpy(), q_stride=np.array(stride), keys=keys,
            shapes=np.array([",".join(map(str, shapes[k])) for k in keys]),
            meta=np.array([seed, F, hw, L]), **extra)


if __name__ == "__main__":
    main()
