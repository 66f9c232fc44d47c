"""Seeded synthetic inputs shared by the golden generator and the tests.

Only elementwise float32 arithmetic and numpy's RandomState are used (no BLAS),
so the arrays are bit-identical on every machine with the same numpy.
"""
import numpy as np


def synthetic_clip_features(seed, num_frames, h, w, channels, num_objects, n_blocks=3,
                            sigma=0.15, texture=0.6, kind="objects"):
    """Returns a list of ``n_blocks`` arrays [2F, h*w, C] float32 (uncond rows first),
    mimicking the stashed attn1.q of output blocks 8/7/6, plus the ground-truth
    object map [F, h, w].

    kind="objects": moving Voronoi regions, a prototype vector per object and an
    object-anchored texture so that nearest-neighbour tracking is meaningful.
    kind="iid": pure N(0,1) features (stress case: many near-ties).
    """
    r = np.random.RandomState(seed)
    F, C = num_frames, channels
    if kind == "iid":
        blocks = [r.standard_normal((2 * F, h * w, C)).astype(np.float32) for _ in range(n_blocks)]
        return blocks, np.zeros((F, h, w), dtype=np.int64)
    proto = r.standard_normal((num_objects, C)).astype(np.float32)
    tex = r.standard_normal((num_objects, h, w, C)).astype(np.float32)
    pos0 = np.stack([r.randint(0, h, num_objects), r.randint(0, w, num_objects)], 1)
    vel = r.randint(-1, 2, (num_objects, 2))
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    seg = np.zeros((F, h, w), dtype=np.int64)
    base = np.zeros((F, h, w, C), dtype=np.float32)
    for t in range(F):
        pos = pos0 + vel * t
        # toroidal squared distance to every object centre (integer arithmetic)
        dy = (yy[None] - pos[:, 0, None, None]) % h
        dy = np.minimum(dy, h - dy)
        dx = (xx[None] - pos[:, 1, None, None]) % w
        dx = np.minimum(dx, w - dx)
        seg[t] = np.argmin(dy * dy + dx * dx, axis=0)
        for k in range(num_objects):
            m = seg[t] == k
            ty = (yy - pos[k, 0]) % h
            tx = (xx - pos[k, 1]) % w
            base[t][m] = proto[k] + np.float32(texture) * tex[k, ty[m], tx[m]]
    base = base.reshape(F, h * w, C)
    blocks = []
    for _ in range(n_blocks):
        cond = base + np.float32(sigma) * r.standard_normal(base.shape).astype(np.float32)
        uncond = r.standard_normal(base.shape).astype(np.float32)
        scale = np.float32(1.0 + 0.5 * r.rand())
        blocks.append(np.concatenate([uncond, cond * scale], axis=0).astype(np.float32))
    return blocks, seg


# (name, seed, F, h, w, C, K, kind) -- sizes follow BASELINE.json configs 1 and 2/3
CLUSTER_CASES = [
    ("c1_objects", 1, 4, 16, 16, 640, 5, "objects"),
    ("c1_iid", 2, 4, 16, 16, 640, 5, "iid"),
    ("small_odd", 3, 3, 9, 7, 40, 4, "objects"),
    ("mid_objects", 4, 6, 24, 24, 320, 12, "objects"),
    ("c2_objects", 1, 14, 32, 32, 640, 20, "objects"),
]


# ---------------------------------------------------------------------------------------------
# UNet: seeded weights and inputs (numpy RandomState only -> identical on every machine)
# ---------------------------------------------------------------------------------------------
def synthetic_unet_weights(shapes, seed=0):
    """{key: float32 array} for a {key: shape} table with the reference's state-dict names.

    Every tensor is drawn (including the layers the reference zero-initialises, which would make
    the UNet output identically zero): matrices / conv kernels N(0, 1/fan_in), norm scales
    1 + 0.1 N(0,1), biases 0.05 N(0,1).  Keys are visited in sorted order."""
    r = np.random.RandomState(seed)
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        z = r.standard_normal(shape).astype(np.float32)
        if len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            out[key] = z * np.float32(1.0 / np.sqrt(fan_in))
        elif key.endswith(".weight"):
            out[key] = np.float32(1.0) + np.float32(0.1) * z
        else:
            out[key] = np.float32(0.05) * z
    return out


def synthetic_unet_inputs(seed, num_frames, latent_hw, in_channels, context_len, context_dim, timestep=261):
    """The batch the guider builds for one step (guiders.py:33-42): uncond rows first, both halves
    share the noisy latent; the uncond context is zero (force_uc_zero_embeddings)."""
    r = np.random.RandomState(seed)
    lat = r.standard_normal((num_frames, in_channels, latent_hw, latent_hw)).astype(np.float32)
    ctx = r.standard_normal((num_frames, context_len, context_dim)).astype(np.float32)
    x = np.concatenate([lat, lat], 0)
    context = np.concatenate([np.zeros_like(ctx), ctx], 0)
    t = np.full((2 * num_frames,), timestep, dtype=np.int64)
    return x, t, context


def synthetic_video_unet_inputs(seed, num_frames, latent_hw, in_channels, context_dim, adm_channels, c_noise=0.8):
    """The batch the SVD guider builds for one step (guiders.py:60-100): uncond rows first; each row is the noisy
    latent (4 channels) concatenated with the conditioning-frame latent (4 channels, zero in the uncond half);
    context is one CLIP image token per frame (zero in the uncond half); y = the vector conditioning
    (fps / motion bucket / cond_aug embeddings), identical in both halves; timesteps = c_noise = 0.25 ln(sigma)
    (denoiser_scaling.py:58)."""
    r = np.random.RandomState(seed)
    half = in_channels // 2
    lat = r.standard_normal((num_frames, half, latent_hw, latent_hw)).astype(np.float32)
    cond = r.standard_normal((1, half, latent_hw, latent_hw)).astype(np.float32).repeat(num_frames, 0)
    ctx = r.standard_normal((1, 1, context_dim)).astype(np.float32).repeat(num_frames, 0)
    y = r.standard_normal((1, adm_channels)).astype(np.float32).repeat(num_frames, 0)
    x = np.concatenate([np.concatenate([lat, np.zeros_like(cond)], 1), np.concatenate([lat, cond], 1)], 0)
    context = np.concatenate([np.zeros_like(ctx), ctx], 0)
    t = np.full((2 * num_frames,), c_noise, dtype=np.float32)
    return x, t, context, np.concatenate([y, y], 0)


def synthetic_modulate_params(seed, num_frames, tokens):
    """A mask-modulation request as svd_single_video_inference.py:438-500 builds it (values seeded): one feature mask per
    frame at the resolution of output blocks 6-8, modulation of the cross-attention and feed-forward outputs of blocks
    7 and 8 with a linear lambda schedule, frame subsets per block / layer / timestep, unconditional half included."""
    r = np.random.RandomState(seed)
    masks = [(r.rand(tokens) > 0.6).astype(np.float32) for _ in range(num_frames)]
    return dict(modulate_block_idx=[7, 8], modulate_block_frames={8: list(range(0, num_frames, 2))},
                modulate_layer_type=["spatial"], modulate_layer_frames={},
                modulate_attn_type=["cross_attn", "ff_out"], feature_masks=masks,
                modulate_timestep_frames_group=list(range(num_frames)), modulate_lambda_start=3.0,
                modulate_lambda_end=1.0, modulate_schedule="linear", num_frames=num_frames, modulate_uc=True)


def synthetic_modulated_frames(seed, num_masks, num_frames, height, width, fh, fw):
    """Decoded frames of the +lambda / -lambda modulated runs (uint8 [K, F, H, W, 3] each) and the K-means label maps
    [F, fh, fw] they belong to: a shared random base video; run k pushes the pixels of label k's region up (+lambda) or
    down (-lambda), every run carries its own small noise.  Integer / elementwise arithmetic only."""
    r = np.random.RandomState(seed)
    K, F = num_masks, num_frames
    labels = r.randint(0, K, (F, fh, fw)).astype(np.int32)
    # blocky regions: repeat a coarse random map so that every label owns connected patches
    coarse = r.randint(0, K, (F, max(fh // 4, 1), max(fw // 4, 1)))
    labels = np.repeat(np.repeat(coarse, -(-fh // coarse.shape[1]), axis=1), -(-fw // coarse.shape[2]), axis=2)[:, :fh, :fw].astype(np.int32)
    yy = (np.arange(height) * fh) // height
    xx = (np.arange(width) * fw) // width
    region = labels[:, yy][:, :, xx]                                   # [F, H, W] nearest upsampling
    base = r.randint(40, 216, (F, height, width, 3)).astype(np.int64)
    pos = np.zeros((K, F, height, width, 3), dtype=np.uint8)
    neg = np.zeros_like(pos)
    for k in range(K):
        push = (region == k)[..., None] * r.randint(8, 40, (F, 1, 1, 3))
        pos[k] = np.clip(base + push + r.randint(-3, 4, base.shape), 0, 255).astype(np.uint8)
        neg[k] = np.clip(base - push + r.randint(-3, 4, base.shape), 0, 255).astype(np.uint8)
    return pos, neg, labels


# (name, seed, K, F, H, W, fh, fw)
SEGMAP_CASES = [
    ("small", 11, 3, 2, 40, 56, 5, 7),
    ("c1", 12, 5, 4, 256, 256, 16, 16),
]
