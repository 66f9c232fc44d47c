"""Classifier-free guidance (reference sgm/modules/diffusionmodules/guiders.py).  ``prepare_inputs`` doubles the batch
(unconditional half first); the combination itself, x_u + scale * (x_c - x_u), is also available as a per-sample scale
vector (``sample_scales``) so that the fused sampler step (csrc/sampler.cu) can apply it inside its single pass."""
import torch
from einops import rearrange, repeat

from ...util import append_dims, default


class Guider:
    def sample_scales(self, batch, device):
        """Guidance scale of every sample of the (un-doubled) batch, or None if the guider does not double it."""
        raise NotImplementedError


class VanillaCFG(Guider):
    def __init__(self, scale):
        self.scale = scale

    def __call__(self, x, sigma):
        x_u, x_c = x.chunk(2)
        return x_u + self.scale * (x_c - x_u)

    def prepare_inputs(self, x, s, c, uc):
        c_out = dict()
        for k in c:
            if k in ["vector", "crossattn", "concat"]:
                c_out[k] = torch.cat((uc[k], c[k]), 0)
            else:
                assert c[k] == uc[k]
                c_out[k] = c[k]
        return torch.cat([x] * 2), torch.cat([s] * 2), c_out

    def sample_scales(self, batch, device):
        return torch.full((batch,), float(self.scale), dtype=torch.float32, device=device)


class IdentityGuider(Guider):
    def __call__(self, x, sigma):
        return x

    def prepare_inputs(self, x, s, c, uc):
        return x, s, {k: c[k] for k in c}

    def sample_scales(self, batch, device):
        return None


class LinearPredictionGuider(Guider):
    """SVD: the scale grows linearly over the frames of a clip (guiders.py:60-100)."""

    def __init__(self, max_scale, num_frames, min_scale=1.0, additional_cond_keys=None):
        self.min_scale, self.max_scale, self.num_frames = min_scale, max_scale, num_frames
        self.scale = torch.linspace(min_scale, max_scale, num_frames).unsqueeze(0)
        additional_cond_keys = default(additional_cond_keys, [])
        if isinstance(additional_cond_keys, str):
            additional_cond_keys = [additional_cond_keys]
        self.additional_cond_keys = additional_cond_keys

    def __call__(self, x, sigma):
        x_u, x_c = x.chunk(2)
        x_u = rearrange(x_u, "(b t) ... -> b t ...", t=self.num_frames)
        x_c = rearrange(x_c, "(b t) ... -> b t ...", t=self.num_frames)
        scale = repeat(self.scale, "1 t -> b t", b=x_u.shape[0])
        scale = append_dims(scale, x_u.ndim).to(x_u.device)
        return rearrange(x_u + scale * (x_c - x_u), "b t ... -> (b t) ...")

    def prepare_inputs(self, x, s, c, uc):
        c_out = dict()
        for k in c:
            if k in ["vector", "crossattn", "concat"] + self.additional_cond_keys:
                c_out[k] = torch.cat((uc[k], c[k]), 0)
            else:
                assert c[k] == uc[k]
                c_out[k] = c[k]
        return torch.cat([x] * 2), torch.cat([s] * 2), c_out

    def sample_scales(self, batch, device):
        assert batch % self.num_frames == 0
        return self.scale.reshape(-1).repeat(batch // self.num_frames).to(device=device, dtype=torch.float32)
