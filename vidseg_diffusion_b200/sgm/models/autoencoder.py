"""First-stage models on B200 (reference: sgm/models/autoencoder.py).

``AutoencoderKL`` (:495-506, the ``first_stage_config`` of sd_2_1.yaml) = ``AutoencodingEngineLegacy`` (:440-492):
``Encoder`` -> ``quant_conv`` -> ``DiagonalGaussianRegularizer`` on the way in, ``post_quant_conv`` -> ``Decoder`` on the
way out, with the optional ``max_batch_size`` chunking.  ``AutoencodingEngine`` (:100-200, the ``first_stage_config`` of
svd.yaml:98-133) is the same without the two 1x1 convolutions and with a ``VideoDecoder`` that takes ``timesteps``.
Constructor kwargs, attribute names and state-dict keys are the reference's; inference only (no loss, EMA, optimiser).
"""
import math

import torch
import torch.nn as nn

from ..util import instantiate_from_config
from ... import kernels as K
from ..modules.diffusionmodules.model import Decoder, Encoder


class DiagonalGaussianDistribution:
    """reference sgm/modules/distributions/distributions.py:24-41.  The noise of ``sample`` is drawn on the CPU from
    torch's global generator and moved to the device, exactly like the reference, so seeded runs agree."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape)
        return self.mean + self.std * noise.to(device=self.parameters.device)

    def mode(self):
        return self.mean


class DiagonalGaussianRegularizer(nn.Module):
    """reference sgm/modules/autoencoding/regularizers/__init__.py:13-31 (the KL term is a training log only)."""

    def __init__(self, sample=True):
        super().__init__()
        self.sample = sample

    def forward(self, z):
        posterior = DiagonalGaussianDistribution(z)
        return (posterior.sample() if self.sample else posterior.mode()), {}


class AutoencodingEngine(nn.Module):
    """reference :100-200 (encode / decode / forward)."""

    def __init__(self, *args, encoder_config=None, decoder_config=None, loss_config=None, regularizer_config=None,
                 encoder=None, decoder=None, regularization=None, **kwargs):
        super().__init__()
        self.encoder = encoder if encoder is not None else instantiate_from_config(encoder_config)
        self.decoder = decoder if decoder is not None else instantiate_from_config(decoder_config)
        self.regularization = regularization if regularization is not None else (
            instantiate_from_config(regularizer_config) if regularizer_config is not None else DiagonalGaussianRegularizer())

    def get_last_layer(self):
        return self.decoder.get_last_layer()

    def encode(self, x, return_reg_log=False, unregularized=False):
        z = self.encoder(x)
        if unregularized:
            return z, dict()
        z, reg_log = self.regularization(z)
        return (z, reg_log) if return_reg_log else z

    def decode(self, z, **kwargs):
        return self.decoder(z, **kwargs)

    def forward(self, x, **additional_decode_kwargs):
        z, reg_log = self.encode(x, return_reg_log=True)
        return z, self.decode(z, **additional_decode_kwargs), reg_log


class AutoencodingEngineLegacy(AutoencodingEngine):
    """reference :440-492."""

    def __init__(self, embed_dim, **kwargs):
        self.max_batch_size = kwargs.pop("max_batch_size", None)
        ddconfig = kwargs.pop("ddconfig")
        kwargs.pop("ckpt_path", None)
        kwargs.pop("ckpt_engine", None)
        kwargs.pop("loss_config", None)
        super().__init__(encoder=Encoder(**ddconfig), decoder=Decoder(**ddconfig), **kwargs)
        self.quant_conv = nn.Conv2d((1 + ddconfig["double_z"]) * ddconfig["z_channels"], (1 + ddconfig["double_z"]) * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim = embed_dim

    def _chunks(self, n):
        bs = n if self.max_batch_size is None else self.max_batch_size
        return [(i * bs, min(n, (i + 1) * bs)) for i in range(int(math.ceil(n / max(bs, 1))))]

    def encode(self, x, return_reg_log=False):
        z = torch.cat([K.conv2d(K.image_split(self.encoder(x[a:b])), self.quant_conv) for a, b in self._chunks(x.shape[0])], 0)
        z, reg_log = self.regularization(z)
        return (z, reg_log) if return_reg_log else z

    def decode(self, z, **decoder_kwargs):
        return torch.cat([self.decoder(K.conv2d(K.image_split(z[a:b].float()), self.post_quant_conv), **decoder_kwargs)
                          for a, b in self._chunks(z.shape[0])], 0)


class AutoencoderKL(AutoencodingEngineLegacy):
    """reference :495-506."""

    def __init__(self, **kwargs):
        if "lossconfig" in kwargs:
            kwargs["loss_config"] = kwargs.pop("lossconfig")
        super().__init__(regularization=DiagonalGaussianRegularizer(), **kwargs)
