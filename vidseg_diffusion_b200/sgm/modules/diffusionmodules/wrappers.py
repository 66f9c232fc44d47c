"""reference sgm/modules/diffusionmodules/wrappers.py: adapts the conditioning dict to the UNet's signature."""
import torch
import torch.nn as nn

OPENAIUNETWRAPPER = "sgm.modules.diffusionmodules.wrappers.OpenAIWrapper"


class IdentityWrapper(nn.Module):
    def __init__(self, diffusion_model, compile_model=False):
        super().__init__()
        if compile_model:
            raise NotImplementedError("torch.compile is not used on this path: the UNet runs on hand-written kernels")
        self.diffusion_model = diffusion_model

    def forward(self, *args, **kwargs):
        return self.diffusion_model(*args, **kwargs)


class OpenAIWrapper(IdentityWrapper):
    def forward(self, x, t, c, **kwargs):
        x = torch.cat((x, c.get("concat", torch.Tensor([]).type_as(x))), dim=1)
        return self.diffusion_model(x, timesteps=t, context=c.get("crossattn", None), y=c.get("vector", None), **kwargs)
