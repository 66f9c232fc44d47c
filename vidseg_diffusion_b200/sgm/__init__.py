"""Drop-in counterparts of the reference's ``sgm`` operators on the hot path.

Import paths mirror the reference (``sgm.modules.attention`` -> ``vidseg_diffusion_b200.sgm.modules.attention``),
so a config's ``target:`` strings only change their prefix.  Class names, constructor kwargs, forward
signatures, side-effect attributes (``attn.q`` / ``attn.k``) and state-dict keys are the reference's;
the arithmetic runs in libvidseg_b200 (tcgen05 GEMMs / attention) on the GPU."""
