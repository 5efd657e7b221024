"""Differentiable denoiser forward (training path).  Filled in by the training milestone; until then a grad-mode
call fails loudly instead of silently running PyTorch eager code."""


def denoiser_forward_with_grad(module, x, timesteps, length, xf_proj, xf_out):
    raise NotImplementedError(
        "hig_b200: backward kernels are not built yet — call the denoiser under torch.no_grad() (sampling)")
