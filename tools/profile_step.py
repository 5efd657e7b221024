"""Run a few denoiser+posterior steps at the C2 shape (S=128, T=196) — the target of the ncu captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
model = bench.build_model(dev)
S, T, C = 128, 196, 263
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(S, T, C, device=dev, generator=g)
xf_proj = torch.randn(S, 2048, device=dev, generator=g) * 0.5
xf_out = torch.randn(S, 77, 256, device=dev, generator=g)
length = torch.full((S,), T, device=dev, dtype=torch.long)
from hig_b200.gaussian_diffusion import GaussianDiffusion, LossType, ModelMeanType, ModelVarType, get_named_beta_schedule
diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", steps), model_mean_type=ModelMeanType.EPSILON,
                         model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE) if steps >= 20 else None
with torch.no_grad():
    for i in range(steps):
        t = torch.full((S,), 999 - i, device=dev, dtype=torch.long)
        torch.cuda.nvtx.range_push(f"step{i}")
        eps = model(x, t, length=length, xf_proj=xf_proj, xf_out=xf_out)
        torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("ok", eps.float().abs().mean().item())
