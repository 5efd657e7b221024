"""ncu target: the posterior-update kernel, q_sample and sample -> joints at the C2 shape (S = 128, T = 196)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hig_b200  # noqa
from hig_b200 import ops
from hig_b200.gaussian_diffusion import GaussianDiffusion, LossType, ModelMeanType, ModelVarType, get_named_beta_schedule
from hig_b200.motion_process import joints_from_samples
dev = torch.device("cuda:0")
S, T, C = 128, 196, 263
diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000), model_mean_type=ModelMeanType.EPSILON,
                         model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
coef = diff._tables(dev)["coef"]
x = torch.randn(S, T, C, device=dev)
eps16 = torch.randn(S * T, 512, device=dev).half()
t = torch.full((S,), 500, device=dev, dtype=torch.long)
xa = torch.zeros(S * T, 272, device=dev, dtype=torch.bfloat16)
seed = torch.zeros(1, device=dev, dtype=torch.long)
for _ in range(3):
    ops.ddpm_step(x, eps16, t, coef, noise=None, seed_dev=seed, packed=xa, t_next=t.clone())
x0 = torch.randn(S, T, C, device=dev)
diff.q_sample(x0, t, noise=torch.randn_like(x0))
joints_from_samples(x0, None, None, None, None)
torch.cuda.synchronize()
print("ok")
