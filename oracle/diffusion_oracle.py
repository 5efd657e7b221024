"""TEST INFRASTRUCTURE — CPU restatement of the reference's DDPM arithmetic around the denoiser.  NOT a product
path (same import rule as denoiser_oracle.py).

Follows codes/models/gaussian_diffusion.py: get_named_beta_schedule('linear') :229-246, GaussianDiffusion.__init__
:329-380, q_sample :399-417, p_mean_variance (EPSILON, FIXED_SMALL, clip_denoised=False) :443-537,
_predict_xstart_from_eps :539-544, q_posterior_mean_variance :419-441, p_sample :606-666,
p_sample_loop_progressive :718-769, and the masked MSE of codes/trainers/mul_ddpm_trainer.py:223-247.
Pinned by tests/test_oracle_cpu.py against goldens from the real reference.
"""
import numpy as np
import torch


class Schedule:
    """float64 tables exactly as GaussianDiffusion.__init__ builds them (:329-380)."""

    def __init__(self, num_steps=1000):
        scale = 1000 / num_steps
        self.betas = np.linspace(scale * 0.0001, scale * 0.02, num_steps, dtype=np.float64)
        self.num_timesteps = num_steps
        alphas = 1.0 - self.betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = self.betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)


def _extract(arr, t, ndim):
    """_extract_into_tensor (:1137-1150): index the float64 table, THEN cast to fp32, then broadcast."""
    res = torch.from_numpy(arr).to(t.device)[t].float()
    return res.view(-1, *([1] * (ndim - 1)))


def q_sample(sch, x_start, t, noise):
    return (_extract(sch.sqrt_alphas_cumprod, t, x_start.dim()) * x_start
            + _extract(sch.sqrt_one_minus_alphas_cumprod, t, x_start.dim()) * noise)


def p_sample_step(sch, x, eps, t, noise):
    """x_{t-1} from x_t and the predicted noise; op order as the reference evaluates it in fp32."""
    n = x.dim()
    pred_xstart = _extract(sch.sqrt_recip_alphas_cumprod, t, n) * x - _extract(sch.sqrt_recipm1_alphas_cumprod, t, n) * eps
    mean = _extract(sch.posterior_mean_coef1, t, n) * pred_xstart + _extract(sch.posterior_mean_coef2, t, n) * x
    log_var = _extract(sch.posterior_log_variance_clipped, t, n)
    nonzero = (t != 0).float().view(-1, *([1] * (n - 1)))
    return mean + nonzero * torch.exp(0.5 * log_var) * noise


def p_sample_loop(sch, model_fn, noise_seq, device="cpu"):
    """noise_seq [steps+1, S, T, C]: [0] is x_T, [1+k] the noise of the k-th reverse step.  model_fn(x, t) -> eps."""
    img = noise_seq[0].to(device)
    S = img.shape[0]
    trace = []
    for k, i in enumerate(range(sch.num_timesteps - 1, -1, -1)):
        t = torch.full((S,), i, dtype=torch.long, device=device)
        eps = model_fn(img, t)
        img = p_sample_step(sch, img, eps, t, noise_seq[1 + k].to(device))
        trace.append(img)
    return img, trace


def masked_mse_loss(pred, target, src_mask, pit=False):
    """DDPMMulTrainer.backward_G (:223-247).  Frame 0 only scores its first 4 dims, other frames all dims.
    pit=False: labelled branch.  pit=True: permutation-invariant branch over the 4B stacked sequences."""
    l0 = ((pred[:, 0, :4] - target[:, 0, :4]) ** 2).mean(dim=-1)
    l1 = ((pred[:, 1:] - target[:, 1:]) ** 2).mean(dim=-1)
    loss = torch.cat([l0.unsqueeze(1), l1], dim=1)
    if not pit:
        return (loss * src_mask).sum() / src_mask.sum()
    B = loss.shape[0]
    loss = (loss * src_mask).sum(dim=1).view(2, B // 2).sum(dim=0)
    return loss.view(2, B // 4).min(dim=0).values.sum() / (src_mask.sum() / 2)
