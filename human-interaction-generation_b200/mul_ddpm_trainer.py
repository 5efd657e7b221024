"""DDPMMulTrainer — drop-in for codes/trainers/mul_ddpm_trainer.py:50-341 on the B200 path.

Kept: the constructor contract `DDPMMulTrainer(args, encoder)` (args.device / multi / label_path / cap_id /
diffusion_steps / is_train), `generate`, `generate_batch`, `forward`, `backward_G`, `update`, `save`, `load`,
`train_mode`, `eval_mode`, `to`.  Sampling goes through GaussianDiffusion.p_sample_loop's CUDA-graph path; the
masked loss follows backward_G (:223-247).  Not re-implemented (out of scope): the
dataloader-driven `train()` epoch loop with its matplotlib loss plots, and label_batch/eval_data.
Deliberate fix, documented: `generate` takes caption2 from caption2 for non-final chunks (the reference reads
caption1 there, :212 — flagged as a bug in SURVEY.md §3.1); pass `reference_chunk_bug=True` to reproduce it.
"""
from collections import OrderedDict

import torch
from torch.nn.utils import clip_grad_norm_

from .gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType,
                                 create_named_schedule_sampler, get_named_beta_schedule)


class DDPMMulTrainer(object):

    def __init__(self, args, encoder):
        self.opt = args
        self.device = args.device
        self.multi = getattr(args, "multi", True)
        self.with_label = getattr(args, "label_path", None) is not None
        self.cap_id = getattr(args, "cap_id", False)
        self.encoder = encoder
        self.diffusion_steps = args.diffusion_steps
        self.diffusion = GaussianDiffusion(
            betas=get_named_beta_schedule("linear", self.diffusion_steps), model_mean_type=ModelMeanType.EPSILON,
            model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
        self.sampler = create_named_schedule_sampler("uniform", self.diffusion)
        self.sampler_name = "uniform"
        if getattr(args, "is_train", False):
            self.mse_criterion = torch.nn.MSELoss(reduction="none")
        self.to(self.device)

    # ------------------------------------------------------------------------------------------ plumbing
    def _net(self):
        return self.encoder.module if hasattr(self.encoder, "module") else self.encoder

    def to(self, device):
        if getattr(self.opt, "is_train", False):
            self.mse_criterion.to(device)
        self.encoder = self.encoder.to(device)

    def train_mode(self):
        self.encoder.train()

    def eval_mode(self):
        self.encoder.eval()

    @staticmethod
    def zero_grad(opt_list):
        for opt in opt_list:
            opt.zero_grad()

    @staticmethod
    def clip_norm(network_list):
        for network in network_list:
            clip_grad_norm_(network.parameters(), 0.5)

    @staticmethod
    def step(opt_list):
        for opt in opt_list:
            opt.step()

    # ------------------------------------------------------------------------------------------ sampling (:164-221)
    def generate_batch(self, caption1, caption2, m_lens, dim_pose, noise=None, noise_seq=None):
        net = self._net()
        m_lens = torch.cat([m_lens, m_lens], dim=0)
        T = int(min(int(m_lens.max()), net.num_frames))
        if self.cap_id:
            kw = {"text": [torch.as_tensor(caption1).reshape(-1), torch.as_tensor(caption2).reshape(-1)],
                  "length": m_lens}
            B = len(caption1) + len(caption2)
        else:
            caption = list(caption1) + list(caption2)
            B = len(caption)
            with torch.no_grad():
                xf_proj, xf_out = net.encode_text(caption, self.device)
            kw = {"xf_proj": xf_proj, "xf_out": xf_out, "length": m_lens}
        return self.diffusion.p_sample_loop(self.encoder, (B, T, dim_pose), noise=noise, clip_denoised=False,
                                            progress=False, model_kwargs=kw, noise_seq=noise_seq)

    def generate(self, caption1, caption2, m_lens, dim_pose, batch_size=512, reference_chunk_bug=False):
        N = len(caption1)
        self.encoder.eval()
        all_output = []
        for lo in range(0, N, batch_size):
            hi = min(lo + batch_size, N)
            c1 = caption1[lo:hi]
            c2 = caption1[lo:hi] if (reference_chunk_bug and hi < N) else caption2[lo:hi]
            output = self.generate_batch(c1, c2, m_lens[lo:hi], dim_pose)
            B = hi - lo
            all_output.extend([output[i], output[B + i]] for i in range(B))
        return all_output

    def generate_joints(self, caption1, caption2, m_lens, dim_pose, mean=None, std=None, init_mean=None, init_std=None,
                        joints_num=22, batch_size=512):
        """caption -> 3-D joints without leaving the GPU: generate_batch followed by the fused de-normalisation +
        recover_from_ric2 kernel (what tools/visualization.py:143-155 + plot_t2m2 :54-56 do per sample on the CPU).
        Returns a list of [joints1, joints2] per pair, each [m_len-1, joints_num, 3]."""
        from .motion_process import joints_from_samples
        N = len(caption1)
        self.encoder.eval()
        out = []
        for lo in range(0, N, batch_size):
            hi = min(lo + batch_size, N)
            lens = torch.as_tensor(m_lens[lo:hi]).reshape(-1)
            x = self.generate_batch(caption1[lo:hi], caption2[lo:hi], lens, dim_pose)
            B, T = hi - lo, x.shape[1]
            l2 = torch.cat([lens, lens]).clamp(max=T)
            j = joints_from_samples(x, mean, std, init_mean, init_std, length=l2, joints_num=joints_num)
            for i in range(B):
                n = max(int(l2[i]) - 1, 0)
                out.append([j[i, :n], j[B + i, :n]])
        return out

    # ------------------------------------------------------------------------------------------ training (:91-162, :223-256)
    def forward(self, batch_data, eval_mode=False):
        if not self.multi:
            raise NotImplementedError("single-person batches belong to the reference's DDPMTrainer (out of scope)")
        caption1, caption2, motion1, motion2, m_lens, _ = batch_data
        motion = torch.cat([motion1.detach().to(self.device).float(), motion2.detach().to(self.device).float()], dim=0)
        caption = list(caption1) + list(caption2)
        B, T = motion1.shape[0], motion.shape[1]
        cur_len = torch.as_tensor([min(T, int(m)) for m in m_lens], dtype=torch.long, device=self.device)
        t, _ = self.sampler.sample(B, motion.device)
        t = torch.cat([t, t], dim=0)
        if not self.with_label:
            # PIT: (m1, m1, m2, m2) x (c1, c2, c2, c1), :110-120
            caption = caption + list(caption2) + list(caption1)
            cur_len = torch.cat([cur_len] * 4, dim=0)
            forward_twice = True
        else:
            cur_len = torch.cat([cur_len, cur_len], dim=0)
            forward_twice = False
        if self.cap_id:
            half = len(caption) // 2
            text = [torch.as_tensor(caption[:half]).reshape(-1), torch.as_tensor(caption[half:]).reshape(-1)]
        else:
            text = caption
        output = self.diffusion.training_losses(model=self.encoder, x_start=motion, t=t,
                                                model_kwargs={"text": text, "length": cur_len},
                                                forward_twice=forward_twice)
        self.real_noise, self.fake_noise = output["target"], output["pred"]
        self.src_mask = self._net().generate_src_mask(T, cur_len).to(motion.device)

    def backward_G(self):
        """Masked MSE: frame 0 scores its first 4 dims only, the other frames all dims (:223-247)."""
        pred, tgt, mask = self.fake_noise, self.real_noise, self.src_mask
        l0 = ((pred[:, 0, :4] - tgt[:, 0, :4]) ** 2).mean(dim=-1)
        l1 = ((pred[:, 1:] - tgt[:, 1:]) ** 2).mean(dim=-1)
        loss = torch.cat([l0.unsqueeze(1), l1], dim=1)
        if self.with_label:
            loss = (loss * mask).sum() / mask.sum()
        else:
            n = loss.shape[0]
            loss = (loss * mask).sum(dim=1).view(2, n // 2).sum(dim=0)
            loss = loss.view(2, n // 4).min(dim=0).values.sum() / (mask.sum() / 2)
        self.loss_mot_rec = loss
        return OrderedDict({"loss_mot_rec": self.loss_mot_rec.item()})

    def update(self):
        self.zero_grad([self.opt_encoder])
        loss_logs = self.backward_G()
        self.loss_mot_rec.backward()
        self.clip_norm([self.encoder])
        self.step([self.opt_encoder])
        return loss_logs

    # ------------------------------------------------------------------------------------------ checkpoints (:269-287)
    def save(self, file_name, ep, total_it):
        state = {"opt_encoder": self.opt_encoder.state_dict(), "ep": ep, "total_it": total_it,
                 "encoder": self._net().state_dict()}
        torch.save(state, file_name)

    def load(self, model_dir):
        checkpoint = torch.load(model_dir, map_location=self.device)
        if getattr(self.opt, "is_train", False):
            self.opt_encoder.load_state_dict(checkpoint["opt_encoder"])
        self._net().load_state_dict(checkpoint["encoder"], strict=True)
        return checkpoint["ep"], checkpoint.get("total_it", 0)
