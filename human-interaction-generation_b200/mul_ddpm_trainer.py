"""DDPMMulTrainer — drop-in for codes/trainers/mul_ddpm_trainer.py:50-341 on the B200 path.

Kept: the constructor contract `DDPMMulTrainer(args, encoder)` (args.device / multi / label_path / cap_id /
diffusion_steps / is_train), `generate`, `generate_batch`, `forward`, `backward_G`, `update`, `save`, `load`,
`train_mode`, `eval_mode`, `to`, `train`.  Sampling goes through GaussianDiffusion.p_sample_loop's CUDA-graph path; the
masked loss follows backward_G (:223-247) — as one fused kernel (hig_masked_mse) when the optimizer is the flat fused Adam
that `train()` builds, as the reference's eager formulation otherwise.  `train()` (:289-341) keeps the reference's control
flow (Adam at opt.lr, --is_continue, log_every, save_latest, save_every_e) and writes its loss curve as JSON lines instead
of a matplotlib figure.  Not re-implemented (out of scope): label_batch / eval_data.
Deliberate fix, documented: `generate` takes caption2 from caption2 for non-final chunks (the reference reads
caption1 there, :212 — flagged as a bug in SURVEY.md §3.1); pass `reference_chunk_bug=True` to reproduce it.
"""
import json
import os
import time
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
    def generate_batch(self, caption1, caption2, m_lens, dim_pose, noise=None, noise_seq=None, pair_noise=None):
        """pair_noise ([pairs, 2, >= T, dim_pose], optional): x_T given PER PAIR (rows beyond the batch's T are dropped) instead
        of one th.randn over the padded batch (:176,192) — makes a pair's sample independent of how pairs are batched."""
        net = self._net()
        m_lens = torch.cat([m_lens, m_lens], dim=0)
        T = int(min(int(m_lens.max()), net.num_frames))
        if pair_noise is not None:
            pn = pair_noise.to(self.device)
            noise = torch.cat([pn[:, 0, :T], pn[:, 1, :T]], dim=0).contiguous()
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

    def generate(self, caption1, caption2, m_lens, dim_pose, batch_size=512, reference_chunk_bug=False, pair_noise=None):
        N = len(caption1)
        self.encoder.eval()
        all_output = []
        for lo in range(0, N, batch_size):
            hi = min(lo + batch_size, N)
            c1 = caption1[lo:hi]
            c2 = caption1[lo:hi] if (reference_chunk_bug and hi < N) else caption2[lo:hi]
            output = self.generate_batch(c1, c2, m_lens[lo:hi], dim_pose,
                                         pair_noise=None if pair_noise is None else pair_noise[lo:hi])
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
        dev = self.device
        motion = torch.cat([motion1.detach().to(dev, non_blocking=True).float(),
                            motion2.detach().to(dev, non_blocking=True).float()], dim=0)
        caption = list(caption1) + list(caption2)
        B, T = motion1.shape[0], motion.shape[1]
        from .staging import stage
        cur_len = stage(torch.as_tensor(m_lens).reshape(-1), dev, torch.long).clamp(max=T)
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
            # the dataset yields caption ids as one-element lists (datasets/mul_dataset.py:214-216), which the default
            # collate turns into [LongTensor[B]]; plain int lists work too.  The reference concatenates them (:561-566).
            ids = lambda seq: torch.cat([torch.as_tensor(c).reshape(-1) for c in seq])
            half = len(caption) // 2
            text = [ids(caption[:half]), ids(caption[half:])]
        else:
            text = caption
        output = self.diffusion.training_losses(model=self.encoder, x_start=motion, t=t,
                                                model_kwargs={"text": text, "length": cur_len},
                                                forward_twice=forward_twice)
        self.real_noise, self.fake_noise = output["target"], output["pred"]
        self.cur_len = cur_len
        self.src_mask = (torch.arange(T, device=dev)[None, :] < cur_len[:, None]).float()   # generate_src_mask (:135-139)

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

    def _fused(self):
        from .optim import FusedAdam
        return isinstance(getattr(self, "opt_encoder", None), FusedAdam) and self.fake_noise.is_cuda

    def update_async(self):
        """One optimisation step without a host synchronisation; returns {'loss_mot_rec': device scalar}.
        Fused path (FusedAdam): hig_masked_mse (loss + d loss/d pred) -> backward graphs -> hig_sumsq + hig_adam_flat."""
        from . import ops
        if not self._fused():
            logs = self.update()
            return OrderedDict((k, torch.as_tensor(v)) for k, v in logs.items())
        self.opt_encoder.zero_grad()
        pred = self.fake_noise
        loss, d_pred = ops.masked_mse(pred.detach().float().contiguous(), self.real_noise.float().contiguous(),
                                      self.cur_len.to(torch.int32), pit=not self.with_label)
        pred.backward(d_pred.to(pred.dtype))
        self.opt_encoder.step(clip_norm=0.5)
        self.loss_mot_rec = loss
        return OrderedDict({"loss_mot_rec": loss})

    def update(self):
        if self._fused():
            return OrderedDict((k, v.item()) for k, v in self.update_async().items())
        self.zero_grad([self.opt_encoder])
        loss_logs = self.backward_G()
        self.loss_mot_rec.backward()
        self.clip_norm([self.encoder])
        self.step([self.opt_encoder])
        return loss_logs

    def train(self, train_dataset, rank, world_size):
        """The reference's epoch loop (:289-341).  The optimizer is the flat fused Adam (same hyper-parameters and the same
        state_dict format as `optim.Adam(self.encoder.parameters(), lr=self.opt.lr)`); losses are accumulated on the device
        and read back every `log_every` iterations only."""
        from .datasets import DevicePrefetcher, build_dataloader
        from .optim import FusedAdam
        opt = self.opt
        self.to(self.device)
        self.opt_encoder = FusedAdam(self.encoder, lr=opt.lr)
        it, cur_epoch = 0, 0
        if getattr(opt, "is_continue", False):
            cur_epoch, it = self.load(os.path.join(opt.model_dir, "latest.tar"))
        start_time = time.time()
        loader = train_dataset if hasattr(train_dataset, "__iter__") and not hasattr(train_dataset, "__getitem__") else \
            build_dataloader(train_dataset, rank, world_size, samples_per_gpu=opt.batch_size, drop_last=True,
                             workers_per_gpu=getattr(opt, "workers_per_gpu", 4), shuffle=True)
        if torch.device(self.device).type == "cuda":
            loader = DevicePrefetcher(loader, self.device)      # host -> device copies of batch i+1 overlap iteration i
        log_path = getattr(opt, "log_file", None) or os.path.join(getattr(opt, "model_dir", "."), "train_log.jsonl")
        max_iters = getattr(opt, "max_iters", None)
        acc, n_acc = None, 0
        frozen, it0 = False, it
        self.mean_losses = []
        for epoch in range(cur_epoch, opt.num_epochs):
            self.train_mode()
            if hasattr(getattr(loader, "sampler", None), "set_epoch"):
                loader.sampler.set_epoch(epoch)
            for i, batch_data in enumerate(loader):
                self.forward(batch_data)
                loss = self.update_async()["loss_mot_rec"].detach()
                acc = loss.clone() if acc is None else acc + loss
                n_acc += 1
                it += 1
                if not frozen and it - it0 >= 3:
                    # everything the loop keeps alive exists now (graphs, plans, optimizer state: ~10^5 Python objects): take
                    # it out of the cyclic collector's reach, or a full collection stalls the host for ~100 ms every few
                    # hundred iterations — longer than the work queued on the GPU, which then idles
                    import gc
                    gc.collect()
                    gc.freeze()
                    frozen = True
                if it % opt.log_every == 0:
                    mean = float(acc) / n_acc            # the only host synchronisation of the loop
                    acc, n_acc = None, 0
                    if rank == 0:
                        self.mean_losses.append(mean)
                        rec = {"it": it, "epoch": epoch, "inner_iter": i, "loss_mot_rec": mean,
                               "elapsed_s": time.time() - start_time}
                        print("epoch: %3d niter: %6d inner_iter: %4d %.1fs loss_mot_rec: %.4f"
                              % (epoch, it, i, rec["elapsed_s"], mean), flush=True)
                        with open(log_path, "a") as f:
                            f.write(json.dumps(rec) + "\n")
                if it % opt.save_latest == 0 and rank == 0:
                    self.save(os.path.join(opt.model_dir, "latest.tar"), epoch, it)
                if max_iters is not None and it >= max_iters:
                    break
            if rank == 0:
                self.save(os.path.join(opt.model_dir, "latest.tar"), epoch, it)
            if epoch % opt.save_every_e == 0 and rank == 0:
                self.save(os.path.join(opt.model_dir, "ckpt_e%03d.tar" % epoch), epoch, total_it=it)
            if max_iters is not None and it >= max_iters:
                break
        return it

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
