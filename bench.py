#!/usr/bin/env python
"""bench.py — interactions/sec of full-DDPM two-person sampling (BASELINE.json metric) on N B200s.

One bench "step" = one pass of the hot path over one batch: a complete 1000-step DDPM sample of B=64 pairs
(S=128 sequences, T=196 frames, 263 features; BASELINE config 2), i.e. 1000 x (denoiser forward + posterior update).
  value  : whole-job interactions/s with the text state, x_T and lengths already in HBM (CUDA events, max over ranks)
  e2e    : the same through the reference-facing API  DDPMMulTrainer.generate(captions, lengths)  with host inputs
           (caption strings + host lengths) and the sampled motions copied back to host memory
  roofline / cpu_baseline / clocks / gpu_launches: see the module docstrings of the helpers below.
  train  : BASELINE configs[3], the DDP training step (128 pairs per GPU, T = 91, labelled, synthetic motion + caption ids):
           ms per iteration through DDPMMulTrainer.forward / update on the graph-replayed training engine, achieved
           denoiser fwd+bwd TFLOP/s against the sustained bf16 peak; at N > 1 also the NCCL gradient-equality check, the same
           loop with the all-reduce disabled (exposed communication) and the all-reduce alone (overlap fraction)
  reference_gpu_eager : the reference's formulation (oracle port: eager PyTorch fp32, TF32 off) on the SAME B200 at the bench
           shape — the "what a user gets by running the reference on this GPU" comparator
`--impl reference` times the reference's algorithm on the host CPU (the oracle port, all host threads).
Multi-GPU: one process per GPU (torchrun), pairs sharded across ranks, no collective in the loop.  N = 1 runs configs[1]
(64 pairs); N > 1 runs configs[2], a FIXED batch of 512 pairs split over the N GPUs (strong scaling; 64 pairs per GPU at
N = 8).  The e2e arm lands every rank's samples in one shared pinned host buffer (ddp.HostGather): no gather collective.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(pairs=64, frames=196, feats=263, diffusion_steps=1000, layers=8, latent=512, ff=1024, heads=8,
           text_tokens=77, text_dim=256)

CAPTIONS = [
    ("a person pushes the other person", "a person is pushed by the other person"),
    ("a person kicks the other person", "a person is kicked by the other person"),
    ("a person pats the other person on the back", "a person is patted on the back by the other person"),
    ("a person points a finger at the other person", "a person is pointed at by the other person"),
    ("a person hugs the other person", "a person hugs the other person"),
    ("a person gives an object to the other person", "a person receives an object from the other person"),
    ("a person touches the pocket of the other person", "a person has the pocket touched by the other person"),
    ("two people shake hands", "two people shake hands"),
    ("a person walks towards the other person", "a person walks towards the other person"),
    ("a person punches the other person", "a person is punched by the other person"),
]


def flops_per_denoiser_step(S, T, L=8, D=512, F=1024, hd=64, E=2048, Dt=256, N=77, C=263):
    """Algorithmic FLOPs of one reference denoiser forward (SURVEY.md §8d formula; multiply-add = 2)."""
    tok = S * T
    per_layer = 2 * tok * (11 * D * D + 2 * D * F) + 10 * tok * D * hd + S * (4 * N * Dt * D + 2 * N * D * hd) \
        + 16 * S * E * D
    return L * per_layer + 4 * S * T * C * D + 2 * S * (D * E + E * E)


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            d = json.load(open(f))
            p.update({k: d[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in d})
            p["source"] = "measured"
        except Exception:
            pass
    return p


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(device, precision="bf16"):
    """Random-init denoiser of the repo-default architecture with a random-init CLIP-shaped text encoder; the
    reference zero-initialises 34 projections (identity network), so those are re-drawn N(0, 0.02^2)."""
    import hig_b200  # noqa: F401
    from hig_b200.interaction_transformer import MotionInteractionTransformer
    torch.manual_seed(0)
    m = MotionInteractionTransformer(CFG["feats"], num_frames=CFG["frames"], num_layers=CFG["layers"],
                                     latent_dim=CFG["latent"], cap_id=False, precision=precision)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if not name.startswith("clip.") and p.abs().max() == 0 and "norm.bias" not in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return m.to(device).eval()


def make_trainer(model, device, steps):
    from hig_b200.mul_ddpm_trainer import DDPMMulTrainer
    opt = argparse.Namespace(device=device, multi=True, label_path=None, cap_id=False, diffusion_steps=steps,
                             is_train=False)
    return DDPMMulTrainer(opt, model)


def time_qkv_kernel(device, peak_burst):
    """The dominant kernel alone: the LayerNorm-folded SA/IC QKV projection [25088 x 1536 x 512] exactly as the step
    launches it (hig_gemm_stream, HIG_GS_LN_QSM -> resident-W CTA-pair tcgen05 kernel, query block softmaxed in the epilogue).  30 back-to-back launches
    between two CUDA events on the launching stream, operands rotating through 6 sets (617 MB, 5x the L2) so that no
    launch finds its input or output lines cached.  `traffic` = DRAM bytes per launch of the same kernel from the
    committed ncu --set full capture (profiles/r01d_ncu_full_summary.txt: 30.6 MB read + 31.1 MB written, cold L2)."""
    from hig_b200 import ops
    M, N, K, R = CFG["pairs"] * 2 * CFG["frames"], 1536, 512, 6
    a = [torch.randn(M, K, device=device).half() for _ in range(R)]
    w = (torch.randn(N, K, device=device) / 22.6).half()
    b = torch.randn(N, device=device)
    wsum = w.float().sum(1).contiguous()
    stats = torch.empty(M, 8, device=device)
    ops.row_stats(a[0], stats)
    o = [torch.empty(M, N, device=device, dtype=torch.bfloat16) for _ in range(R)]
    run = lambda i: ops.gemm_stream(ops.GS_LN_QSM, a[i % R], w, b, o[i % R], wsum=wsum, stats_in=stats, ln_width=K)
    for i in range(R):
        run(i)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3 / 30)
    t = sorted(ts)[len(ts) // 2]
    ach = 2.0 * M * N * K / t / 1e12
    tr = kernel_traffic()
    return {"name": "gemm_wres_kernel<LN_QSM> (QKV 25088x1536x512, LayerNorm folded, query softmax)", "us": t * 1e6, "achieved": ach,
            "peak": peak_burst, "unit": "TFLOP/s", "frac": ach / peak_burst,
            "l2": "30 back-to-back launches over 6 rotating operand sets (617 MB > L2)",
            "traffic": tr["bytes_per_launch"], "traffic_source": tr["source"]}


def kernel_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of THIS kernel kind
    (profiles/dominant_kernel_traffic.json, refreshed whenever the kernel changes; it names the .csv it was read from).
    ncu cannot run inside the timed bench, so the number is a citation, not a live measurement — null if the file is gone."""
    f = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    try:
        d = json.load(open(f))
        return {"bytes_per_launch": float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]),
                "source": f"profiles/dominant_kernel_traffic.json <- {d.get('source')}"}
    except Exception as ex:  # noqa
        return {"bytes_per_launch": None, "source": f"unavailable: {ex}"}


def cpu_reference_arm(state_dict, B, T, n_denoiser_steps, warm):
    """The reference's algorithm on the host cores: oracle port (CPU fp32, all threads); one step = one denoiser
    forward + posterior update at the bench batch.  Returns seconds per denoiser step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import denoiser_oracle as DO
    import diffusion_oracle as DF
    torch.set_num_threads(os.cpu_count() or 1)
    S = 2 * B
    g = torch.Generator().manual_seed(0)
    sd = {k: v.detach().float().cpu() for k, v in state_dict.items() if not k.startswith(("clip.", "textTrans", "text_pre", "text_ln"))}
    x = torch.randn(S, T, CFG["feats"], generator=g)
    xf_proj = torch.randn(S, 4 * CFG["latent"], generator=g) * 0.5
    xf_out = torch.randn(S, CFG["text_tokens"], CFG["text_dim"], generator=g)
    length = torch.full((S,), T, dtype=torch.long)
    sch = DF.Schedule(CFG["diffusion_steps"])
    times = []
    with torch.no_grad():
        for i in range(warm + n_denoiser_steps):
            t = torch.full((S,), CFG["diffusion_steps"] - 1 - i, dtype=torch.long)
            t0 = time.perf_counter()
            eps = DO.denoiser_forward(sd, x, t, length, xf_proj, xf_out)
            x = DF.p_sample_step(sch, x, eps, t, torch.randn(x.shape, generator=g))
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    return sum(times) / len(times), times


def gpu_eager_reference(model, device, B, T, n_steps=3):
    """The reference's own formulation on this GPU: the oracle port (plain eager PyTorch, fp32, TF32 off — what running
    line/Human-Interaction-Generation's denoiser + posterior step unmodified on a B200 costs), `n_steps` denoiser steps at the
    bench shape, CUDA events.  It is a comparator, not the product: nothing of hig_b200's kernels is on this path."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import denoiser_oracle as DO
    import diffusion_oracle as DF
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        S = 2 * B
        g = torch.Generator(device=device).manual_seed(0)
        sd = {k: v.detach().float() for k, v in model.state_dict().items()
              if not k.startswith(("clip.", "textTrans", "text_pre", "text_ln"))}
        x = torch.randn(S, T, CFG["feats"], device=device, generator=g)
        xf_proj = torch.randn(S, 4 * CFG["latent"], device=device, generator=g) * 0.5
        xf_out = torch.randn(S, CFG["text_tokens"], CFG["text_dim"], device=device, generator=g)
        length = torch.full((S,), T, dtype=torch.long, device=device)
        sch = DF.Schedule(CFG["diffusion_steps"])
        ev = []
        with torch.no_grad():
            for i in range(n_steps + 1):
                t = torch.full((S,), CFG["diffusion_steps"] - 1 - i, dtype=torch.long, device=device)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eps = DO.denoiser_forward(sd, x, t, length, xf_proj, xf_out)
                x = DF.p_sample_step(sch, x, eps, t, torch.randn(x.shape, device=device, generator=g))
                e1.record()
                ev.append((e0, e1))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev[1:])[len(ev[1:]) // 2]
        return {"us_per_denoiser_step": ms * 1e3, "interactions_per_s": B / (ms * 1e-3 * CFG["diffusion_steps"]),
                "what": f"oracle port of the reference denoiser + posterior step, eager PyTorch fp32 (TF32 off) on this GPU, "
                        f"median of {n_steps} steps at {B} pairs x {T} frames, extrapolated to {CFG['diffusion_steps']} steps; "
                        f"the reference's per-step host mask loop (0.5 s at S=1024) and its text encoder are not included"}
    except Exception as ex:  # noqa
        return {"error": f"{type(ex).__name__}: {ex}"}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


def train_block(device, rank, world, iters, pk):
    """BASELINE configs[3]: the DDP training step — 128 pairs per GPU x 91 frames (the rows the dataset produces,
    datasets/mul_dataset.py:186-201), labelled mode, caption ids (cap_id: denoiser only), through
    DDPMMulTrainer.forward / update_async on the graph-replayed engine with the fused loss / clip / Adam kernels.
    N > 1: NCCL gradient-equality check first, then the same loop with the all-reduce disabled (exposed communication) and
    the all-reduce of the flat gradient buffer alone."""
    import numpy as np
    import torch.distributed as dist
    from hig_b200 import ops
    from hig_b200.interaction_transformer import MotionInteractionTransformer
    from hig_b200.mul_ddpm_trainer import DDPMMulTrainer
    from hig_b200.optim import FusedAdam
    B, T = 128, 91
    torch.manual_seed(0)
    m = MotionInteractionTransformer(CFG["feats"], num_frames=CFG["frames"], num_layers=CFG["layers"], latent_dim=CFG["latent"],
                                     cap_id=True)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.abs().max() == 0 and "norm.bias" not in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    m = m.to(device)
    enc = m
    if world > 1:
        from hig_b200.ddp import DataParallel
        enc = DataParallel(m)
    opt = argparse.Namespace(device=device, multi=True, label_path="labels", cap_id=True, diffusion_steps=1000, is_train=True)
    tr = DDPMMulTrainer(opt, enc)
    tr.opt_encoder = FusedAdam(m, lr=2e-4)
    tr.train_mode()
    rs = np.random.RandomState(100 + rank)
    gen = torch.Generator().manual_seed(100 + rank)
    batch = (list(rs.randint(0, 43, B)), list(rs.randint(0, 43, B)), torch.randn(B, T, CFG["feats"], generator=gen).pin_memory(),
             torch.randn(B, T, CFG["feats"], generator=gen).pin_memory(), torch.from_numpy(rs.randint(20, 200, B)), None)
    hooks = (getattr(m, "_grad_segment_hook", None), getattr(m, "_grad_finish_hook", None))

    def set_hooks(on):
        m._grad_segment_hook, m._grad_finish_hook = hooks if on else (None, None)

    from hig_b200.datasets import DevicePrefetcher

    class _Repeat:        # the same pinned batch through the prefetcher DDPMMulTrainer.train uses (H2D on a side stream)
        def __iter__(self):
            while True:
                yield batch
    feed = iter(DevicePrefetcher(_Repeat(), device))

    def run(n):
        for _ in range(3):
            tr.forward(next(feed))
            tr.update_async()
        torch.cuda.synchronize()
        import gc
        gc.collect()
        gc.freeze()          # as DDPMMulTrainer.train does after its first iterations: no full collections inside the loop
        if world > 1:
            dist.barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i in range(n):
            tr.forward(next(feed))
            loss = tr.update_async()["loss_mot_rec"]
            evs[i + 1].record()
        torch.cuda.synchronize()
        ms = evs[0].elapsed_time(evs[n]) / n
        # the median iteration: what the comparisons between two runs use (a single allocator or clock hiccup of ~100 ms in
        # one of the runs otherwise decides the sign of their difference)
        med = float(np.median([evs[i].elapsed_time(evs[i + 1]) for i in range(n)]))
        if world > 1:
            tt = torch.tensor([ms, med], device=device, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms, med = tt.tolist()
        medians.append(med)
        return ms, float(loss)

    medians = []

    out = {"workload": f"configs[3]: DDP training step, {B} pairs/GPU x {T} frames, labelled, caption ids (denoiser + loss + "
                       f"clip + Adam; text encoder not on this path)", "iters": iters, "optimizer": "hig_b200.optim.FusedAdam"}
    fp = tr.opt_encoder.fp
    # ---- gradient equality across ranks (what tools/ddp_check.py asserts), on the production engine
    if world > 1:
        def grads(on):
            set_hooks(on)
            np.random.seed(7); torch.manual_seed(7)
            tr.forward(batch)
            tr.opt_encoder.zero_grad()
            _, d_pred = ops.masked_mse(tr.fake_noise.detach().float().contiguous(), tr.real_noise.float().contiguous(),
                                       tr.cur_len.to(torch.int32))
            tr.fake_noise.backward(d_pred)
            torch.cuda.synchronize()
            return fp.grad[:fp.n_den].clone()
        local = grads(False)
        want = local.clone()
        dist.all_reduce(want)
        want /= world
        got = grads(True)
        err = ((got - want).norm() / want.norm().clamp_min(1e-20)).item()
        same = got.clone()
        dist.broadcast(same, src=0)
        identical = bool(torch.equal(same, got))
        flags = torch.tensor([1.0 if (err < 2e-3 and identical) else 0.0], device=device)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        out["ddp_check"] = {"ok": bool(flags.item() > 0), "rel_err_vs_mean_of_local_grads": err,
                            "identical_on_all_ranks": identical,
                            "what": "flat denoiser gradient after the overlapped NCCL all-reduce vs the all-reduced mean of the "
                                    "ranks' local gradients (bf16 split-K atomics reorder fp32 sums between two backward runs: "
                                    "tolerance 2e-3, not equality)"}
        set_hooks(True)
    else:
        out["ddp_check"] = None
    ms, loss = run(iters)
    med_on = medians[-1]
    fl = flops_per_denoiser_step(2 * B, T) * 3
    tf = fl / (ms * 1e-3) / 1e12
    out.update({"ms_per_iter": ms, "ms_per_iter_median": med_on, "pairs_per_s": world * B / ms * 1e3, "loss": loss, "denoiser_fwd_bwd_tflops": tf,
                "roofline_frac": tf / pk["bf16_tflops_sustained"],
                "roofline_basis": f"3 x {fl / 3e9:.1f} algorithmic GFLOP (forward + 2x backward, SURVEY §8d) per iteration and GPU "
                                  f"over the CUDA-event time, against the {pk['source']} sustained bf16 peak",
                "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30})
    if world > 1:
        red = enc.reducer
        b0, c0 = red.bytes_reduced, red.calls
        ms1, _ = run(max(2, iters // 4))
        n_it = max(2, iters // 4) + 3
        out["allreduce_bytes"] = (red.bytes_reduced - b0) / n_it
        out["allreduce_calls_per_iter"] = (red.calls - c0) / n_it
        set_hooks(False)
        ms_off, _ = run(iters)
        med_off = medians[-1]
        set_hooks(True)
        # the all-reduce alone, same buffer, same collective
        flat = fp.grad[:fp.n_den]
        for _ in range(2):
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        e1.record()
        torch.cuda.synchronize()
        ms_ar = e0.elapsed_time(e1) / 5
        exposed = max(med_on - med_off, 0.0)         # medians, see run()
        out.update({"ms_per_iter_allreduce_off": ms_off, "ms_per_iter_allreduce_off_median": med_off,
                    "exposed_allreduce_ms": exposed, "allreduce_alone_ms": ms_ar,
                    "overlap_frac": max(0.0, min(1.0, 1.0 - exposed / ms_ar)) if ms_ar > 0 else None,
                    "allreduce_busbw_gbs": 2 * (world - 1) / world * flat.numel() * 4 / (ms_ar * 1e-3) / 1e9})
        out["nccl_reserved_sms"] = int(os.environ.get("HIG_DDP_NCCL_SMS", "0"))
        out["exchange"] = ("copy-engine reduce-scatter + all-gather over symmetric peer memory (HIG_DDP_EXCHANGE=peer)"
                           if getattr(fp, "grad_symmetric", False) else
                           "NCCL all-reduce (AVG), one coalesced launch per gradient segment, issued between the backward graphs")
        out["exposed_note"] = ("per-segment timeline in profiles/r02_n2_peer_exchange_timeline.txt: the transfers finish long before "
                               "the next segment is ready; the exposed time is the tail after the last backward kernel plus the "
                               "backward kernels running slower while the exchange is in flight (same figure with an SM-free "
                               "copy-engine exchange)")
    else:
        out.update({"allreduce_bytes": 0, "overlap_frac": None})
    # ---- the same step with captions as TEXT: (random-init) CLIP features cached per caption, the trainable 4-layer text
    # encoder through torch.autograd (forward / backward each one CUDA graph), 77 text tokens per caption on the K/V side of
    # every layer's text cross attention, shared between the sequences that carry the same caption
    try:
        out["with_captions"] = train_with_captions(device, rank, world, max(5, iters // 2), B, T)
    except Exception as e:  # noqa: BLE001 — reported, the caption-id numbers above stand on their own
        out["with_captions"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    ops.set_sm_limit(0)      # DataParallel reserved SMs for NCCL: give them back to whatever runs after this block
    return out


# the 26 two-person classes of NTU RGB+D 120 (the reference's dataset, datasets/evaluator.py:133 `range(26)`): one caption per
# role, so a batch of 128 pairs carries at most 52 distinct captions
_ACTIONS = ["punches", "kicks", "pushes", "pats the back of", "points a finger at", "hugs", "gives an object to",
            "touches the pocket of", "shakes hands with", "walks towards", "walks away from", "hits", "threatens",
            "knocks over", "grabs something from", "aims at", "steps on the foot of", "high-fives", "drinks a toast with",
            "carries something with", "takes a photo of", "follows", "whispers to", "exchanges things with", "supports",
            "plays rock-paper-scissors with"]
TRAIN_CAPTIONS = [(f"a person {a} the other person", f"the other person {a} a person") for a in _ACTIONS]


def train_with_captions(device, rank, world, iters, B, T):
    import torch.distributed as dist
    from hig_b200.datasets import DevicePrefetcher
    from hig_b200.interaction_transformer import MotionInteractionTransformer
    from hig_b200.mul_ddpm_trainer import DDPMMulTrainer
    from hig_b200.optim import FusedAdam
    os.environ.setdefault("HIG_CLIP_STUB", "1")
    torch.manual_seed(0)
    m = MotionInteractionTransformer(CFG["feats"], num_frames=CFG["frames"], num_layers=CFG["layers"], latent_dim=CFG["latent"])
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if not name.startswith("clip.") and p.abs().max() == 0 and "norm.bias" not in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    m = m.to(device)
    enc = m
    if world > 1:
        from hig_b200.ddp import DataParallel
        enc = DataParallel(m)
    opt = argparse.Namespace(device=device, multi=True, label_path="labels", cap_id=False, diffusion_steps=1000, is_train=True)
    tr = DDPMMulTrainer(opt, enc)
    tr.opt_encoder = FusedAdam(m, lr=2e-4)
    tr.train_mode()
    gen = torch.Generator().manual_seed(200 + rank)
    rs = __import__("numpy").random.RandomState(200 + rank)
    pick = rs.randint(0, len(TRAIN_CAPTIONS), B)
    batch = ([TRAIN_CAPTIONS[i][0] for i in pick], [TRAIN_CAPTIONS[i][1] for i in pick],
             torch.randn(B, T, CFG["feats"], generator=gen).pin_memory(), torch.randn(B, T, CFG["feats"], generator=gen).pin_memory(),
             torch.from_numpy(rs.randint(20, 200, B)), None)

    class _Repeat:
        def __iter__(self):
            while True:
                yield batch
    feed = iter(DevicePrefetcher(_Repeat(), device))
    for _ in range(4):
        tr.forward(next(feed))
        tr.update_async()
    torch.cuda.synchronize()
    import gc
    gc.collect()
    gc.freeze()
    if world > 1:
        dist.barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    evs[0].record()
    for i in range(iters):
        tr.forward(next(feed))
        loss = tr.update_async()["loss_mot_rec"]
        evs[i + 1].record()
    torch.cuda.synchronize()
    ms = evs[0].elapsed_time(evs[iters]) / iters
    med = float(__import__("numpy").median([evs[i].elapsed_time(evs[i + 1]) for i in range(iters)]))
    if world > 1:
        tt = torch.tensor([ms, med], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, med = tt.tolist()
    return {"workload": f"{B} pairs/GPU x {T} frames, labelled, captions as text ({len(set(batch[0] + batch[1]))} distinct among "
                        f"{2 * B}; text encoder: {m.text_encoder_kind})", "iters": iters, "ms_per_iter": ms,
            "ms_per_iter_median": med, "pairs_per_s": world * B / ms * 1e3, "loss": float(loss)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=None, help="pairs per GPU (default: 64 at N=1 = configs[1]; 512 / N at N>1 = configs[2])")
    ap.add_argument("--train-iters", type=int, default=20, help="timed iterations of the configs[3] training block (0 = skip)")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reference-formulation-on-this-GPU comparator")
    ap.add_argument("--diffusion-steps", type=int, default=CFG["diffusion_steps"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    ngpu = max(world, args.gpus if args.impl == "reference" else world)
    # N = 1: configs[1] (64 pairs).  N > 1: configs[2], a fixed batch of 512 pairs split over the N GPUs (strong scaling).
    strong = ngpu > 1 and args.pairs is None
    CFG["pairs"] = args.pairs if args.pairs is not None else (512 // ngpu if ngpu > 1 else 64)
    CFG["diffusion_steps"] = args.diffusion_steps
    B, T, C, NS = CFG["pairs"], CFG["frames"], CFG["feats"], CFG["diffusion_steps"]
    which = f"configs[2]: batch of {B * ngpu} pairs sharded over {ngpu} GPUs ({B} per GPU)" if ngpu > 1 else "configs[1]"
    workload = {"workload": f"{which}: full {NS}-step DDPM sampling, {B} pairs/GPU x {T} frames x {C} feats, "
                            f"8-layer role-aware denoiser, random-init CLIP-shaped text encoder",
                "pairs_per_gpu": B, "total_pairs": B * ngpu, "frames": T, "diffusion_steps": NS,
                "parallelism": f"batch-shard x{ngpu}",
                "l2": "working set per denoiser step (~1.5 GB of activations + 214 MB weights) exceeds the 126 MB L2"}
    unit = "interactions/s"

    if args.impl == "reference":
        if rank != 0:
            return
        # weights of the same architecture (CPU construction only; no CUDA kernel is involved in this arm)
        import hig_b200  # noqa: F401
        from hig_b200.interaction_transformer import MotionInteractionTransformer
        torch.manual_seed(0)
        m = MotionInteractionTransformer(C, num_frames=T, num_layers=CFG["layers"], latent_dim=CFG["latent"], cap_id=True)
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for name, p in m.named_parameters():
                if p.abs().max() == 0 and "norm.bias" not in name:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
        s_per, times = cpu_reference_arm(m.state_dict(), B, T, args.steps, args.warmup)
        val = B / (s_per * NS)
        cores = os.cpu_count() or 1
        print(json.dumps({
            "impl": "reference", "metric": "interactions/sec (full DDPM sample, 196f)", "value": val, "unit": unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per * NS * 1e3,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload, "extrapolated": True,
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "port", "extrapolated": True,
                             "sample": f"{args.steps} of {NS} denoiser + posterior-update steps at {B} pairs x {T} frames "
                                       f"({s_per:.3f} s each, oracle port of the reference, torch fp32, {cores} host threads), "
                                       f"extrapolated x{NS}; excluded: the text encoder (once per sample) and the reference's "
                                       f"per-step host mask loop — both would make the reference slower"},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hig_b200 product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    from hig_b200 import _lib

    model = build_model(device)
    trainer = make_trainer(model, device, NS)
    diff = trainer.diffusion
    S = 2 * B
    caps1 = [CAPTIONS[(i + rank * B) % len(CAPTIONS)][0] for i in range(B)]
    caps2 = [CAPTIONS[(i + rank * B) % len(CAPTIONS)][1] for i in range(B)]
    m_lens_host = torch.full((B,), T, dtype=torch.long)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: text state, lengths and x_T live in HBM ----------------
    with torch.no_grad():
        xf_proj, xf_out = model.encode_text(caps1 + caps2, device)
    kw = {"xf_proj": xf_proj, "xf_out": xf_out, "length": torch.full((S,), T, device=device, dtype=torch.long)}
    x_T = torch.randn(S, T, C, device=device)

    def resident_step():
        return diff.p_sample_loop(model, (S, T, C), noise=x_T, clip_denoised=False, model_kwargs=kw)

    for _ in range(args.warmup):
        resident_step()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = _lib.launch_count()
    launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        resident_step()
        launches += diff.last_launches
    e1.record()
    barrier()
    clk = clocks.stop()
    dt = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        tt = torch.tensor([dt], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = tt.item()
    value = world * B * args.steps / dt

    # ---------------- end-to-end arm: host captions + host lengths in, host motions out ----------------
    # N > 1: the whole batch's captions / lengths go to hig_b200.ddp.generate_sharded, every rank samples its slice and
    # copies it (asynchronous DMA) into ONE host buffer shared by the ranks (ddp.HostGather: /dev/shm mapping registered with
    # CUDA) — no gather collective, no single-rank D2H funnel; rank 0 holds all samples in host memory after the barrier.
    hg = None
    if world > 1:
        from hig_b200.ddp import HostGather, generate_sharded
        all1 = [CAPTIONS[i % len(CAPTIONS)][0] for i in range(B * world)]
        all2 = [CAPTIONS[i % len(CAPTIONS)][1] for i in range(B * world)]
        all_lens = torch.full((B * world,), T, dtype=torch.long)
        hg = HostGather(B * world, T, C, tag="bench")

    def e2e_step():
        if world > 1:
            return generate_sharded(trainer, all1, all2, all_lens, C, batch_size=512, host_gather=hg)
        out = trainer.generate(caps1, caps2, m_lens_host, C)
        res = torch.stack([torch.stack(p) for p in out])          # [B, 2, T, C] on device
        return res.cpu()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    dt_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt_e2e], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt_e2e = tt.item()
    e2e_val = world * B * args.steps / dt_e2e
    if hg is not None:
        hg.close()
    h2d = (S * 77 * 8 + B * 8) * world            # token ids [S,77] int64 + lengths [B] int64, per rank
    d2h = B * 2 * T * C * 4 * world               # every rank's samples, each by its own DMA

    # ---------------- configs[3]: the DDP training step (all ranks take part) ----------------
    pk = peaks()
    train = None
    if args.train_iters > 0:
        try:
            torch.cuda.empty_cache()
            train = train_block(device, rank, world, args.train_iters, pk)
        except Exception as ex:  # noqa
            import traceback
            train = {"error": f"{type(ex).__name__}: {ex}", "trace": traceback.format_exc()[-1500:]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    f_step = flops_per_denoiser_step(S, T)
    ach = f_step * NS * args.steps / dt / 1e12      # per GPU: each rank runs the same per-GPU workload
    roof = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": ach / pk["bf16_tflops_sustained"], "traffic": None,
            "basis": f"{f_step / 1e9:.1f} algorithmic GFLOP per denoiser step (SURVEY §8d) x {NS} steps per sample, "
                     f"per GPU, over the CUDA-event time of the timed region; peak = {pk['source']} sustained bf16",
            "us_per_denoiser_step": dt / (args.steps * NS) * 1e6}
    try:
        roof["kernel"] = time_qkv_kernel(device, pk["bf16_tflops"])
        roof["traffic"] = roof["kernel"]["traffic"]
    except Exception as ex:  # noqa
        roof["kernel"] = {"error": str(ex)}

    gpu_eager = None
    if not args.no_gpu_eager:
        gpu_eager = gpu_eager_reference(model, device, B, T)

    cpu = None
    if not args.no_cpu_baseline:
        sd = {k: v for k, v in model.state_dict().items()}
        s_per, _ = cpu_reference_arm(sd, B, T, 2, 1)
        cpu = {"value": B / (s_per * NS), "unit": unit, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"2 of {NS} denoiser+posterior steps at {B} pairs x {T} frames on the host CPU "
                         f"({s_per:.2f} s each, oracle port, torch fp32, all threads), extrapolated to {NS} steps",
               "extrapolated": True}

    print(json.dumps({
        "metric": "interactions/sec (full DDPM sample, 196f)", "value": value, "unit": unit, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload, "us_per_denoiser_step": dt / (args.steps * NS) * 1e6,
        "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": ("hig_b200.ddp.generate_sharded(trainer, caption strings, host lengths) -> one pinned host tensor shared "
                        "by the ranks (per-rank D2H DMA, no gather collective)") if world > 1 else
                       "DDPMMulTrainer.generate(caption strings, host lengths) -> host tensors",
                "text_cache": "the frozen-CLIP feature cache is warm (same captions as the warm-up call): the timed text work "
                              "is tokenisation + the 4-layer text encoder + text_proj, not the 12-layer CLIP stack"},
        "gpu_launches": launches, "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "train": train,
        "reference_gpu_eager": gpu_eager}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
