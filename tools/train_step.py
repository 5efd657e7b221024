"""BASELINE config 4: DDP training step, 128 pairs per GPU (S=256 sequences, T=91 rows as the dataset crops them,
datasets/mul_dataset.py:186-201), synthetic two-person motion + captions, labelled mode (forward_twice=False).
    python tools/train_step.py [--pairs 128] [--frames 91] [--iters 10] [--denoiser-only] [--pit]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step.py ...
Prints one JSON line: ms per iteration (CUDA events, max over ranks), pairs/s over all ranks, phase split."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=128)
    ap.add_argument("--frames", type=int, default=91)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cprofile", action="store_true", help="cProfile the timed loop on rank 0 (host-side cost per call site)")
    ap.add_argument("--sync-debug", action="store_true", help="warn (with a stack) on every host-synchronising CUDA call in the timed loop")
    ap.add_argument("--denoiser-only", action="store_true", help="cap_id model: no CLIP / text-encoder forward")
    ap.add_argument("--pit", action="store_true", help="unlabelled (PIT) mode: 4B sequences per iteration")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="fused: hig_b200.optim.FusedAdam + fused loss (product path); torch: reference sequence on torch.optim.Adam")
    ap.add_argument("--profile", action="store_true", help="cudaProfilerStart/Stop around the timed iterations (ncu --profile-from-start off)")
    ap.add_argument("--no-reduce", action="store_true", help="DataParallel with the gradient all-reduce disabled (exposed-comm A/B)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import hig_b200  # noqa: F401
    from hig_b200 import _lib
    from hig_b200.interaction_transformer import MotionInteractionTransformer
    from hig_b200.mul_ddpm_trainer import DDPMMulTrainer
    torch.manual_seed(0)
    m = MotionInteractionTransformer(263, num_frames=196, num_layers=8, latent_dim=512, cap_id=args.denoiser_only)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if not name.startswith("clip.") and p.abs().max() == 0 and "norm.bias" not in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    m = m.to(dev)
    enc = m
    if world > 1:
        from hig_b200.ddp import DataParallel
        enc = DataParallel(m)
        if args.no_reduce:
            m._grad_segment_hook = None
            m._grad_finish_hook = None
    opt = argparse.Namespace(device=dev, multi=True, label_path=None if args.pit else "labels", cap_id=args.denoiser_only,
                             diffusion_steps=1000, is_train=True)
    tr = DDPMMulTrainer(opt, enc)
    if args.optimizer == "fused":
        from hig_b200.optim import FusedAdam
        tr.opt_encoder = FusedAdam(m, lr=2e-4)
    else:
        tr.opt_encoder = torch.optim.Adam(m.parameters(), lr=2e-4, fused=True)
    tr.train_mode()
    B, T = args.pairs, args.frames
    rs = np.random.RandomState(rank)
    if args.denoiser_only:
        c1, c2 = list(rs.randint(0, 43, B)), list(rs.randint(0, 43, B))
    else:
        pick = rs.randint(0, len(bench.TRAIN_CAPTIONS), B)          # 26 classes x 2 roles, as NTU RGB+D 120's two-person set
        c1 = [bench.TRAIN_CAPTIONS[i][0] for i in pick]
        c2 = [bench.TRAIN_CAPTIONS[i][1] for i in pick]
    # pinned host memory, as hig_b200.datasets.build_dataloader hands the batches over
    batch = (c1, c2, torch.randn(B, T, 263).pin_memory(), torch.randn(B, T, 263).pin_memory(),
             torch.from_numpy(rs.randint(20, 200, B)), None)

    from hig_b200.datasets import DevicePrefetcher

    class _Repeat:        # the same pinned batch over and over, through the prefetcher train() uses (H2D on a side stream)
        def __iter__(self):
            while True:
                yield batch
    feed = iter(DevicePrefetcher(_Repeat(), dev))

    def it():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        tr.forward(next(feed))
        ev[1].record()
        logs = tr.update_async()
        ev[2].record()
        return ev, logs

    for _ in range(args.warmup):
        it()
    torch.cuda.synchronize()
    import gc
    gc.collect()
    gc.freeze()
    if world > 1:
        dist.barrier()
    # host time to ISSUE one iteration while the launch queue is empty (later iterations block on the queue depth: the host
    # runs ~80 ms ahead of the GPU and then moves in lock-step with it)
    import time as _t
    h_ = _t.perf_counter()
    for _ in range(3):
        it()
    host_free_ms = (_t.perf_counter() - h_) * 1e3 / 3
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile:
        torch.cuda.cudart().cudaProfilerStart()
    import time
    if args.sync_debug:
        import traceback
        import warnings
        warnings.simplefilter("always")
        warnings.showwarning = lambda m, c, f, l, *a, **k: (print(f"SYNC: {m}"), traceback.print_stack(limit=12))
        torch.cuda.set_sync_debug_mode(1)
    e0.record()
    h0 = time.perf_counter()
    prof = None
    if args.cprofile and rank == 0:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    evs = [it() for _ in range(args.iters)]
    if prof is not None:
        prof.disable()
        import io
        import pstats
        buf = io.StringIO()
        pstats.Stats(prof, stream=buf).sort_stats("cumulative").print_stats(45)
        print(buf.getvalue()[:9000], file=sys.stderr)
    host_ms = (time.perf_counter() - h0) * 1e3 / args.iters      # time the host needs to ISSUE an iteration (no sync inside)
    e1.record()
    if args.sync_debug:
        torch.cuda.set_sync_debug_mode(0)
    torch.cuda.synchronize()
    if args.profile:
        torch.cuda.cudart().cudaProfilerStop()
    ms = e0.elapsed_time(e1) / args.iters
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    fwd = sum(e[0][0].elapsed_time(e[0][1]) for e in evs) / args.iters
    bwd = sum(e[0][1].elapsed_time(e[0][2]) for e in evs) / args.iters
    S = (4 if args.pit else 2) * B
    fl = bench.flops_per_denoiser_step(S, T) * 3
    if rank == 0:
        print(json.dumps({"workload": f"training step, {B} pairs/GPU x {T} frames, {'PIT' if args.pit else 'labelled'}, "
                                      f"{'denoiser only (cap_id)' if args.denoiser_only else 'with CLIP + text encoder'}",
                          "n_gpus": world, "ms_per_iter": ms, "pairs_per_s": world * B / ms * 1e3,
                          "host_issue_ms_per_iter": host_ms, "host_issue_ms_queue_empty": host_free_ms, "forward_ms": fwd, "backward_plus_adam_ms": bwd, "loss": float(evs[-1][1]["loss_mot_rec"]), "optimizer": args.optimizer,
                          "hig_launches_per_iter": (_lib.launch_count() - l0) / args.iters,
                          "denoiser_fwd_bwd_tflops": fl / (ms * 1e-3) / 1e12,
                          "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1 and os.environ.get("HIG_DDP_TRACE") == "1" and rank == 0:
        for px in getattr(enc.reducer, "_peer", {}).values():
            tr_ = px.trace if px else None
            if not tr_:
                continue
            n_last = next(i for i in range(len(tr_) - 1, -1, -1) if tr_[i][0] == "join")
            start = max(i for i in range(n_last) if tr_[i][0] == "join") + 1 if any(t[0] == "join" for t in tr_[:n_last]) else 0
            t0 = tr_[start][1]
            print("peer exchange timeline of the last iteration (ms after the first segment was ready):", file=sys.stderr)
            for kind, a, b, c in tr_[start:n_last + 1]:
                print(f"  {kind:8s} ready {t0.elapsed_time(a):8.3f}  barrier passed {t0.elapsed_time(b):8.3f}  done {t0.elapsed_time(c):8.3f}",
                      file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
