"""NCCL data-parallel check (run under torchrun, >= 2 GPUs): gradients after DataParallel backward equal the mean of
the ranks' local gradients for EVERY parameter, are identical on all ranks, and sharded sampling gathers to rank 0."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import hig_b200  # noqa: F401
    import weights
    from hig_b200.ddp import DataParallel, generate_sharded
    from hig_b200.interaction_transformer import MotionInteractionTransformer
    from hig_b200.mul_ddpm_trainer import DDPMMulTrainer
    L, S, T = 2, 8, 40
    m = MotionInteractionTransformer(263, num_frames=196, num_layers=L, cap_id=True)
    m.load_state_dict(weights.make_state_dict(seed=0, num_layers=L), strict=True)
    m = m.to(dev).train()
    inp = weights.make_inputs(100 + rank, S, T, n_text=1)
    tgt = weights.make_noise(200 + rank, 0, S, T)[0].to(dev)
    g = lambda k: inp[k].to(dev)

    def run(net):
        net.zero_grad()
        out = net(g("x"), g("t"), length=g("length"), text=[g("cap1"), g("cap2")])
        ((out - tgt) ** 2).mean().backward()
        return {n: p.grad.clone() for n, p in m.named_parameters()}

    local_g = run(m)
    ddp = DataParallel(m)
    got = run(ddp)
    worst = 0.0
    for n in local_g:
        want = local_g[n].clone()
        dist.all_reduce(want)
        want /= world
        err = ((got[n] - want).norm() / want.norm().clamp_min(1e-20)).item()
        if want.norm() > 1e-6:
            worst = max(worst, err)
        same = got[n].clone()
        dist.broadcast(same, src=0)
        assert torch.equal(same, got[n]), f"{n} differs across ranks"
    # bf16 atomics reorder fp32 sums between the two backward runs: tolerance, not equality
    assert worst < 2e-3, worst
    assert ddp.reducer.calls >= L + 2
    # sharded sampling: 6 pairs over the ranks, 50-step schedule, gathered on rank 0
    opt = argparse.Namespace(device=dev, multi=True, label_path=None, cap_id=True, diffusion_steps=50, is_train=False)
    tr = DDPMMulTrainer(opt, m.eval())
    n = 6
    out = generate_sharded(tr, list(range(n)), list(range(n, 2 * n)), torch.tensor([30, 24, 30, 12, 30, 18]), 263)
    if rank == 0:
        assert len(out) == n and all(a.shape == b.shape and a.shape[1] == 263 for a, b in out)
        assert all(torch.isfinite(a).all() and torch.isfinite(b).all() for a, b in out)
        print(f"ddp_check ok: world={world}, worst grad rel err {worst:.2e}, {ddp.reducer.calls} all-reduces, "
              f"{ddp.reducer.bytes_reduced / 2**20:.0f} MiB")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
