"""Host-side logic of the multi-GPU paths on CPU, world_size 2, gloo backend:
gradient-segment reducer, the trailing 'other parameters' bucket, parameter broadcast, pair sharding and the final
gather.  The kernels themselves need a GPU (tests/test_train_gpu.py); nothing here launches one."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn_name, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world)
        ret[rank] = "ok"
    except Exception as ex:  # noqa
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn_name, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)


# ------------------------------------------------------------------------------------------------ workers
def _w_segments(rank, world):
    import hig_b200  # noqa: F401
    from hig_b200.ddp import GradReducer
    red = GradReducer()
    flat = torch.arange(100, dtype=torch.float32) * (rank + 1)
    bounds = [(0, 10), (10, 64), (64, 100)]
    for i, (lo, hi) in enumerate(bounds):     # segments become final one after the other
        red.segment_ready(i, flat[lo:hi])
    red.finish()
    want = torch.arange(100, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    assert torch.allclose(flat, want)
    assert red.calls == 3 and red.bytes_reduced == 400
    # several ranges that became final together (a layer's parameters + its share of the emb-linears): on gloo every range
    # is reduced on its own (NCCL coalesces them into one launch); a FlatParams without a symmetric gradient buffer — or
    # none at all — takes this path too
    flat2 = torch.arange(60, dtype=torch.float32) * (rank + 2)
    red.segments_ready(7, [flat2[0:16], flat2[40:60], flat2[20:20]], None)
    red.finish()
    mean_scale = sum(r + 2 for r in range(world)) / world
    assert torch.allclose(flat2[0:16], torch.arange(0, 16, dtype=torch.float32) * mean_scale)
    assert torch.allclose(flat2[40:60], torch.arange(40, 60, dtype=torch.float32) * mean_scale)
    assert torch.equal(flat2[16:40], torch.arange(16, 40, dtype=torch.float32) * (rank + 2))      # untouched


def _w_other_bucket_and_broadcast(rank, world):
    import hig_b200  # noqa: F401
    from hig_b200.ddp import DataParallel
    torch.manual_seed(100 + rank)             # different init per rank: construction must broadcast rank 0's
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    ddp = DataParallel(net)
    ref = [torch.empty_like(p) for p in net.parameters()]
    for r, p in zip(ref, net.parameters()):
        r.copy_(p.data)
        dist.broadcast(r, src=0)
        assert torch.equal(r, p.data)
    assert hasattr(ddp, "module") and ddp.module is net
    for it in range(2):                       # two iterations: the hook counter must re-arm
        net.zero_grad()
        torch.manual_seed(7 + rank + 10 * it)
        x = torch.randn(4, 5)
        ddp(x).pow(2).sum().backward()
        got = [p.grad.clone() for p in net.parameters()]
        # expected: mean over ranks of the local gradients
        net.zero_grad()
        net(x).pow(2).sum().backward()
        for g, p in zip(got, net.parameters()):
            loc = p.grad.clone()
            dist.all_reduce(loc)
            assert torch.allclose(g, loc / world, atol=1e-6)


def _w_sampling_shards(rank, world):
    import hig_b200  # noqa: F401
    from hig_b200.ddp import gather_pairs, shard_pairs
    n = 7
    cover = []
    for r in range(world):
        lo, hi = shard_pairs(n, world, r)
        cover += list(range(lo, hi))
    assert cover == list(range(n))
    lo, hi = shard_pairs(n, world, rank)
    # pair i has T = 3 + i frames; motion values encode (pair, person)
    local = [[torch.full((3 + i, 5), float(10 * i)), torch.full((3 + i, 5), float(10 * i + 1))] for i in range(lo, hi)]
    out = gather_pairs(local, n, 5)
    if rank == 0:
        assert len(out) == n
        for i, (a, b) in enumerate(out):
            assert a.shape == (3 + i, 5) and b.shape == (3 + i, 5)
            assert float(a[0, 0]) == 10 * i and float(b[-1, -1]) == 10 * i + 1
    else:
        assert out is None


class _FakeNet:
    num_frames = 196


class _FakeTrainer:
    """Host-logic stand-in for DDPMMulTrainer: generate_batch returns [2B, T, C] with (caption id, person, T) encoded."""
    encoder = torch.nn.Identity()

    def __init__(self):
        self.batches = []

    def _net(self):
        return _FakeNet()

    def generate_batch(self, c1, c2, m_lens, dim_pose):
        T = int(min(int(torch.as_tensor(m_lens).max()), 196))
        self.batches.append((T, len(c1)))
        out = torch.zeros(2 * len(c1), T, dim_pose)
        for k, (a, b) in enumerate(zip(c1, c2)):
            out[k, :, 0], out[len(c1) + k, :, 0] = float(a), float(b)
            out[k, :, 1] = out[len(c1) + k, :, 1] = float(T)
        return out


def _w_bucketed_generation(rank, world):
    import hig_b200  # noqa: F401
    from hig_b200.ddp import generate_bucketed, plan_buckets
    rs = torch.Generator().manual_seed(5)
    n = 37
    lens = torch.randint(20, 240, (n, 1), generator=rs)
    c1, c2 = list(range(n)), [1000 + i for i in range(n)]
    tr = _FakeTrainer()
    out = generate_bucketed(tr, c1, c2, lens, 6, batch_size=8)
    plan = plan_buckets(lens, 8, world)
    assert [(T, len(ix)) for T, ix in plan[rank]] == tr.batches          # this rank ran exactly its planned batches
    if rank == 0:
        assert len(out) == n
        for i, (a, b) in enumerate(out):
            want = min(int(lens[i]), 196)
            assert a.shape == (want, 6) and b.shape == (want, 6)          # trimmed to the pair's own length
            assert float(a[0, 0]) == i and float(b[0, 0]) == 1000 + i     # caller's order, persons not swapped
            assert float(a[0, 1]) >= want                                # its batch was long enough
    else:
        assert out is None


# ------------------------------------------------------------------------------------------------ tests
def test_plan_buckets_covers_balances_and_saves_frames():
    import hig_b200  # noqa: F401
    from hig_b200.ddp import plan_buckets
    g = torch.Generator().manual_seed(1)
    lens = torch.randint(20, 200, (1000,), generator=g)
    for world in (1, 2, 8):
        plan = plan_buckets(lens, 64, world)
        seen = sorted(i for r in plan for _, ix in r for i in ix)
        assert seen == list(range(1000))
        for r in plan:
            for T, ix in r:
                assert len(ix) <= 64 and T == min(int(lens[ix].max()), 196) and T == min(int(lens[ix[0]]), 196)
        cost = [sum(T * len(ix) for T, ix in r) for r in plan]
        assert max(cost) - min(cost) <= 196 * 64                          # LPT: within one batch of each other
        # frames actually sampled vs the reference's chunking in caller order (every chunk padded to its longest)
        ref = sum(min(int(lens[i:i + 64].max()), 196) * len(lens[i:i + 64]) for i in range(0, 1000, 64))
        assert sum(cost) < 0.65 * ref
    assert plan_buckets(torch.tensor([[5], [300]]), 4, 1)[0] == [(196, [1, 0])]   # [N,1] lengths, clamped to num_frames


def test_bucketed_generation_single_process():
    import hig_b200  # noqa: F401
    from hig_b200.ddp import generate_bucketed
    tr = _FakeTrainer()
    out = generate_bucketed(tr, [3, 4, 5], [30, 40, 50], [100, 20, 60], 4, batch_size=2)
    assert [a.shape[0] for a, _ in out] == [100, 20, 60] and tr.batches == [(100, 2), (20, 1)]
    assert [float(a[0, 0]) for a, _ in out] == [3, 4, 5] and [float(b[0, 0]) for _, b in out] == [30, 40, 50]


def test_bucketed_generation_world2():
    _run("_w_bucketed_generation")


def test_gradient_segments_are_averaged_world2():
    _run("_w_segments")


def test_other_parameter_bucket_and_broadcast_world2():
    _run("_w_other_bucket_and_broadcast")


def test_sampling_shards_and_gather_world2():
    _run("_w_sampling_shards")


def test_flat_gradient_layout_covers_every_denoiser_parameter():
    """autograd._Grads: L+2 segments in backward-completion order, adjacent regions where a fused kernel writes several
    parameters at once (Q|K|V weights, LayerNorm weight|bias, all stylization emb-linears)."""
    import hig_b200  # noqa: F401
    from hig_b200.autograd import _Grads, denoiser_param_names
    from hig_b200.interaction_transformer import MotionInteractionTransformer
    m = MotionInteractionTransformer(263, num_frames=196, num_layers=2, cap_id=True)
    segs = denoiser_param_names(m)
    assert len(segs) == 2 + 2
    names = [n for s in segs for n in s]
    assert len(names) == len(set(names))
    outside = {n for n, _ in m.named_parameters()} - set(names)
    assert outside == {"cap_embedding", "text_proj.0.weight", "text_proj.0.bias"}
    gr = _Grads(m, "cpu")
    assert gr.flat.numel() == sum(p.numel() for n, p in m.named_parameters() if n in set(names))
    assert gr.seg_bounds[0][0] == 0 and gr.seg_bounds[-1][1] == gr.flat.numel()
    for (lo, hi), (lo2, _) in zip(gr.seg_bounds, gr.seg_bounds[1:]):
        assert hi == lo2
    p = "temporal_decoder_blocks.1.sa_block."
    r = gr.region(p + "query.weight", 3 * 512 * 512, (1536, 512))
    r[512:1024].fill_(2.0)
    assert float(gr.views[p + "key.weight"].min()) == 2.0 and float(gr.views[p + "query.weight"].max()) == 0.0
    r = gr.region(p + "norm.weight", 1024, (1024,))
    r[512:].fill_(3.0)
    assert float(gr.views[p + "norm.bias"].min()) == 3.0
    first = segs[-1][0]
    n_styl = 2 * 4
    r = gr.region(first, n_styl * 1024 * 2048, (n_styl * 1024, 2048))
    r[1024 * 5:1024 * 6].fill_(5.0)
    assert float(gr.views["temporal_decoder_blocks.1.ca_block.proj_out.emb_layers.1.weight"].min()) == 5.0


def test_peer_exchange_slices_partition_every_range():
    """PeerGradExchange._cuts: every element of a gradient range belongs to exactly one rank, slices start on 16-byte
    boundaries relative to the range start (copy-engine transfers), empty slices are allowed for tiny ranges."""
    import hig_b200  # noqa: F401
    from hig_b200.ddp import PeerGradExchange
    for lo, hi in ((0, 1), (8, 8 + 13), (1000, 1000 + 4096), (64, 64 + 9_437_187), (5, 5 + 7)):
        for world in (2, 3, 8):
            cuts = PeerGradExchange._cuts(lo, hi, world)
            assert len(cuts) == world + 1 and cuts[0] == lo and cuts[-1] == hi
            assert all(a <= b for a, b in zip(cuts, cuts[1:]))
            assert all((c - lo) % 4 == 0 or c == hi for c in cuts)
            sizes = [b - a for a, b in zip(cuts, cuts[1:])]
            assert sum(sizes) == hi - lo and max(sizes) <= -(-(hi - lo) // world) + 3
