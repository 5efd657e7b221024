"""Data parallelism for the two places the path shards.

Training — `DataParallel` replaces `MMDistributedDataParallel` over Gloo (codes/tools/train.py:55,78-82): one process per
GPU, gradients averaged across ranks.  The denoiser's backward (autograd.py) writes its gradients into one flat fp32
buffer whose segments become final in a known order (heads | layer L-1 | ... | layer 0 | embeddings); each segment is
handed to `GradReducer.segment_ready` the moment it is final and all-reduced asynchronously (NCCL over
NVLink/NVSwitch runs on its own stream, ordered after the kernels that produced the segment), so the exchange of
layer i overlaps the backward kernels of layers i-1..0.  Parameters outside the denoiser kernels (text encoder,
text_proj, cap_embedding — differentiated by torch.autograd) are reduced as one flat bucket when the last of them
has accumulated its gradient.  No collective is issued in forward; buffers are not broadcast
(`broadcast_buffers=False` in the reference).

Sampling — pairs are independent (both persons of a pair stay on one rank): `shard_pairs` cuts the batch
contiguously, every rank samples its slice with no collective in the loop, `gather_pairs` brings the results to
rank 0 (the reference samples on a single GPU, codes/trainers/mul_ddpm_trainer.py:200-221).

The host logic is backend-agnostic: tests/test_ddp_cpu.py runs it with world_size 2 on `gloo`.
"""
import os

import torch
import torch.distributed as dist
from torch import nn


def _avg_supported(group):
    return dist.get_backend(group) == "nccl"


class PeerGradExchange:
    """Mean of the flat gradient buffer over the ranks WITHOUT an SM-resident collective kernel.

    OPT-IN (HIG_DDP_EXCHANGE=peer); the default is the coalesced NCCL all-reduce.  It was written to test a hypothesis and
    the measurement refuted it.  Hypothesis: every compute kernel of the backward fills the machine (persistent tcgen05 GEMM
    grids with one CTA per SM, attention and LayerNorm backward at their register / shared-memory occupancy limit), so an
    NCCL all-reduce kernel launched beside them either waits for SMs or takes them from a statically partitioned GEMM grid —
    hence 0.6-0.7 of a 0.83 ms all-reduce exposed at N = 2 and 0.9-1.2 of 1.18 ms at N = 8, whatever the number of NCCL
    channels or of SMs set aside for them (profiles/r02_n2_nccl_channels_sm_reservation_sweep.txt).  Result: this exchange
    uses no SMs for the transfers and exposes the SAME time (N = 2: 14.5-14.9 ms per iteration against 14.8-15.0 with NCCL
    and 14.2 without any exchange; N = 8: 15.24 against 15.27 and 14.33).  Its timeline (HIG_DDP_TRACE=1,
    profiles/r02_n2_peer_exchange_timeline.txt) shows why: every segment's exchange ends 0.3-0.5 ms after the segment is
    ready, long before the next one, so the transfers ARE overlapped; what remains is the tail after the last backward kernel
    (layer 0's and the embeddings' segments, 0.26-0.35 ms) and a ~0.4 ms slowdown of the backward kernels themselves while
    850 MB cross NVLink and HBM under the board's power cap.  Neither is a scheduling problem.

    How: the gradient buffer is symmetric memory (train_engine._alloc_flat_grad), every rank maps its peers' buffers over
    NVLink / NVSwitch, and a segment that became final is averaged as reduce-scatter + all-gather made of COPY-ENGINE
    transfers on a side stream:
        barrier                      every rank's segment is final
        pull                         rank r copies slice r of the segment from each peer into a local staging area
        reduce (one small kernel)    slice r = (own + staged) / world, written in place
        barrier                      every slice is reduced
        pull                         rank r copies the reduced slices of the others from their owners
    The copy engines move the bytes while the SMs run the next layer's backward; the only kernels are the signal-pad
    barriers (one block) and the slice reduction.  Every slice is reduced by exactly one rank, so all ranks end up with
    bit-identical gradients.  finish() adds a last barrier (no rank may zero its buffer for the next iteration while a peer
    still reads it) and makes the compute stream wait for the side stream."""

    def __init__(self, fp, group=None):
        import torch.distributed._symmetric_memory as symm
        self.grad = fp.grad
        self.group = group if group is not None else dist.group.WORLD
        self.hdl = symm.rendezvous(self.grad, self.group)
        if self.hdl is None:
            raise RuntimeError("flat gradient buffer is not symmetric memory")
        self.rank, self.world = self.hdl.rank, self.hdl.world_size
        n = self.grad.numel()
        self.peers = [self.grad if p == self.rank else self.hdl.get_buffer(p, (n,), torch.float32, 0)
                      for p in range(self.world)]
        # (alternating two exchange streams so that the small last segment does not queue behind the layer before it was
        #  measured SLOWER: 15.04 vs 14.87 ms per iteration at N = 2)
        self.stream = torch.cuda.Stream(device=self.grad.device, priority=-1)
        self.pull_streams = [torch.cuda.Stream(device=self.grad.device, priority=-1)
                             for _ in range(max(self.world - 1, int(os.environ.get("HIG_DDP_COPY_STREAMS", "6"))))]
        # staging for the slices this rank owns: (world - 1) contributions of at most ceil(n / world) + slack elements
        self.stage = torch.empty((self.world - 1) * (n // self.world + 64 * 8), device=self.grad.device, dtype=torch.float32)
        self.bytes_moved = 0
        self.calls = 0
        self._dirty = False
        # HIG_DDP_TRACE=1: timing events per segment (ready on the compute stream, done on the exchange stream) and at the
        # join; tools/train_step.py prints the last iteration's timeline
        self.trace = [] if os.environ.get("HIG_DDP_TRACE", "0") == "1" else None

    @staticmethod
    def _cuts(lo, hi, world):
        """Slice boundaries of [lo, hi) — multiples of 4 elements (16-byte copies) except the end."""
        n = hi - lo
        per = -(-n // world)
        per = -(-per // 4) * 4
        return [min(hi, lo + k * per) for k in range(world + 1)]

    def _fan_out(self, jobs):
        """jobs[i]: list of (dst, src) copies from peer step i + 1.  The copies are spread over several streams — one copy
        engine moves ~130 GB/s over NVLink, far below the link — forked from and joined back into the exchange stream; with
        few peers a large copy is cut into pieces so that every stream has work."""
        flat = [c for todo in jobs for c in todo]
        if not flat:
            return
        ns = len(self.pull_streams)
        pieces = max(1, ns // max(1, len(flat)))
        work = []
        for dst, src in flat:
            n = dst.numel()
            k = pieces if n >= pieces * (1 << 18) else 1           # pieces of at least 1 MB
            step = -(-(-(-n // k)) // 4) * 4
            for a in range(0, n, step):
                work.append((dst[a:a + step], src[a:a + step]))
        fork = torch.cuda.Event()
        fork.record(self.stream)
        for q, st in enumerate(self.pull_streams):
            mine = work[q::ns]
            if not mine:
                continue
            st.wait_event(fork)
            with torch.cuda.stream(st):
                for dst, src in mine:
                    dst.copy_(src)
            done = torch.cuda.Event()
            done.record(st)
            self.stream.wait_event(done)

    def segment_ready(self, ranges):
        """`ranges`: [(lo, hi)] element ranges of the flat buffer that are final on this rank."""
        from . import ops
        r, W = self.rank, self.world
        cur = torch.cuda.current_stream(self.grad.device)
        ev = torch.cuda.Event(enable_timing=self.trace is not None)
        ev.record(cur)
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            self.hdl.barrier(channel=0, timeout_ms=60000)
            if self.trace is not None:
                b1 = torch.cuda.Event(enable_timing=True)
                b1.record(self.stream)
            off, owned = 0, []
            pulls = [[] for _ in range(W - 1)]
            for lo, hi in ranges:
                cuts = self._cuts(lo, hi, W)
                a, b = cuts[r], cuts[r + 1]
                n = b - a
                if n <= 0:
                    continue
                st = self.stage[off:off + (W - 1) * n].view(W - 1, n)
                off += (W - 1) * n
                for i in range(1, W):
                    pulls[i - 1].append((st[i - 1], self.peers[(r - i) % W][a:b]))
                owned.append((a, b, st))
                self.bytes_moved += (W - 1) * n * 4
            self._fan_out(pulls)
            for a, b, st in owned:
                ops.mean_slices(self.grad[a:b], st, 1.0 / W)
            self.hdl.barrier(channel=0, timeout_ms=60000)
            pulls = [[] for _ in range(W - 1)]
            for lo, hi in ranges:
                cuts = self._cuts(lo, hi, W)
                for i in range(1, W):
                    p = (r - i) % W
                    a, b = cuts[p], cuts[p + 1]
                    if b > a:
                        pulls[i - 1].append((self.grad[a:b], self.peers[p][a:b]))
                        self.bytes_moved += (b - a) * 4
            self._fan_out(pulls)
            if self.trace is not None:
                done = torch.cuda.Event(enable_timing=True)
                done.record(self.stream)
                self.trace.append(("segment", ev, b1, done))
        self.calls += 1
        self._dirty = True

    def finish(self):
        if not self._dirty:
            return
        cur = torch.cuda.current_stream(self.grad.device)
        if self.trace is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record(cur)                       # the last backward kernel is done
        with torch.cuda.stream(self.stream):
            self.hdl.barrier(channel=0, timeout_ms=60000)
        cur.wait_stream(self.stream)
        if self.trace is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(cur)                       # the exchange has joined: the optimizer may start
            self.trace.append(("join", e0, e0, e1))
        self._dirty = False


class GradReducer:
    """Asynchronous mean all-reduce of gradient segments + one trailing bucket of 'other' parameters."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self._pending = []          # (work, tensor, needs_div)
        self._peer = {}             # id(FlatParams) -> PeerGradExchange (or False: unavailable)
        self._px_active = None
        self.bytes_reduced = 0
        self.calls = 0

    def segment_ready(self, index, flat):
        """`flat` (a contiguous 1-D view) is final: start its all-reduce; the result is the mean over ranks."""
        if self.world == 1:
            return
        if _avg_supported(self.group):
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._pending.append((work, flat, False))
        else:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, flat, True))
        self.bytes_reduced += flat.numel() * flat.element_size()
        self.calls += 1

    def _peer_exchange(self, fp):
        """The copy-engine exchange for this FlatParams (created on first use: a collective rendezvous), or None."""
        if fp is None or not getattr(fp, "grad_symmetric", False) or not _avg_supported(self.group):
            return None
        px = self._peer.get(id(fp))
        if px is None:
            try:
                px = PeerGradExchange(fp, self.group)
            except Exception as e:  # noqa: BLE001 — every rank fails alike (same software, same topology): NCCL path
                import warnings
                warnings.warn(f"hig_b200: peer-memory gradient exchange unavailable ({type(e).__name__}: {e}); using NCCL")
                px = False
            self._peer[id(fp)] = px
        return px or None

    def segments_ready(self, index, flats, fp=None):
        """Several disjoint ranges became final together (a layer's own parameters + its share of the stylization
        emb-linears).  Symmetric gradient buffer: the copy-engine exchange (PeerGradExchange); otherwise ONE coalesced NCCL
        launch (ncclGroupStart/End) instead of one per range."""
        flats = [f for f in flats if f.numel()]
        if self.world == 1 or not flats:
            return
        px = self._peer_exchange(fp)
        if px is not None:
            base = fp.grad.storage_offset()
            px.segment_ready([(f.storage_offset() - base, f.storage_offset() - base + f.numel()) for f in flats])
            self.bytes_reduced += sum(f.numel() * f.element_size() for f in flats)
            self.calls += 1
            self._px_active = px
            return
        if len(flats) == 1 or not _avg_supported(self.group) or os.environ.get("HIG_DDP_COALESCE", "1") == "0":
            for f in flats:
                self.segment_ready(index, f)
            return
        with dist._coalescing_manager(self.group, async_ops=True) as cm:
            for f in flats:
                dist.all_reduce(f, op=dist.ReduceOp.AVG, group=self.group)
        self._pending.append((cm, None, False))
        self.bytes_reduced += sum(f.numel() * f.element_size() for f in flats)
        self.calls += 1

    def finish(self):
        """Make the current stream wait for every outstanding all-reduce (no host synchronisation on NCCL)."""
        for work, flat, needs_div in self._pending:
            work.wait()
            if needs_div:
                flat.div_(self.world)
        self._pending = []
        if self._px_active is not None:
            self._px_active.finish()
            self._px_active = None

    def reduce_params(self, params):
        """Mean all-reduce of p.grad for `params` as ONE flat bucket (blocking on the stream, not the host)."""
        grads = [p.grad for p in params if p.grad is not None]
        if self.world == 1 or not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        self.segment_ready(-1, flat)
        self.finish()
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()


class DataParallel(nn.Module):
    """`DataParallel(encoder)` — exposes `.module` like torch DDP (DDPMMulTrainer.backward_G reads
    `self.encoder.module.two_embed`, trainers/mul_ddpm_trainer.py:225)."""

    def __init__(self, module, process_group=None, broadcast_parameters=True):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("hig_b200.ddp.DataParallel needs torch.distributed.init_process_group first")
        self.module = module
        self.group = process_group
        self.reducer = GradReducer(process_group)
        if broadcast_parameters and self.reducer.world > 1:
            with torch.no_grad():
                for p in module.parameters():
                    dist.broadcast(p.data, src=0, group=process_group)
        # the denoiser's flat gradient segments (autograd.py / train_engine.py)
        module._grad_segment_hook = self.reducer.segment_ready
        module._grad_segments_hook = self.reducer.segments_ready
        module._grad_finish_hook = self.reducer.finish
        # Persistent GEMM grids own every SM, so an NCCL kernel launched beside them only runs in the gaps between kernels
        # (measured at N = 2: 0.70 ms of a 0.82 ms all-reduce exposed).  Reserve a few SMs for NCCL while training in parallel:
        # HIG_DDP_NCCL_SMS=n (default 0 = no reservation: measured at N = 2, reserving 8 or 16 SMs changed nothing — 16.42 / 16.47 / 16.50 ms).  Graphs captured afterwards use the reduced grids.
        import os
        reserve = int(os.environ.get("HIG_DDP_NCCL_SMS", "0"))
        if self.reducer.world > 1 and reserve > 0 and torch.cuda.is_available() and next(module.parameters()).is_cuda:
            from . import ops
            sms = torch.cuda.get_device_properties(next(module.parameters()).device).multi_processor_count
            ops.set_sm_limit(max(2, sms - reserve))
            self.sm_limit = sms - reserve
        # everything torch.autograd differentiates outside the denoiser kernels
        self._other = []
        if hasattr(module, "temporal_decoder_blocks"):
            from .autograd import denoiser_param_names
            mine = {n for seg in denoiser_param_names(module) for n in seg}
        else:
            mine = set()
        for n, p in module.named_parameters():
            if p.requires_grad and n not in mine:
                self._other.append(p)
                p.register_post_accumulate_grad_hook(self._other_hook)
        self._other_seen = 0

    def _other_hook(self, p):
        self._other_seen += 1
        if self._other_seen == len(self._other):
            self._other_seen = 0
            self.reducer.reduce_params(self._other)

    def forward(self, *args, **kwargs):
        self._other_seen = 0
        return self.module(*args, **kwargs)


# ---------------------------------------------------------------------------------------------- sampling shards
def shard_pairs(n_pairs, world, rank):
    """Contiguous [lo, hi) slice of the pairs for `rank`; sizes differ by at most one."""
    base, extra = divmod(n_pairs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_pairs(local, n_pairs, dim_pose, group=None, dst=0):
    """local: list of [motion1 [T,C], motion2 [T,C]] for this rank's slice (T may differ between pairs).
    Returns the full list in pair order on `dst`, None elsewhere.  One collective, after sampling has finished."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_pairs + world - 1) // world
    dev = local[0][0].device if local else torch.device("cpu")
    t_loc = max([m.shape[0] for pair in local for m in pair], default=0)
    t_all = torch.tensor([t_loc], device=dev, dtype=torch.long)
    dist.all_reduce(t_all, op=dist.ReduceOp.MAX, group=group)
    Tm = int(t_all.item())
    buf = torch.zeros(per, 2, Tm, dim_pose, device=dev)
    lens = torch.zeros(per, dtype=torch.long, device=dev)
    for i, pair in enumerate(local):
        lens[i] = pair[0].shape[0]
        for j in (0, 1):
            buf[i, j, :pair[j].shape[0]] = pair[j]
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    lls = [torch.empty_like(lens) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst, group=group)
    dist.gather(lens, lls, dst=dst, group=group)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        lo, hi = shard_pairs(n_pairs, world, r)
        for i in range(hi - lo):
            T = int(lls[r][i])
            out.append([bufs[r][i, 0, :T], bufs[r][i, 1, :T]])
    return out


class HostGather:
    """Results of a sharded sampling run straight into ONE host buffer shared by the ranks of a node: a /dev/shm file mapped
    by every rank and registered with CUDA (cudaHostRegister), so each rank's device -> host copy is an asynchronous DMA
    into its own slice — no NCCL gather to rank 0, no pageable staging, no single-rank D2H funnel (round 1: dist.gather of
    211 MB + one pageable .cpu() cost 3.3 % of the 8-GPU end-to-end rate).  `dst` reads the whole buffer after a barrier."""

    def __init__(self, n_pairs, T, dim_pose, group=None, tag="gather"):
        import mmap
        import os
        self.group = group
        self.world, self.rank = (dist.get_world_size(group), dist.get_rank(group)) if dist.is_initialized() else (1, 0)
        self.shape = (n_pairs, 2, T, dim_pose)
        nbytes = max(4 * n_pairs * 2 * T * dim_pose, 4096)
        port = os.environ.get("MASTER_PORT", "0")
        self.path = f"/dev/shm/hig_b200_{tag}_{port}_{n_pairs}x{T}x{dim_pose}.bin"
        if self.rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        if self.world > 1:
            dist.barrier(group=group)
        self._f = open(self.path, "r+b")
        self._mm = mmap.mmap(self._f.fileno(), nbytes)
        self.host = torch.frombuffer(self._mm, dtype=torch.float32, count=n_pairs * 2 * T * dim_pose).view(self.shape)
        self._registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.host.data_ptr(), nbytes, 0)
            self._registered = int(rc) == 0
        self._nbytes = nbytes

    def put(self, lo, pairs_dev):
        """pairs_dev [n, 2, T, C] (device) -> host rows [lo, lo + n); asynchronous on the current stream."""
        self.host[lo:lo + pairs_dev.shape[0]].copy_(pairs_dev, non_blocking=self._registered)

    def finish(self, dst=0):
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)
        return self.host if self.rank == dst else None

    def close(self):
        import os
        try:
            if self._registered:
                torch.cuda.cudart().cudaHostUnregister(self.host.data_ptr())
            self.host = None
            self._mm.close()
            self._f.close()
            if self.world > 1:
                dist.barrier(group=self.group)
            if self.rank == 0 and os.path.exists(self.path):
                os.unlink(self.path)
        except Exception:
            pass


def generate_sharded(trainer, caption1, caption2, m_lens, dim_pose, batch_size=512, group=None, dst=0, host_gather=None):
    """`DDPMMulTrainer.generate` over all ranks of `group`: every rank samples its contiguous slice of the pairs.
    host_gather (a HostGather sized for all pairs at a common T): the samples go device -> shared pinned host memory per
    rank, and `dst` gets ONE host tensor [n_pairs, 2, T, C]; otherwise the padded NCCL gather to `dst` (device lists)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_pairs(len(caption1), world, rank)
    if host_gather is not None:
        T = host_gather.shape[2]
        for b0 in range(lo, hi, batch_size):
            b1 = min(b0 + batch_size, hi)
            out = trainer.generate_batch(caption1[b0:b1], caption2[b0:b1], torch.as_tensor(m_lens[b0:b1]), dim_pose)
            n = b1 - b0
            if out.shape[1] != T:
                raise ValueError(f"HostGather was sized for T={T}, this batch sampled T={out.shape[1]}")
            host_gather.put(b0, torch.stack([out[:n], out[n:]], dim=1))
        return host_gather.finish(dst)
    local = trainer.generate(caption1[lo:hi], caption2[lo:hi], m_lens[lo:hi], dim_pose, batch_size=batch_size) \
        if hi > lo else []
    return gather_pairs(local, len(caption1), dim_pose, group=group, dst=dst)


# ---------------------------------------------------------------------------------------------- length-bucketed generation
def plan_buckets(m_lens, batch_size, world=1, num_frames=196):
    """Schedule for generating N pairs of very different lengths (the evaluation driver, datasets/evaluator.py:26-127 ->
    trainer.generate over the whole test split): the reference pads every chunk of 512 to its longest motion, so with
    shuffled lengths every batch runs at ~196 frames.  Here the pairs are sorted by length, cut into batches of at most
    `batch_size`, and the batches are dealt to the ranks longest-processing-time-first on cost = T_batch * pairs (the
    denoiser is linear in tokens).  Deterministic: every rank computes the same plan.
    Returns per rank a list of (T_batch, [pair indices])."""
    lens = [max(1, min(int(v), num_frames)) for v in torch.as_tensor(m_lens).reshape(-1).tolist()]
    order = sorted(range(len(lens)), key=lambda i: (-lens[i], i))
    batches = [order[i:i + batch_size] for i in range(0, len(order), batch_size)]
    plan = [[] for _ in range(world)]
    load = [0] * world
    for b in sorted(batches, key=lambda b: (-lens[b[0]] * len(b), b[0])):
        r = min(range(world), key=lambda r: (load[r], r))
        plan[r].append((lens[b[0]], b))
        load[r] += lens[b[0]] * len(b)
    return plan


def gather_indexed(local, indices, n_pairs, dim_pose, group=None, dst=0):
    """local[k] = [motion1 [T_k,C], motion2 [T_k,C]] is pair indices[k]; returns the list of all n_pairs in index order on
    `dst` (None elsewhere).  One padded gather after sampling has finished; no collective inside the sampling loop."""
    if not (dist.is_available() and dist.is_initialized()):
        out = [None] * n_pairs
        for i, pair in zip(indices, local):
            out[i] = pair
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = local[0][0].device if local else torch.device("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    meta = torch.tensor([len(local), max([m.shape[0] for pair in local for m in pair], default=0)], device=dev)
    dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=group)
    per, Tm = int(meta[0]), int(meta[1])
    buf = torch.zeros(per, 2, Tm, dim_pose, device=dev)
    info = torch.full((per, 2), -1, dtype=torch.long, device=dev)      # (pair index, rows)
    for k, (i, pair) in enumerate(zip(indices, local)):
        info[k, 0], info[k, 1] = i, pair[0].shape[0]
        for j in (0, 1):
            buf[k, j, :pair[j].shape[0]] = pair[j]
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    infos = [torch.empty_like(info) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst, group=group)
    dist.gather(info, infos, dst=dst, group=group)
    if rank != dst:
        return None
    out = [None] * n_pairs
    for r in range(world):
        for k in range(per):
            i, T = int(infos[r][k, 0]), int(infos[r][k, 1])
            if i >= 0:
                out[i] = [bufs[r][k, 0, :T], bufs[r][k, 1, :T]]
    return out


def generate_bucketed(trainer, caption1, caption2, m_lens, dim_pose, batch_size=512, group=None, dst=0, pair_noise=None):
    """Length-bucketed, rank-sharded `DDPMMulTrainer.generate`: same arguments, returns (on `dst`; everywhere when
    torch.distributed is not initialised) the list of [motion1, motion2] in the callers' order, each trimmed to its own
    length — what the evaluator keeps anyway (datasets/evaluator.py:94-97)."""
    on = dist.is_available() and dist.is_initialized()
    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if on else (1, 0)
    lens = torch.as_tensor(m_lens).reshape(-1)
    if not (len(caption1) == len(caption2) == lens.numel()):
        raise ValueError("caption1, caption2 and m_lens must have one entry per pair")
    plan = plan_buckets(lens, batch_size, world, trainer._net().num_frames)
    trainer.encoder.eval()
    local, indices = [], []
    for T_b, idx in plan[rank]:
        ml = lens[idx]
        extra = {} if pair_noise is None else {"pair_noise": pair_noise[idx]}
        out = trainer.generate_batch([caption1[i] for i in idx], [caption2[i] for i in idx], ml, dim_pose, **extra)
        B = len(idx)
        for k, i in enumerate(idx):
            n = max(1, min(int(ml[k]), out.shape[1]))
            local.append([out[k, :n], out[B + k, :n]])
            indices.append(i)
    return gather_indexed(local, indices, len(caption1), dim_pose, group=group, dst=dst)
