"""TrainEngine — the bf16 training step of the denoiser as replayed CUDA graphs over static buffers.

What it replaces: torch.autograd through MotionInteractionTransformer.forward (codes/models/interaction_transformer.py:
577-616) inside DDPMMulTrainer.forward / update (codes/trainers/mul_ddpm_trainer.py:91-162, 249-256), data-parallel under
tools/train.py:78-82.  Round 1 scheduled ~1 700 launches per iteration from Python (709 of this library + ~1 000 torch
fills / casts / optimizer kernels): 54 ms per iteration of which only 28 ms was GPU work.  This module keeps the same
kernels-only contract (no PyTorch math on the path) and removes the host from the loop:

  * FlatParams: every trainable parameter is a view of ONE flat fp32 buffer (names, shapes and state_dict unchanged), with a
    flat fp32 gradient buffer laid out in backward-completion order (heads | layer L-1 .. 0 | embeddings | non-denoiser
    parameters) and a flat bf16 mirror that IS the GEMM operand set: Q|K|V, key|value and the 4L stylization emb-linears
    are contiguous slices of it, so nothing is concatenated or cast per iteration except the mirror itself — and that is
    written by the fused Adam kernel (hig_adam_flat) in the same sweep as the parameter update.
  * the forward is one captured graph, the backward L + 2 graphs cut at the gradient-segment boundaries, so the
    data-parallel reducer (ddp.py) can start the all-reduce of a segment between two replays while later segments run.
  * weight gradients  dW = dY^T X  and data gradients  dX = dY W  run on the tcgen05 GEMM with MN-major operands
    (hig_gemm_bf16_t): no transposed copies of activations, gradients or weights exist (round 1: 91 transposes per iteration).
  * one memset of the flat gradient buffer replaces the per-tensor zero fills.
"""
import os

import torch

from . import ops
from .autograd import denoiser_param_names

HEAD_DIM = 64


def _rup(n, m):
    return (n + m - 1) // m * m


# ====================================================================================================== flat parameters
def _alloc_flat_grad(n, dev):
    """The flat gradient buffer.  HIG_DDP_EXCHANGE=peer (opt-in, see ddp.PeerGradExchange for why it is not the default):
    under NCCL data parallelism it is allocated as SYMMETRIC memory (same allocation on every rank, mappable by the peers over
    NVLink / NVSwitch) so that the ranks can average it with copy-engine transfers between the backward graphs; anything
    that goes wrong here simply leaves an ordinary tensor (NCCL all-reduce)."""
    import torch.distributed as dist
    if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl"
            and os.environ.get("HIG_DDP_EXCHANGE", "nccl") == "peer"):
        try:
            import torch.distributed._symmetric_memory as symm
            g = symm.empty(n, dtype=torch.float32, device=dev)
            g.zero_()
            return g, True
        except Exception as e:  # noqa: BLE001
            import warnings
            warnings.warn(f"hig_b200: symmetric memory unavailable ({type(e).__name__}: {e}); gradients go through NCCL")
    return torch.zeros(n, device=dev, dtype=torch.float32), False


class FlatParams:
    """fp32 parameters / gradients of a module as flat buffers (+ bf16 mirror); p.data become views (names unchanged)."""

    ALIGN = 8     # elements: every parameter starts on a 32-byte (fp32) / 16-byte (bf16) boundary — TMA's operand rule

    def __init__(self, module):
        named = dict(module.named_parameters())
        self.segments = denoiser_param_names(module)
        den = [n for seg in self.segments for n in seg]
        mine = set(den)
        self.other_names = [n for n, p in named.items() if n not in mine and p.requires_grad]
        dev = named[den[0]].device
        if dev.type != "cuda":
            raise RuntimeError("hig_b200: training runs on CUDA only (no CPU fallback); move the module to a GPU")
        self.offsets, self.seg_bounds = {}, []
        off = 0
        for seg in self.segments + [self.other_names]:
            lo = off
            for n in seg:
                off = _rup(off, self.ALIGN)
                self.offsets[n] = off
                off += named[n].numel()
            off = _rup(off, self.ALIGN)
            self.seg_bounds.append((lo, off))
        self.n_den = self.seg_bounds[len(self.segments) - 1][1]
        self.n_total = off
        self.other_bounds = self.seg_bounds.pop()
        # What leaves with each gradient segment (data-parallel all-reduce ranges).  The emb-linears of the StylizationBlocks
        # live in the LAST segment's region (one contiguous forward operand), but a layer's share of them is final when the
        # layer's backward is: it is handed out with that layer (train_engine._Plan._backward).
        L = module.num_layers
        last = self.segments[-1]
        n_styl = sum(1 for n in last if n.endswith("emb_layers.1.weight"))
        npl = n_styl // L
        per_w = named[last[0]].numel()                    # 2D * E
        per_b = named[last[n_styl]].numel()               # 2D
        w0, b0 = self.offsets[last[0]], self.offsets[last[n_styl]]
        rest0 = self.offsets[last[2 * n_styl]]
        self.seg_ranges = [[self.seg_bounds[0]]]
        for k in range(1, L + 1):
            li = L - k
            self.seg_ranges.append([self.seg_bounds[k], (w0 + li * npl * per_w, w0 + (li + 1) * npl * per_w),
                                    (b0 + li * npl * per_b, b0 + (li + 1) * npl * per_b)])
        self.seg_ranges.append([(rest0, self.seg_bounds[-1][1])])
        self.param = torch.zeros(self.n_total, device=dev, dtype=torch.float32)
        self.grad, self.grad_symmetric = _alloc_flat_grad(self.n_total, dev)
        self.mirror = torch.zeros(self.n_den, device=dev, dtype=torch.bfloat16)
        self.names = den + self.other_names
        self.shapes = {n: tuple(named[n].shape) for n in self.names}
        with torch.no_grad():
            for n in self.names:
                p = named[n]
                v = self.view32(n)
                v.copy_(p.data)
                p.data = v
        self.gviews = {n: self._view(self.grad, n) for n in self.names}
        self.params = [named[n] for n in self.names]          # the Parameter objects, in flat order (identity is stable)
        self.n_den_params = len(den)
        self._all_trainable = [p for p in module.parameters() if p.requires_grad]
        self._mirror_key = None
        self.direct = False          # True: gradients are exposed as persistent p.grad views (FusedAdam path)

    def _view(self, flat, n):
        o = self.offsets[n]
        shape = self.shapes[n]
        k = 1
        for d in shape:
            k *= d
        return flat[o:o + k].view(shape)

    def view32(self, n):
        return self._view(self.param, n)

    def view16(self, n):
        return self._view(self.mirror, n)

    def region(self, flat, first, count, shape):
        o = self.offsets[first]
        return flat[o:o + count].view(shape)

    def owns(self, module):
        """True while the module's parameters are still the views this object created (a later .to() / .cuda() / re-creation
        of a parameter breaks that)."""
        base = self.param.data_ptr()
        for i in (0, self.n_den_params - 1, len(self.names) - 1):
            n, p = self.names[i], self.params[i]
            if p.data_ptr() != base + 4 * self.offsets[n] or tuple(p.shape) != self.shapes[n]:
                return False
        return getattr(module, "_hig_flat", None) is self

    def refresh_mirror(self, module, force=False):
        """bf16 operand mirror <- fp32 parameters (one cast kernel) when a parameter changed behind our back (a third-party
        optimizer, load_state_dict, ...).  hig_adam_flat writes the mirror itself and calls mark_mirror_fresh()."""
        key = (getattr(module, "_hig_param_generation", 0), sum(p._version for p in self._all_trainable))
        if force or key != self._mirror_key:
            ops.act_fwd(self.param[:self.n_den], ops.ACT_NONE, self.mirror)
            self._mirror_key = key
            return True
        return False

    def mark_mirror_fresh(self, module):
        module._hig_param_generation = getattr(module, "_hig_param_generation", 0) + 1
        self._mirror_key = (module._hig_param_generation, sum(p._version for p in self._all_trainable))


def flat_params(module):
    fp = getattr(module, "_hig_flat", None)
    if fp is None or not fp.owns(module):
        fp = FlatParams(module)
        module._hig_flat = fp
        module._hig_train_engine = None
    return fp


# ====================================================================================================== operand set
def build_operands(module, fp, eng):
    """The W dictionary of denoiser_engine.DenoiserEngine.packed(), as VIEWS: bf16 weights into the mirror, fp32 biases /
    LayerNorm parameters into the live parameter buffer.  Only the K-padded motion-embedding operand and the positional
    table are separate buffers (refresh_special)."""
    D, E, C = eng.D, eng.E, eng.C
    dev = fp.param.device
    W = {}
    w16, f32 = fp.view16, fp.view32
    r16 = lambda first, rows, cols: fp.region(fp.mirror, first, rows * cols, (rows, cols))
    r32 = lambda first, n: fp.region(fp.param, first, n, (n,))
    half = D // 2
    import math
    W["freqs"] = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(dev)
    W["te0.w"], W["te0.b"] = w16("time_embed.0.weight"), f32("time_embed.0.bias")
    W["te2.w"], W["te2.b"] = w16("time_embed.2.weight"), f32("time_embed.2.bias")
    subs = [("sa", "sa_block"), ("ca", "ca_block")] + ([] if module.no_cross_attn else [("ic", "int_ca_block")])
    Dt = module.text_latent_dim
    n_styl = 0
    for i in range(module.num_layers):
        p, mp = f"l{i}.", f"temporal_decoder_blocks.{i}."
        for name, sub in subs:
            q = mp + sub + "."
            W[p + name + ".ln.w"], W[p + name + ".ln.b"] = f32(q + "norm.weight"), f32(q + "norm.bias")
            if name == "ca":
                W[p + "ca.tln.w"], W[p + "ca.tln.b"] = f32(q + "text_norm.weight"), f32(q + "text_norm.bias")
                W[p + "ca.q.w"], W[p + "ca.q.b"] = w16(q + "query.weight"), f32(q + "query.bias")
                W[p + "ca.kv.w"], W[p + "ca.kv.b"] = r16(q + "key.weight", 2 * D, Dt), r32(q + "key.bias", 2 * D)
            else:
                W[p + name + ".qkv.w"] = r16(q + "query.weight", 3 * D, D)
                W[p + name + ".qkv.b"] = r32(q + "query.bias", 3 * D)
        q = mp + "ffn."
        W[p + "ffn.w1"], W[p + "ffn.b1"] = w16(q + "linear1.weight"), f32(q + "linear1.bias")
        W[p + "ffn.w2"], W[p + "ffn.b2"] = w16(q + "linear2.weight"), f32(q + "linear2.bias")
        for name, sub in subs + [("ffn", "ffn")]:
            q = mp + sub + ".proj_out."
            W[p + name + ".po.ln.w"], W[p + name + ".po.ln.b"] = f32(q + "norm.weight"), f32(q + "norm.bias")
            W[p + name + ".po.w"], W[p + name + ".po.b"] = w16(q + "out_layers.2.weight"), f32(q + "out_layers.2.bias")
            W[p + name + ".ss"] = n_styl
            n_styl += 1
    first_w = fp.segments[-1][0]
    first_b = fp.segments[-1][n_styl]
    W["emb.w"] = r16(first_w, n_styl * 2 * D, E)
    W["emb.b"] = r32(first_b, n_styl * 2 * D)
    W["n_styl"] = n_styl
    W["out.w"], W["out.b"] = w16("out.weight"), f32("out.bias")
    W["out2.w"], W["out2.b"] = w16("out2.weight"), f32("out2.bias")
    W["in.w"] = torch.zeros(D, eng.CP, device=dev, dtype=torch.bfloat16)
    W["in.pos"] = torch.zeros(module.num_frames, D, device=dev, dtype=torch.float32)
    W["zero.b"] = {}
    return W


def refresh_special(module, W, C):
    """[joint_embed.W | joint_embed2.W] K-padded operand and the positional rows (:593-602), in place."""
    with torch.no_grad():
        W["in.w"][:, :C].copy_(module.joint_embed.weight)
        W["in.w"][:, C:C + 4].copy_(module.joint_embed2.weight)
        pos = W["in.pos"]
        pos[0].copy_(module.joint_embed2.bias)
        torch.add(module.sequence_embedding[:module.num_frames - 1], module.joint_embed.bias[None], out=pos[1:])


# ====================================================================================================== one shape's plan
class _Plan:
    """Static buffers + captured graphs of one (S, T, N) training shape."""

    def __init__(self, te, S, T, N, U=None):
        # U: rows of the text side when the batch's captions were deduplicated (xf_out holds U distinct captions, text_idx maps
        # every sequence to its caption) — None: one text row per sequence
        self.te, self.S, self.T, self.N, self.U = te, S, T, N, U
        eng, dev = te.eng, te.fp.param.device
        self.x = torch.zeros(S, T, eng.C, device=dev)
        self.t = torch.zeros(S, device=dev, dtype=torch.int64)
        self.len = torch.zeros(S, device=dev, dtype=torch.int32)
        self.xf_proj = torch.zeros(S, eng.E, device=dev)
        self.xf_out = torch.zeros(U or S, N, te.module.text_latent_dim, device=dev)
        self.text_idx = torch.zeros(S, device=dev, dtype=torch.int64) if U else None
        self.d_eps = torch.zeros(S, T, eng.C, device=dev)
        self.saved = None
        self.fwd_graph = None
        self.bwd_graphs = None
        self.results = None
        self.use_graph = os.environ.get("HIG_TRAIN_GRAPH", "1") != "0"
        self.launches_fwd = self.launches_bwd = 0

    # ------------------------------------------------------------------------------------------ forward schedule
    def _forward(self):
        te = self.te
        eng, W = te.eng, te.W
        S, T, N = self.S, self.T, self.N
        D, F_, E, H, L = eng.D, eng.F, eng.E, eng.H, eng.L
        dev, tok = self.x.device, S * T
        bf, f32 = torch.bfloat16, torch.float32
        new = lambda *shape, dtype=bf: torch.empty(*shape, device=dev, dtype=dtype)
        G = ops.gemm
        st = type("Saved", (), {})()
        # ---- embedding MLP (:591) and every StylizationBlock's (scale | shift) (:88-90)
        st.temb = ops.timestep_embed(self.t, W["freqs"], new(S, D))
        st.h0 = new(S, E)
        G(st.temb, W["te0.w"], bias=W["te0.b"], out_bf16=st.h0)
        st.te_h = ops.act_fwd(st.h0, ops.ACT_SILU, new(S, E))
        st.emb = new(S, E, dtype=f32)
        G(st.te_h, W["te2.w"], bias=W["te2.b"], out_f32=st.emb, residual=self.xf_proj)
        st.semb = ops.act_fwd(st.emb, ops.ACT_SILU, new(S, E))
        st.ss = new(S, W["n_styl"] * 2 * D, dtype=f32)
        G(st.semb, W["emb.w"], bias=W["emb.b"], out_f32=st.ss)
        # ---- text K/V side of every layer's cross attention (:155-161)
        # (a caption's K/V side does not depend on the motion: with deduplicated captions it is computed once per distinct
        #  caption — SU rows instead of S — and its A matrices are handed to the sequences by index)
        Dt = self.xf_out.shape[2]
        SU = self.U or S
        st.xf = self.xf_out.view(SU * N, Dt)
        st.tn, st.kv, st.a_text = [], [], []
        for i in range(L):
            p = f"l{i}.ca."
            tn = ops.ln_film_silu(st.xf, W[p + "tln.w"], W[p + "tln.b"], new(SU * N, Dt))
            kv = new(SU * N, 2 * D)
            G(tn, W[p + "kv.w"], bias=W[p + "kv.b"], out_bf16=kv)
            a = new(SU, H, HEAD_DIM, HEAD_DIM)
            ops.eff_attn(ops.ATTN_KV_ONLY, SU, N, H, k=kv[:, :D], v=kv[:, D:], a_out=a)
            if self.U:
                a = torch.index_select(a, 0, self.text_idx)
            st.tn.append(tn); st.kv.append(kv); st.a_text.append(a)
        # ---- motion embedding (:593-602)
        st.xa = torch.zeros(tok, eng.CP, device=dev, dtype=bf)
        ops.pack_motion(self.x, st.xa)
        xres = new(tok, D, dtype=f32)
        G(st.xa, W["in.w"], residual=W["in.pos"], res_row_mod=T, out_f32=xres)
        st.blocks = []

        def project(blk, sact, xres_in, p, want_xb):
            """x += out_layers(sact) (:92-97 + the block's residual add); xb = bf16 copy for the next block's GEMM operand."""
            xres_out = new(tok, D, dtype=f32)
            xb = new(tok, D) if want_xb else None
            G(sact, W[p + ".po.w"], bias=W[p + ".po.b"], residual=xres_in, out_f32=xres_out, out_bf16=xb)
            return xres_out, xb

        def ss_of(p):
            i = W[p + ".ss"]
            return i, st.ss[:, i * 2 * D:(i + 1) * 2 * D]

        fused_attn = os.environ.get("HIG_TRAIN_FUSED_ATTN", "1") != "0"
        # HIG_TRAIN_FUSED_FFN=1: linear1's epilogue writes h1 and GELU(h1), linear2's input-gradient GEMM applies GELU'(h1).
        # Correct (bit-identical forward) and SLOWER: 16.5 vs 16.0 ms per iteration on the same box — the K = 512 GEMMs are
        # paced by their 8 epilogue warps, and an erf / exp per element there costs more than the 143 MB round trip of the
        # 64-warps-per-SM elementwise kernels it removes.  Off by default.
        fused_ffn = os.environ.get("HIG_TRAIN_FUSED_FFN", "0") == "1" and tok >= 512     # CTA-pair GEMM shapes only
        xb = None
        kinds = ["sa", "ca"] + (["ic"] if eng.has_ic else [])
        for li in range(L):
            p = f"l{li}."
            for kind in kinds:
                blk = {"kind": kind, "xres_in": xres, "li": li}
                n = ops.ln_film_silu(xres, W[p + kind + ".ln.w"], W[p + kind + ".ln.b"], new(tok, D))
                y, sact = new(tok, D), new(tok, D)
                i_ss, ss = ss_of(p + kind)
                if kind == "ca":
                    qv = new(tok, D)
                    G(n, W[p + "ca.q.w"], bias=W[p + "ca.q.b"], out_bf16=qv)
                    a_blk = st.a_text[li]
                    blk.update(n=n, q=qv)
                else:
                    qkv = new(tok, 3 * D)
                    G(n, W[p + kind + ".qkv.w"], bias=W[p + kind + ".qkv.b"], out_bf16=qkv)
                    qv, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
                    blk.update(n=n, qkv=qkv)
                    a_blk = None
                if fused_attn:
                    # K/V half, then the query half fused with LayerNorm + FiLM + SiLU (the sampling path's kernels); the
                    # attention output y is written too — the LayerNorm backward reads it
                    if a_blk is None:
                        a_blk = new(S, H, HEAD_DIM, HEAD_DIM)
                        ops.attn_kv(k, v, a_blk, S, T, H, length=self.len, pair_shift=S // 2 if kind == "ic" else 0)
                    ops.attn_apply_stylize(qv, a_blk, W[p + kind + ".po.ln.w"], W[p + kind + ".po.ln.b"], sact, S, T, H,
                                           scale_shift=ss, silu=True, y_out=y)
                else:
                    if kind == "ca":
                        ops.eff_attn(ops.ATTN_Q_ONLY, S, T, H, q=qv, a_in=a_blk, y=y)
                    elif kind == "sa":
                        ops.eff_attn(ops.ATTN_SELF, S, T, H, q=qv, k=k, v=v, y=y, length=self.len, mask_v=True)
                    else:
                        ops.eff_attn(ops.ATTN_INTER, S, T, H, q=qv, k=k, v=v, y=y, length=self.len, pair_shift=S // 2,
                                     mask_v=False)
                    ops.ln_film_silu(y, W[p + kind + ".po.ln.w"], W[p + kind + ".po.ln.b"], sact, rows_per_seq=T,
                                     scale_shift=ss, silu=True)
                blk.update(y=y, sact=sact, ss_index=i_ss)
                xres, xb = project(blk, sact, xres, p + kind, kind == kinds[-1])
                st.blocks.append(blk)
            # FFN (:261-264)
            blk = {"kind": "ffn", "xres_in": xres, "li": li, "xb_in": xb}
            h1, g = new(tok, F_), new(tok, F_)
            if fused_ffn:      # linear1 with both h1 (saved for GELU') and GELU(h1) written by its epilogue
                ops.gemm_fused(xb, W[p + "ffn.w1"], W[p + "ffn.b1"], g, act=ops.ACT_GELU, out_pre=h1)
            else:
                G(xb, W[p + "ffn.w1"], bias=W[p + "ffn.b1"], out_bf16=h1)
                ops.act_fwd(h1, ops.ACT_GELU, g)
            y = new(tok, D)
            G(g, W[p + "ffn.w2"], bias=W[p + "ffn.b2"], out_bf16=y)
            i_ss, ss = ss_of(p + "ffn")
            sact = ops.ln_film_silu(y, W[p + "ffn.po.ln.w"], W[p + "ffn.po.ln.b"], new(tok, D), rows_per_seq=T, scale_shift=ss,
                                    silu=True)
            blk.update(h1=h1, g=g, y=y, sact=sact, ss_index=i_ss)
            xres, xb = project(blk, sact, xres, p + "ffn", True)
            st.blocks.append(blk)
        # ---- output heads (:613-616)
        st.xb_final = xb
        st.eps = new(tok, eng.LD_EPS, dtype=f32)
        G(xb, W["out.w"], bias=W["out.b"], out_f32=st.eps[:, :eng.C])
        G(xb.view(S, T * D)[:, :D], W["out2.w"], bias=W["out2.b"], out_f32=st.eps.view(S, T * eng.LD_EPS)[:, :eng.C])
        return st

    # ------------------------------------------------------------------------------------------ backward schedule
    def _backward(self):
        """Generator: runs the backward kernels and yields after each gradient segment is final (heads, layers L-1..0,
        embeddings).  The last value (StopIteration.value) is (d_xf_proj, d_xf_out)."""
        te, st = self.te, self.saved
        eng, W, fp = te.eng, te.W, te.fp
        S, T, N, C = self.S, self.T, self.N, eng.C
        D, F_, E, H = eng.D, eng.F, eng.E, eng.H
        dev, tok = self.x.device, S * T
        bf, f32 = torch.bfloat16, torch.float32
        new = lambda *shape, dtype=bf: torch.empty(*shape, device=dev, dtype=dtype)
        gv = fp.gviews
        greg = lambda first, count, shape: fp.region(fp.grad, first, count, shape)
        # HIG_DETERMINISTIC=1 (bit-reproducible gradients, a debugging / regression mode): weight gradients in ONE pass per
        # output tile instead of K-slices combined with fp32 atomics, column sums and the gradient norm with one contributor
        # per address (the C side reads the same variable), LayerNorm parameter gradients through per-sequence partials.
        det = os.environ.get("HIG_DETERMINISTIC", "0") not in ("", "0")
        fused_ffn = os.environ.get("HIG_TRAIN_FUSED_FFN", "0") == "1" and tok >= 512

        def zero_bias(n):
            zb = W["zero.b"].get(n)
            if zb is None:
                zb = W["zero.b"][n] = torch.zeros(n, device=dev, dtype=f32)
            return zb

        def dgrad(dy, w, out=None, out_f32=None, accumulate=False):
            """dx[M,K] (+)= dy[M,N] . W[N,K] with W as stored (MN-major B operand)."""
            ops.gemm_t(dy, w, trans_b=True, bias=zero_bias(w.shape[1]), residual=out_f32 if accumulate else None,
                       out_f32=out_f32, out_bf16=out)

        def wgrad(dy, x, w_grad, split=True):
            """w_grad[N,K] += dy[M,N]^T . x[M,K]: both operands token-major as they lie; K = tokens split over CTA pairs
            (fp32 atomics into the zeroed gradient buffer).  split=False: enough output tiles to fill the machine — plain
            stores (the region has a single writer)."""
            ops.gemm_t(dy, x, trans_a=True, trans_b=True, out_f32=w_grad, split_k=-1 if (split and not det) else 0)

        pending_gb = []

        def bcast(region2w, S=S):
            """[2W] parameter-gradient region as a stride-0 [S, 2W] view: ln_film_silu_bwd accumulates every sequence's
            (dgamma | dbeta) partials straight into the parameter gradient.  Deterministic mode: per-sequence partials,
            column-summed in a fixed order by flush_gb()."""
            if not det:
                return region2w.view(1, -1).expand(S, -1)
            part = torch.zeros(S, region2w.numel(), device=dev, dtype=f32)
            pending_gb.append((part, region2w))
            return part

        def flush_gb():
            for part, region in pending_gb:
                ops.colsum(part, region)
            pending_gb.clear()

        fp.grad[:fp.n_den].zero_()
        d_ss = torch.zeros_like(st.ss)
        SU = self.U or S
        d_xf = torch.zeros(SU * N, st.xf.shape[1], device=dev, dtype=f32)

        # ---------------- output heads: eps = out(h[:,1:]) / out2(h[:,0])  (:613-616)
        d_eps = self.d_eps.view(tok, C)
        LDE = _rup(C, 8)
        de_a = torch.zeros(tok, LDE, device=dev, dtype=bf)               # frames >= 1 (frame-0 rows zeroed)
        de_0 = torch.zeros(S, LDE, device=dev, dtype=bf)                  # frame 0 of every sequence
        if det:
            ops.transpose(d_eps, copy=de_a, rows_zero_mod=T)
            ops.transpose(d_eps.view(S, T * C)[:, :C], copy=de_0)
            ops.colsum(de_a[:, :C], gv["out.bias"])
            ops.colsum(de_0[:, :C], gv["out2.bias"])
        else:
            ops.transpose(d_eps, copy=de_a, colsum=gv["out.bias"], rows_zero_mod=T)
            ops.transpose(d_eps.view(S, T * C)[:, :C], copy=de_0, colsum=gv["out2.bias"])
        dres = new(tok, D, dtype=f32)
        ops.gemm_t(de_a[:, :C], W["out.w"], trans_b=True, out_f32=dres)
        ops.gemm_t(de_0[:, :C], W["out2.w"], trans_b=True, out_f32=dres.view(S, T * D)[:, :D])
        wgrad(de_a[:, :C], st.xb_final, gv["out.weight"])
        wgrad(de_0[:, :C], st.xb_final.view(S, T * D)[:, :D], gv["out2.weight"])
        yield 0

        def stylize_project_bwd(blk, pfx, mp):
            dres_c = new(tok, D)
            ops.transpose(dres, copy=dres_c, colsum=gv[mp + "proj_out.out_layers.2.bias"])
            d_sact = new(tok, D)
            dgrad(dres_c, W[pfx + ".po.w"], out=d_sact)
            wgrad(dres_c, blk["sact"], gv[mp + "proj_out.out_layers.2.weight"])
            i = blk["ss_index"]
            ss = st.ss[:, i * 2 * D:(i + 1) * 2 * D]
            d_y = new(tok, D)
            ops.ln_film_silu_bwd(blk["y"], W[pfx + ".po.ln.w"], W[pfx + ".po.ln.b"], d_sact, d_y, T, scale_shift=ss,
                                 silu=True, d_ss=d_ss[:, i * 2 * D:(i + 1) * 2 * D],
                                 d_gb=bcast(greg(mp + "proj_out.norm.weight", 2 * D, (2 * D,))))
            return d_y

        def linear_bwd(dy, x_saved, w, w_grad, b_grad, dx_into=None):
            """b_grad None: the kernel that produced dy already left its column sums in the bias gradient."""
            if b_grad is not None:
                ops.colsum(dy, b_grad)
            dx = None
            if dx_into is not None:
                dgrad(dy, w, out_f32=dx_into, accumulate=True)
            else:
                dx = new(dy.shape[0], w.shape[1])
                dgrad(dy, w, out=dx)
            wgrad(dy, x_saved, w_grad)
            return dx

        def pre_ln_bwd(blk, d_n, pfx, mp):
            ops.ln_film_silu_bwd(blk["xres_in"], W[pfx + ".ln.w"], W[pfx + ".ln.b"], d_n, dres, T, dx_accumulate=True,
                                 d_gb=bcast(greg(mp + "norm.weight", 2 * D, (2 * D,))))

        modname = {"sa": "sa_block.", "ca": "ca_block.", "ic": "int_ca_block.", "ffn": "ffn."}
        n_styl = W["n_styl"]
        npl = n_styl // eng.L                                      # StylizationBlocks per layer (sa, ca, [ic], ffn)
        emb_w_grad = greg(fp.segments[-1][0], n_styl * 2 * D * E, (n_styl * 2 * D, E))
        emb_b_grad = greg(fp.segments[-1][n_styl], n_styl * 2 * D, (n_styl * 2 * D,))
        d_ss_c = new(S, n_styl * 2 * D)
        seg = 1
        for blk in reversed(st.blocks):
            li, kind = blk["li"], blk["kind"]
            pfx = f"l{li}.{kind}"
            mp = f"temporal_decoder_blocks.{li}.{modname[kind]}"
            d_y = stylize_project_bwd(blk, pfx, mp)
            if kind == "ffn":
                if fused_ffn:
                    # d_h1 = (d_y . W2) * GELU'(h1) straight from the input-gradient GEMM's epilogue: d_g never reaches HBM
                    w2 = W[f"l{li}.ffn.w2"]
                    ops.colsum(d_y, gv[mp + "linear2.bias"])
                    d_h1 = ops.gemm_fused(d_y, w2, zero_bias(w2.shape[1]), new(tok, F_), trans_b=True, gate=blk["h1"],
                                          gate_act=ops.ACT_GELU)
                    wgrad(d_y, blk["g"], gv[mp + "linear2.weight"])
                else:
                    d_g = linear_bwd(d_y, blk["g"], W[f"l{li}.ffn.w2"], gv[mp + "linear2.weight"], gv[mp + "linear2.bias"])
                    d_h1 = ops.act_bwd(blk["h1"], d_g, ops.ACT_GELU, new(tok, F_))
                # no pre-norm in the FFN: d(xres_in) = dres (skip path) + d_h1 . W1, accumulated by the GEMM epilogue
                linear_bwd(d_h1, blk["xb_in"], W[f"l{li}.ffn.w1"], gv[mp + "linear1.weight"], gv[mp + "linear1.bias"],
                           dx_into=dres)
            elif kind == "ca":
                d_q = new(tok, D)
                dA = new(S, H, HEAD_DIM, HEAD_DIM, dtype=f32)
                ops.eff_attn_bwd(ops.ATTN_Q_ONLY, S, T, H, q=blk["q"], a_in=st.a_text[li], dy=d_y, dq=d_q, dA=dA,
                                 q_sum=gv[mp + "query.bias"])
                d_n = linear_bwd(d_q, blk["n"], W[f"l{li}.ca.q.w"], gv[mp + "query.weight"], None)
                pre_ln_bwd(blk, d_n, pfx, mp)
                if self.U:       # sequences that share a caption: their dA add up
                    dA = torch.zeros(SU, H, HEAD_DIM, HEAD_DIM, device=dev, dtype=f32).index_add_(0, self.text_idx, dA)
                kv = st.kv[li]
                kv_b = greg(mp + "key.bias", 2 * D, (2 * D,))
                if N == 1 and not det:
                    # one text token (caption ids, :561-566): its time softmax is exactly 1, so dK == 0 and
                    # dV[l] = sum_d dA[d, l] — a reduction of dA instead of a launch of the attention backward that spends
                    # 58 us (ncu) loading 64 x 64 fp32 matrices to multiply them with a column of ones
                    d_kv = torch.zeros(SU, 2 * D, device=dev, dtype=bf)
                    d_kv[:, D:].copy_(dA.view(SU, H, HEAD_DIM, HEAD_DIM).sum(dim=2).view(SU, D))
                    ops.colsum(d_kv[:, D:], kv_b[D:])
                else:
                    d_kv = new(SU * N, 2 * D)
                    ops.eff_attn_bwd(ops.ATTN_KV_ONLY, SU, N, H, k=kv[:, :D], v=kv[:, D:], dk=d_kv[:, :D], dv=d_kv[:, D:],
                                     dA=dA, k_sum=kv_b[:D], v_sum=kv_b[D:])
                Dt = st.xf.shape[1]
                d_tn = linear_bwd(d_kv, st.tn[li], W[f"l{li}.ca.kv.w"], greg(mp + "key.weight", 2 * D * Dt, (2 * D, Dt)), None)
                ops.ln_film_silu_bwd(st.xf, W[pfx + ".tln.w"], W[pfx + ".tln.b"], d_tn, d_xf, N, dx_accumulate=True,
                                     d_gb=bcast(greg(mp + "text_norm.weight", 2 * Dt, (2 * Dt,)), SU))
            else:
                qkv = blk["qkv"]
                d_qkv = new(tok, 3 * D)
                mode = ops.ATTN_SELF if kind == "sa" else ops.ATTN_INTER
                qkv_b = greg(mp + "query.bias", 3 * D, (3 * D,))
                ops.eff_attn_bwd(mode, S, T, H, q=qkv[:, :D], k=qkv[:, D:2 * D], v=qkv[:, 2 * D:], dy=d_y,
                                 dq=d_qkv[:, :D], dk=d_qkv[:, D:2 * D], dv=d_qkv[:, 2 * D:], length=self.len,
                                 pair_shift=S // 2 if kind == "ic" else 0, q_sum=qkv_b[:D], k_sum=qkv_b[D:2 * D],
                                 v_sum=qkv_b[2 * D:])
                d_n = linear_bwd(d_qkv, blk["n"], W[f"l{li}.{kind}.qkv.w"], greg(mp + "query.weight", 3 * D * D, (3 * D, D)),
                                 None)
                pre_ln_bwd(blk, d_n, pfx, mp)
            if kind == "sa":
                # every parameter of layer li is final — including its StylizationBlocks' emb-linears: their (scale | shift)
                # gradient slabs receive nothing from the layers below, so their weight / bias gradients are taken now and
                # leave with this segment instead of waiting for the end of backward (a third of all parameters)
                lo_c, hi_c = li * npl * 2 * D, (li + 1) * npl * 2 * D
                ops.transpose(d_ss[:, lo_c:hi_c], copy=d_ss_c[:, lo_c:hi_c], colsum=emb_b_grad[lo_c:hi_c])
                wgrad(d_ss_c[:, lo_c:hi_c], st.semb, emb_w_grad[lo_c:hi_c], split=False)
                flush_gb()
                yield seg
                seg += 1

        # ---------------- motion embedding (:593-602)
        dres_c = new(tok, D)
        ops.transpose(dres, copy=dres_c)
        w_in = torch.zeros(D, eng.CP, device=dev, dtype=f32)
        wgrad(dres_c, st.xa, w_in)
        gv["joint_embed.weight"].copy_(w_in[:, :C])
        gv["joint_embed2.weight"].copy_(w_in[:, C:C + 4])
        dpos = torch.zeros(T * D, device=dev, dtype=f32)
        ops.colsum(dres.view(S, T * D), dpos)
        dpos = dpos.view(T, D)
        gv["joint_embed2.bias"].copy_(dpos[0])
        if T > 1:
            gv["sequence_embedding"][:T - 1].copy_(dpos[1:])
            ops.colsum(dpos[1:], gv["joint_embed.bias"])

        # ---------------- stylization emb-linears + time-embedding MLP (:88-90, :474-478, :591)
        d_semb = torch.zeros(S, E, device=dev, dtype=f32)      # K = 4L * 1024 is long, the output 8 tiles: split-K
        ops.gemm_t(d_ss_c, W["emb.w"], trans_b=True, out_f32=d_semb, split_k=0 if det else -1)
        d_emb = ops.act_bwd(st.emb, d_semb, ops.ACT_SILU, new(S, E, dtype=f32))
        d_emb_c = new(S, E)
        ops.transpose(d_emb, copy=d_emb_c, colsum=gv["time_embed.2.bias"])
        d_te_h = new(S, E)
        dgrad(d_emb_c, W["te2.w"], out=d_te_h)
        wgrad(d_emb_c, st.te_h, gv["time_embed.2.weight"])
        d_h0 = ops.act_bwd(st.h0, d_te_h, ops.ACT_SILU, new(S, E))
        ops.colsum(d_h0, gv["time_embed.0.bias"])
        wgrad(d_h0, st.temb, gv["time_embed.0.weight"])
        self.results = (d_emb, d_xf.view(SU, N, -1))
        yield seg

    # ------------------------------------------------------------------------------------------ execution
    def _pool(self):
        return self.te.pool

    def forward(self):
        from . import _lib
        if not self.use_graph:
            self.saved = self._forward()
            return self.saved.eps
        if self.fwd_graph is None:
            c0 = _lib.launch_count()
            self.saved = self._forward()          # eager once: lazy initialisation (kernel attributes, tensor maps) + warm-up
            self.launches_fwd = _lib.launch_count() - c0
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._pool()):
                self.saved = self._forward()
            self.fwd_graph = g
        self.fwd_graph.replay()
        return self.saved.eps

    def backward(self, hook):
        """Runs (replays) the backward; hook(segment index) is called as soon as a gradient segment is final."""
        from . import _lib
        if not self.use_graph:
            for seg in self._backward():
                hook(seg)
            return self.results
        if self.bwd_graphs is None:
            c0 = _lib.launch_count()
            for _ in self._backward():             # eager once (warm-up); its gradients are discarded by the replay below
                pass
            self.launches_bwd = _lib.launch_count() - c0
            torch.cuda.synchronize()
            graphs, gen = [], self._backward()
            n_seg = len(self.te.fp.seg_bounds)
            for _ in range(n_seg):                 # one graph per gradient segment (the generator yields after each)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self._pool()):
                    seg = next(gen)
                graphs.append((g, seg))
            gen.close()
            self.bwd_graphs = graphs
        for g, seg in self.bwd_graphs:
            g.replay()
            hook(seg)
        return self.results


# ====================================================================================================== engine
class TrainEngine:
    def __init__(self, module):
        self.module = module
        self.eng = module.engine("bf16")
        self.fp = flat_params(module)
        self.W = build_operands(module, self.fp, self.eng)
        self.plans = {}
        self.pool = torch.cuda.graph_pool_handle()
        self.last = None

    def plan(self, S, T, N, U=None):
        # the kernel-selection knobs are baked into the captured graphs
        key = (S, T, N, U) + tuple(os.environ.get(k, "") for k in ("HIG_DETERMINISTIC", "HIG_TRAIN_FUSED_ATTN", "HIG_TRAIN_FUSED_FFN", "HIG_TRAIN_GRAPH"))
        p = self.plans.get(key)
        if p is None:
            if len(self.plans) >= 3:      # each plan pins its saved activations: keep a few shapes only
                self.plans.pop(next(iter(self.plans)))
            p = self.plans[key] = _Plan(self, S, T, N, U)
        return p

    def refresh(self):
        """Operand mirror up to date with the fp32 parameters (no-op after hig_adam_flat, one cast kernel after a third-party
        optimizer step) + the two derived operands of the motion embedding."""
        self.fp.refresh_mirror(self.module)
        refresh_special(self.module, self.W, self.eng.C)


def train_engine(module):
    fp = flat_params(module)
    te = getattr(module, "_hig_train_engine", None)
    if te is None or te.fp is not fp:
        te = module._hig_train_engine = TrainEngine(module)
    return te


class DenoiserGraphFn(torch.autograd.Function):
    """eps = denoiser(x, t, length, xf_proj, xf_out; params) on the captured graphs.  Gradients: xf_proj, xf_out and every
    denoiser parameter (views of the flat gradient buffer, or — FlatParams.direct — written in place and not returned)."""

    @staticmethod
    def forward(ctx, module, x, timesteps, length, xf_proj, xf_out, text_index, *params):
        te = train_engine(module)
        S, T, C = x.shape
        if S % 2:
            raise ValueError("the batch stacks person 1 and person 2 on dim 0: S must be even")
        if T > module.num_frames:
            raise ValueError(f"T={T} exceeds num_frames={module.num_frames}")
        te.refresh()
        U = None
        if text_index is not None:           # xf_out holds distinct captions; pad their count to a bucket of 16 (graph shapes)
            U = -(-xf_out.shape[0] // 16) * 16
        elif xf_out.shape[0] != S:
            raise ValueError("xf_out must have one row per sequence unless text_index maps the sequences to its rows")
        plan = te.plan(S, T, xf_out.shape[1], U)
        plan.x.copy_(x.detach())
        plan.t.copy_(timesteps.detach().to(torch.int64))
        plan.len.copy_(length.to(device=x.device, dtype=torch.int32).clamp(min=0, max=T))
        plan.xf_proj.copy_(xf_proj.detach())
        if U:
            plan.xf_out[:xf_out.shape[0]].copy_(xf_out.detach())
            plan.text_idx.copy_(text_index)
            ctx.n_text = xf_out.shape[0]
        else:
            plan.xf_out.copy_(xf_out.detach())
            ctx.n_text = S
        eps = plan.forward()
        ctx.plan, ctx.te = plan, te
        te.last = plan
        return eps.view(S, T, te.eng.LD_EPS)[:, :, :C].contiguous()

    @staticmethod
    def backward(ctx, d_eps):
        plan, te = ctx.plan, ctx.te
        fp, module = te.fp, te.module
        first = fp.params[0]
        # a p.grad that aliases the flat gradient buffer (left by a previous backward) would be doubled by autograd's
        # in-place accumulation: preserve accumulate semantics explicitly
        carry = None
        if not fp.direct and first.grad is not None and first.grad.data_ptr() == fp.gviews[fp.names[0]].data_ptr():
            carry = fp.grad[:fp.n_den].clone()
            for p in fp.params[:fp.n_den_params]:
                p.grad = None
        plan.d_eps.copy_(d_eps.detach())
        hook = getattr(module, "_grad_segment_hook", None)
        hook_many = getattr(module, "_grad_segments_hook", None) if hook is not None else None

        def seg_done(k):
            if hook_many is not None:
                hook_many(k, [fp.grad[lo:hi] for lo, hi in fp.seg_ranges[k]], fp)
            elif hook is not None:
                for lo, hi in fp.seg_ranges[k]:
                    hook(k, fp.grad[lo:hi])

        d_xf_proj, d_xf_out = plan.backward(seg_done)
        fin = getattr(module, "_grad_finish_hook", None)
        if fin is not None:
            fin()
        if carry is not None:
            fp.grad[:fp.n_den].add_(carry)
        ctx.plan = None
        n_den = fp.n_den_params
        if fp.direct:
            grads = [None] * n_den
        else:
            grads = [fp._view(fp.grad, n) for n in fp.names[:n_den]]
        return (None, None, None, None, d_xf_proj.clone(), d_xf_out[:ctx.n_text].clone(), None, *grads)


def denoiser_forward_graph(module, x, timesteps, length, xf_proj, xf_out, text_index=None):
    """text_index (LongTensor [S], optional): xf_out holds the batch's DISTINCT captions, sequence s uses row text_index[s]."""
    if not x.is_cuda:
        raise RuntimeError("hig_b200: the denoiser runs on CUDA only (no CPU fallback)")
    fp = flat_params(module)
    params = fp.params[:fp.n_den_params]
    ln = torch.as_tensor(length).reshape(-1)
    if text_index is not None:
        text_index = torch.as_tensor(text_index, device=x.device, dtype=torch.int64).reshape(-1)
        deterministic = os.environ.get("HIG_DETERMINISTIC", "0") not in ("", "0")
        if text_index.numel() != x.shape[0]:
            raise ValueError("text_index needs one entry per sequence")
        if deterministic or os.environ.get("HIG_TRAIN_TEXT_DEDUP", "1") == "0" or -(-xf_out.shape[0] // 16) * 16 >= x.shape[0]:
            # nothing to share (or bit-reproducible mode: the shared dA would be summed with atomics): one row per sequence
            xf_out, text_index = xf_out.index_select(0, text_index), None
    return DenoiserGraphFn.apply(module, x.float(), timesteps, ln, xf_proj.float(), xf_out.float(), text_index,
                                 *params).to(x.dtype)
