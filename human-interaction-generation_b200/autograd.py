"""Differentiable denoiser forward (training path): ONE torch.autograd.Function whose forward and backward are both
hand-scheduled sequences of C-ABI kernels (include/hig_b200.h) — no PyTorch math on the path.

What it replaces: torch.autograd's backward through MotionInteractionTransformer.forward
(codes/models/interaction_transformer.py:577-616), reached from `loss.backward()` in DDPMMulTrainer.update
(codes/trainers/mul_ddpm_trainer.py:249-256) via GaussianDiffusion.training_losses (models/gaussian_diffusion.py:1017).

Forward = the sampling schedule of denoiser_engine.py, except that every intermediate a derivative needs is kept
(per sub-block: residual input fp32, LayerNorm output, Q|K|V, attention output, SiLU(FiLM(LN)) output; FFN
pre-activation and GELU output; the embedding MLP's pre-activations) — ~25 [tok,512] tensors per layer, nothing is
recomputed except the attention softmaxes (recomputed in shared memory by eff_attn_bwd).
Backward, per linear y = x W^T + b:   dx = dy W  (same tcgen05 GEMM, B operand = W^T kept as a packed copy),
dW = dy^T x  (same GEMM on transposed operands, K = tokens split over CTAs with fp32 atomic accumulation),
db = column sums of dy (fused into the transposition pass).  The residual-stream gradient stays fp32.
Gradients land in ONE flat fp32 buffer laid out in backward-completion order (heads | layer L-1 | ... | layer 0 |
embeddings), so the data-parallel reducer (ddp.py) can all-reduce each segment as soon as it is final, overlapping
NCCL with the remaining backward kernels.
"""
import torch

from . import ops

HEAD_DIM = 64


def _rup(n, m):
    return (n + m - 1) // m * m


def denoiser_param_names(module):
    """Parameters the denoiser kernels consume, in flat-gradient-buffer order (segments complete in this order)."""
    segs = [["out.weight", "out.bias", "out2.weight", "out2.bias"]]
    subs = ["sa_block", "ca_block"] + ([] if module.no_cross_attn else ["int_ca_block"])
    emb_names = []
    for li in reversed(range(module.num_layers)):
        p = f"temporal_decoder_blocks.{li}."
        names = []
        # reverse execution order inside a layer: ffn, ic, ca, sa
        names += [p + "ffn.proj_out.out_layers.2.weight", p + "ffn.proj_out.out_layers.2.bias",
                  p + "ffn.proj_out.norm.weight", p + "ffn.proj_out.norm.bias",
                  p + "ffn.linear2.weight", p + "ffn.linear2.bias", p + "ffn.linear1.weight", p + "ffn.linear1.bias"]
        for sub in reversed(subs):
            q = p + sub + "."
            names += [q + "proj_out.out_layers.2.weight", q + "proj_out.out_layers.2.bias",
                      q + "proj_out.norm.weight", q + "proj_out.norm.bias"]
            if sub == "ca_block":
                names += [q + "query.weight", q + "query.bias", q + "key.weight", q + "value.weight",
                          q + "key.bias", q + "value.bias", q + "text_norm.weight", q + "text_norm.bias"]
            else:
                names += [q + "query.weight", q + "key.weight", q + "value.weight",
                          q + "query.bias", q + "key.bias", q + "value.bias"]
            names += [q + "norm.weight", q + "norm.bias"]
        segs.append(names)
    # stylization emb-linears in the engine's slab order (layer ascending; sa, ca, ic, ffn): ONE contiguous region
    for li in range(module.num_layers):
        p = f"temporal_decoder_blocks.{li}."
        for sub in subs + ["ffn"]:
            emb_names.append(p + sub + ".proj_out.emb_layers.1.weight")
    for li in range(module.num_layers):
        p = f"temporal_decoder_blocks.{li}."
        for sub in subs + ["ffn"]:
            emb_names.append(p + sub + ".proj_out.emb_layers.1.bias")
    segs.append(emb_names + ["time_embed.2.weight", "time_embed.2.bias", "time_embed.0.weight", "time_embed.0.bias",
                             "joint_embed.weight", "joint_embed.bias", "joint_embed2.weight", "joint_embed2.bias",
                             "sequence_embedding"])
    return segs


class _Grads:
    """Flat fp32 gradient buffer + named views; segment k is final once backward has passed it."""

    def __init__(self, module, device):
        named = dict(module.named_parameters())
        self.segments = denoiser_param_names(module)
        self.names = [n for seg in self.segments for n in seg]
        sizes = [named[n].numel() for n in self.names]
        self.flat = torch.zeros(sum(sizes), device=device, dtype=torch.float32)
        self.views, self.seg_bounds = {}, []
        off = 0
        for seg in self.segments:
            lo = off
            for n in seg:
                k = named[n].numel()
                self.views[n] = self.flat[off:off + k].view(named[n].shape)
                off += k
            self.seg_bounds.append((lo, off))

    def region(self, first, count_elems, shape):
        """Contiguous region starting at parameter `first` spanning several adjacent parameters."""
        v = self.views[first]
        off = v.storage_offset()
        return self.flat[off:off + count_elems].view(shape)


class _Ctx:
    pass


class DenoiserFn(torch.autograd.Function):
    """eps = denoiser(x, t, length, xf_proj, xf_out; params).  Gradients: xf_proj, xf_out and every parameter."""

    @staticmethod
    def forward(ctx, module, x, timesteps, length, xf_proj, xf_out, *params):
        eng = module.engine()
        S, T, C = x.shape
        if S % 2:
            raise ValueError("the batch stacks person 1 and person 2 on dim 0: S must be even")
        if T > module.num_frames:
            raise ValueError(f"T={T} exceeds num_frames={module.num_frames}")
        st = _Ctx()
        st.module, st.eng, st.S, st.T, st.C = module, eng, S, T, C
        _forward(st, x, timesteps, length, xf_proj, xf_out)
        ctx.st = st
        ctx.n_params = len(params)
        ctx.param_names = module._denoiser_param_order
        return st.eps.view(S, T, eng.LD_EPS)[:, :, :C].contiguous()

    @staticmethod
    def backward(ctx, d_eps):
        st = ctx.st
        grads, d_xf_proj, d_xf_out = _backward(st, d_eps)
        out = [grads.views[n] for n in ctx.param_names]
        ctx.st = None
        return (None, None, None, None, d_xf_proj, d_xf_out, *out)


def denoiser_forward_with_grad(module, x, timesteps, length, xf_proj, xf_out):
    if not x.is_cuda:
        raise RuntimeError("hig_b200: the denoiser runs on CUDA only (no CPU fallback)")
    segs = denoiser_param_names(module)
    order = [n for seg in segs for n in seg]
    module._denoiser_param_order = order
    named = dict(module.named_parameters())
    params = [named[n] for n in order]
    ln = torch.as_tensor(length).reshape(-1)
    return DenoiserFn.apply(module, x, timesteps, ln, xf_proj, xf_out, *params).to(x.dtype)


# ====================================================================================================== forward
def _forward(st, x, timesteps, length, xf_proj, xf_out):
    eng, S, T = st.eng, st.S, st.T
    W = eng.packed()
    D, F_, E, H, L = eng.D, eng.F, eng.E, eng.H, eng.L
    dt, dev, tok = eng.act_dtype, x.device, S * T
    f32 = torch.float32
    new = lambda *shape, dtype=dt: torch.empty(*shape, device=dev, dtype=dtype)
    G = eng._gemm
    st.len = torch.empty(S, device=dev, dtype=torch.int32)
    st.len.copy_(length.to(device=dev, dtype=torch.int32).clamp(min=0, max=T))
    t_dev = timesteps.to(device=dev, dtype=torch.int64).contiguous()

    # ---- embedding MLP (:591) and every StylizationBlock's (scale | shift) (:88-90)
    st.temb = ops.timestep_embed(t_dev, W["freqs"], new(S, D))
    st.h0 = new(S, E)
    G(st.temb, W["te0.w"], W["te0.b"], out=st.h0)
    st.te_h = ops.act_fwd(st.h0, ops.ACT_SILU, new(S, E))
    st.emb = new(S, E, dtype=f32)
    G(st.te_h, W["te2.w"], W["te2.b"], out_f32=st.emb, residual=xf_proj.detach().to(f32).contiguous())
    st.semb = ops.act_fwd(st.emb, ops.ACT_SILU, new(S, E))
    st.ss = new(S, W["n_styl"] * 2 * D, dtype=f32)
    G(st.semb, W["emb.w"], W["emb.b"], out_f32=st.ss)

    # ---- text K/V side of every layer's cross attention (:155-161)
    Sx, N, Dt = xf_out.shape
    st.N = N
    st.xf = xf_out.detach().to(f32).contiguous().view(S * N, Dt)
    st.tn, st.kv, st.a_text = [], [], []
    for i in range(L):
        p = f"l{i}.ca."
        tn = ops.ln_film_silu(st.xf, W[p + "tln.w"], W[p + "tln.b"], new(S * N, Dt))
        kv = new(S * N, 2 * D)
        G(tn, W[p + "kv.w"], W[p + "kv.b"], out=kv)
        a = new(S, H, HEAD_DIM, HEAD_DIM)
        ops.eff_attn(ops.ATTN_KV_ONLY, S, N, H, k=kv[:, :D], v=kv[:, D:], a_out=a)
        st.tn.append(tn); st.kv.append(kv); st.a_text.append(a)

    # ---- motion embedding (:593-602)
    st.xa = torch.zeros(tok, eng.CP, device=dev, dtype=dt)
    ops.pack_motion(x.detach().to(f32).contiguous(), st.xa)
    xres = new(tok, D, dtype=f32)
    G(st.xa, W["in.w"], None, residual=W["in.pos"], res_row_mod=T, out_f32=xres)

    fp32_mode = eng.precision == "fp32"
    st.blocks = []          # per sub-block saved tensors, execution order

    def stylize_project(blk, y, xres_in, p, want_xb):
        i = W[p + ".ss"]
        ss = st.ss[:, i * 2 * D:(i + 1) * 2 * D]
        sact = ops.ln_film_silu(y, W[p + ".po.ln.w"], W[p + ".po.ln.b"], new(tok, D), rows_per_seq=T, scale_shift=ss,
                                silu=True)
        xres_out = new(tok, D, dtype=f32)
        xb = None
        if fp32_mode:
            G(sact, W[p + ".po.w"], W[p + ".po.b"], residual=xres_in, out_f32=xres_out)
            xb = xres_out
        else:
            xb = new(tok, D) if want_xb else None
            G(sact, W[p + ".po.w"], W[p + ".po.b"], residual=xres_in, out_f32=xres_out, out=xb)
        blk.update(y=y, sact=sact, ss_index=i, p=p)
        return xres_out, xb

    xb = None
    for li in range(L):
        p = f"l{li}."
        kinds = ["sa", "ca"] + (["ic"] if eng.has_ic else [])
        for kind in kinds:
            blk = {"kind": kind, "xres_in": xres, "li": li}
            n = ops.ln_film_silu(xres, W[p + kind + ".ln.w"], W[p + kind + ".ln.b"], new(tok, D))
            y = new(tok, D)
            if kind == "ca":
                qc = new(tok, D)
                G(n, W[p + "ca.q.w"], W[p + "ca.q.b"], out=qc)
                ops.eff_attn(ops.ATTN_Q_ONLY, S, T, H, q=qc, a_in=st.a_text[li], y=y)
                blk.update(n=n, q=qc)
            else:
                qkv = new(tok, 3 * D)
                G(n, W[p + kind + ".qkv.w"], W[p + kind + ".qkv.b"], out=qkv)
                q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
                if kind == "sa":
                    ops.eff_attn(ops.ATTN_SELF, S, T, H, q=q, k=k, v=v, y=y, length=st.len, mask_v=True)
                else:
                    ops.eff_attn(ops.ATTN_INTER, S, T, H, q=q, k=k, v=v, y=y, length=st.len, pair_shift=S // 2,
                                 mask_v=False)
                blk.update(n=n, qkv=qkv)
            last_attn = kind == kinds[-1]
            xres, xb = stylize_project(blk, y, xres, p + kind, last_attn)
            st.blocks.append(blk)
        # FFN (:261-264)
        blk = {"kind": "ffn", "xres_in": xres, "li": li, "xb_in": xb}
        h1 = new(tok, F_)
        G(xb, W[p + "ffn.w1"], W[p + "ffn.b1"], out=h1)
        g = ops.act_fwd(h1, ops.ACT_GELU, new(tok, F_))
        y = new(tok, D)
        G(g, W[p + "ffn.w2"], W[p + "ffn.b2"], out=y)
        blk.update(h1=h1, g=g)
        xres, xb = stylize_project(blk, y, xres, p + "ffn", True)
        st.blocks.append(blk)

    # ---- output heads (:613-616)
    st.xb_final = xb
    st.eps = new(tok, eng.LD_EPS, dtype=f32)
    G(xb, W["out.w"], W["out.b"], out_f32=st.eps[:, :eng.C])
    a0 = xb.view(S, T * D)[:, :D]
    o0 = st.eps.view(S, T * eng.LD_EPS)[:, :eng.C]
    G(a0, W["out2.w"], W["out2.b"], out_f32=o0)


# ====================================================================================================== backward
class _Bwd:
    """Shared helpers of the backward schedule: operand transposition, dgrad / wgrad GEMMs in either precision."""

    def __init__(self, st):
        self.st, self.eng = st, st.eng
        self.dt = st.eng.act_dtype
        self.dev = st.eps.device
        self.bf16 = st.eng.precision == "bf16"
        self._zero_bias = st.eng.__dict__.setdefault("_zero_bias", {})

    def new(self, *shape, dtype=None):
        return torch.empty(*shape, device=self.dev, dtype=dtype or self.dt)

    def tr(self, x, colsum=None, want_copy=False, rows_zero_mod=0, want_t=True):
        """x [M,N] -> (x^T [N, M] as a view of an [N, rup(M,8)] buffer, optional [M,N] copy in the operand dtype)."""
        M, N = x.shape
        xt = self.new(N, _rup(M, 8)) if want_t else None
        cp = self.new(M, N) if want_copy else None
        ops.transpose(x, out_t=xt, copy=cp, colsum=colsum, rows_zero_mod=rows_zero_mod)
        return (xt[:, :M] if want_t else None), cp

    def dgrad(self, dy, wT, out=None, out_f32=None, accumulate=False):
        """dx[M,K] (+)= dy[M,N] . W[N,K]   (wT = W^T [K,N] packed in the operand dtype).  A zero bias vector keeps the
        GEMM on its specialised coalesced epilogues."""
        n = wT.shape[0]
        zb = self._zero_bias.get(n)
        if zb is None:
            zb = self._zero_bias[n] = torch.zeros(n, device=self.dev, dtype=torch.float32)
        self.eng._gemm(dy, wT, zb, out=out, out_f32=out_f32, residual=out_f32 if accumulate else None)

    def wgrad(self, dyT, xT, w_grad):
        """w_grad[N,K] += dy^T[N,M] . x[M,K]  with dyT [N,M], xT [K,M] (views with 16-byte-aligned leading dims)."""
        if self.bf16:
            ops.gemm_splitk(dyT, xT, w_grad)
        else:
            ops.gemm(dyT, xT, residual=w_grad, out_f32=w_grad)


def _backward(st, d_eps):
    eng, S, T, C = st.eng, st.S, st.T, st.C
    W, WT = eng.packed(), eng.packed_T()
    D, F_, E, H, L = eng.D, eng.F, eng.E, eng.H, eng.L
    tok, N = S * T, st.N
    f32 = torch.float32
    B = _Bwd(st)
    dt, dev = B.dt, B.dev
    gr = _Grads(st.module, dev)
    gv = gr.views
    hook = getattr(st.module, "_grad_segment_hook", None)
    seg_id = [0]

    def segment_done():
        if hook is not None:
            lo, hi = gr.seg_bounds[seg_id[0]]
            hook(seg_id[0], gr.flat[lo:hi])
        seg_id[0] += 1

    d_ss = torch.zeros_like(st.ss)
    d_xf = torch.zeros_like(st.xf)

    # ---------------- output heads: eps = out(h[:,1:]) / out2(h[:,0])  (:613-616) ----------------
    d_eps = d_eps.detach().to(f32).contiguous().view(tok, C)
    LDE = _rup(C, 8)
    de_a = torch.zeros(tok, LDE, device=dev, dtype=dt)               # frames >= 1 (frame-0 rows zeroed)
    de_aT = torch.zeros(LDE, _rup(tok, 8), device=dev, dtype=dt)
    ops.transpose(d_eps, out_t=de_aT, copy=de_a, colsum=gv["out.bias"], rows_zero_mod=T)
    de_0 = torch.zeros(S, LDE, device=dev, dtype=dt)                  # frame 0 of every sequence
    de_0T = torch.zeros(LDE, _rup(S, 8), device=dev, dtype=dt)
    ops.transpose(d_eps.view(S, T * C)[:, :C], out_t=de_0T, copy=de_0, colsum=gv["out2.bias"])
    dres = B.new(tok, D, dtype=f32)
    B.dgrad(de_a, WT["out.w"], out_f32=dres)
    B.dgrad(de_0, WT["out2.w"], out_f32=dres.view(S, T * D)[:, :D])
    xbT, _ = B.tr(st.xb_final)
    wtmp = torch.zeros(LDE, D, device=dev, dtype=f32)
    B.wgrad(de_aT[:, :tok], xbT, wtmp)
    gv["out.weight"].copy_(wtmp[:C])
    xb0T, _ = B.tr(st.xb_final.view(S, T * D)[:, :D])
    wtmp.zero_()
    B.wgrad(de_0T[:, :S], xb0T, wtmp)
    gv["out2.weight"].copy_(wtmp[:C])
    segment_done()

    # ---------------- layers, reverse order ----------------
    def stylize_project_bwd(blk, pfx, mod_prefix):
        """dres is d(xres_out).  Returns d_y; accumulates po weight/bias, po-norm and d_ss slab gradients."""
        dresT, dres_c = B.tr(dres, colsum=gv[mod_prefix + "proj_out.out_layers.2.bias"], want_copy=True)
        d_sact = B.new(tok, D)
        B.dgrad(dres_c, WT[pfx + ".po.w"], out=d_sact)
        sactT, _ = B.tr(blk["sact"])
        B.wgrad(dresT, sactT, gv[mod_prefix + "proj_out.out_layers.2.weight"])
        i = blk["ss_index"]
        ss = st.ss[:, i * 2 * D:(i + 1) * 2 * D]
        d_y = B.new(tok, D)
        d_gb = torch.zeros(S, 2 * D, device=dev, dtype=f32)
        ops.ln_film_silu_bwd(blk["y"], W[pfx + ".po.ln.w"], W[pfx + ".po.ln.b"], d_sact, d_y, T, scale_shift=ss,
                             silu=True, d_ss=d_ss[:, i * 2 * D:(i + 1) * 2 * D], d_gb=d_gb)
        ops.colsum(d_gb, gr.region(mod_prefix + "proj_out.norm.weight", 2 * D, (2 * D,)))
        return d_y

    def linear_bwd(dy, x_saved, wT_key, w_first, w_elems, w_shape, b_first, b_elems, dx_into=None):
        """Backward of y = x W^T + b: returns dx (operand dtype), or accumulates it into the fp32 `dx_into`."""
        dyT, _ = B.tr(dy, colsum=gr.region(b_first, b_elems, (b_elems,)))
        dx = None
        if dx_into is not None:
            B.dgrad(dy, WT[wT_key], out_f32=dx_into, accumulate=True)
        else:
            dx = B.new(dy.shape[0], w_shape[1])
            B.dgrad(dy, WT[wT_key], out=dx)
        xT, _ = B.tr(x_saved)
        B.wgrad(dyT, xT, gr.region(w_first, w_elems, w_shape))
        return dx

    def pre_ln_bwd(blk, d_n, pfx, mod_prefix):
        d_gb = torch.zeros(S, 2 * D, device=dev, dtype=f32)
        ops.ln_film_silu_bwd(blk["xres_in"], W[pfx + ".ln.w"], W[pfx + ".ln.b"], d_n, dres, T, dx_accumulate=True,
                             d_gb=d_gb)
        ops.colsum(d_gb, gr.region(mod_prefix + "norm.weight", 2 * D, (2 * D,)))

    modname = {"sa": "sa_block.", "ca": "ca_block.", "ic": "int_ca_block.", "ffn": "ffn."}
    for blk in reversed(st.blocks):
        li, kind = blk["li"], blk["kind"]
        pfx = f"l{li}.{kind}"
        mp = f"temporal_decoder_blocks.{li}.{modname[kind]}"
        d_y = stylize_project_bwd(blk, pfx, mp)
        if kind == "ffn":
            d_g = linear_bwd(d_y, blk["g"], f"l{li}.ffn.w2", mp + "linear2.weight", D * F_, (D, F_),
                             mp + "linear2.bias", D)
            d_h1 = ops.act_bwd(blk["h1"], d_g, ops.ACT_GELU, B.new(tok, F_))
            # no pre-norm in the FFN: d(xres_in) = dres (skip path) + d_h1 . W1, accumulated by the GEMM epilogue
            linear_bwd(d_h1, blk["xb_in"], f"l{li}.ffn.w1", mp + "linear1.weight", F_ * D, (F_, D),
                       mp + "linear1.bias", F_, dx_into=dres)
        elif kind == "ca":
            d_q = B.new(tok, D)
            dA = B.new(S, H, HEAD_DIM, HEAD_DIM, dtype=f32)
            ops.eff_attn_bwd(ops.ATTN_Q_ONLY, S, T, H, q=blk["q"], a_in=st.a_text[li], dy=d_y, dq=d_q, dA=dA)
            d_n = linear_bwd(d_q, blk["n"], f"l{li}.ca.q.w", mp + "query.weight", D * D, (D, D), mp + "query.bias", D)
            pre_ln_bwd(blk, d_n, pfx, mp)
            # text K/V side
            kv = st.kv[li]
            d_kv = B.new(S * N, 2 * D)
            ops.eff_attn_bwd(ops.ATTN_KV_ONLY, S, N, H, k=kv[:, :D], v=kv[:, D:], dk=d_kv[:, :D], dv=d_kv[:, D:], dA=dA)
            Dt = st.xf.shape[1]
            d_tn = linear_bwd(d_kv, st.tn[li], f"l{li}.ca.kv.w", mp + "key.weight", 2 * D * Dt, (2 * D, Dt),
                              mp + "key.bias", 2 * D)
            d_gb = torch.zeros(S, 2 * Dt, device=dev, dtype=f32)
            ops.ln_film_silu_bwd(st.xf, W[pfx + ".tln.w"], W[pfx + ".tln.b"], d_tn, d_xf, N, dx_accumulate=True,
                                 d_gb=d_gb)
            ops.colsum(d_gb, gr.region(mp + "text_norm.weight", 2 * Dt, (2 * Dt,)))
        else:
            qkv = blk["qkv"]
            d_qkv = B.new(tok, 3 * D)
            mode = ops.ATTN_SELF if kind == "sa" else ops.ATTN_INTER
            ops.eff_attn_bwd(mode, S, T, H, q=qkv[:, :D], k=qkv[:, D:2 * D], v=qkv[:, 2 * D:], dy=d_y,
                             dq=d_qkv[:, :D], dk=d_qkv[:, D:2 * D], dv=d_qkv[:, 2 * D:], length=st.len,
                             pair_shift=S // 2 if kind == "ic" else 0)
            d_n = linear_bwd(d_qkv, blk["n"], f"l{li}.{kind}.qkv.w", mp + "query.weight", 3 * D * D, (3 * D, D),
                             mp + "query.bias", 3 * D)
            pre_ln_bwd(blk, d_n, pfx, mp)
        if kind == "sa":
            segment_done()   # every parameter of layer li (except its emb-linears) is final

    # ---------------- motion embedding (:593-602) ----------------
    dresT, _ = B.tr(dres)
    xaT, _ = B.tr(st.xa)
    w_in = torch.zeros(D, eng.CP, device=dev, dtype=f32)
    B.wgrad(dresT, xaT, w_in)
    gv["joint_embed.weight"].copy_(w_in[:, :C])
    gv["joint_embed2.weight"].copy_(w_in[:, C:C + 4])
    dpos = torch.zeros(T * D, device=dev, dtype=f32)
    ops.colsum(dres.view(S, T * D), dpos)
    dpos = dpos.view(T, D)
    gv["joint_embed2.bias"].copy_(dpos[0])
    if T > 1:
        gv["sequence_embedding"][:T - 1].copy_(dpos[1:])
        ops.colsum(dpos[1:], gv["joint_embed.bias"])

    # ---------------- stylization emb-linears + time-embedding MLP (:88-90, :474-478, :591) ----------------
    n_styl = W["n_styl"]
    first_w = gr.segments[-1][0]
    first_b = gr.segments[-1][n_styl]
    d_ssT, d_ss_c = B.tr(d_ss, colsum=gr.region(first_b, n_styl * 2 * D, (n_styl * 2 * D,)), want_copy=True)
    d_semb = B.new(S, E, dtype=f32)
    B.dgrad(d_ss_c, WT["emb.w"], out_f32=d_semb)
    sembT, _ = B.tr(st.semb)
    B.wgrad(d_ssT, sembT, gr.region(first_w, n_styl * 2 * D * E, (n_styl * 2 * D, E)))
    d_emb = ops.act_bwd(st.emb, d_semb, ops.ACT_SILU, B.new(S, E, dtype=f32))
    d_embT, d_emb_c = B.tr(d_emb, colsum=gv["time_embed.2.bias"], want_copy=True)
    d_te_h = B.new(S, E)
    B.dgrad(d_emb_c, WT["te2.w"], out=d_te_h)
    te_hT, _ = B.tr(st.te_h)
    B.wgrad(d_embT, te_hT, gv["time_embed.2.weight"])
    d_h0 = ops.act_bwd(st.h0, d_te_h, ops.ACT_SILU, B.new(S, E))
    d_h0T, _ = B.tr(d_h0, colsum=gv["time_embed.0.bias"])
    tembT, _ = B.tr(st.temb)
    B.wgrad(d_h0T, tembT, gv["time_embed.0.weight"])
    segment_done()
    fin = getattr(st.module, "_grad_finish_hook", None)
    if fin is not None:
        fin()
    return gr, d_emb, d_xf.view(S, N, -1)
