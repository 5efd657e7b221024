"""DenoiserEngine — the kernel schedule of one role-aware denoiser forward on a B200.

Owns (a) the packed weights (fp32 master parameters -> bf16 GEMM operands, Q/K/V and all 4L stylization
emb-linears concatenated, joint_embed/joint_embed2 fused into one K-padded operand) and (b) persistent
activation workspaces in HBM, and issues the C-ABI kernels in order on the current stream.  No torch math on
the path: torch only allocates the buffers.  The whole schedule is capturable in a CUDA graph (no allocation,
no host sync, tensor maps are kernel parameters) — gaussian_diffusion.p_sample_loop does exactly that.

HBM layout for S sequences x T frames (tok = S*T rows, D = 512):
  xres  fp16 [tok, 512]   residual stream (fp16: half the out-proj epilogue traffic of fp32; bf16 would cost 8e-3
                          relative error per step over its 32 rounded adds, fp16 costs 1e-3; fp32 in fp32 mode)
  xb    act  [tok, 512]   copy of the residual stream in the GEMM operand type (FFN / output heads read it)
  n     act  [tok, 512]   LayerNorm output feeding the Q/K/V projections
  qkv   act  [tok, 1536]  projections (Q | K | V), also reused as [tok, 512] for the text-CA query
  y     act  [tok, 512]   FFN branch output (fp32 mode: also the attention output)
  a_blk act  [S, 8, 64, 64] softmax_time(K)^T V of the current attention block (8 MB: L2-resident hand-off)
  sact  act  [tok, 512]   SiLU(FiLM(LN(y))) feeding the block's output projection
  g     act  [tok, 1024]  GELU(linear1)
  ss    fp32 [S, 4L*1024] (scale | shift) of every StylizationBlock, one GEMM per step
  eps   fp32 [tok, 264]   network output (leading dim padded 263 -> 264 for 16-byte rows)
"act" = bf16 in the product path, fp32 in fp32 mode.
Reference: MotionInteractionTransformer.forward, codes/models/interaction_transformer.py:577-616.
"""
import math
import os

import torch

from . import ops

HEAD_DIM = 64


class DenoiserEngine:
    def __init__(self, module, precision="bf16"):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.m = module
        self.precision = precision
        self.act_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        # residual stream storage: fp16 on the product path (saturating stores; 1e-3 relative per step (CPU experiment),
        # against 8e-3 for bf16 — the stream takes 32 rounded adds per step), fp32 in fp32 mode
        self.res_dtype = torch.float16 if precision == "bf16" else torch.float32
        self.D = module.latent_dim
        self.F = module.ff_size
        self.E = module.time_embed_dim
        self.C = module.input_feats
        self.H = module.num_heads
        self.L = module.num_layers
        self.has_ic = not module.no_cross_attn
        if self.D != 512 or self.D // self.H != HEAD_DIM:
            raise NotImplementedError("hig_b200 kernels are built for latent_dim=512, head_dim=64 (the repo default)")
        self.CP = (self.C + 4 + 7) // 8 * 8          # 263 + 4 -> 272: K padded for TMA's 16-byte rule
        self.LD_EPS = (self.C + 3) // 4 * 4          # 264
        self._packed = None
        self._packed_T = None
        self._packed_key = None
        self.packed_generation = 0   # bumped whenever the operand copies are rebuilt (captured graphs go stale)
        self._ws = {}
        self._text_cache = None
        self._time_table = None
        # product path: projections with the TMA-staged epilogue, pre-attention LayerNorms folded into them
        self.stream = precision == "bf16"
        # output heads through the resident-W kernel with fp16 eps (HIG_HEADS16=0: the general GEMM with fp32 eps)
        self.heads16 = self.stream and os.environ.get("HIG_HEADS16", "1") != "0"

    @property
    def qsm(self):
        """Feature softmax of the queries in the Q / Q|K|V projection's epilogue (resident-W kernel only); read at call
        time so that HIG_QSM / HIG_WRES can be A/B-toggled between captures."""
        return self.stream and os.environ.get("HIG_QSM", "1") != "0" and os.environ.get("HIG_WRES", "1") != "0"

    @property
    def apply_tc(self):
        """Query half of the attention on tcgen05 / TMEM (attn_apply_tc.cu): needs the queries softmaxed by the projection
        (qsm) and A^T from the K/V half.  HIG_APPLY_TC=0 keeps the mma.sync kernel (A/B runs)."""
        return self.qsm and os.environ.get("HIG_APPLY_TC", "1") != "0"

    # ------------------------------------------------------------------------------------------ weights
    def _param_key(self):
        # _hig_param_generation: bumped by the fused Adam kernel, which updates the parameters through raw pointers
        return (getattr(self.m, "_hig_param_generation", 0),) + tuple((p.data_ptr(), p._version) for p in self.m.parameters())

    def packed(self):
        key = self._param_key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        m, dt = self.m, self.act_dtype
        dev = m.joint_embed.weight.device
        if dev.type != "cuda":
            raise RuntimeError("hig_b200: the denoiser runs on CUDA only (no CPU fallback); move the module to a GPU")
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        op = lambda t: t.detach().to(dt).contiguous()
        W = {}
        with torch.no_grad():
            w_in = torch.zeros(self.D, self.CP, device=dev, dtype=torch.float32)
            w_in[:, :self.C] = m.joint_embed.weight
            w_in[:, self.C:self.C + 4] = m.joint_embed2.weight
            W["in.w"] = op(w_in)
            pos = torch.empty(m.num_frames, self.D, device=dev, dtype=torch.float32)
            pos[0] = m.joint_embed2.bias
            pos[1:] = m.joint_embed.bias[None] + m.sequence_embedding[:m.num_frames - 1]
            W["in.pos"] = pos.contiguous()
            W["zero.b"] = torch.zeros(self.D, device=dev, dtype=torch.float32)
            W["te0.w"], W["te0.b"] = op(m.time_embed[0].weight), f32(m.time_embed[0].bias)
            W["te2.w"], W["te2.b"] = op(m.time_embed[2].weight), f32(m.time_embed[2].bias)
            emb_w, emb_b = [], []
            for i, blk in enumerate(m.temporal_decoder_blocks):
                p = f"l{i}."
                subs = [("sa", blk.sa_block), ("ca", blk.ca_block)]
                if self.has_ic:
                    subs.append(("ic", blk.int_ca_block))
                for name, a in subs:
                    W[p + name + ".ln.w"], W[p + name + ".ln.b"] = f32(a.norm.weight), f32(a.norm.bias)
                    if name == "ca":
                        W[p + "ca.tln.w"], W[p + "ca.tln.b"] = f32(a.text_norm.weight), f32(a.text_norm.bias)
                        W[p + "ca.q.w"], W[p + "ca.q.b"] = op(a.query.weight), f32(a.query.bias)
                        W[p + "ca.kv.w"] = op(torch.cat([a.key.weight, a.value.weight], 0))
                        W[p + "ca.kv.b"] = f32(torch.cat([a.key.bias, a.value.bias], 0))
                        if self.stream:
                            self._fold_ln(W, p + "ca.q", a.norm, a.query.weight, a.query.bias)
                    else:
                        w_cat = torch.cat([a.query.weight, a.key.weight, a.value.weight], 0)
                        b_cat = torch.cat([a.query.bias, a.key.bias, a.value.bias], 0)
                        W[p + name + ".qkv.w"], W[p + name + ".qkv.b"] = op(w_cat), f32(b_cat)
                        if self.stream:
                            self._fold_ln(W, p + name + ".qkv", a.norm, w_cat, b_cat)
                W[p + "ffn.w1"], W[p + "ffn.b1"] = op(blk.ffn.linear1.weight), f32(blk.ffn.linear1.bias)
                W[p + "ffn.w2"], W[p + "ffn.b2"] = op(blk.ffn.linear2.weight), f32(blk.ffn.linear2.bias)
                if self.stream:   # linear1 reads the fp16 residual stream directly
                    W[p + "ffn.w1h"] = blk.ffn.linear1.weight.detach().to(torch.float16).contiguous()
                for name, a in subs + [("ffn", blk.ffn)]:
                    st = a.proj_out
                    W[p + name + ".po.ln.w"], W[p + name + ".po.ln.b"] = f32(st.norm.weight), f32(st.norm.bias)
                    W[p + name + ".po.w"], W[p + name + ".po.b"] = op(st.out_layers[2].weight), f32(st.out_layers[2].bias)
                    W[p + name + ".ss"] = len(emb_w)      # index of this block's (scale|shift) slab
                    emb_w.append(st.emb_layers[1].weight)
                    emb_b.append(st.emb_layers[1].bias)
            W["emb.w"] = op(torch.cat(emb_w, 0))          # [n_styl * 2D, E]
            W["emb.b"] = f32(torch.cat(emb_b, 0))
            W["n_styl"] = len(emb_w)
            W["out.w"], W["out.b"] = op(m.out.weight), f32(m.out.bias)
            W["out2.w"], W["out2.b"] = op(m.out2.weight), f32(m.out2.bias)
            if self.stream:
                # output heads on the resident-W kernel: fp16 weights padded to 512 rows (N % 256 == 0), reading the fp16
                # stream directly (no bf16 copy of the stream), eps stored as fp16 [tok, 512]
                for nm, lin in (("out", m.out), ("out2", m.out2)):
                    wp = torch.zeros(self.D, self.D, device=dev, dtype=torch.float16)
                    wp[:self.C] = lin.weight.detach().to(torch.float16)
                    bp = torch.zeros(self.D, device=dev, dtype=torch.float32)
                    bp[:self.C] = lin.bias.detach().float()
                    W[nm + ".w16"], W[nm + ".b16"] = wp, bp
            half = self.D // 2
            W["freqs"] = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(dev)
        self._packed, self._packed_key = W, key
        self.packed_generation += 1
        self._packed_T = None
        self._text_cache = None
        return W

    @staticmethod
    def _fold_ln(W, key, norm, weight, bias):
        """nn.LayerNorm folded into the Linear that consumes it (hig_gemm_stream, HIG_GS_LN_BF16):
        LN(x) W^T + b = rstd (x (gamma o W)^T - mu rowsum(gamma o W)) + (b + W beta).  The operand is stored in fp16 (the
        residual stream's type); rowsum is taken over the ROUNDED operand so the mean term cancels exactly."""
        w32 = weight.detach().to(torch.float32)
        wg = (w32 * norm.weight.detach().to(torch.float32)[None, :]).to(torch.float16).contiguous()
        W[key + ".wg"] = wg
        W[key + ".wsum"] = wg.to(torch.float32).sum(1).contiguous()
        W[key + ".bg"] = (bias.detach().to(torch.float32) + w32 @ norm.bias.detach().to(torch.float32)).contiguous()

    def packed_T(self):
        """W^T copies in the operand dtype: the B operand of the data-gradient GEMMs dx = dy . W (training only).
        The 263-row output heads are zero-padded to 264 columns (TMA 16-byte rule)."""
        W = self.packed()
        if self._packed_T is not None:
            return self._packed_T
        WT = {}
        with torch.no_grad():
            for k, v in W.items():
                if not (torch.is_tensor(v) and v.dim() == 2 and v.dtype == self.act_dtype):
                    continue
                if not (k.endswith(".w") or k.endswith(".w1") or k.endswith(".w2")) or k in ("in.w", "te0.w"):
                    continue
                t = v.t().contiguous()
                if k in ("out.w", "out2.w"):
                    pad = torch.zeros(t.shape[0], (t.shape[1] + 7) // 8 * 8, device=t.device, dtype=t.dtype)
                    pad[:, :t.shape[1]] = t
                    t = pad
                WT[k] = t
        self._packed_T = WT
        return WT

    # ------------------------------------------------------------------------------------------ workspaces
    def workspace(self, S, T, slot=0):
        """Activation buffers for S sequences x T frames; `slot` distinguishes concurrent branches of the same shape."""
        key = (S, T, slot)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        dev = self.m.joint_embed.weight.device
        tok, D, dt = S * T, self.D, self.act_dtype
        n_styl = self.packed()["n_styl"]
        e = lambda *shape, dtype=dt: torch.empty(*shape, device=dev, dtype=dtype)
        ws = {
            "xa": torch.zeros(tok, self.CP, device=dev, dtype=dt),
            "xres": e(tok, D, dtype=self.res_dtype), "xb": e(tok, D), "n": e(tok, D), "qkv": e(tok, 3 * D),
            "y": e(tok, D), "sact": e(tok, D), "g": e(tok, self.F),
            "temb": e(S, D), "te_h": e(S, self.E), "semb": e(S, self.E),
            "ss": e(S, n_styl * 2 * D, dtype=torch.float32),
            "eps": e(tok, self.LD_EPS, dtype=torch.float32),
            "eps16": e(tok, D, dtype=torch.float16) if self.stream else None,   # product path: fp16 eps, 512-wide rows
            "len": torch.empty(S, device=dev, dtype=torch.int32),
            "a_blk": e(S, self.H, HEAD_DIM, HEAD_DIM),
            "stats": e(tok, 8, dtype=torch.float32),   # LayerNorm row statistics of the stream (4 partials per row)
        }
        if self.precision == "fp32":
            ws["xb"] = ws["xres"]            # fp32 mode: the operand copy IS the residual stream
        if len(self._ws) > 8:
            self._ws.clear()
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------ text (step-invariant)
    def text_state(self, xf_out):
        """A_text[l] = softmax_tokens(K_l)^T V_l for every layer: depends only on xf_out, so it is computed once
        per batch and reused by all 1000 steps (the reference recomputes it every step, :145-161)."""
        # keyed on the tensor OBJECT (kept alive by the cache, so its address cannot be recycled for another batch's text),
        # its version counter and the weights it was projected with
        W = self.packed()
        c = self._text_cache
        transposed = self.apply_tc      # the tcgen05 apply kernel takes A^T (its K-major B operand)
        if c is not None and c[0] is xf_out and c[1] == (xf_out._version, self.packed_generation, transposed):
            return c[2]
        S, N, Dt = xf_out.shape
        dt, dev = self.act_dtype, xf_out.device
        xf = xf_out.detach().to(torch.float32).contiguous().view(S * N, Dt)
        tn = torch.empty(S * N, Dt, device=dev, dtype=dt)
        kv = torch.empty(S * N, 2 * self.D, device=dev, dtype=dt)
        a_all = torch.empty(self.L, S, self.H, HEAD_DIM, HEAD_DIM, device=dev, dtype=dt)
        for i in range(self.L):
            p = f"l{i}.ca."
            ops.ln_film_silu(xf, W[p + "tln.w"], W[p + "tln.b"], tn)
            self._gemm(tn, W[p + "kv.w"], W[p + "kv.b"], out=kv)
            if transposed:
                ops.attn_kv(kv[:, :self.D], kv[:, self.D:], a_all[i], S, N, self.H, transposed=True)
            else:
                ops.eff_attn(ops.ATTN_KV_ONLY, S, N, self.H, k=kv[:, :self.D], v=kv[:, self.D:], a_out=a_all[i])
        self._text_cache = (xf_out, (xf_out._version, self.packed_generation, transposed), a_all)
        return a_all

    # ------------------------------------------------------------------------------------------ helpers
    def _gemm(self, a, w, bias, out=None, residual=None, res_row_mod=0, out_f32=None, act=ops.ACT_NONE):
        """out: activation-typed output; out_f32: fp32 output (residual stream / eps)."""
        if self.precision == "bf16":
            ops.gemm(a, w, bias=bias, residual=residual, res_row_mod=res_row_mod, out_f32=out_f32, out_bf16=out, act=act)
        else:
            if out_f32 is not None:
                ops.gemm(a, w, bias=bias, residual=residual, res_row_mod=res_row_mod, out_f32=out_f32, act=act)
                if out is not None and out.data_ptr() != out_f32.data_ptr():
                    raise RuntimeError("fp32 mode aliases the operand copy with the fp32 output")
            else:
                ops.gemm(a, w, bias=bias, residual=residual, res_row_mod=res_row_mod, out_f32=out, act=act)

    def set_lengths(self, ws, length, S, T):
        if length is None:
            ws["len"].fill_(T)
            return
        ln = torch.as_tensor(length).reshape(-1)
        if ln.numel() != S:
            raise ValueError(f"length must have {S} entries, got {ln.numel()}")
        ws["len"].copy_(ln.to(device=ws["len"].device, dtype=torch.int32, non_blocking=True).clamp(min=0, max=T))

    # ------------------------------------------------------------------------------------------ forward
    def time_table(self, n_steps):
        """time_embed(timestep_embedding(n)) for n = 0 .. n_steps-1, fp32 [n_steps, E]: the time MLP depends on nothing
        but the integer timestep, so the sampling loop reads it from this table (built with the same kernels; a row's
        value does not depend on how many rows the GEMM has).  Rebuilt when the weights change."""
        W = self.packed()
        key = (self.packed_generation, n_steps)
        if not isinstance(self._time_table, dict) or self._time_table.get("gen") != self.packed_generation:
            self._time_table = {"gen": self.packed_generation}      # one table per schedule length, all for these weights
        if n_steps in self._time_table:
            return self._time_table[n_steps]
        dev, dt = W["te0.w"].device, self.act_dtype
        t_all = torch.arange(n_steps, device=dev, dtype=torch.int64)
        temb = torch.empty(n_steps, self.D, device=dev, dtype=dt)
        te_h = torch.empty(n_steps, self.E, device=dev, dtype=dt)
        table = torch.empty(n_steps, self.E, device=dev, dtype=torch.float32)
        ops.timestep_embed(t_all, W["freqs"], temb)
        for lo in range(0, n_steps, 256):   # <= 256 rows per call: the same single-CTA kernel the per-step path (M = S) takes
            hi = min(lo + 256, n_steps)
            self._gemm(temb[lo:hi], W["te0.w"], W["te0.b"], out=te_h[lo:hi], act=ops.ACT_SILU)
            self._gemm(te_h[lo:hi], W["te2.w"], W["te2.b"], out_f32=table[lo:hi])
        self._time_table[n_steps] = table
        return table

    def embed(self, ws, t_dev, xf_proj, S, time_table=None):
        """emb = time_embed(timestep_embedding(t)) + xf_proj (:591), then every block's (scale|shift) in one GEMM."""
        W = self.packed()
        if time_table is not None:
            ops.time_table_silu(time_table, t_dev, xf_proj, ws["semb"])
        else:
            ops.timestep_embed(t_dev, W["freqs"], ws["temb"])
            self._gemm(ws["temb"], W["te0.w"], W["te0.b"], out=ws["te_h"], act=ops.ACT_SILU)
            # StylizationBlock applies SiLU(emb) before its linear (:74-77): fold it into this epilogue
            self._gemm(ws["te_h"], W["te2.w"], W["te2.b"], out=ws["semb"], residual=xf_proj, act=ops.ACT_SILU)
        self._gemm(ws["semb"], W["emb.w"], W["emb.b"], out_f32=ws["ss"])

    def _project(self, ws, W, p, write_xb):
        """x += out_layers(sact)  (StylizationBlock's Linear, :92-97, and the block's residual add)."""
        self._gemm(ws["sact"], W[p + ".po.w"], W[p + ".po.b"], residual=ws["xres"], out_f32=ws["xres"],
                   out=ws["xb"] if write_xb else None)

    def _ss(self, ws, W, p):
        i = W[p + ".ss"]
        return ws["ss"][:, i * 2 * self.D:(i + 1) * 2 * self.D]

    def _stylize_and_project(self, ws, W, p, T, write_xb):
        ops.ln_film_silu(ws["y"], W[p + ".po.ln.w"], W[p + ".po.ln.b"], ws["sact"], rows_per_seq=T,
                         scale_shift=self._ss(ws, W, p), silu=True)
        self._project(ws, W, p, write_xb)

    def _attend(self, ws, W, p, S, T, q, k=None, v=None, a_in=None, pair_shift=0, mask_v=True, q_softmaxed=False):
        """Attention output -> SiLU(FiLM(LN(.))) in ws['sact'].
        bf16: K/V half (A = softmax_time(K)^T V, 8 MB, stays in L2) then the fused query half + stylization front end:
        the attention output never reaches HBM.  fp32 mode: the unfused validation kernels."""
        H = self.H
        if self.precision == "bf16":
            tc = q_softmaxed and self.apply_tc
            if a_in is None:
                a_in = ws["a_blk"]
                if tc:
                    ops.attn_kv(k, v, a_in, S, T, H, length=ws["len"], pair_shift=pair_shift, transposed=True)
                else:
                    ops.eff_attn(ops.ATTN_KV_ONLY, S, T, H, k=k, v=v, a_out=a_in, length=ws["len"],
                                 pair_shift=pair_shift, mask_v=mask_v)
            if tc:
                ops.attn_apply_stylize_tc(q, a_in, W[p + ".po.ln.w"], W[p + ".po.ln.b"], ws["sact"], S, T, H,
                                          scale_shift=self._ss(ws, W, p), silu=True)
            else:
                ops.attn_apply_stylize(q, a_in, W[p + ".po.ln.w"], W[p + ".po.ln.b"], ws["sact"], S, T, H,
                                       scale_shift=self._ss(ws, W, p), silu=True, q_softmaxed=q_softmaxed)
            return
        if a_in is not None:
            ops.eff_attn(ops.ATTN_Q_ONLY, S, T, H, q=q, a_in=a_in, y=ws["y"])
        elif pair_shift:
            ops.eff_attn(ops.ATTN_INTER, S, T, H, q=q, k=k, v=v, y=ws["y"], length=ws["len"], pair_shift=pair_shift,
                         mask_v=False)
        else:
            ops.eff_attn(ops.ATTN_SELF, S, T, H, q=q, k=k, v=v, y=ws["y"], length=ws["len"], mask_v=True)
        ops.ln_film_silu(ws["y"], W[p + ".po.ln.w"], W[p + ".po.ln.b"], ws["sact"], rows_per_seq=T,
                         scale_shift=self._ss(ws, W, p), silu=True)

    def layers_stream(self, ws, a_text, S, T):
        """bf16 product schedule.  Per layer: 7 tcgen05 projections (hig_gemm_stream), 2 K/V-half kernels, 3 fused
        query-half + stylization kernels, 1 LayerNorm+FiLM+SiLU (FFN branch).  The three pre-attention LayerNorms
        (:119,153,190) are folded into the Q/K/V projections: each out-projection epilogue leaves the row statistics
        of the stream it has just updated in ws['stats'], the next projection reads the raw fp16 stream."""
        W, D = self.packed(), self.D
        gs = ops.gemm_stream
        ln_kind = ops.GS_LN_QSM if self.qsm else ops.GS_LN_BF16   # queries leave the projection already softmaxed
        qkv, xres, sact, stats = ws["qkv"], ws["xres"], ws["sact"], ws["stats"]
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        tok = S * T
        q_ca = qkv.view(-1)[:tok * D].view(tok, D)
        for li in range(self.L):
            p = f"l{li}."
            last = li == self.L - 1
            # --- self attention (:112-130)
            gs(ln_kind, xres, W[p + "sa.qkv.wg"], W[p + "sa.qkv.bg"], qkv, wsum=W[p + "sa.qkv.wsum"],
               stats_in=stats, ln_width=D)
            self._attend(ws, W, p + "sa", S, T, q, k, v, mask_v=True, q_softmaxed=self.qsm)
            gs(ops.GS_RES_H, sact, W[p + "sa.po.w"], W[p + "sa.po.b"], xres, stats_out=stats)
            # --- text cross attention (:145-165), K/V side precomputed in text_state()
            gs(ln_kind, xres, W[p + "ca.q.wg"], W[p + "ca.q.bg"], q_ca, wsum=W[p + "ca.q.wsum"],
               stats_in=stats, ln_width=D)
            self._attend(ws, W, p + "ca", S, T, q_ca, a_in=a_text[li], q_softmaxed=self.qsm)
            gs(ops.GS_RES_H, sact, W[p + "ca.po.w"], W[p + "ca.po.b"], xres, stats_out=stats if self.has_ic else None)
            # --- inter-person cross attention (:181-207): K,V of the partner, mask of the query side
            if self.has_ic:
                gs(ln_kind, xres, W[p + "ic.qkv.wg"], W[p + "ic.qkv.bg"], qkv, wsum=W[p + "ic.qkv.wsum"],
                   stats_in=stats, ln_width=D)
                self._attend(ws, W, p + "ic", S, T, q, k, v, pair_shift=S // 2, mask_v=False, q_softmaxed=self.qsm)
                gs(ops.GS_RES_H, sact, W[p + "ic.po.w"], W[p + "ic.po.b"], xres)
            # --- FFN (:261-264): no pre-norm, linear1 reads the fp16 stream, GELU fused in its epilogue
            gs(ops.GS_BF16_GELU, xres, W[p + "ffn.w1h"], W[p + "ffn.b1"], ws["g"])
            gs(ops.GS_BF16, ws["g"], W[p + "ffn.w2"], W[p + "ffn.b2"], ws["y"])
            ops.ln_film_silu(ws["y"], W[p + "ffn.po.ln.w"], W[p + "ffn.po.ln.b"], sact, rows_per_seq=T,
                             scale_shift=self._ss(ws, W, p + "ffn"), silu=True)
            if last and not self.heads16:
                self._project(ws, W, p + "ffn", True)        # also writes the bf16 copy the output heads read
            else:
                gs(ops.GS_RES_H, sact, W[p + "ffn.po.w"], W[p + "ffn.po.b"], xres, stats_out=None if last else stats)

    def layers(self, ws, a_text, S, T):
        if self.stream:
            return self.layers_stream(ws, a_text, S, T)
        W, D = self.packed(), self.D
        qkv = ws["qkv"]
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        tok = S * T
        q_ca = qkv.view(-1)[:tok * D].view(tok, D)   # dense [tok, 512] alias for the text-CA query
        for li in range(self.L):
            p = f"l{li}."
            # --- self attention (:112-130)
            ops.ln_film_silu(ws["xres"], W[p + "sa.ln.w"], W[p + "sa.ln.b"], ws["n"])
            self._gemm(ws["n"], W[p + "sa.qkv.w"], W[p + "sa.qkv.b"], out=qkv)
            self._attend(ws, W, p + "sa", S, T, q, k, v, mask_v=True)
            self._project(ws, W, p + "sa", False)
            # --- text cross attention (:145-165), K/V side precomputed in text_state()
            ops.ln_film_silu(ws["xres"], W[p + "ca.ln.w"], W[p + "ca.ln.b"], ws["n"])
            self._gemm(ws["n"], W[p + "ca.q.w"], W[p + "ca.q.b"], out=q_ca)
            self._attend(ws, W, p + "ca", S, T, q_ca, a_in=a_text[li])
            self._project(ws, W, p + "ca", not self.has_ic)
            # --- inter-person cross attention (:181-207): K,V of the partner, mask of the query side
            if self.has_ic:
                ops.ln_film_silu(ws["xres"], W[p + "ic.ln.w"], W[p + "ic.ln.b"], ws["n"])
                self._gemm(ws["n"], W[p + "ic.qkv.w"], W[p + "ic.qkv.b"], out=qkv)
                self._attend(ws, W, p + "ic", S, T, q, k, v, pair_shift=S // 2, mask_v=False)
                self._project(ws, W, p + "ic", True)
            # --- FFN (:261-264): no pre-norm, exact GELU fused in the linear1 epilogue
            self._gemm(ws["xb"], W[p + "ffn.w1"], W[p + "ffn.b1"], out=ws["g"], act=ops.ACT_GELU)
            self._gemm(ws["g"], W[p + "ffn.w2"], W[p + "ffn.b2"], out=ws["y"])
            self._stylize_and_project(ws, W, p + "ffn", T, True)

    def heads(self, ws, S, T):
        """out on frames 1.., out2 on frame 0 (:613-616) into eps [tok, 264] fp32 — or, on the product path, into
        eps16 [tok, 512] fp16 (columns >= 263 are zero) by two resident-W projections that read the fp16 stream."""
        W = self.packed()
        if self.heads16:
            eps16, xres = ws["eps16"], ws["xres"]
            ops.gemm_stream(ops.GS_F16, xres, W["out.w16"], W["out.b16"], eps16)
            a0 = xres.view(S, T * self.D)[:, :self.D]
            ops.gemm_stream(ops.GS_F16, a0, W["out2.w16"], W["out2.b16"], eps16.view(S, T * self.D)[:, :self.D])
            return eps16
        eps, xb = ws["eps"], ws["xb"]
        self._gemm(xb, W["out.w"], W["out.b"], out_f32=eps[:, :self.C])
        a0 = xb.view(S, T * self.D)[:, :self.D]
        o0 = eps.view(S, T * self.LD_EPS)[:, :self.C]
        self._gemm(a0, W["out2.w"], W["out2.b"], out_f32=o0)
        return eps

    def embed_motion(self, ws, T):
        W = self.packed()
        if self.stream and os.environ.get("HIG_EMBED_STREAM", "1") != "0":
            # stream <- positional rows, then stream += xa . W_in^T on the resident-W kernel, which also leaves the row
            # statistics the first LayerNorm-folded projection needs (no separate row_stats pass)
            ops.tile_rows(W["in.pos"], T, ws["xres"])
            ops.gemm_stream(ops.GS_RES_H, ws["xa"], W["in.w"], W["zero.b"], ws["xres"], stats_out=ws["stats"])
            return
        self._gemm(ws["xa"], W["in.w"], None, residual=W["in.pos"], res_row_mod=T, out_f32=ws["xres"])
        if self.stream:
            ops.row_stats(ws["xres"], ws["stats"])

    def run_packed(self, ws, t_dev, xf_proj, a_text, S, T, time_table=None):
        """Everything after pack_motion: ws['xa'] must hold the packed motion, ws['len'] the lengths."""
        self.embed(ws, t_dev, xf_proj, S, time_table)
        self.embed_motion(ws, T)
        self.layers(ws, a_text, S, T)
        return self.heads(ws, S, T)

    def forward(self, x, timesteps, length, xf_proj, xf_out):
        S, T, C = x.shape
        if S % 2:
            raise ValueError("the batch stacks person 1 and person 2 on dim 0: S must be even")
        if T > self.m.num_frames:
            raise ValueError(f"T={T} exceeds num_frames={self.m.num_frames}")
        ws = self.workspace(S, T)
        self.set_lengths(ws, length, S, T)
        a_text = self.text_state(xf_out)
        ops.pack_motion(x.detach().to(torch.float32).contiguous(), ws["xa"])
        t_dev = timesteps.to(device=x.device, dtype=torch.int64).contiguous()
        xfp = xf_proj.detach().to(torch.float32).contiguous()
        eps = self.run_packed(ws, t_dev, xfp, a_text, S, T)
        return eps.view(S, T, eps.shape[1])[:, :, :C].float().contiguous()
