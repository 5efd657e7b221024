"""TextEncoderEngine — MotionInteractionTransformer.encode_text (codes/models/interaction_transformer.py:533-559) on the
library's kernels, for the no-grad paths (sampling, evaluation): SURVEY §8(f) rank 1.

  CLIP text transformer (12 x [LN -> causal 8-head MHA -> +res -> LN -> Linear 2048 / QuickGELU / Linear -> +res], ln_final)
  -> text_pre_proj (512 -> 256) -> 4 x post-norm nn.TransformerEncoderLayer (4-head MHA, erf-GELU FFN 2048) (:446-455)
  -> text_ln -> xf_out [U, 77, 256];  xf_proj = text_proj(xf_out[EOT]) (:552-557)

Every projection / MLP runs on the tcgen05 GEMM (bf16 operands, fp32 accumulate and fp32 residual stream; `precision="fp32"`
uses the fp32 GEMM), LayerNorms on hig_ln_film_silu, attention on hig_mha_attention, activations on hig_act_fwd; torch only
allocates, gathers the token embeddings and indexes the EOT rows.  Training keeps torch.autograd for the (small) trainable
text encoder — it needs a backward the library does not provide; the frozen CLIP stack is cached per caption either way.
"""
import torch

from . import ops


class TextEncoderEngine:
    def __init__(self, module):
        self.m = module
        self._key = None
        self._W = None

    # ------------------------------------------------------------------------------------------ operand copies
    def _packed(self):
        m = self.m
        params = list(m.clip.parameters()) + list(m.textTransEncoder.parameters()) + list(m.text_ln.parameters()) + \
            list(m.text_proj.parameters()) + (list(m.text_pre_proj.parameters()) if hasattr(m.text_pre_proj, "weight") else [])
        key = (m.precision, getattr(m, "_hig_param_generation", 0), tuple((p.data_ptr(), p._version) for p in params))
        if key == self._key:
            return self._W
        dt = torch.bfloat16 if m.precision == "bf16" else torch.float32
        op = lambda t: t.detach().to(dt).contiguous()
        f32 = lambda t: t.detach().float().contiguous()
        W = {"dt": dt, "clip": [], "enc": []}
        for blk in m.clip.transformer.resblocks:
            W["clip"].append({"ln1": (f32(blk.ln_1.weight), f32(blk.ln_1.bias)), "ln2": (f32(blk.ln_2.weight), f32(blk.ln_2.bias)),
                              "in": (op(blk.attn.in_proj_weight), f32(blk.attn.in_proj_bias)),
                              "out": (op(blk.attn.out_proj.weight), f32(blk.attn.out_proj.bias)),
                              "fc": (op(blk.mlp.c_fc.weight), f32(blk.mlp.c_fc.bias)),
                              "proj": (op(blk.mlp.c_proj.weight), f32(blk.mlp.c_proj.bias)),
                              "heads": blk.attn.num_heads})
        W["ln_final"] = (f32(m.clip.ln_final.weight), f32(m.clip.ln_final.bias))
        W["pre"] = (op(m.text_pre_proj.weight), f32(m.text_pre_proj.bias)) if hasattr(m.text_pre_proj, "weight") else None
        for lyr in m.textTransEncoder.layers:
            if getattr(lyr, "norm_first", False):
                raise NotImplementedError("the reference's text encoder is post-norm (nn.TransformerEncoderLayer default)")
            W["enc"].append({"in": (op(lyr.self_attn.in_proj_weight), f32(lyr.self_attn.in_proj_bias)),
                             "out": (op(lyr.self_attn.out_proj.weight), f32(lyr.self_attn.out_proj.bias)),
                             "l1": (op(lyr.linear1.weight), f32(lyr.linear1.bias)),
                             "l2": (op(lyr.linear2.weight), f32(lyr.linear2.bias)),
                             "n1": (f32(lyr.norm1.weight), f32(lyr.norm1.bias)), "n2": (f32(lyr.norm2.weight), f32(lyr.norm2.bias)),
                             "heads": lyr.self_attn.num_heads})
        W["text_ln"] = (f32(m.text_ln.weight), f32(m.text_ln.bias))
        W["proj"] = (op(m.text_proj[0].weight), f32(m.text_proj[0].bias))
        self._key, self._W = key, W
        return W

    # ------------------------------------------------------------------------------------------ helpers
    @staticmethod
    def _lin(x, wb, dt, residual=None, want_f32=False):
        """y = x W^T + b (+ residual).  Returns (operand-typed y or None, fp32 y or None)."""
        w, b = wb
        M, N = x.shape[0], w.shape[0]
        if dt == torch.float32:
            y = torch.empty(M, N, device=x.device, dtype=torch.float32)
            ops.gemm(x, w, bias=b, residual=residual, out_f32=y)
            return y, y
        y16 = None if want_f32 else torch.empty(M, N, device=x.device, dtype=dt)
        y32 = torch.empty(M, N, device=x.device, dtype=torch.float32) if want_f32 else None
        ops.gemm(x, w, bias=b, residual=residual, out_f32=y32, out_bf16=y16)
        return y16, y32

    @staticmethod
    def _cast(x, dt):
        """fp32 -> operand dtype through the library's cast kernel (hig_act_fwd with HIG_ACT_NONE)."""
        if x.dtype == dt:
            return x
        return ops.act_fwd(x.contiguous(), ops.ACT_NONE, torch.empty(x.shape, device=x.device, dtype=dt))

    @staticmethod
    def _ln(x32, gb, dt):
        out = torch.empty(x32.shape, device=x32.device, dtype=dt)
        return ops.ln_film_silu(x32, gb[0], gb[1], out)

    # ------------------------------------------------------------------------------------------ CLIP text transformer
    def clip_features(self, tokens):
        """tokens [U, 77] int64 (device) -> ln_final(transformer(token_embedding + positional_embedding)) as [U, 77, 512] fp32
        (what the reference computes in LND layout at :536-550)."""
        m, W = self.m, self._packed()
        dt = W["dt"]
        U, N = tokens.shape
        with torch.no_grad():
            x = (m.clip.token_embedding(tokens).float() + m.clip.positional_embedding.float()[:N]).reshape(U * N, -1).contiguous()
        D = x.shape[1]
        for blk in W["clip"]:
            n = self._ln(x, blk["ln1"], dt)
            qkv, _ = self._lin(n, blk["in"], dt)
            a = ops.mha_attention(qkv, torch.empty(U * N, D, device=x.device, dtype=dt), U, N, blk["heads"], causal=True)
            _, x = self._lin(a, blk["out"], dt, residual=x, want_f32=True)
            n = self._ln(x, blk["ln2"], dt)
            h, _ = self._lin(n, blk["fc"], dt)
            h = ops.act_fwd(h, ops.ACT_QUICKGELU, torch.empty_like(h))
            _, x = self._lin(h, blk["proj"], dt, residual=x, want_f32=True)
        out = torch.empty_like(x)
        ops.ln_film_silu(x, W["ln_final"][0], W["ln_final"][1], out)
        return out.view(U, N, D)

    # ------------------------------------------------------------------------------------------ 4-layer encoder + heads
    def encode(self, feats, eot):
        """feats [U, 77, 512] fp32 (CLIP features), eot [U] int64 -> (xf_proj [U, 2048], xf_out [U, 77, 256]) fp32."""
        W = self._packed()
        dt = W["dt"]
        U, N, _ = feats.shape
        dev = feats.device
        x = self._cast(feats.reshape(U * N, -1).float().contiguous(), dt)
        if W["pre"] is not None:
            _, x = self._lin(x, W["pre"], dt, want_f32=True)
        else:
            x = x.float()
        D = x.shape[1]
        for lyr in W["enc"]:
            xo = self._cast(x, dt)
            qkv, _ = self._lin(xo, lyr["in"], dt)
            a = ops.mha_attention(qkv, torch.empty(U * N, D, device=dev, dtype=dt), U, N, lyr["heads"], causal=False)
            _, r = self._lin(a, lyr["out"], dt, residual=x, want_f32=True)                   # x + SA(x)
            x = torch.empty_like(r)
            ops.ln_film_silu(r, lyr["n1"][0], lyr["n1"][1], x)                              # norm1, fp32 stream
            xo = self._cast(x, dt)
            h, _ = self._lin(xo, lyr["l1"], dt)
            h = ops.act_fwd(h, ops.ACT_GELU, torch.empty_like(h))
            _, r = self._lin(h, lyr["l2"], dt, residual=x, want_f32=True)                   # x + FF(x)
            x = torch.empty_like(r)
            ops.ln_film_silu(r, lyr["n2"][0], lyr["n2"][1], x)                              # norm2
        xf_out = torch.empty_like(x)
        ops.ln_film_silu(x, W["text_ln"][0], W["text_ln"][1], xf_out)
        xf_out = xf_out.view(U, N, D)
        rows = xf_out[torch.arange(U, device=dev), eot].contiguous()
        rows = self._cast(rows, dt)
        _, xf_proj = self._lin(rows, W["proj"], dt, want_f32=True)
        return xf_proj, xf_out
