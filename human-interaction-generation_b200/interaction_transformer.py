"""MotionInteractionTransformer — drop-in for the reference's role-aware denoiser
(codes/models/interaction_transformer.py:397-616) whose forward runs on hand-written sm_100a kernels.

What is kept identical to the reference (the Python callable contract):
  * constructor keywords, `forward(x, timesteps, length=None, text=None, xf_proj=None, xf_out=None)`,
    `encode_text`, `get_class_embedding`, `generate_src_mask`, attributes `num_frames`, `two_embed`;
  * every parameter name and shape, so reference checkpoints load with strict=True (cap_id models) and the
    optimizer / DDP wrappers see the same parameter list.
What is different: the nn.Linear / nn.LayerNorm children are parameter CONTAINERS only — they are never
called.  forward() hands the raw tensors to DenoiserEngine, which issues the C-ABI kernels; there is no
PyTorch-eager or CPU fallback (a CPU tensor raises).  `precision` selects bf16 tensor-core GEMMs (product) or
the fp32 validation mode.
"""
import torch
from torch import nn

from .denoiser_engine import DenoiserEngine


def _zeroed(module):
    """zero_module, interaction_transformer.py:62-68."""
    for p in module.parameters():
        p.detach().zero_()
    return module


def _stylization(latent_dim, time_embed_dim, dropout):
    """Parameter layout of StylizationBlock (:71-84): emb_layers.1, norm, out_layers.2 (zero-init)."""
    blk = nn.Module()
    blk.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(time_embed_dim, 2 * latent_dim))
    blk.norm = nn.LayerNorm(latent_dim)
    blk.out_layers = nn.Sequential(nn.SiLU(), nn.Dropout(p=dropout), _zeroed(nn.Linear(latent_dim, latent_dim)))
    return blk


def _attention(latent_dim, kv_dim, time_embed_dim, dropout, text_norm):
    """Parameter layout shared by the three efficient-attention blocks (:100-110, :132-143, :167-179)."""
    blk = nn.Module()
    blk.norm = nn.LayerNorm(latent_dim)
    if text_norm:
        blk.text_norm = nn.LayerNorm(kv_dim)
    blk.query = nn.Linear(latent_dim, latent_dim)
    blk.key = nn.Linear(kv_dim, latent_dim)
    blk.value = nn.Linear(kv_dim, latent_dim)
    blk.proj_out = _stylization(latent_dim, time_embed_dim, dropout)
    return blk


def _decoder_layer(latent_dim, text_latent_dim, time_embed_dim, ffn_dim, dropout, no_cross_attn):
    """LinearTemporalDiffusionTransformerDecoderLayer (:334-355)."""
    layer = nn.Module()
    layer.sa_block = _attention(latent_dim, latent_dim, time_embed_dim, dropout, text_norm=False)
    layer.ca_block = _attention(latent_dim, text_latent_dim, time_embed_dim, dropout, text_norm=True)
    if not no_cross_attn:
        layer.int_ca_block = _attention(latent_dim, latent_dim, time_embed_dim, dropout, text_norm=False)
    ffn = nn.Module()
    ffn.linear1 = nn.Linear(latent_dim, ffn_dim)
    ffn.linear2 = _zeroed(nn.Linear(ffn_dim, latent_dim))
    ffn.proj_out = _stylization(latent_dim, time_embed_dim, dropout)
    layer.ffn = ffn
    return layer


class _TextStack(nn.Module):
    """text_pre_proj -> textTransEncoder -> text_ln; xf_proj = text_proj(EOT token) (:551-559).  feats [77, U, 512] LND."""

    def __init__(self, net):
        super().__init__()
        self.pre, self.enc, self.ln, self.proj = net.text_pre_proj, net.textTransEncoder, net.text_ln, net.text_proj

    def forward(self, feats, eot):
        out = self.ln(self.enc(self.pre(feats)))
        proj = self.proj(out[eot, torch.arange(out.shape[1], device=out.device)])
        return proj, out.permute(1, 0, 2)


class MotionInteractionTransformer(nn.Module):
    def __init__(self, input_feats, num_frames=240, latent_dim=512, ff_size=1024, num_layers=8, num_heads=8,
                 dropout=0, activation="gelu", num_text_layers=4, text_latent_dim=256, text_ff_size=2048,
                 text_num_heads=4, no_clip=False, no_eff=False, no_cross_attn=False, cap_id=False,
                 precision="bf16", **kargs):
        super().__init__()
        if no_eff:
            raise NotImplementedError("--no_eff is broken in the reference for this model (the quadratic layer's "
                                      "forward takes 4 arguments but is called with 6, :389 vs :611); only the "
                                      "efficient-attention denoiser is implemented")
        if dropout != 0:
            raise NotImplementedError("the reference never sets dropout != 0 for this model (:405); not implemented")
        if activation != "gelu":
            raise NotImplementedError("only the reference's GELU FFN is implemented")
        self.num_frames, self.latent_dim, self.ff_size = num_frames, latent_dim, ff_size
        self.num_layers, self.num_heads, self.dropout, self.activation = num_layers, num_heads, dropout, activation
        self.input_feats = input_feats
        self.time_embed_dim = latent_dim * 4
        self.cap_id = cap_id
        self.no_cross_attn = no_cross_attn
        self.text_latent_dim = text_latent_dim

        if cap_id:
            self.cap_embedding = nn.Parameter(torch.randn(43, text_latent_dim))
            self.text_proj = nn.Sequential(nn.Linear(text_latent_dim, self.time_embed_dim))
        else:
            from . import clip_text
            self.clip, self._tokenize = clip_text.load_clip()
            self.text_encoder_kind = clip_text.LOADED        # "openai-clip" or "stub" (random-init, see clip_text.load_clip)
            # CLIP's `dtype` property reads visual.conv1.weight, which no_clip deletes below: keep it (the reference's d_type)
            self._clip_dtype = self.clip.dtype
            self.no_clip = no_clip
            if no_clip:
                self.clip.initialize_parameters()
                for name in ("visual", "logit_scale", "text_projection"):
                    if hasattr(self.clip, name):
                        delattr(self.clip, name)
            else:
                for p in self.clip.parameters():
                    p.requires_grad = False
            self.text_pre_proj = nn.Linear(512, text_latent_dim) if text_latent_dim != 512 else nn.Identity()
            enc_layer = nn.TransformerEncoderLayer(d_model=text_latent_dim, nhead=text_num_heads,
                                                   dim_feedforward=text_ff_size, dropout=dropout,
                                                   activation=activation)
            self.textTransEncoder = nn.TransformerEncoder(enc_layer, num_layers=num_text_layers)
            self.text_ln = nn.LayerNorm(text_latent_dim)
            self.text_proj = nn.Sequential(nn.Linear(text_latent_dim, self.time_embed_dim))

        self.sequence_embedding = nn.Parameter(torch.randn(num_frames, latent_dim))
        self.two_embed = True
        self.joint_embed = nn.Linear(input_feats, latent_dim)
        self.joint_embed2 = nn.Linear(4, latent_dim)
        self.time_embed = nn.Sequential(nn.Linear(latent_dim, self.time_embed_dim), nn.SiLU(),
                                        nn.Linear(self.time_embed_dim, self.time_embed_dim))
        self.temporal_decoder_blocks = nn.ModuleList(
            _decoder_layer(latent_dim, text_latent_dim, self.time_embed_dim, ff_size, dropout, no_cross_attn)
            for _ in range(num_layers))
        self.out = _zeroed(nn.Linear(latent_dim, input_feats))
        self.out2 = _zeroed(nn.Linear(latent_dim, input_feats))

        self._engines = {}
        self.precision = precision

    # ------------------------------------------------------------------------------------------ engine
    def engine(self, precision=None):
        precision = precision or self.precision
        eng = self._engines.get(precision)
        if eng is None:
            eng = self._engines[precision] = DenoiserEngine(self, precision)
        return eng

    # ------------------------------------------------------------------------------------------ text side (PyTorch)
    def encode_text(self, text, device):
        """:533-559 — CLIP text transformer -> text_pre_proj -> 4-layer encoder -> text_ln; xf_proj from the EOT token.

        Same values as the reference, less work: every distinct caption is encoded once per call and
        the result is indexed back to the batch (NTU RGB+D has 43 distinct captions, a training batch 256-512), and
        the FROZEN CLIP features of a caption are cached across calls (invalidated when a CLIP parameter changes)."""
        xf_proj, xf_out, idx = self._encode_unique(text, device)
        if idx is not None:
            xf_proj, xf_out = xf_proj.index_select(0, idx), xf_out.index_select(0, idx)
        return xf_proj, xf_out

    def _encode_unique(self, text, device):
        """(xf_proj [U, E], xf_out [U, 77, Dt], idx) for the U distinct captions of `text`; idx (LongTensor [len(text)], or None
        when every caption is distinct) maps each entry of `text` to its row."""
        text = list(text)
        uniq = list(dict.fromkeys(text))
        if self._text_on_kernels(device):
            # no-grad paths (sampling, evaluation): the whole text stack on the library's kernels (text_engine.py)
            feats = self._clip_features(uniq, device)           # [77, U, 512], frozen CLIP cached per caption
            eot = self._eot_index(uniq, device)
            xf_proj, xf_out = self.text_engine().encode(feats.permute(1, 0, 2).contiguous(), eot)
        elif self._text_graphed(device):
            # training: the trainable half of the text stack (pre-projection, 4-layer encoder, LayerNorm, EOT projection)
            # runs through torch.autograd, forward and backward each replayed as ONE CUDA graph per caption-count bucket
            # (about 350 launch-bound kernels otherwise); captions are independent rows, so padding rows change nothing.
            bucket = -(-len(uniq) // 16) * 16
            padded = uniq + [uniq[0]] * (bucket - len(uniq))
            feats = self._clip_features(padded, device)
            eot = self._eot_index(padded, device)
            xf_proj, xf_out = self._text_stack_graphed(bucket, feats, eot)
            xf_proj, xf_out = xf_proj[:len(uniq)], xf_out[:len(uniq)]
        else:
            feats = self._clip_features(uniq, device)           # [77, U, 512]
            eot = self._eot_index(uniq, device)
            xf_proj, xf_out = self._text_stack()(feats, eot)
        idx = None
        if len(uniq) != len(text):
            where = {c: i for i, c in enumerate(uniq)}
            from .staging import stage
            idx = stage(torch.tensor([where[c] for c in text], dtype=torch.long), device)
        return xf_proj, xf_out, idx

    def _text_stack(self):
        st = self.__dict__.get("_text_stack_mod")
        if st is None:
            st = _TextStack(self)
            object.__setattr__(self, "_text_stack_mod", st)     # not a registered submodule: no duplicate state_dict keys
        return st

    def _text_graphed(self, device):
        import os
        return torch.is_grad_enabled() and torch.device(device).type == "cuda" and not self.no_clip and \
            self.training and os.environ.get("HIG_TEXT_GRAPH", "1") != "0" and not torch.cuda.is_current_stream_capturing()

    def _text_stack_graphed(self, bucket, feats, eot):
        st = self._text_stack()
        # the graphs read the parameters in place: only their addresses (FusedAdam re-flattens) and grad flags key the capture
        import os
        key = (bucket, tuple(p.data_ptr() for p in st.parameters()), tuple(p.requires_grad for p in st.parameters()),
               os.environ.get("HIG_TEXT_TF32", "1"))
        cache = self.__dict__.setdefault("_text_graphs", {})
        fn = cache.get(key)
        if fn is None:
            for k in [k for k in cache if k[0] == bucket]:
                del cache[k]
            # make_graphed_callables rebinds .forward of the module it is given: hand it its own wrapper (same submodules).
            # bf16 mode: the encoder's fp32 GEMMs may use TF32 tensor cores (cuBLAS picks its kernels at capture time, so the
            # switch only needs to hold here) — as fp32 SIMT GEMMs, which is what torch runs by default, they cost 4.5 ms of a
            # 21 ms iteration; TF32's 10-bit mantissa is inside the mode's 1e-2 tolerance.  HIG_TEXT_TF32=0 keeps fp32.
            tf32 = self.precision == "bf16" and os.environ.get("HIG_TEXT_TF32", "1") != "0"
            before = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32 or before
            try:
                fn = torch.cuda.make_graphed_callables(_TextStack(self), (feats.detach().clone(), eot.clone()))
            finally:
                torch.backends.cuda.matmul.allow_tf32 = before
            cache[key] = fn
        return fn(feats, eot)

    def text_engine(self):
        eng = getattr(self, "_text_engine", None)
        if eng is None:
            from .text_engine import TextEncoderEngine
            eng = self._text_engine = TextEncoderEngine(self)
        return eng

    def _text_on_kernels(self, device):
        import os
        return (not torch.is_grad_enabled()) and torch.device(device).type == "cuda" and \
            os.environ.get("HIG_TEXT_ENGINE", "1") != "0" and self.text_latent_dim in (256, 512)

    def _eot_index(self, captions, device):
        """Position of the end-of-text token of every caption (the argmax of the token ids, :556) — remembered per caption: the
        BPE tokeniser is host Python, 0.1 ms per caption, and a training batch repeats a few dozen captions every iteration."""
        cache = self.__dict__.setdefault("_eot_cache", {})
        missing = [c for c in dict.fromkeys(captions) if c not in cache]
        if missing:
            pos = self._tokenize(missing, truncate=True).argmax(dim=-1).tolist()
            if len(cache) > 65536:
                cache.clear()
            cache.update(zip(missing, pos))
        from .staging import stage
        return stage(torch.tensor([cache[c] for c in captions], dtype=torch.long), device)

    def _clip_features(self, captions, device):
        """clip.ln_final(clip.transformer(token_embedding + positional_embedding)) per caption, LND layout (:536-550)."""
        clip = self.clip

        def run(caps):
            tokens = self._tokenize(caps, truncate=True).to(device)
            if self._text_on_kernels(device):
                return self.text_engine().clip_features(tokens).permute(1, 0, 2)      # LND like the reference
            dt = self._clip_dtype
            x = clip.token_embedding(tokens).type(dt)
            x = x + clip.positional_embedding.type(dt)
            x = clip.transformer(x.permute(1, 0, 2))
            return clip.ln_final(x).type(dt)

        if self.no_clip:                  # trainable CLIP-shaped encoder: nothing can be cached
            with torch.enable_grad():
                return run(captions)
        key = (str(device), sum(p._version for p in clip.parameters()))
        if getattr(self, "_clip_cache_key", None) != key:
            self._clip_cache_key, self._clip_cache = key, {}
        cache = self._clip_cache
        missing = [c for c in captions if c not in cache]
        if missing:
            with torch.no_grad():
                f = run(missing)
            for i, c in enumerate(missing):
                cache[c] = f[:, i].clone()
            if len(cache) > 8192:
                for c in list(cache)[:len(cache) - 8192]:
                    del cache[c]
        return torch.stack([cache[c] for c in captions], dim=1)

    def get_class_embedding(self, text):
        """:561-566 — text = [LongTensor of person-1 caption ids, LongTensor of person-2 caption ids]."""
        from .staging import stage
        ids = torch.cat([torch.as_tensor(t).reshape(-1) for t in text])
        ids = stage(ids, self.cap_embedding.device, torch.long)
        e = self.cap_embedding[ids]
        return self.text_proj(e), e.unsqueeze(1)

    def load_my_state_dict(self, state_dict, opt):
        """:511-531 — partial checkpoint load used by tools/train.py:50 (--pretrained) and tools/label_data.py:85: copies the
        entries that exist here, restricted to the language side (opt.only_language) or the motion side (opt.only_motion);
        names it skips are printed, as in the reference (cap_id models stay silent about CLIP / text-encoder keys)."""
        own_state = self.state_dict()
        only_language, only_motion = getattr(opt, "only_language", False), getattr(opt, "only_motion", False)
        with torch.no_grad():
            for name, param in state_dict.items():
                lang = "clip" in name or "text" in name
                if only_language:
                    if name not in own_state or not lang:
                        print(name)
                        continue
                elif only_motion:
                    if name not in own_state or lang:
                        print(name)
                        continue
                elif name not in own_state:
                    if not (getattr(opt, "cap_id", False) and lang):
                        print(name)
                    continue
                if isinstance(param, torch.nn.parameter.Parameter):
                    param = param.data
                own_state[name].copy_(param)
        self._hig_param_generation = getattr(self, "_hig_param_generation", 0) + 1   # packed operands are stale

    def generate_src_mask(self, T, length):
        """:568-575 — CPU FloatTensor [len(length), T]; vectorised (the reference loops in Python, 0.53 s at S=1024)."""
        ln = torch.as_tensor(length).reshape(-1).cpu()
        return (torch.arange(T)[None, :] < ln[:, None]).float()

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, x, timesteps, length=None, text=None, xf_proj=None, xf_out=None):
        """x [2B, T, input_feats] with persons stacked on dim 0 -> predicted noise, same shape (:577-616)."""
        if not x.is_cuda:
            raise RuntimeError("hig_b200.MotionInteractionTransformer runs on CUDA only: there is no CPU fallback")
        text_index = None
        if self.cap_id:
            xf_proj, xf_out = self.get_class_embedding(text)
        elif xf_proj is None or xf_out is None:
            # (xf_out stays one row per DISTINCT caption here; the training engine shares a caption's K/V side between the
            #  sequences that carry it, every other path expands it below)
            xf_proj, xf_out, text_index = self._encode_unique(text, x.device)
            if text_index is not None:
                xf_proj = xf_proj.index_select(0, text_index)
        if length is None:
            length = [x.shape[1]] * x.shape[0]
        needs_grad = torch.is_grad_enabled() and (
            x.requires_grad or xf_proj.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            import os
            if self.precision == "bf16" and os.environ.get("HIG_TRAIN_ENGINE", "1") != "0":
                # product training path: captured graphs over static buffers (train_engine.py)
                from .train_engine import denoiser_forward_graph
                return denoiser_forward_graph(self, x, timesteps, length, xf_proj, xf_out, text_index=text_index)
            if text_index is not None:
                xf_out = xf_out.index_select(0, text_index)
            from .autograd import denoiser_forward_with_grad     # fp32 validation mode: eager kernel schedule
            return denoiser_forward_with_grad(self, x, timesteps, length, xf_proj, xf_out)
        if text_index is not None:
            xf_out = xf_out.index_select(0, text_index)
        out = self.engine().forward(x, timesteps, length, xf_proj, xf_out)
        return out.to(x.dtype)
