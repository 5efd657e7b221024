"""CLIP-text-shaped encoder used when OpenAI's `clip` package is not installed (it is un-vendored and unpinned in
the reference: codes/requirements.txt:6, and its weights need network access).

`load_clip()` returns the real `clip.load('ViT-B/32')` model when the package imports, otherwise a random-init
module with the same text-side attribute names the reference touches in encode_text
(codes/models/interaction_transformer.py:533-559): token_embedding, positional_embedding, transformer (LND
layout, causal mask, QuickGELU MLP), ln_final, dtype — this is BASELINE.json's "random-init CLIP text encoder".
Text encoding is OFF the per-step hot path (once per batch when sampling); it stays plain PyTorch.
"""
import hashlib

import torch
from torch import nn

CONTEXT, VOCAB, WIDTH, LAYERS, HEADS = 77, 49408, 512, 12, 8
SOT, EOT = 49406, 49407


class _QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class _ResBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.attn = nn.MultiheadAttention(WIDTH, HEADS)
        self.ln_1 = nn.LayerNorm(WIDTH)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(WIDTH, 4 * WIDTH))
        self.mlp.add_module("gelu", _QuickGELU())
        self.mlp.add_module("c_proj", nn.Linear(4 * WIDTH, WIDTH))
        self.ln_2 = nn.LayerNorm(WIDTH)

    def forward(self, x, mask):
        n = self.ln_1(x)
        x = x + self.attn(n, n, n, need_weights=False, attn_mask=mask.to(x.dtype))[0]
        return x + self.mlp(self.ln_2(x))


class _TextTransformer(nn.Module):
    def __init__(self):
        super().__init__()
        self.resblocks = nn.Sequential(*[_ResBlock() for _ in range(LAYERS)])
        mask = torch.full((CONTEXT, CONTEXT), float("-inf")).triu_(1)
        self.register_buffer("attn_mask", mask, persistent=False)

    def forward(self, x):  # [L, N, D]
        m = self.attn_mask[:x.shape[0], :x.shape[0]]
        for blk in self.resblocks:
            x = blk(x, m)
        return x


class ClipTextStub(nn.Module):
    """Random-init stand-in exposing the attributes encode_text uses."""

    def __init__(self):
        super().__init__()
        self.token_embedding = nn.Embedding(VOCAB, WIDTH)
        self.positional_embedding = nn.Parameter(torch.empty(CONTEXT, WIDTH))
        self.transformer = _TextTransformer()
        self.ln_final = nn.LayerNorm(WIDTH)
        self.initialize_parameters()

    def initialize_parameters(self):
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (WIDTH ** -0.5) * ((2 * LAYERS) ** -0.5)
        attn_std, fc_std = WIDTH ** -0.5, (2 * WIDTH) ** -0.5
        for blk in self.transformer.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(blk.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(blk.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(blk.mlp.c_proj.weight, std=proj_std)

    @property
    def dtype(self):
        return self.positional_embedding.dtype


def tokenize(texts, truncate=True):
    """Deterministic stand-in for clip.tokenize: [SOT, hashed word ids..., EOT, 0...] so that, as in CLIP,
    argmax over the row finds the EOT position.  (The real BPE vocabulary file is not available offline.)"""
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), CONTEXT, dtype=torch.long)
    for i, s in enumerate(texts):
        words = str(s).lower().replace(".", " .").replace(",", " ,").split()
        ids = [SOT] + [int(hashlib.md5(w.encode()).hexdigest(), 16) % (SOT - 1) + 1 for w in words] + [EOT]
        if len(ids) > CONTEXT:
            if not truncate:
                raise RuntimeError(f"Input {s!r} is too long for context length {CONTEXT}")
            ids = ids[:CONTEXT]
            ids[-1] = EOT
        out[i, :len(ids)] = torch.tensor(ids)
    return out


LOADED = None     # "openai-clip" or "stub": which encoder the last load_clip() call returned


def load_clip(stub=None):
    """(model, tokenize_fn).  The real OpenAI CLIP ViT-B/32 (`clip.load`, models/interaction_transformer.py:436) when the
    `clip` package is installed; the random-init CLIP-shaped stub ONLY when asked for — `stub=True` or HIG_CLIP_STUB=1 — or
    when the package itself is absent (no network on the benchmark boxes; a warning says so).  A `clip` package that is
    present but fails to load its weights raises: a text-conditioned model must never silently run on a random encoder."""
    global LOADED
    import os
    import warnings
    if stub is None:
        stub = os.environ.get("HIG_CLIP_STUB", "") not in ("", "0")
    if not stub:
        try:
            import clip as _clip  # noqa
        except ImportError:
            _clip = None
        if _clip is not None and hasattr(_clip, "load") and getattr(_clip, "__file__", None):
            model, _ = _clip.load("ViT-B/32", "cpu")      # errors (missing weights, no network) propagate
            LOADED = "openai-clip"
            return model, _clip.tokenize
        warnings.warn("hig_b200: the `clip` package is not installed — using the RANDOM-INIT CLIP-shaped text encoder "
                      "(architecture-faithful, untrained).  Text conditioning is meaningless until real CLIP weights are "
                      "loaded; set HIG_CLIP_STUB=1 to silence this warning.", RuntimeWarning, stacklevel=2)
    LOADED = "stub"
    return ClipTextStub(), tokenize
