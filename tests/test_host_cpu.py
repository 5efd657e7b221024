"""Host-side logic of the drop-in module that needs no GPU: caption de-duplication + frozen-CLIP feature cache give
the same (xf_proj, xf_out) as the reference's straight-line encode_text (interaction_transformer.py:533-559);
generate_src_mask equals the reference's Python loop (:568-575); parameter names match the reference's state_dict."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _model(**kw):
    import hig_b200  # noqa: F401
    from hig_b200.interaction_transformer import MotionInteractionTransformer
    torch.manual_seed(0)
    return MotionInteractionTransformer(263, num_frames=196, num_layers=1, **kw)


def test_encode_text_dedup_and_cache_match_straight_line():
    m = _model(cap_id=False).eval()
    caps = ["a person pushes the other person", "a person kicks the other person", "a person pushes the other person",
            "two people shake hands", "a person kicks the other person"]
    with torch.no_grad():
        p1, o1 = m.encode_text(caps, "cpu")
        clip = m.clip
        tokens = m._tokenize(caps, truncate=True)
        x = clip.token_embedding(tokens).type(clip.dtype) + clip.positional_embedding.type(clip.dtype)
        x = clip.ln_final(clip.transformer(x.permute(1, 0, 2))).type(clip.dtype)
        xo = m.text_ln(m.textTransEncoder(m.text_pre_proj(x)))
        pr = m.text_proj(xo[tokens.argmax(dim=-1), torch.arange(xo.shape[1])])
        xo = xo.permute(1, 0, 2)
        p2, o2 = m.encode_text(list(reversed(caps)), "cpu")       # served from the cache, different order
    assert p1.shape == (5, 2048) and o1.shape == (5, 77, 256)
    assert (p1 - pr).abs().max() < 1e-5 and (o1 - xo).abs().max() < 1e-5
    assert (p2.flip(0) - p1).abs().max() < 1e-5 and len(m._clip_cache) == 3
    # the cache is dropped when a CLIP parameter changes
    with torch.no_grad():
        next(m.clip.parameters()).add_(1e-3)
        m.encode_text(caps[:1], "cpu")
    assert len(m._clip_cache) == 1
    # gradients reach the trainable text encoder through the index_select
    p3, o3 = m.encode_text(caps, "cpu")
    (p3.sum() + o3.sum()).backward()
    assert m.text_proj[0].weight.grad is not None and m.textTransEncoder.layers[0].linear1.weight.grad is not None
    assert all(p.grad is None for p in m.clip.parameters())


def test_generate_src_mask_matches_reference_loop():
    m = _model(cap_id=True)
    T, lengths = 9, [3, 9, 1, 0]
    ref = torch.ones(len(lengths), T)
    for i in range(len(lengths)):          # the reference's double loop, :570-574
        for j in range(lengths[i], T):
            ref[i, j] = 0
    assert torch.equal(m.generate_src_mask(T, lengths), ref)
    assert torch.equal(m.generate_src_mask(T, torch.tensor(lengths).view(-1, 1)), ref)


def test_state_dict_names_match_reference_inventory():
    import weights
    m = _model(cap_id=True)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == weights.param_shapes(num_layers=1)
