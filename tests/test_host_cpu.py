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


# ------------------------------------------------------------------------------------------------ round 2: training seams
def test_load_my_state_dict_follows_the_reference(capsys):
    """interaction_transformer.py:511-531: partial copy with the only_language / only_motion filters; unknown names are
    printed, except the CLIP / text-encoder keys of a text-conditioned checkpoint loaded into a cap_id model."""
    import argparse
    m = _model(cap_id=True)
    src = {k: torch.full_like(v, 0.25) for k, v in m.state_dict().items()}
    src["clip.positional_embedding"] = torch.zeros(77, 512)       # from a text-conditioned checkpoint
    src["textTransEncoder.layers.0.linear1.weight"] = torch.zeros(4, 4)
    src["not_a_key"] = torch.zeros(1)
    opt = argparse.Namespace(only_language=False, only_motion=False, cap_id=True)
    m.load_my_state_dict(src, opt)
    out = capsys.readouterr().out.split()
    assert out == ["not_a_key"]
    assert all(bool((v == 0.25).all()) for v in m.state_dict().values())
    # only_motion: language-side entries ("text" / "clip" in the name) are left alone — and printed, as in the reference
    m2 = _model(cap_id=True)
    before = {k: v.clone() for k, v in m2.state_dict().items()}
    m2.load_my_state_dict({k: v for k, v in src.items() if k in before}, argparse.Namespace(only_language=False, only_motion=True, cap_id=True))
    printed = set(capsys.readouterr().out.split())
    for k, v in m2.state_dict().items():
        lang = "clip" in k or "text" in k
        assert bool((v == 0.25).all()) != lang or bool((before[k] == 0.25).all()), k
        assert (k in printed) == lang, k
    # only_language: the mirror image
    m3 = _model(cap_id=True)
    m3.load_my_state_dict({k: v for k, v in src.items() if k in before}, argparse.Namespace(only_language=True, only_motion=False, cap_id=True))
    capsys.readouterr()
    for k, v in m3.state_dict().items():
        lang = "clip" in k or "text" in k
        assert bool((v == 0.25).all()) == lang, k


def test_synthetic_dataset_mirrors_the_reference_item_format():
    """datasets/mul_dataset.py:180-253: (caption1, caption2, motion1 [91,263], motion2 [91,263], m_length, file_id); row 0 is the
    LAST raw frame, rows 1.. a contiguous window or the whole motion padded with its last frame; caption ids are one-element
    lists that the default collate turns into [LongTensor[B]] (what DDPMMulTrainer.forward's cap_id branch must accept)."""
    import numpy as np
    from hig_b200.datasets import SyntheticText2MotionMulDataset, build_dataloader
    ds = SyntheticText2MotionMulDataset(n_items=12, cap_id=True, with_label=True, seed=3, min_len=30, max_len=150)
    for i in range(len(ds)):
        c1, c2, m1, m2, n, fid = ds[i]
        raw, lab = ds.items[i]["motion"], ds.items[i]["label"]
        a, b = (raw[1], raw[0]) if lab else (raw[0], raw[1])
        assert m1.shape == (91, 263) and m2.shape == (91, 263) and isinstance(c1, list) and len(c1) == 1
        nfr = raw.shape[1] - 1
        assert np.array_equal(m1[0], a[nfr]) and np.array_equal(m2[0], b[nfr])         # row 0 = the initialisation frame
        if nfr < 90:
            assert np.array_equal(m1[1:nfr + 1], a[:nfr]) and np.array_equal(m1[nfr + 1:], np.repeat(a[nfr - 1:nfr], 90 - nfr, 0))
        else:
            starts = [s for s in range(nfr - 89) if np.array_equal(m1[1], a[s])]
            assert starts and np.array_equal(m1[1:], a[starts[0]:starts[0] + 90])
        assert n == nfr
    batch = next(iter(build_dataloader(ds, 0, 1, samples_per_gpu=4, workers_per_gpu=0, shuffle=False)))
    c1, c2, m1, m2, lens, fid = batch
    assert isinstance(c1, list) and len(c1) == 1 and c1[0].shape == (4,) and c1[0].dtype == torch.int64
    assert m1.shape == (4, 91, 263) and m1.dtype == torch.float32 and lens.shape == (4,)
    text_ds = SyntheticText2MotionMulDataset(n_items=4, cap_id=False, with_label=False)
    assert isinstance(text_ds[0][0], str)


def test_sampling_seed_follows_torch_rng_and_differs_per_rank(monkeypatch):
    """the Philox key of the in-kernel posterior noise comes from torch's generator (torch.manual_seed controls
    sampling) and is mixed with the rank, so two ranks with the same seed and shard shape draw different noise."""
    import hig_b200  # noqa: F401
    import torch.distributed as dist
    from hig_b200.gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType,
                                             get_named_beta_schedule)
    mk = lambda: GaussianDiffusion(betas=get_named_beta_schedule("linear", 50), model_mean_type=ModelMeanType.EPSILON,
                                   model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
    torch.manual_seed(123)
    a = [mk()._next_seed() for _ in range(2)]
    torch.manual_seed(123)
    b = [mk()._next_seed() for _ in range(2)]
    torch.manual_seed(124)
    c = mk()._next_seed()
    assert a == b and a[0] != a[1] and c != a[0]
    d = mk()
    s0, s1 = d._next_seed(), d._next_seed()
    assert s0 != s1 and 0 <= s1 < 2 ** 64                         # successive calls of one object advance the key
    seeds = {}
    for rank in (0, 1):
        monkeypatch.setattr(dist, "is_initialized", lambda: True)
        monkeypatch.setattr(dist, "get_rank", lambda r=rank: r)
        torch.manual_seed(7)
        seeds[rank] = mk()._next_seed()
    assert seeds[0] != seeds[1]


def test_clip_stub_is_explicit(monkeypatch):
    """the random-init CLIP-shaped encoder is used when asked for (HIG_CLIP_STUB=1) or, with a warning, when the
    `clip` package is absent; the module records which encoder it holds."""
    import warnings
    import hig_b200  # noqa: F401
    from hig_b200 import clip_text
    monkeypatch.setenv("HIG_CLIP_STUB", "1")
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        model, tok = clip_text.load_clip()
    assert clip_text.LOADED == "stub" and isinstance(model, clip_text.ClipTextStub)
    monkeypatch.delenv("HIG_CLIP_STUB")
    try:
        import clip  # noqa: F401
        have = getattr(clip, "__file__", None) is not None
    except ImportError:
        have = False
    if not have:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            clip_text.load_clip()
        assert any("RANDOM-INIT" in str(x.message) for x in w)
        m = _model(cap_id=False)
        assert m.text_encoder_kind == "stub" and m._clip_dtype == torch.float32


def test_encode_unique_returns_distinct_captions_and_an_index():
    """forward() hands the training engine one text row per DISTINCT caption + an index per sequence; expanding it must give
    encode_text's output, and the EOT positions come from a per-caption cache that agrees with the tokeniser."""
    m = _model(cap_id=False).eval()
    caps = ["a person pushes the other person", "two people shake hands", "a person pushes the other person",
            "a person pushes the other person", "two people shake hands", "a person hugs the other person"]
    with torch.no_grad():
        pu, ou, idx = m._encode_unique(caps, "cpu")
        p, o = m.encode_text(caps, "cpu")
    assert pu.shape[0] == ou.shape[0] == 3 and idx.tolist() == [0, 1, 0, 0, 1, 2]
    assert torch.equal(pu.index_select(0, idx), p) and torch.equal(ou.index_select(0, idx), o)
    want = m._tokenize(caps, truncate=True).argmax(dim=-1)
    assert torch.equal(m._eot_index(caps, "cpu"), want) and torch.equal(m._eot_index(caps, "cpu"), want)     # cold, then cached
    assert set(m._eot_cache) == set(caps)
    # all captions distinct: no index
    with torch.no_grad():
        assert m._encode_unique(caps[:2], "cpu")[2] is None


def test_training_losses_terms_compute_mse_lazily():
    from hig_b200.gaussian_diffusion import _Terms
    noise, pred = torch.randn(4, 5, 6), torch.randn(4, 5, 6)
    t = _Terms(target=noise, pred=pred)
    assert "mse" in t and len(t) == 3 and not dict.__contains__(t, "mse")        # advertised, not computed yet
    assert t["pred"] is pred and t["target"] is noise and not dict.__contains__(t, "mse")
    want = ((noise - pred) ** 2).mean(dim=[1, 2])
    assert torch.equal(t["mse"], want) and dict.__contains__(t, "mse")
    assert sorted(t.keys()) == ["mse", "pred", "target"] and torch.equal(dict(t.items())["mse"], want)
    assert torch.equal(_Terms(target=noise, pred=pred).get("mse"), want)


def test_stage_falls_back_to_plain_copies_off_gpu():
    """staging.stage: pinned asynchronous copies on CUDA; for a CPU target (and for tensors that are not on the host) it is
    an ordinary .to() with the optional cast."""
    from hig_b200.staging import stage
    a = torch.arange(6, dtype=torch.int32)
    b = stage(a, "cpu", torch.long)
    assert b.dtype == torch.long and b.tolist() == list(range(6))
    assert stage([1.5, 2.5], torch.device("cpu"), torch.float32).tolist() == [1.5, 2.5]
    assert stage(torch.empty(0), "cpu").numel() == 0
