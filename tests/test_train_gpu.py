"""Training path on the B200: backward kernels against torch.autograd of the same op (fp32 reference), and the
denoiser's parameter gradients against (a) the REAL reference's gradients (tests/golden/train.npz) and (b) the
differentiable CPU oracle on seeded inputs.  Everything goes through the public module / the C ABI.
Tolerances: fp32 mode 2e-4 relative (fp32 accumulation-order differences over tok-long reductions),
bf16 3e-2 relative per gradient tensor (bf16 operand rounding of activations and activation gradients)."""
import ast
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")
GTOL = {"fp32": 2e-4, "bf16": 3e-2}


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _ops():
    import hig_b200  # noqa: F401
    from hig_b200 import ops
    return ops


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("in_dt,out_dt", [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16),
                                          (torch.float32, torch.float32)])
@pytest.mark.parametrize("M,N", [(130, 263), (64, 64), (1000, 1536), (7, 5)])
def test_transpose_copy_colsum(cuda, M, N, in_dt, out_dt):
    ops = _ops()
    x = torch.randn(M, N, device=cuda).to(in_dt)
    xt = torch.zeros(N, (M + 7) // 8 * 8, device=cuda, dtype=out_dt)
    cp = torch.zeros(M, N, device=cuda, dtype=out_dt)
    cs = torch.zeros(N, device=cuda)
    ops.transpose(x, out_t=xt, copy=cp, colsum=cs)
    assert torch.equal(xt[:, :M], x.to(out_dt).t())
    assert torch.equal(cp, x.to(out_dt))
    assert rel(cs, x.float().sum(0)) < 1e-5
    # frame-0 rows dropped
    cs.zero_()
    ops.transpose(x, out_t=xt, copy=cp, colsum=cs, rows_zero_mod=5)
    keep = (torch.arange(M, device=cuda) % 5 != 0).to(x.dtype)[:, None]
    assert torch.equal(cp, (x * keep).to(out_dt))
    assert rel(cs, (x * keep).float().sum(0)) < 1e-5


@pytest.mark.parametrize("in_dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N", [(23296, 512), (1000, 1536), (37, 8), (256, 32768)])
def test_cast_and_colsum_without_transposition(cuda, M, N, in_dt):
    """hig_transpose with no transposed output = the vectorised cast (+ column sums) kernel of the training backward."""
    ops = _ops()
    x = torch.randn(M, N + 8, device=cuda).to(in_dt)[:, :N]
    cp = torch.zeros(M, N, device=cuda, dtype=torch.bfloat16)
    cs = torch.randn(N, device=cuda)
    base = cs.clone()
    ops.transpose(x, copy=cp, colsum=cs)
    assert torch.equal(cp, x.to(torch.bfloat16))
    assert rel(cs - base, x.float().sum(0)) < 2e-5
    cp2 = torch.zeros_like(cp)
    ops.transpose(x, copy=cp2)
    assert torch.equal(cp2, cp)


def test_colsum(cuda):
    ops = _ops()
    for M, N, dt in [(256, 91 * 512, torch.float32), (4, 1024, torch.float32), (3000, 77, torch.bfloat16)]:
        x = torch.randn(M, N, device=cuda).to(dt)
        out = torch.zeros(N, device=cuda)
        ops.colsum(x, out)
        assert rel(out, x.float().sum(0)) < 1e-5


@pytest.mark.parametrize("act", [1, 2])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_act_fwd_bwd(cuda, act, dt):
    ops = _ops()
    x = (2 * torch.randn(1000, 77, device=cuda)).to(dt)
    dy = torch.randn(1000, 77, device=cuda).to(dt)
    xr = x.float().requires_grad_(True)
    yr = F.gelu(xr) if act == 1 else F.silu(xr)
    yr.backward(dy.float())
    y = ops.act_fwd(x, act, torch.empty_like(x))
    dx = ops.act_bwd(x, dy, act, torch.empty_like(x))
    tol = 1e-6 if dt == torch.float32 else 5e-3
    assert rel(y, yr.detach()) < tol and rel(dx, xr.grad) < tol


@pytest.mark.parametrize("M,N,K", [(512, 512, 23296), (1536, 512, 1000), (264, 512, 784), (1024, 2048, 256),
                                   (2048, 512, 14)])
def test_gemm_splitk(cuda, M, N, K):
    ops = _ops()
    Kp = (K + 7) // 8 * 8
    a = torch.randn(M, Kp, device=cuda).bfloat16()[:, :K]
    w = torch.randn(N, Kp, device=cuda).bfloat16()[:, :K]
    base = torch.randn(M, N, device=cuda)
    out = base.clone()
    ops.gemm_splitk(a, w, out)
    ref = base.double() + a.double() @ w.double().t()
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize("M,N,K", [(512, 512, 23296), (1536, 512, 1000), (264, 512, 784), (512, 272, 23296),
                                   (2048, 2048, 256), (1024, 512, 100)])
def test_gemm_t_wgrad_form(cuda, M, N, K):
    """dW[M,N] += dY[K,M]^T X[K,N]: both operands MN-major (token-major as the activations lie), split-K atomics."""
    ops = _ops()
    Mp, Np = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    dy = torch.randn(K, Mp, device=cuda).bfloat16()[:, :M]
    x = torch.randn(K, Np, device=cuda).bfloat16()[:, :N]
    base = torch.randn(M, N, device=cuda)
    out = base.clone()
    ops.gemm_t(dy, x, trans_a=True, trans_b=True, out_f32=out, split_k=-1)
    ref = base.double() + dy.double().t() @ x.double()
    assert rel(out, ref) < 1e-5
    out2 = torch.empty(M, N, device=cuda)
    ops.gemm_t(dy, x, trans_a=True, trans_b=True, out_f32=out2)       # single pass, no atomics: deterministic
    assert rel(out2, dy.double().t() @ x.double()) < 1e-4           # one fp32 accumulator over all K tokens
    out3 = torch.empty(M, N, device=cuda)
    ops.gemm_t(dy, x, trans_a=True, trans_b=True, out_f32=out3)
    assert torch.equal(out2, out3)


@pytest.mark.parametrize("M,N,K", [(23296, 512, 512), (1000, 512, 1536), (784, 512, 264), (300, 1024, 512),
                                   (23296, 256, 1024), (130, 2048, 2048)])
def test_gemm_t_dgrad_form(cuda, M, N, K):
    """dX[M,N] = dY[M,K] W[K,N]: W consumed as stored ([out, in] row-major = MN-major B operand), bf16 / fp32-accumulate
    outputs."""
    ops = _ops()
    Kp = (K + 7) // 8 * 8
    dy = torch.randn(M, Kp, device=cuda).bfloat16()[:, :K]
    w = (torch.randn(K, N, device=cuda) / K ** 0.5).bfloat16()
    ref = dy.double() @ w.double()
    zb = torch.zeros(N, device=cuda)
    o16 = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    ops.gemm_t(dy, w, trans_b=True, bias=zb, out_bf16=o16)
    assert rel(o16, ref) < 4e-3
    base = torch.randn(M, N, device=cuda)
    acc = base.clone()
    ops.gemm_t(dy, w, trans_b=True, bias=zb, residual=acc, out_f32=acc)
    assert rel(acc, base.double() + ref) < 1e-5


@pytest.mark.parametrize("width", [512, 256])
@pytest.mark.parametrize("x_dt,g_dt,dx_dt", [(torch.float32, torch.float32, torch.float32),
                                             (torch.float32, torch.bfloat16, torch.float32),
                                             (torch.bfloat16, torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("film", [True, False])
def test_ln_film_silu_bwd(cuda, width, x_dt, g_dt, dx_dt, film):
    ops = _ops()
    S, T = 6, 37
    rows = S * T
    x = torch.randn(rows, width, device=cuda).to(x_dt)
    gamma = (1 + 0.1 * torch.randn(width, device=cuda))
    beta = 0.1 * torch.randn(width, device=cuda)
    ss = 0.5 * torch.randn(S, 2 * width + 64, device=cuda)[:, :2 * width] if film else None
    dout = torch.randn(rows, width, device=cuda).to(g_dt)
    # torch reference in fp32
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    h = F.layer_norm(xr, (width,), gr, br, 1e-5)
    if film:
        ssr = ss.clone().requires_grad_(True)
        sc, sh = ssr[:, :width], ssr[:, width:]
        h = h.view(S, T, width) * (1 + sc[:, None]) + sh[:, None]
        h = F.silu(h).view(rows, width)
    h.backward(dout.float())
    base = torch.randn(rows, width, device=cuda).to(dx_dt)
    dx = base.clone()
    d_ss = torch.zeros(S, 2 * width, device=cuda) if film else None
    d_gb = torch.zeros(S, 2 * width, device=cuda)
    acc = dx_dt == torch.float32
    ops.ln_film_silu_bwd(x, gamma, beta, dout, dx, T, scale_shift=ss, silu=film, dx_accumulate=acc, d_ss=d_ss, d_gb=d_gb)
    tol = 2e-5 if dx_dt == torch.float32 else 6e-3
    want = xr.grad + (base.float() if acc else 0)
    assert rel(dx, want) < tol
    assert rel(d_gb[:, :width].sum(0), gr.grad) < 1e-4 and rel(d_gb[:, width:].sum(0), br.grad) < 1e-4
    if film:
        assert rel(d_ss, ssr.grad) < 1e-4


def _attn_ref(q, k, v, mask_k, mask_v):
    """q [S,T,H,64], k/v [S,Tk,H,64], masks [S,Tk,1,1] (1 = valid)."""
    k = k + (1 - mask_k) * -1000000
    v = v * mask_v
    qs = F.softmax(q, dim=-1)
    ks = F.softmax(k, dim=1)
    att = torch.einsum("bnhd,bnhl->bhdl", ks, v)
    return torch.einsum("bnhd,bhdl->bnhl", qs, att)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("S,T,H", [(4, 196, 8), (6, 33, 8), (2, 1, 2)])
def test_eff_attn_bwd_self_and_inter(cuda, dtype, S, T, H):
    ops = _ops()
    D = H * 64
    g = torch.Generator(device=cuda).manual_seed(S * 1000 + T)
    qkv = torch.randn(S * T, 3 * D, device=cuda, generator=g).to(dtype)
    dy = torch.randn(S * T, D, device=cuda, generator=g).to(dtype)
    lens = torch.randint(max(1, T // 3), T + 1, (S,), device=cuda, generator=g).int()
    mask = (torch.arange(T, device=cuda)[None] < lens[:, None]).float().view(S, T, 1, 1)
    tol = 3e-5 if dtype == torch.float32 else 2e-2
    for mode in (ops.ATTN_SELF, ops.ATTN_INTER):
        r = qkv.float().view(S, T, 3, H, 64).clone().requires_grad_(True)
        q, k, v = r[:, :, 0], r[:, :, 1], r[:, :, 2]
        if mode == ops.ATTN_INTER:
            sw = lambda a: torch.cat([a[S // 2:], a[:S // 2]])
            y = _attn_ref(q, sw(k), sw(v), mask, torch.ones_like(mask))
        else:
            y = _attn_ref(q, k, v, mask, mask)
        y.backward(dy.float().view(S, T, H, 64))
        want = r.grad.view(S * T, 3 * D)
        d = torch.empty_like(qkv)
        base = torch.randn(3 * D, device=cuda)                 # bias-gradient column sums are ACCUMULATED
        sums = base.clone()
        ops.eff_attn_bwd(mode, S, T, H, q=qkv[:, :D], k=qkv[:, D:2 * D], v=qkv[:, 2 * D:], dy=dy, dq=d[:, :D],
                         dk=d[:, D:2 * D], dv=d[:, 2 * D:], length=lens, pair_shift=S // 2 if mode == ops.ATTN_INTER else 0,
                         q_sum=sums[:D], k_sum=sums[D:2 * D], v_sum=sums[2 * D:])
        want_sums = d.double().sum(0)
        assert ((sums - base).double() - want_sums).abs().max().item() < 1e-4 * max(1.0, want_sums.abs().max().item())
        for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
            if T == 1 and name != "dv":
                # one key row: Ks == 1, A has identical rows, so dQ and dK are exactly zero analytically
                assert (d[:, sl].float() - want[:, sl]).abs().max().item() < 2e-5
                continue
            e = rel(d[:, sl], want[:, sl])
            assert e < tol, (mode, name, e)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N", [77, 1])
def test_eff_attn_bwd_text(cuda, dtype, N):
    ops = _ops()
    S, T, H = 4, 50, 8
    D = H * 64
    q = torch.randn(S * T, D, device=cuda).to(dtype)
    kv = torch.randn(S * N, 2 * D, device=cuda).to(dtype)
    dy = torch.randn(S * T, D, device=cuda).to(dtype)
    qr = q.float().view(S, T, H, 64).clone().requires_grad_(True)
    kvr = kv.float().view(S, N, 2, H, 64).clone().requires_grad_(True)
    ones = torch.ones(S, N, 1, 1, device=cuda)
    y = _attn_ref(qr, kvr[:, :, 0], kvr[:, :, 1], ones, ones)
    y.backward(dy.float().view(S, T, H, 64))
    a = torch.empty(S, H, 64, 64, device=cuda, dtype=dtype)
    ops.eff_attn(ops.ATTN_KV_ONLY, S, N, H, k=kv[:, :D], v=kv[:, D:], a_out=a)
    dq = torch.empty_like(q)
    dA = torch.empty(S, H, 64, 64, device=cuda)
    sums = torch.zeros(3 * D, device=cuda)
    ops.eff_attn_bwd(ops.ATTN_Q_ONLY, S, T, H, q=q, a_in=a, dy=dy, dq=dq, dA=dA, q_sum=sums[:D])
    dkv = torch.empty_like(kv)
    ops.eff_attn_bwd(ops.ATTN_KV_ONLY, S, N, H, k=kv[:, :D], v=kv[:, D:], dk=dkv[:, :D], dv=dkv[:, D:], dA=dA,
                     k_sum=sums[D:2 * D], v_sum=sums[2 * D:])
    want_sums = torch.cat([dq.double().sum(0), dkv.double().sum(0)])
    assert (sums.double() - want_sums).abs().max().item() < 1e-4 * max(1.0, want_sums.abs().max().item())
    tol = 3e-5 if dtype == torch.float32 else 2e-2
    if N > 1:
        assert rel(dq, qr.grad.view(S * T, D)) < tol
    else:       # one token: A has identical rows, dQ is analytically zero (rounding noise ~1e-9 on both sides)
        assert (dq.float() - qr.grad.view(S * T, D)).abs().max().item() < 1e-5
    assert rel(dkv[:, D:], kvr.grad[:, :, 1].reshape(S * N, D)) < tol
    if N > 1:   # a single token has softmax == 1 and a zero key gradient
        assert rel(dkv[:, :D], kvr.grad[:, :, 0].reshape(S * N, D)) < tol
    else:
        assert dkv[:, :D].float().abs().max().item() < 2e-5     # analytically zero: rounding noise of two summation orders


# ------------------------------------------------------------------------------------------------ model level
def _build(layers, precision, cuda, seed=0):
    import weights
    import hig_b200  # noqa: F401
    from hig_b200.interaction_transformer import MotionInteractionTransformer
    m = MotionInteractionTransformer(263, num_frames=196, num_layers=layers, latent_dim=512, cap_id=True,
                                     precision=precision)
    sd = weights.make_state_dict(seed=seed, num_layers=layers)
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda).train()
    m.cap_id = False
    return m, sd


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_training_grads_match_reference_golden(cuda, precision):
    """The labelled training step of the REAL reference (q_sample -> denoiser -> masked MSE -> backward): prediction,
    loss and the gradient of EVERY parameter (norms) plus four full gradient tensors."""
    import diffusion_oracle as DF
    import weights
    from hig_b200.gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType,
                                             get_named_beta_schedule)
    d = np.load(os.path.join(GOLDEN, "train.npz"))
    cfg = ast.literal_eval(str(d["cfg"]))
    inp = weights.make_inputs(cfg["seed"], cfg["S"], cfg["T"], n_text=cfg["n_text"], lengths=cfg["lengths"])
    m, _ = _build(cfg["layers"], precision, cuda)
    diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", 1000), model_mean_type=ModelMeanType.EPSILON,
                             model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
    noise = weights.make_noise(cfg["seed"] + 100, 0, cfg["S"], cfg["T"])[0].to(cuda)
    g = lambda k: inp[k].to(cuda)
    terms = diff.training_losses(m, g("x"), g("t"), noise=noise,
                                 model_kwargs={"xf_proj": g("xf_proj"), "xf_out": g("xf_out"), "length": g("length")})
    tol = GTOL[precision]
    ftol = 1e-5 if precision == "fp32" else 1e-2
    assert rel(terms["pred"].detach(), d["pred"]) < ftol
    mask = (torch.arange(cfg["T"], device=cuda)[None] < g("length")[:, None]).float()
    loss = DF.masked_mse_loss(terms["pred"], noise, mask)
    assert abs(loss.item() - float(d["loss_label"])) < max(ftol, 1e-5) * abs(float(d["loss_label"])) * 2
    loss.backward()
    named = dict(m.named_parameters())
    for key in d.files:
        if key.startswith("grad:"):
            e = rel(named[key[5:]].grad, d[key])
            assert e < tol, (key, e)
    bad = []
    for n, gn in zip([str(n) for n in d["grad_names"]], d["grad_norms"]):
        if n.startswith(("text_proj", "cap_embedding")):
            continue        # not reached on the xf_proj/xf_out-given branch
        got = 0.0 if named[n].grad is None else named[n].grad.norm().item()
        if gn < 0:
            continue
        # key.bias gradients are analytically zero (the time softmax is shift invariant): absolute floor in bf16
        if abs(got - gn) > tol * max(gn, 1e-9) * 2 + (1e-9 if precision == "fp32" else 2e-5):
            bad.append((n, got, float(gn)))
    assert not bad, bad[:8]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_training_grads_against_oracle(cuda, precision):
    """3 layers, T=64, 77 text tokens, mixed lengths: every parameter gradient and the gradients that flow back into
    the text encoder (xf_proj, xf_out) against torch.autograd through the CPU oracle."""
    import denoiser_oracle as DO
    import diffusion_oracle as DF
    import weights
    L, S, T = 3, 6, 64
    m, sd = _build(L, precision, cuda, seed=4)
    inp = weights.make_inputs(41, S, T, n_text=77, lengths=[64, 40, 17, 64, 40, 17])
    tgt = weights.make_noise(7, 0, S, T)[0]
    # oracle
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xfp, xfo = inp["xf_proj"].clone().requires_grad_(True), inp["xf_out"].clone().requires_grad_(True)
    pred_o = DO.denoiser_forward(sdg, inp["x"], inp["t"], inp["length"], xfp, xfo)
    mask = DO.src_mask_from_length(T, inp["length"], "cpu")
    DF.masked_mse_loss(pred_o, tgt, mask).backward()
    # B200
    g = lambda k: inp[k].to(cuda)
    xfp_c, xfo_c = g("xf_proj").requires_grad_(True), g("xf_out").requires_grad_(True)
    pred = m(g("x"), g("t"), length=g("length"), xf_proj=xfp_c, xf_out=xfo_c)
    DF.masked_mse_loss(pred, tgt.to(cuda), mask.to(cuda)).backward()
    tol = GTOL[precision]
    assert rel(pred.detach(), pred_o.detach()) < (1e-5 if precision == "fp32" else 1e-2)
    assert rel(xfp_c.grad, xfp.grad) < tol and rel(xfo_c.grad, xfo.grad) < tol
    worst = []
    for n, p in m.named_parameters():
        if n.startswith(("text_proj", "cap_embedding")):
            continue
        go = sdg[n].grad
        if n == "sequence_embedding":
            assert p.grad[T - 1:].abs().max().item() == 0.0
        if n.endswith("key.bias"):      # analytically zero (time-softmax shift invariance)
            assert p.grad.abs().max().item() < 1e-4
            continue
        e = rel(p.grad, go)
        worst.append((e, n))
    worst.sort(reverse=True)
    assert worst[0][0] < tol, worst[:6]


def test_trainer_update_step_runs_and_learns(cuda):
    """DDPMMulTrainer.forward/update (labelled and PIT) through Adam: the loss on a fixed batch goes down."""
    import argparse
    from hig_b200.interaction_transformer import MotionInteractionTransformer
    from hig_b200.mul_ddpm_trainer import DDPMMulTrainer
    torch.manual_seed(0)
    np.random.seed(0)
    m = MotionInteractionTransformer(263, num_frames=196, num_layers=2, latent_dim=512, cap_id=True)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.abs().max() == 0 and "norm.bias" not in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    for label_path in ("labels.npy", None):
        opt = argparse.Namespace(device=cuda, multi=True, label_path=label_path, cap_id=True, diffusion_steps=1000,
                                 is_train=True)
        tr = DDPMMulTrainer(opt, m)
        tr.opt_encoder = torch.optim.Adam(m.parameters(), lr=2e-4)
        tr.train_mode()
        B, T = 4, 24
        batch = ([1, 2, 3, 4], [5, 6, 7, 8], torch.randn(B, T, 263), torch.randn(B, T, 263),
                 torch.tensor([24, 20, 9, 24]), None)
        losses = []
        for _ in range(6):
            np.random.seed(3)      # same timesteps every iteration
            torch.manual_seed(3)   # same noise
            tr.forward(batch)
            losses.append(tr.update()["loss_mot_rec"])
        assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def test_nccl_data_parallel_two_gpus(cuda):
    """Gradient all-reduce over NCCL + sharded sampling with 2 ranks (skipped on a 1-GPU box; the host logic is
    covered on CPU by tests/test_ddp_cpu.py)."""
    import subprocess
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tools", "ddp_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ddp_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("M,N,K", [(23296, 1024, 512), (1000, 256, 512), (2048, 512, 1024)])
def test_gemm_fused_forward_and_gated_backward(cuda, M, N, K):
    """linear -> GELU with both the pre-activation and the activation written by the GEMM epilogue, and the input-gradient
    GEMM of the next linear gated by GELU'(saved pre-activation): against the unfused kernels they replace (same arithmetic:
    bit-identical) and against fp64."""
    ops = _ops()
    g = torch.Generator(device=cuda).manual_seed(M + N)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    w = (torch.randn(N, K, device=cuda, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=cuda, generator=g)
    pre, act = torch.empty(M, N, device=cuda, dtype=torch.bfloat16), torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    ops.gemm_fused(a, w, b, act, act=ops.ACT_GELU, out_pre=pre)
    pre_ref = torch.empty_like(pre)
    ops.gemm(a, w, bias=b, out_bf16=pre_ref)
    act_ref = ops.act_fwd(pre_ref, ops.ACT_GELU, torch.empty_like(pre_ref))
    assert torch.equal(pre, pre_ref) and torch.equal(act, act_ref)
    want = torch.nn.functional.gelu(a.double() @ w.double().t() + b.double())
    assert rel(act, want) < 6e-3
    # backward: d_pre = (dy @ w2) * GELU'(pre), w2 [K2, N] as nn.Linear stores it
    K2 = 512
    dy = torch.randn(M, K2, device=cuda, generator=g).bfloat16()
    w2 = (torch.randn(K2, N, device=cuda, generator=g) / K2 ** 0.5).bfloat16()
    zb = torch.zeros(N, device=cuda)
    d_pre = ops.gemm_fused(dy, w2, zb, torch.empty(M, N, device=cuda, dtype=torch.bfloat16), trans_b=True, gate=pre,
                           gate_act=ops.ACT_GELU)
    d_act = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    ops.gemm_t(dy, w2, trans_b=True, bias=zb, out_bf16=d_act)
    d_ref = ops.act_bwd(pre, d_act, ops.ACT_GELU, torch.empty_like(d_act))
    # the unfused path rounds d_act to bf16 before the multiplication: one extra rounding, not a different formula
    assert rel(d_pre, d_ref) < 6e-3
    p64 = pre.double().requires_grad_(True)
    torch.nn.functional.gelu(p64).backward(dy.double() @ w2.double())
    assert rel(d_pre, p64.grad) < 6e-3
