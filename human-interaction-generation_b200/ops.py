"""Tensor-level wrappers over the C ABI (include/hig_b200.h).  PyTorch supplies device memory and the current
stream only; all arithmetic happens in the sm_100a kernels of csrc/.  CPU tensors are rejected — there is no
fallback path.
"""
import torch

from . import _lib

BF16, F32, F16 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_SILU, ACT_QUICKGELU = 0, 1, 2, 3
ATTN_SELF, ATTN_INTER, ATTN_KV_ONLY, ATTN_Q_ONLY = 0, 1, 2, 3


def _dt(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    raise TypeError(f"hig_b200: unsupported dtype {t.dtype}")


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("hig_b200: CUDA tensor required (no CPU fallback)")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rowmajor(t, what):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"hig_b200: {what} must be 2-D with unit inner stride, got {tuple(t.shape)} {t.stride()}")
    return t.stride(0)


def gemm(a, w, bias=None, residual=None, res_row_mod=0, out_f32=None, out_bf16=None, act=ACT_NONE):
    """act(a @ w.T + bias + residual).  a [M,K], w [N,K] (both bf16 -> tcgen05 kernel; both fp32 -> fp32 mode)."""
    lib = _lib.load()
    M, K = a.shape
    N = w.shape[0]
    lda, ldw = _rowmajor(a, "A"), _rowmajor(w, "W")
    if w.shape[1] != K:
        raise ValueError("hig_b200.gemm: K mismatch")
    ldr = _rowmajor(residual, "residual") if residual is not None else 0
    if bias is not None and bias.dtype != torch.float32:
        raise TypeError("hig_b200.gemm: bias must be fp32")
    half = (residual is not None and residual.dtype == torch.float16) or \
        (out_f32 is not None and out_f32.dtype == torch.float16)
    if half:
        # fp16 residual stream (product path): `out_f32` carries the main (stream) output whatever its dtype
        if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
            raise TypeError("hig_b200.gemm: the fp16 residual stream belongs to the bf16 path")
        rc = lib.hig_gemm_bf16_ex(_ptr(a), lda, _ptr(w), ldw, M, N, K, _ptr(bias), _ptr(residual),
                                  _dt(residual) if residual is not None else F32, ldr, res_row_mod, _ptr(out_f32),
                                  _dt(out_f32) if out_f32 is not None else F32,
                                  _rowmajor(out_f32, "out") if out_f32 is not None else 0, _ptr(out_bf16),
                                  _rowmajor(out_bf16, "out_bf16") if out_bf16 is not None else 0, act, _stream())
        _lib.check(rc, "hig_gemm_bf16_ex")
        return out_f32 if out_bf16 is None else (out_bf16 if out_f32 is None else (out_f32, out_bf16))
    if residual is not None and residual.dtype != torch.float32:
        raise TypeError("hig_b200.gemm: residual must be fp32 or fp16")
    if a.dtype == torch.bfloat16:
        if w.dtype != torch.bfloat16:
            raise TypeError("hig_b200.gemm: A/W dtype mismatch")
        if out_f32 is None and out_bf16 is None:
            out_bf16 = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
        rc = lib.hig_gemm_bf16(_ptr(a), lda, _ptr(w), ldw, M, N, K, _ptr(bias), _ptr(residual), ldr, res_row_mod,
                               _ptr(out_f32), _rowmajor(out_f32, "out_f32") if out_f32 is not None else 0,
                               _ptr(out_bf16), _rowmajor(out_bf16, "out_bf16") if out_bf16 is not None else 0,
                               act, _stream())
        _lib.check(rc, "hig_gemm_bf16")
        return out_f32 if out_bf16 is None else (out_bf16 if out_f32 is None else (out_f32, out_bf16))
    if a.dtype == torch.float32:
        if w.dtype != torch.float32 or out_bf16 is not None:
            raise TypeError("hig_b200.gemm: fp32 mode takes fp32 W and fp32 output")
        if out_f32 is None:
            out_f32 = torch.empty((M, N), device=a.device, dtype=torch.float32)
        rc = lib.hig_gemm_f32(_ptr(a), lda, _ptr(w), ldw, M, N, K, _ptr(bias), _ptr(residual), ldr, res_row_mod,
                              _ptr(out_f32), _rowmajor(out_f32, "out_f32"), act, _stream())
        _lib.check(rc, "hig_gemm_f32")
        return out_f32
    raise TypeError(f"hig_b200.gemm: unsupported dtype {a.dtype}")


GS_BF16, GS_BF16_GELU, GS_RES_H, GS_LN_BF16, GS_F16, GS_LN_QSM = 0, 1, 2, 3, 4, 5


def gemm_stream(kind, a, w, bias, out, wsum=None, stats_in=None, stats_out=None, ln_width=0):
    """Token-sized projection with the TMA-staged epilogue (csrc/gemm_stream.cu).  a [M,K], w [N,K] both bf16 or both
    fp16; out bf16 [M,N] (GS_BF16 / GS_BF16_GELU / GS_LN_BF16) or the fp16 residual stream updated in place (GS_RES_H)."""
    lib = _lib.load()
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K or a.dtype != w.dtype or a.dtype not in (torch.bfloat16, torch.float16):
        raise TypeError("hig_b200.gemm_stream: A and W must both be bf16 or both fp16, with matching K")
    want = torch.float16 if kind in (GS_RES_H, GS_F16) else torch.bfloat16
    if kind == GS_LN_QSM and (wsum is None or stats_in is None):
        raise ValueError("hig_b200.gemm_stream: GS_LN_QSM needs wsum and stats_in like GS_LN_BF16")
    if out.dtype != want or tuple(out.shape) != (M, N):
        raise TypeError(f"hig_b200.gemm_stream: out must be {want} [{M},{N}]")
    for t, nm in ((bias, "bias"), (wsum, "wsum"), (stats_in, "stats_in"), (stats_out, "stats_out")):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
            raise TypeError(f"hig_b200.gemm_stream: {nm} must be contiguous fp32")
    for t, nm in ((stats_in, "stats_in"), (stats_out, "stats_out")):
        if t is not None and tuple(t.shape) != (M, 8):
            raise ValueError(f"hig_b200.gemm_stream: {nm} must be [M, 8]")
    rc = lib.hig_gemm_stream(kind, _ptr(a), _rowmajor(a, "A"), _ptr(w), _rowmajor(w, "W"), _dt(a), M, N, K, _ptr(bias),
                             _ptr(wsum), _ptr(stats_in), _ptr(stats_out), ln_width, _ptr(out), _rowmajor(out, "out"),
                             _stream())
    _lib.check(rc, "hig_gemm_stream")
    return out


def row_stats(x, stats):
    """stats[row] = (sum, sum of squares, 0 x 6) of the fp16 rows of x [rows, 512] (layout of gemm_stream's partials)."""
    lib = _lib.load()
    rows, width = x.shape
    if not x.is_contiguous() or stats.dtype != torch.float32 or tuple(stats.shape) != (rows, 8) or not stats.is_contiguous():
        raise ValueError("hig_b200.row_stats: x contiguous [rows, 512], stats contiguous fp32 [rows, 8]")
    rc = lib.hig_row_stats(_ptr(x), _dt(x), rows, width, _ptr(stats), _stream())
    _lib.check(rc, "hig_row_stats")
    return stats


def ln_film_silu(x, gamma, beta, out, rows_per_seq=1, scale_shift=None, silu=False):
    """out = [SiLU](LN(x) * (1 + scale) + shift); x [rows, W] (W in {256, 512}); scale_shift fp32 view [S, >=2W]."""
    lib = _lib.load()
    rows, width = x.shape
    if not x.is_contiguous() or not out.is_contiguous():
        raise ValueError("hig_b200.ln_film_silu: contiguous tensors required")
    ss_stride = 0
    if scale_shift is not None:
        if scale_shift.dtype != torch.float32 or scale_shift.stride(1) != 1:
            raise ValueError("hig_b200.ln_film_silu: scale_shift must be fp32 with unit inner stride")
        ss_stride = scale_shift.stride(0)
    rc = lib.hig_ln_film_silu(_ptr(x), _dt(x), rows, width, rows_per_seq, _ptr(gamma), _ptr(beta),
                              _ptr(scale_shift), ss_stride, 1 if silu else 0, _ptr(out), _dt(out), _stream())
    _lib.check(rc, "hig_ln_film_silu")
    return out


def eff_attn(mode, S, T, H, q=None, k=None, v=None, a_in=None, a_out=None, y=None, length=None, pair_shift=0,
             mask_v=True):
    """Fused efficient attention; q/k/v/y are 2-D views [S*T, ld] positioned at head 0."""
    lib = _lib.load()
    ref = q if q is not None else k
    dt = _dt(ref)
    ldq = q.stride(0) if q is not None else 0
    ldkv = k.stride(0) if k is not None else 0
    if k is not None and v.stride(0) != ldkv:
        raise ValueError("hig_b200.eff_attn: K and V must share a leading dimension")
    ldy = y.stride(0) if y is not None else 0
    if length is not None and length.dtype != torch.int32:
        raise TypeError("hig_b200.eff_attn: length must be int32")
    rc = lib.hig_eff_attn(mode, _ptr(q), ldq, _ptr(k), _ptr(v), ldkv, _ptr(a_in), _ptr(a_out), _ptr(y), ldy,
                          _ptr(length), S, T, H, pair_shift, 1 if mask_v else 0, dt, _stream())
    _lib.check(rc, "hig_eff_attn")
    return y if y is not None else a_out


def attn_apply_stylize(q, a_in, gamma, beta, out, S, T, H, scale_shift=None, silu=True, q_softmaxed=False, y_out=None):
    """out = [SiLU](LN(concat_h softmax_feat(q_h) @ a_in[s,h]) * (1 + scale) + shift), bf16; q is a [S*T, ld] view.
    q_softmaxed: q already holds softmax_feat(Q) (written by a GS_LN_QSM projection)."""
    lib = _lib.load()
    if q.dtype != torch.bfloat16 or a_in.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        raise TypeError("hig_b200.attn_apply_stylize: bf16 storage only (fp32 mode uses eff_attn + ln_film_silu)")
    if not out.is_contiguous() or not a_in.is_contiguous():
        raise ValueError("hig_b200.attn_apply_stylize: contiguous out / a_in required")
    ss_stride = 0
    if scale_shift is not None:
        if scale_shift.dtype != torch.float32 or scale_shift.stride(1) != 1:
            raise ValueError("hig_b200.attn_apply_stylize: scale_shift must be fp32 with unit inner stride")
        ss_stride = scale_shift.stride(0)
    flags = (1 if silu else 0) | (2 if q_softmaxed else 0)
    if y_out is not None:
        if y_out.dtype != torch.bfloat16 or not y_out.is_contiguous() or y_out.shape != out.shape:
            raise ValueError("hig_b200.attn_apply_stylize: y_out must be a contiguous bf16 tensor shaped like out")
        rc = lib.hig_attn_apply_stylize_y(_ptr(q), q.stride(0), _ptr(a_in), _ptr(gamma), _ptr(beta), _ptr(scale_shift),
                                          ss_stride, flags, _ptr(out), _ptr(y_out), S, T, H, _stream())
    else:
        rc = lib.hig_attn_apply_stylize(_ptr(q), q.stride(0), _ptr(a_in), _ptr(gamma), _ptr(beta), _ptr(scale_shift),
                                        ss_stride, flags, _ptr(out), S, T, H, _stream())
    _lib.check(rc, "hig_attn_apply_stylize")
    return out


def attn_kv(k, v, a_out, S, T, H, length=None, pair_shift=0, transposed=False):
    """a_out[s,h] = softmax_time(K_masked)^T V (bf16 [S,H,64,64]); transposed: A^T for attn_apply_stylize_tc."""
    lib = _lib.load()
    if k.dtype != torch.bfloat16 or v.dtype != torch.bfloat16 or a_out.dtype != torch.bfloat16 or not a_out.is_contiguous():
        raise TypeError("hig_b200.attn_kv: bf16 storage, contiguous a_out")
    if v.stride(0) != k.stride(0):
        raise ValueError("hig_b200.attn_kv: K and V must share a leading dimension")
    if length is not None and length.dtype != torch.int32:
        raise TypeError("hig_b200.attn_kv: length must be int32")
    rc = lib.hig_attn_kv(_ptr(k), _ptr(v), k.stride(0), _ptr(a_out), _ptr(length), S, T, H, pair_shift,
                         1 if transposed else 0, _stream())
    _lib.check(rc, "hig_attn_kv")
    return a_out


def attn_apply_stylize_tc(q, a_t, gamma, beta, out, S, T, H, scale_shift=None, silu=True):
    """tcgen05 / TMEM variant of attn_apply_stylize: q already softmaxed, a_t = A^T [S,H,64,64] (attn_kv transposed)."""
    lib = _lib.load()
    if q.dtype != torch.bfloat16 or a_t.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        raise TypeError("hig_b200.attn_apply_stylize_tc: bf16 storage only")
    if not out.is_contiguous() or not a_t.is_contiguous():
        raise ValueError("hig_b200.attn_apply_stylize_tc: contiguous out / a_t required")
    ss_stride = 0
    if scale_shift is not None:
        if scale_shift.dtype != torch.float32 or scale_shift.stride(1) != 1:
            raise ValueError("hig_b200.attn_apply_stylize_tc: scale_shift must be fp32 with unit inner stride")
        ss_stride = scale_shift.stride(0)
    rc = lib.hig_attn_apply_stylize_tc(_ptr(q), q.stride(0), _ptr(a_t), _ptr(gamma), _ptr(beta), _ptr(scale_shift),
                                       ss_stride, 1 if silu else 0, _ptr(out), S, T, H, _stream())
    _lib.check(rc, "hig_attn_apply_stylize_tc")
    return out


def timestep_embed(t, freqs, out):
    lib = _lib.load()
    if t.dtype != torch.int64:
        raise TypeError("hig_b200.timestep_embed: t must be int64")
    S, half = t.shape[0], freqs.shape[0]
    rc = lib.hig_timestep_embed(_ptr(t), _ptr(freqs), S, half, _ptr(out), _dt(out), _stream())
    _lib.check(rc, "hig_timestep_embed")
    return out


def pack_motion(x, out):
    lib = _lib.load()
    S, T, C = x.shape
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("hig_b200.pack_motion: x must be contiguous fp32")
    rc = lib.hig_pack_motion(_ptr(x), S, T, C, out.stride(0), _ptr(out), _dt(out), _stream())
    _lib.check(rc, "hig_pack_motion")
    return out


def ddpm_step(x, eps, t, coef, noise=None, seed=0, packed=None, t_next=None, seed_dev=None):
    """In-place x <- posterior sample; eps [S*T, ld_eps] fp32 view, coef fp32 [5, n_steps].  seed_dev: int64 device
    scalar holding the Philox key (overrides `seed`; lets a captured graph be reused with a fresh seed)."""
    lib = _lib.load()
    S, T, C = x.shape
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("hig_b200.ddpm_step: x must be contiguous fp32")
    if noise is not None and (noise.dtype != torch.float32 or not noise.is_contiguous()):
        raise ValueError("hig_b200.ddpm_step: noise must be contiguous fp32")
    eps2 = eps.reshape(S * T, -1) if eps.dim() == 3 else eps
    if eps2.dtype not in (torch.float32, torch.float16) or eps2.stride(1) != 1:
        raise ValueError("hig_b200.ddpm_step: eps must be fp32 or fp16 with unit inner stride")
    rc = lib.hig_ddpm_step(_ptr(x), _ptr(eps2), eps2.stride(0), _dt(eps2), _ptr(noise), _ptr(t), _ptr(coef), coef.shape[1],
                           S, T, C, int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(seed_dev), _ptr(packed),
                           packed.stride(0) if packed is not None else 0,
                           _dt(packed) if packed is not None else F32, _ptr(t_next), _stream())
    _lib.check(rc, "hig_ddpm_step")
    return x


def time_table_silu(table, t, xf_proj, out):
    """out[s] = SiLU(table[t[s]] + xf_proj[s]); table fp32 [n_steps,E], t int64 [S], xf_proj fp32 [S,E], out bf16/fp32."""
    lib = _lib.load()
    S, E = xf_proj.shape
    if table.dtype != torch.float32 or xf_proj.dtype != torch.float32 or t.dtype != torch.int64:
        raise ValueError("hig_b200.time_table_silu: table / xf_proj fp32, t int64")
    if not (table.is_contiguous() and xf_proj.is_contiguous() and out.is_contiguous()) or table.shape[1] != E:
        raise ValueError("hig_b200.time_table_silu: contiguous [*, E] operands required")
    rc = lib.hig_time_table_silu(_ptr(table), table.shape[0], _ptr(t), _ptr(xf_proj), S, E, _ptr(out), _dt(out), _stream())
    _lib.check(rc, "hig_time_table_silu")
    return out


def tile_rows(table, period, out):
    """out[r] = fp16(table[r % period]); table fp32 [>= period, W], out fp16 [rows, W]."""
    lib = _lib.load()
    if table.dtype != torch.float32 or out.dtype != torch.float16 or not table.is_contiguous() or not out.is_contiguous() \
            or table.shape[1] != out.shape[1] or not 0 < period <= table.shape[0]:
        raise ValueError("hig_b200.tile_rows: table contiguous fp32 [>= period, W], out contiguous fp16 [rows, W]")
    rc = lib.hig_tile_rows(_ptr(table), period, table.shape[1], out.shape[0], _ptr(out), _stream())
    _lib.check(rc, "hig_tile_rows")
    return out


def recover_joints(x, mean=None, std=None, init_mean=None, init_std=None, length=None, joints_num=22, init_row=0,
                   out=None):
    """x fp32 [S,T,C] on CUDA -> joints fp32 [S,T-1,joints_num,3] (hig_recover_joints); init_row 0 or -1 / T-1."""
    lib = _lib.load()
    if x.dim() != 3 or x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("hig_b200.recover_joints: x must be contiguous fp32 [S, T, C]")
    S, T, C = x.shape
    if init_row < 0:
        init_row += T
    f = lambda v, n: None if v is None else torch.as_tensor(v, dtype=torch.float32).to(x.device).contiguous().reshape(n)
    mean, std, init_mean, init_std = f(mean, C), f(std, C), f(init_mean, 4), f(init_std, 4)
    if length is not None:
        length = torch.as_tensor(length).reshape(-1).to(device=x.device, dtype=torch.int32).contiguous()
        if length.numel() != S:
            raise ValueError(f"hig_b200.recover_joints: length must have {S} entries")
    if out is None:
        out = torch.empty(S, max(T - 1, 0), joints_num, 3, device=x.device, dtype=torch.float32)
    rc = lib.hig_recover_joints(_ptr(x), S, T, C, init_row, _ptr(mean), _ptr(std), _ptr(init_mean), _ptr(init_std),
                                _ptr(length), joints_num, _ptr(out), _stream())
    _lib.check(rc, "hig_recover_joints")
    return out


def q_sample(x0, noise, t, sqrt_ac, sqrt_1mac, out=None):
    lib = _lib.load()
    S = x0.shape[0]
    TC = x0.numel() // S
    if out is None:
        out = torch.empty_like(x0)
    rc = lib.hig_q_sample(_ptr(x0), _ptr(noise), _ptr(t), _ptr(sqrt_ac), _ptr(sqrt_1mac), S, TC, _ptr(out), _stream())
    _lib.check(rc, "hig_q_sample")
    return out


def set_sm_limit(n):
    """Persistent kernels size their grids for n SMs (0 = all): leaves SMs to concurrent NCCL kernels."""
    _lib.check(_lib.load().hig_set_sm_limit(int(n)), "hig_set_sm_limit")


def debug_saturation(counter):
    """counter: int64 / uint64 device scalar that counts saturating fp16 stream stores (None switches the counter off)."""
    lib = _lib.load()
    if counter is not None and (not counter.is_cuda or counter.element_size() != 8):
        raise ValueError("hig_b200.debug_saturation: an 8-byte CUDA scalar is required")
    _lib.check(lib.hig_debug_saturation(_ptr(counter)), "hig_debug_saturation")


def l2_persist(t, hit_ratio=1.0):
    """Pin tensor `t` in L2 for kernels launched/captured on the current stream (None clears the window)."""
    lib = _lib.load()
    if t is None:
        rc = lib.hig_l2_persist(None, 0, 0.0, _stream())
    else:
        rc = lib.hig_l2_persist(_ptr(t), t.numel() * t.element_size(), float(hit_ratio), _stream())
    _lib.check(rc, "hig_l2_persist")


# ---------------------------------------------------------------------------------------------- training path
def gemm_splitk(a, w, out_f32, k_splits=0):
    """out_f32[M,N] += a[M,K] @ w[N,K].T, bf16 operands, K split over CTAs and combined with fp32 atomics."""
    lib = _lib.load()
    M, K = a.shape
    N = w.shape[0]
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or out_f32.dtype != torch.float32:
        raise TypeError("hig_b200.gemm_splitk: bf16 operands and an fp32 accumulator are required")
    if w.shape[1] != K or tuple(out_f32.shape) != (M, N):
        raise ValueError("hig_b200.gemm_splitk: shape mismatch")
    rc = lib.hig_gemm_bf16_splitk(_ptr(a), _rowmajor(a, "A"), _ptr(w), _rowmajor(w, "W"), M, N, K, _ptr(out_f32),
                                  _rowmajor(out_f32, "out_f32"), int(k_splits), _stream())
    _lib.check(rc, "hig_gemm_bf16_splitk")
    return out_f32


def gemm_t(a, w, trans_a=False, trans_b=False, bias=None, residual=None, out_f32=None, out_bf16=None, split_k=0):
    """C[M,N] (+)= opA(a) @ opB(w).T on tcgen05 with MN-major operands read in place (no transposed copies).
    trans_a: a is [K, M]; trans_b: w is [K, N].  split_k != 0: fp32 atomic accumulation into out_f32."""
    lib = _lib.load()
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise TypeError("hig_b200.gemm_t: bf16 operands required")
    K, M = (a.shape if trans_a else a.shape[::-1])
    Kw, N = (w.shape if trans_b else w.shape[::-1])
    if K != Kw:
        raise ValueError(f"hig_b200.gemm_t: contraction mismatch {K} vs {Kw}")
    for t, nm in ((bias, "bias"), (residual, "residual"), (out_f32, "out_f32")):
        if t is not None and t.dtype != torch.float32:
            raise TypeError(f"hig_b200.gemm_t: {nm} must be fp32")
    rc = lib.hig_gemm_bf16_t(1 if trans_a else 0, 1 if trans_b else 0, _ptr(a), _rowmajor(a, "A"), _ptr(w),
                             _rowmajor(w, "W"), M, N, K, _ptr(bias), _ptr(residual),
                             _rowmajor(residual, "residual") if residual is not None else 0, _ptr(out_f32),
                             _rowmajor(out_f32, "out_f32") if out_f32 is not None else 0, _ptr(out_bf16),
                             _rowmajor(out_bf16, "out_bf16") if out_bf16 is not None else 0, int(split_k), _stream())
    _lib.check(rc, "hig_gemm_bf16_t")
    return out_f32 if out_bf16 is None else out_bf16


def gemm_fused(a, w, bias, out, trans_b=False, act=0, out_pre=None, gate=None, gate_act=0):
    """nn.Linear next to an activation in one tcgen05 GEMM (training path):
    forward  — pre = a @ w.T + bias; out_pre = pre (bf16); out = act(pre as stored)        (act: ACT_GELU / ACT_SILU / 0)
    backward — trans_b=True, w = the next layer's weight [K, N] as stored: out = (a @ w + bias) * act'(gate)."""
    lib = _lib.load()
    for t in (a, w, out, out_pre, gate):
        if t is not None and t.dtype != torch.bfloat16:
            raise TypeError("hig_b200.gemm_fused: bf16 operands / outputs required")
    M, K = a.shape
    Kw, N = (w.shape if trans_b else w.shape[::-1])
    if K != Kw or bias is None or bias.dtype != torch.float32 or bias.numel() < N:
        raise ValueError("hig_b200.gemm_fused: shape mismatch or missing fp32 bias")
    rc = lib.hig_gemm_bf16_fused(1 if trans_b else 0, _ptr(a), _rowmajor(a, "A"), _ptr(w), _rowmajor(w, "W"), M, N, K,
                                 _ptr(bias), int(act), _ptr(out), _rowmajor(out, "out"), _ptr(out_pre),
                                 _rowmajor(out_pre, "out_pre") if out_pre is not None else 0, _ptr(gate),
                                 _rowmajor(gate, "gate") if gate is not None else 0, int(gate_act), _stream())
    _lib.check(rc, "hig_gemm_bf16_fused")
    return out


def transpose(x, out_t=None, copy=None, colsum=None, rows_zero_mod=0):
    """x [M,N] -> out_t [N,>=M] (transposed), copy [M,N] (cast), colsum[N] += column sums; any subset."""
    lib = _lib.load()
    M, N = x.shape
    ref = out_t if out_t is not None else copy
    odt = _dt(ref) if ref is not None else _dt(x)
    if out_t is not None and copy is not None and out_t.dtype != copy.dtype:
        raise TypeError("hig_b200.transpose: out_t and copy must share a dtype")
    if colsum is not None and (colsum.dtype != torch.float32 or colsum.numel() < N):
        raise ValueError("hig_b200.transpose: colsum must be fp32 [N]")
    rc = lib.hig_transpose(_ptr(x), _dt(x), M, N, _rowmajor(x, "x"), _ptr(out_t),
                           _rowmajor(out_t, "out_t") if out_t is not None else 0, _ptr(copy),
                           _rowmajor(copy, "copy") if copy is not None else 0, odt, _ptr(colsum), int(rows_zero_mod),
                           _stream())
    _lib.check(rc, "hig_transpose")


def colsum(x, out):
    """out[N] += x[M,N].sum(0)  (fp32 atomics; zero `out` first for a plain sum)."""
    lib = _lib.load()
    M, N = x.shape
    if out.dtype != torch.float32 or out.numel() < N:
        raise ValueError("hig_b200.colsum: out must be fp32 [N]")
    rc = lib.hig_colsum(_ptr(x), _dt(x), M, N, _rowmajor(x, "x"), _ptr(out), _stream())
    _lib.check(rc, "hig_colsum")
    return out


def act_fwd(x, act, out):
    lib = _lib.load()
    if not x.is_contiguous() or not out.is_contiguous() or x.numel() != out.numel():
        raise ValueError("hig_b200.act_fwd: contiguous tensors of equal size required")
    rc = lib.hig_act_fwd(_ptr(x), _dt(x), x.numel(), act, _ptr(out), _dt(out), _stream())
    _lib.check(rc, "hig_act_fwd")
    return out


def act_bwd(x, dy, act, dx):
    lib = _lib.load()
    if not (x.is_contiguous() and dy.is_contiguous() and dx.is_contiguous()) or x.numel() != dy.numel():
        raise ValueError("hig_b200.act_bwd: contiguous tensors of equal size required")
    rc = lib.hig_act_bwd(_ptr(x), _dt(x), _ptr(dy), _dt(dy), x.numel(), act, _ptr(dx), _dt(dx), _stream())
    _lib.check(rc, "hig_act_bwd")
    return dx


def ln_film_silu_bwd(x, gamma, beta, dout, dx, rows_per_seq, scale_shift=None, silu=False, dx_accumulate=False,
                     d_ss=None, d_gb=None):
    """Backward of ln_film_silu.  d_ss: fp32 view [S, >=2W] (+=), d_gb: fp32 [S, 2W] per-sequence partials (+=)."""
    lib = _lib.load()
    rows, width = x.shape
    if not (x.is_contiguous() and dout.is_contiguous() and dx.is_contiguous()):
        raise ValueError("hig_b200.ln_film_silu_bwd: contiguous tensors required")
    ss_stride = scale_shift.stride(0) if scale_shift is not None else 0
    rc = lib.hig_ln_film_silu_bwd(_ptr(x), _dt(x), rows, width, rows_per_seq, _ptr(gamma), _ptr(beta),
                                  _ptr(scale_shift), ss_stride, 1 if silu else 0, _ptr(dout), _dt(dout), _ptr(dx),
                                  _dt(dx), 1 if dx_accumulate else 0, _ptr(d_ss),
                                  d_ss.stride(0) if d_ss is not None else 0, _ptr(d_gb),
                                  d_gb.stride(0) if d_gb is not None else 0, _stream())
    _lib.check(rc, "hig_ln_film_silu_bwd")
    return dx


def eff_attn_bwd(mode, S, T, H, q=None, k=None, v=None, a_in=None, dy=None, dq=None, dk=None, dv=None, dA=None,
                 length=None, pair_shift=0, q_sum=None, k_sum=None, v_sum=None):
    """q_sum / k_sum / v_sum (fp32 [H*64], optional): += column sums of dq / dk / dv — the bias gradients of the projections."""
    lib = _lib.load()
    for t in (q_sum, k_sum, v_sum):
        if t is not None and (t.dtype != torch.float32 or t.numel() < H * 64 or not t.is_contiguous()):
            raise ValueError("hig_b200.eff_attn_bwd: column-sum outputs must be contiguous fp32 [H*64]")
    ref = q if q is not None else k
    if dA is not None and dA.dtype != torch.float32:
        raise TypeError("hig_b200.eff_attn_bwd: dA must be fp32")
    if dk is not None and dv.stride(0) != dk.stride(0):
        raise ValueError("hig_b200.eff_attn_bwd: dK and dV must share a leading dimension")
    rc = lib.hig_eff_attn_bwd_sums(mode, _ptr(q), q.stride(0) if q is not None else 0, _ptr(k), _ptr(v),
                                   k.stride(0) if k is not None else 0, _ptr(a_in), _ptr(dy),
                                   dy.stride(0) if dy is not None else 0, _ptr(dq), dq.stride(0) if dq is not None else 0,
                                   _ptr(dk), _ptr(dv), dk.stride(0) if dk is not None else 0, _ptr(dA), _ptr(length), S, T,
                                   H, pair_shift, _dt(ref), _ptr(q_sum), _ptr(k_sum), _ptr(v_sum), _stream())
    _lib.check(rc, "hig_eff_attn_bwd_sums")


def mha_attention(qkv, out, B, N, H, causal=False):
    """Softmax MHA over rows [B*N, 3*H*64] laid out (q | k | v); out [B*N, H*64].  bf16 or fp32 storage."""
    lib = _lib.load()
    D = H * 64
    if qkv.shape != (B * N, 3 * D) or out.shape != (B * N, D) or qkv.dtype != out.dtype:
        raise ValueError("hig_b200.mha_attention: qkv [B*N, 3*H*64], out [B*N, H*64] of one dtype")
    rc = lib.hig_mha_attention(_ptr(qkv[:, :D]), _ptr(qkv[:, D:2 * D]), _ptr(qkv[:, 2 * D:]), _rowmajor(qkv, "qkv"), _ptr(out),
                               _rowmajor(out, "out"), B, N, H, 1 if causal else 0, _dt(qkv), _stream())
    _lib.check(rc, "hig_mha_attention")
    return out


def masked_mse(pred, target, length, pit=False, want_grad=True):
    """DDPMMulTrainer.backward_G's loss (labelled or PIT) and d loss / d pred.  Returns (loss device scalar, d_pred)."""
    lib = _lib.load()
    S, T, C = pred.shape
    if pred.dtype != torch.float32 or target.dtype != torch.float32 or not pred.is_contiguous() or not target.is_contiguous():
        raise ValueError("hig_b200.masked_mse: contiguous fp32 pred / target required")
    if length is not None and (length.dtype != torch.int32 or length.numel() != S):
        raise ValueError("hig_b200.masked_mse: length must be int32 [S]")
    scratch = torch.empty(2 * S + 1, device=pred.device, dtype=torch.float32)
    d_pred = torch.empty_like(pred) if want_grad else None
    rc = lib.hig_masked_mse(_ptr(pred), _ptr(target), _ptr(length), S, T, C, 1 if pit else 0, _ptr(scratch[:S]),
                            _ptr(scratch[S:2 * S]), _ptr(scratch[2 * S:]), _ptr(d_pred), _stream())
    _lib.check(rc, "hig_masked_mse")
    return scratch[2 * S], d_pred


def sumsq(x, out):
    """out (fp64 device scalar, caller-zeroed) += sum(x ** 2)."""
    lib = _lib.load()
    if x.dtype != torch.float32 or not x.is_contiguous() or out.dtype != torch.float64:
        raise ValueError("hig_b200.sumsq: contiguous fp32 x, fp64 out")
    rc = lib.hig_sumsq(_ptr(x), x.numel(), _ptr(out), _stream())
    _lib.check(rc, "hig_sumsq")
    return out


def mean_slices(own, staged, scale):
    """own[n] = (own + staged[count, n].sum(0)) * scale in one pass (fp32; the peer-memory gradient exchange's reduction)."""
    lib = _lib.load()
    if own.dtype != torch.float32 or staged.dtype != torch.float32 or staged.dim() != 2 or staged.shape[1] != own.numel():
        raise ValueError("hig_b200.mean_slices: own fp32 [n], staged fp32 [count, n]")
    if not own.is_contiguous() or staged.stride(1) != 1:
        raise ValueError("hig_b200.mean_slices: contiguous rows required")
    rc = lib.hig_mean_slices(_ptr(own), _ptr(staged), own.numel(), staged.stride(0), staged.shape[0], float(scale), _stream())
    _lib.check(rc, "hig_mean_slices")
    return own


def adam_flat(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8, p_bf16=None, gnorm2=None, max_norm=0.0):
    """One fused clip + Adam step over flat fp32 buffers (+ bf16 mirror of the new parameters)."""
    lib = _lib.load()
    n = p.numel()
    for t in (p, g, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n:
            raise ValueError("hig_b200.adam_flat: p, g, m, v must be contiguous fp32 of equal size")
    if p_bf16 is not None and (p_bf16.dtype != torch.bfloat16 or p_bf16.numel() != n or not p_bf16.is_contiguous()):
        raise ValueError("hig_b200.adam_flat: p_bf16 must be a contiguous bf16 mirror of p")
    rc = lib.hig_adam_flat(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p_bf16), n, float(lr), float(betas[0]), float(betas[1]),
                           float(eps), int(step), _ptr(gnorm2), float(max_norm), _stream())
    _lib.check(rc, "hig_adam_flat")
