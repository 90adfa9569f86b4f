"""TEST INFRASTRUCTURE ONLY: torch-CPU stand-ins for the `pcdms_b200.ops` entry points, used to exercise the HOST logic
(weight packing, topology walk, skip bookkeeping, conditioning layout, step tables) in the CPU-only test run.  They
follow the documented semantics of each `pcdm_*` C-ABI function (include/pcdm_b200.h).  The product never imports this
module: without the CUDA library `pcdms_b200.ops` raises.
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from oracle import blocks as OB


def _rt(x, dtype):
    return x.to(dtype)


def _chan_stats_of(y_rows, hw, order=None):
    """pcdm_ext.chan_stats of stored rows y [M, N]: [M / 32, N, 2] (sum, sum of squares) per 32-row slab, or None when
    the shape cannot carry them; `order` permutes the rows first (the slab order of pcdm_conv3x3_up2x)."""
    from pcdms_b200.ops import ChanStats
    M, N = y_rows.shape
    if hw % 32 or M % 32:
        return None
    yf = y_rows.float()
    if order is not None:
        yf = yf[order]
    yf = yf.view(M // 32, 32, N)
    return ChanStats(torch.stack([yf.sum(1), (yf ** 2).sum(1)], dim=-1).contiguous(), hw)


def gemm(a, w, out=None, *, a2=None, bias=None, rowvec=None, rows_per_image=1, residual=None, geglu=False,
         out_f32=False, silu=False, gelu=False, bn=0, w_static=True, cta_group=0, skinny=True, row_stats=False,
         ln=None, chan_stats=False):
    x = a.float() if a2 is None else torch.cat([a.float(), a2.float()], dim=1)
    y = x @ w.float().t()
    if ln is not None:   # pcdm_ext.ln_*: out = rstd * acc + bias (row-centred weight), rstd from the producer's sums
        K = x.shape[1]
        st = ln.stats.buf[: ln.stats.parts].sum(0)
        mean = st[:, 0] / K
        rstd = torch.rsqrt((st[:, 1] / K - mean * mean).clamp_min(0) + ln.eps)
        y = rstd[:, None] * y
    if geglu:
        y = y + (bias if bias is not None else 0)
        g = y.view(y.shape[0], -1, 64)
        y = (g[..., :32] * (F.silu(g[..., 32:]) if silu else F.gelu(g[..., 32:]))).reshape(y.shape[0], -1)
    else:
        if bias is not None:
            y = y + bias
        if rowvec is not None:
            y = y + rowvec.repeat_interleave(rows_per_image, 0)[: y.shape[0]]
        if residual is not None:
            y = y + residual.float()
        if silu:
            y = F.silu(y)
        if gelu:
            y = F.gelu(y)
    y = y if out_f32 else _rt(y, a.dtype)
    if out is not None:
        out.copy_(y)
        y = out
    if row_stats:
        from pcdms_b200.ops import RowStats
        yf = y.float()
        # two slots, as a real launch with one N tile would write (statistics of the values before the 16-bit store
        # differ from these by rounding noise far below the tolerance of any consumer)
        half = yf.shape[1] // 2
        buf = torch.stack([torch.stack([yf[:, :half].sum(1), (yf[:, :half] ** 2).sum(1)], dim=1),
                           torch.stack([yf[:, half:].sum(1), (yf[:, half:] ** 2).sum(1)], dim=1)])
        return y, RowStats(buf, 2)
    if chan_stats:
        return y, _chan_stats_of(y, rows_per_image)
    return y


def ln_gemm(x, gamma, beta, eps, w, out=None, *, bias=None, rowvec=None, rows_per_image=1, residual=None,
            out_f32=False, silu=False, gelu=False):
    n = layernorm(x, gamma, beta, eps)
    return gemm(n, w, out, bias=bias, rowvec=rowvec, rows_per_image=rows_per_image, residual=residual,
                out_f32=out_f32, silu=silu, gelu=gelu)


def conv3x3(x, w_packed, out=None, *, bias=None, rowvec=None, residual=None, stride=1, out_f32=False, silu=False,
            pad_br=False, bn=0, cta_group=0, chan_stats=False):
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    w = w_packed.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    xin = x.float().permute(0, 3, 1, 2)
    if pad_br:
        assert stride == 2
        y = F.conv2d(F.pad(xin, (0, 1, 0, 1)), w, bias, stride=2, padding=0).permute(0, 2, 3, 1)
    else:
        y = F.conv2d(xin, w, bias, stride=stride, padding=1).permute(0, 2, 3, 1)
    if rowvec is not None:
        y = y + rowvec[:, None, None, :]
    if residual is not None:
        y = y + residual.float()
    if silu:
        y = F.silu(y)
    y = y.contiguous()
    y = y if out_f32 else _rt(y, x.dtype)
    if chan_stats:
        return y, (None if out_f32 else _chan_stats_of(y.reshape(-1, Cout), y.shape[1] * y.shape[2]))
    return y


def conv3x3_up2x(x, w_up, out=None, *, bias=None, silu=False, bn=0, cta_group=0, chan_stats=False):
    """Four 2x2 convolutions over the low-resolution input, one per output parity (the documented semantics of
    pcdm_conv3x3_up2x): plane p = 2 py + px, tap (ty, tx) reads source pixel (i + ty - 1 + py, j + tx - 1 + px)."""
    B, H, W, Cin = x.shape
    Cout = w_up.shape[1]
    xin = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1))            # source rows / columns -1 .. H / W
    y = x.new_zeros((B, 2 * H, 2 * W, Cout), dtype=torch.float32)
    for py in (0, 1):
        for px in (0, 1):
            wk = w_up[2 * py + px].float().view(Cout, 2, 2, Cin).permute(0, 3, 1, 2)
            o = F.conv2d(xin[:, :, py:py + H + 1, px:px + W + 1], wk, bias)
            y[:, py::2, px::2] = o.permute(0, 2, 3, 1)
    if silu:
        y = F.silu(y)
    y = _rt(y.contiguous(), x.dtype)
    if chan_stats:
        if (H * W) % 32:
            return y, None
        # slab order of the kernel: [image][parity plane 2 py + px][32 low-resolution pixels]
        planes = torch.stack([y[:, py::2, px::2] for py in (0, 1) for px in (0, 1)], dim=1)   # [B, 4, H, W, Cout]
        return y, _chan_stats_of(planes.reshape(-1, Cout), 4 * H * W)
    return y


def groupnorm(x1, gamma, beta, eps, *, x2=None, groups=32, silu=False, out=None, workspace=None, path=None,
              stats=None):
    x = x1 if x2 is None else torch.cat([x1, x2], dim=-1)
    shp = x.shape
    B, Ct = shp[0], shp[-1]
    HW = x.numel() // (B * Ct)
    if stats is not None and stats[0] is not None and (x2 is None or stats[1] is not None) and HW % 32 == 0:
        # pcdm_groupnorm_apply: (mean, rstd) per (image, group) folded from the PRODUCERS' slab statistics — the mock
        # really uses them, so a host-side mix-up (stale or foreign statistics) shows up in the CPU tests
        srcs = [stats[0]] + ([stats[1]] if x2 is not None else [])
        assert all(s_.hw == HW and s_.buf.shape[0] == B * HW // 32 for s_ in srcs)
        per_img = torch.cat([s_.buf.view(B, HW // 32, -1, 2).double().sum(1) for s_ in srcs], dim=1)   # [B, Ct, 2]
        g = per_img.view(B, groups, Ct // groups, 2).sum(2)
        n = (Ct // groups) * HW
        mean = g[..., 0] / n
        rstd = torch.rsqrt((g[..., 1] / n - mean * mean).clamp_min(0) + eps)
        cpg = Ct // groups
        xf = x.float().reshape(B, HW, groups, cpg)
        y = (xf - mean.float()[:, None, :, None]) * rstd.float()[:, None, :, None]
        y = y.reshape(B, HW, Ct) * gamma + beta
        if silu:
            y = F.silu(y)
        return _rt(y.reshape(shp).contiguous(), x1.dtype)
    y = F.group_norm(x.float().reshape(shp[0], -1, shp[-1]).transpose(1, 2), groups, gamma, beta, eps)
    if silu:
        y = F.silu(y)
    return _rt(y.transpose(1, 2).reshape(shp).contiguous(), x1.dtype)


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    return _rt(F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps), x.dtype)


def attention(q, k, v, B, heads, out=None, scale=0.125, head_dim=64):
    Sq, Skv = q.shape[0] // B, k.shape[0] // B
    C = heads * head_dim
    qh = q[:, :C].float().reshape(B, Sq, heads, head_dim).transpose(1, 2)
    kh = k[:, :C].float().reshape(B, Skv, heads, head_dim).transpose(1, 2)
    vh = v[:, :C].float().reshape(B, Skv, heads, head_dim).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return _rt((p @ vh).transpose(1, 2).reshape(B * Sq, C), q.dtype)


def nchw_to_nhwc_pad(x, cpad, dtype, out=None):
    B, C, H, W = x.shape
    y = torch.zeros(B, H, W, cpad, dtype=dtype)
    y[..., :C] = x.permute(0, 2, 3, 1).to(dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def nhwc_to_nchw(x, channels, out_dtype, out=None):
    return x[..., :channels].permute(0, 3, 1, 2).contiguous().to(out_dtype)


def timestep_embedding(t, B, dim, dtype, out=None):
    tt = t.float().reshape(-1)
    return OB.get_timestep_embedding(tt.expand(B) if tt.numel() == 1 else tt, dim, flip_sin_to_cos=True,
                                     downscale_freq_shift=0).to(dtype)


def upsample_nearest2x(x, out=None):
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).contiguous()


def _cfg(e, n, guidance_scale, ratio, guidance_rescale):
    """e: [2n, C, ...] fp32 -> CFG combine (+ the reference's rescale with the given per-sample std ratio)."""
    cfg = e[:n] + guidance_scale * (e[n:] - e[:n])
    if ratio is not None:
        r = ratio.view(-1, *([1] * (cfg.dim() - 1)))
        cfg = guidance_rescale * (cfg * r) + (1 - guidance_rescale) * cfg
    return cfg


def cfg_rescale_ratio(eps, guidance_scale, out=None, *, nhwc_channels=None):
    e = eps.float()
    if nhwc_channels is not None:
        e = e[..., :nhwc_channels]
    n = e.shape[0] // 2
    cfg = e[:n] + guidance_scale * (e[n:] - e[:n])
    dims = list(range(1, e.dim()))
    ratio = e[n:].std(dim=dims) / cfg.std(dim=dims)
    if out is not None:
        out.copy_(ratio)
        return out
    return ratio


def cfg_combine(eps, guidance_scale, guidance_rescale=0.0, out=None):
    n = eps.shape[0] // 2
    ratio = cfg_rescale_ratio(eps, guidance_scale) if guidance_rescale > 0.0 else None
    return _cfg(eps.float(), n, guidance_scale, ratio, guidance_rescale).to(eps.dtype)


def cfg_ddim_step(eps_rows, latents, x9, coef_table, step_counter, guidance_scale, t_table=None, t_cur=None,
                  ratio=None, guidance_rescale=0.0):
    n = latents.shape[0]
    step = int(step_counter[0])
    c = coef_table[step]
    e = eps_rows[..., :4].float().permute(0, 3, 1, 2)
    e = _cfg(e, n, guidance_scale, ratio, guidance_rescale)
    x0 = (latents - c[1] * e) * c[0]
    latents.copy_(c[2] * x0 + c[3] * e)
    x9[..., :4] = torch.cat([latents, latents]).permute(0, 2, 3, 1).to(x9.dtype)
    step_counter[0] = step + 1
    if t_table is not None and t_cur is not None:
        t_cur[0] = t_table[step + 1]


def add_noise(x0, noise, alphas_cumprod, timesteps, out=None):
    a = alphas_cumprod[timesteps].view(-1, *([1] * (x0.dim() - 1)))
    return (a.sqrt() * x0.float() + (1 - a).sqrt() * noise.float()).to(x0.dtype)


def ddim_step(model_output, sample, coefs, out=None):
    x0 = (sample.float() - coefs[1] * model_output.float()) * coefs[0]
    return (coefs[2] * x0 + coefs[3] * model_output.float()).to(sample.dtype)


def _unipc_update(r, e, x, last, ma, mb):
    """include/pcdm_b200.h, pcdm_cfg_unipc_step: one UniPC step on fp32 tensors, the reference's operation order."""
    r = [torch.tensor(v, dtype=torch.float32) for v in r]
    mt = (x - r[0] * e) / r[1]
    xc = x
    if float(r[2]) != 0.0:
        xt_ = r[3] * last - r[4] * ma
        corr = r[8] * (mt - ma)
        if float(r[9]) >= 2:
            corr = r[7] * ((mb - ma) / r[6]) + corr
        xc = xt_ - r[5] * corr
    xn = r[10] * xc - r[11] * mt
    if float(r[15]) >= 2:
        xn = xn - r[12] * (r[14] * ((ma - mt) / r[13]))
    return xn, xc, mt, ma


def cfg_unipc_step(eps_rows, state, x9, coef_table, step_counter, guidance_scale, t_table=None, t_cur=None,
                   ratio=None, guidance_rescale=0.0):
    n = state.shape[1]
    step = int(step_counter[0])
    e = eps_rows[..., :4].float().permute(0, 3, 1, 2)
    e = _cfg(e, n, guidance_scale, ratio, guidance_rescale)
    xn, xc, mt, ma_old = _unipc_update(coef_table[step].tolist(), e, state[0], state[1], state[2], state[3])
    state[3].copy_(ma_old)
    state[2].copy_(mt)
    state[1].copy_(xc)
    state[0].copy_(xn)
    x9[..., :4] = torch.cat([state[0], state[0]]).permute(0, 2, 3, 1).to(x9.dtype)
    step_counter[0] = step + 1
    if t_table is not None and t_cur is not None:
        t_cur[0] = t_table[step + 1]


def unipc_step(model_output, sample, last_sample, m0, m1, coef_row, out=None):
    shp = sample.shape
    xn, xc, mt, ma_old = _unipc_update(coef_row, model_output.float().reshape(-1), sample.float().reshape(-1),
                                       last_sample, m0, m1)
    m1.copy_(ma_old)
    m0.copy_(mt)
    last_sample.copy_(xc)
    return xn.reshape(shp).to(sample.dtype)


def _unclip_update(r, mo, x, noise):
    """include/pcdm_b200.h, pcdm_unclip_step: fp32, the reference's operation order."""
    r = [torch.tensor(v, dtype=torch.float32) for v in r]
    x0 = mo
    if float(r[6]) != 0.0:
        x0 = (x - r[5] * mo) / r[4]
    x0 = torch.clamp(x0, -r[3], r[3])
    xp = r[0] * x0 + r[1] * x
    if float(r[2]) != 0.0:
        xp = xp + r[2] * noise
    return xp


def cfg_unclip_step(pred, latents, xin, coef_table, noise_table, step_counter, guidance_scale, use_cfg, t_table=None,
                    t_cur=None):
    n = latents.shape[0]
    step = int(step_counter[0])
    mo = pred[:n]
    if use_cfg:
        mo = mo + guidance_scale * (pred[n:] - mo)
    latents.copy_(_unclip_update(coef_table[step].tolist(), mo, latents, noise_table[step]))
    xin.copy_(torch.cat([latents, latents]) if use_cfg else latents)
    step_counter[0] = step + 1
    if t_table is not None and t_cur is not None:
        t_cur[0] = t_table[step + 1]


def unclip_step(model_output, sample, noise, coef_row, out=None):
    return _unclip_update(coef_row, model_output.float(), sample.float(),
                          None if noise is None else noise.float()).to(sample.dtype)


def softmax_rows(x, scale, dtype, out=None):
    y = torch.softmax(x.float() * scale, dim=-1).to(dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def gaussian_sample(moments, B, channels, HW, noise=None, scale=1.0, out=None):
    m = moments.view(B, HW, -1).permute(0, 2, 1)
    v = m[:, :channels]
    if noise is not None:
        v = v + torch.exp(0.5 * m[:, channels:2 * channels].clamp(-30.0, 20.0)) * noise.view(B, channels, HW)
    return (v * scale).contiguous()


def row_stats(x):
    from pcdms_b200.ops import RowStats
    xf = x.float()
    return RowStats(torch.stack([xf.sum(1), (xf ** 2).sum(1)], dim=1)[None], 1)


def ensure_workspace(device, nbytes=0):
    return None


def require_cuda(t, what):
    return None


_NAMES = ["ensure_workspace", "require_cuda", "row_stats", "gemm", "ln_gemm", "conv3x3", "conv3x3_up2x", "groupnorm", "layernorm", "attention", "nchw_to_nhwc_pad", "nhwc_to_nchw",
          "timestep_embedding", "upsample_nearest2x", "cfg_ddim_step", "cfg_rescale_ratio", "cfg_combine", "add_noise", "ddim_step", "cfg_unipc_step",
          "unipc_step", "softmax_rows", "gaussian_sample", "cfg_unclip_step", "unclip_step"]


@contextlib.contextmanager
def patched():
    """Temporarily route pcdms_b200.ops.<fn> to the CPU stand-ins above (tests only)."""
    from pcdms_b200 import ops
    saved = {n: getattr(ops, n) for n in _NAMES}
    try:
        for n in _NAMES:
            setattr(ops, n, globals()[n])
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
