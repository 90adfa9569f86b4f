"""Thin torch-tensor wrappers over the C ABI (one Python function per `pcdm_*` entry point).

Torch is plumbing here (device memory + streams); all arithmetic happens inside libpcdm_b200.so.  Activations are
NHWC: a tensor of logical shape [B, H, W, C] or [rows, C], contiguous in C.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import threading

import torch

from . import lib as _l


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float16:
        return _l.DT_F16
    if t.dtype == torch.bfloat16:
        return _l.DT_BF16
    raise TypeError(f"pcdm_b200 kernels take fp16 or bf16 activations, got {t.dtype}")


def _stream(t: torch.Tensor) -> C.c_void_p:
    if not t.is_cuda:
        raise RuntimeError("pcdm_b200 ops need CUDA tensors (there is no CPU path)")
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda(t: torch.Tensor, what: str) -> None:
    """The product has no CPU compute path: boundary classes call this before dispatching to the kernels."""
    if not t.is_cuda:
        raise RuntimeError(f"{what} runs on CUDA tensors only (no CPU fallback)")


def _f32(t):
    if t is not None and t.dtype != torch.float32:
        raise TypeError("bias / rowvec vectors must be fp32")
    return t


# ---------------------------------------------------------------------------------------------------------------
# split-K scratch.  The C ABI takes it per call (pcdm_ext.workspace): the library itself holds no state.  This layer
# keeps ONE default buffer per device — allocated once, never replaced or shrunk, because captured CUDA graphs freeze
# its address — and lets a caller that drives several streams concurrently hand each of them its own buffer.
# ---------------------------------------------------------------------------------------------------------------
WORKSPACE_BYTES = 64 << 20   # covers every BASELINE configuration (see pcdm_gemm_workspace_bytes)
_default_ws = {}             # device index -> uint8 tensor
_tls = threading.local()


def _dev_index(device) -> int:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("pcdm_b200 ops need a CUDA device (there is no CPU path)")
    return device.index if device.index is not None else torch.cuda.current_device()


def ensure_workspace(device, nbytes: int = WORKSPACE_BYTES):
    """The default split-K scratch of `device` (idempotent: 'cuda' and 'cuda:<current>' name the same buffer; an
    existing buffer is never freed, so graphs captured earlier stay valid)."""
    if torch.device(device).type != "cuda":
        return None
    idx = _dev_index(device)
    ws = _default_ws.get(idx)
    if ws is None:
        ws = torch.zeros(max(int(nbytes), WORKSPACE_BYTES), dtype=torch.uint8, device=torch.device("cuda", idx))
        _default_ws[idx] = ws
    return ws


@contextlib.contextmanager
def use_workspace(ws):
    """Route the split-K scratch of every op issued by this host thread inside the block to `ws` (a zero-initialised
    uint8 CUDA tensor): one buffer per stream that runs concurrently with another."""
    prev = getattr(_tls, "ws", None)
    _tls.ws = ws
    try:
        yield ws
    finally:
        _tls.ws = prev


def _ext(t: torch.Tensor, force_cta_group: int = 0):
    if not t.is_cuda:
        raise RuntimeError("pcdm_b200 ops need CUDA tensors (there is no CPU path)")
    ws = getattr(_tls, "ws", None)
    if ws is None or ws.device != t.device:
        ws = ensure_workspace(t.device)
    e = _l.Ext()
    e.size, e.force_cta_group, e.workspace, e.workspace_bytes = C.sizeof(_l.Ext), int(force_cta_group), ws.data_ptr(), ws.numel()
    return e


class RowStats:
    """Per-row (sum, sum of squares) partials a producing GEMM wrote for the LayerNorm that follows it
    (pcdm_ext.row_stats): `buf` [cap, M, 2] fp32, of which the first `parts` slots are live."""
    __slots__ = ("buf", "parts")

    def __init__(self, buf, parts):
        self.buf, self.parts = buf, parts


class ChanStats:
    """Per 32-row slab and per channel (sum, sum of squares) a producing conv / GEMM wrote for the GroupNorm that
    follows it (pcdm_ext.chan_stats): `buf` [rows / 32, N, 2] fp32; `hw` = rows per image."""
    __slots__ = ("buf", "hw")

    def __init__(self, buf, hw):
        self.buf, self.hw = buf, hw


def _chan_stats_buf(t, M, N, hw):
    """The buffer a launch with pcdm_ext.chan_stats fills, or None when the shape cannot carry slab statistics (the
    rows of a 32-row slab must belong to one image) — the consumer then runs the stand-alone GroupNorm."""
    if hw % 32 or M % 32:
        return None
    return ChanStats(torch.empty((M // 32, N, 2), device=t.device, dtype=torch.float32), hw)


class FoldedLN:
    """A LayerNorm folded into the GEMM that consumes it: the producer's RowStats and eps.  The GEMM's weight must be
    `fold_layernorm_weight(...)` (gamma-scaled, rows centred) and its `bias` must carry W . beta."""
    __slots__ = ("stats", "eps")

    def __init__(self, stats: RowStats, eps=1e-5):
        self.stats, self.eps = stats, eps


def fold_layernorm_weight(W: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias, dtype):
    """Weight / bias of Linear(LayerNorm(x; gamma, beta)) for the folded form `rstd[m] * (x[m] . W'^T) + bias'`:
    W'[n, k] = W[n, k] gamma[k] - mean_k(W[n, :] gamma) — with centred rows the row mean of x cancels inside the
    accumulation, only rstd is left for the epilogue — and bias'[n] = W[n, :] . beta (+ bias[n]).  fp64 arithmetic,
    W' rounded once to `dtype`."""
    Wg = W.double() * gamma.double()[None, :]
    Wc = (Wg - Wg.mean(dim=1, keepdim=True)).to(dtype)
    cb = W.double() @ beta.double()
    if bias is not None:
        cb = cb + bias.double()
    return Wc.contiguous(), cb.float().contiguous()


# ---------------------------------------------------------------------------------------------------------------
# weight packing (host side, done once at load time)
# ---------------------------------------------------------------------------------------------------------------
def pack_conv3x3_weight(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """[Cout, Cin, 3, 3] (torch Conv2d layout) -> [Cout, 3, 3, Cin] flattened to [Cout, 9*Cin] (tap-major K)."""
    assert w.dim() == 4 and w.shape[2:] == (3, 3)
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(dtype).contiguous()


def pack_upsample_conv_weight(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """[Cout, Cin, 3, 3] of the conv that follows a nearest-2x upsample -> [4, Cout, 4*Cin]: per output parity
    p = 2*py + px a 2x2 filter over the LOW-resolution input whose tap (ty, tx) is the sum of the 3x3 taps that read
    the same source pixel (py = 0: rows {0} | {1, 2}; py = 1: rows {0, 1} | {2}; columns alike).  Summed in fp64,
    rounded once."""
    assert w.dim() == 4 and w.shape[2:] == (3, 3)
    w = w.double()
    collapse = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    planes = []
    for py in (0, 1):
        for px in (0, 1):
            taps = [w[:, :, collapse[py][ty]][:, :, :, collapse[px][tx]].sum(dim=(2, 3))
                    for ty in (0, 1) for tx in (0, 1)]
            planes.append(torch.stack(taps, dim=1).reshape(w.shape[0], -1))
    return torch.stack(planes).to(dtype).contiguous()


def geglu_row_permutation(n_out: int) -> torch.Tensor:
    """Row order for a GEGLU projection weight [2*n_out, K]: groups of [32 value rows | 32 gate rows]."""
    assert n_out % 32 == 0
    idx = torch.arange(n_out).view(-1, 32)
    return torch.cat([idx, idx + n_out], dim=1).reshape(-1)


# ---------------------------------------------------------------------------------------------------------------
# K1/K2
# ---------------------------------------------------------------------------------------------------------------
def gemm(a, w, out=None, *, a2=None, bias=None, rowvec=None, rows_per_image=1, residual=None, geglu=False,
         out_f32=False, silu=False, gelu=False, bn=0, w_static=True, cta_group=0, skinny=True, row_stats=False,
         ln: "FoldedLN | None" = None, chan_stats=False):
    """out[M, N] = [a | a2][M, K] @ w[N, K]^T (+bias) (+rowvec[m // rows_per_image]) (+residual).
    row_stats=True: also returns the RowStats of the output rows (-> `(out, stats)`), the LayerNorm statistics of the
    next op for free.  ln=FoldedLN(...): `a` holds RAW rows, `w` is gamma-scaled, the LayerNorm happens in the epilogue.
    chan_stats=True (with rows_per_image = H*W): also returns the ChanStats of the output (-> `(out, stats)`, stats None
    when the shape cannot carry them), the GroupNorm statistics of the next op for free.
    w_static: `w` holds model weights (not written by the kernel launched just before on this stream); pass False when
    `w` is an activation (the VAE's QK^T / PV products written as GEMMs)."""
    lib = _l.load()
    _dt(a)   # dtype errors first (TypeError), then device errors
    M, k1 = a.shape
    N, K = w.shape
    k2 = 0
    if a2 is not None:
        assert a2.shape[0] == M
        k2 = a2.shape[1]
    assert k1 + k2 == K, (k1, k2, K)
    assert a.stride(1) == 1 and w.is_contiguous()
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_f32 else a.dtype)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    flags = ((_l.FLAG_GEGLU if geglu else 0) | (_l.FLAG_OUT_F32 if out_f32 else 0) | (_l.FLAG_SILU if silu else 0) |
             (_l.FLAG_GELU if gelu else 0) | (_l.FLAG_W_STATIC if w_static else 0) |
             (0 if skinny else _l.FLAG_NO_SKINNY))
    ext = _ext(a, cta_group)
    stats = None
    if row_stats:
        cap = 2 * ((N + 63) // 64)
        stats = RowStats(torch.empty((cap, M, 2), device=a.device, dtype=torch.float32), 0)
        ext.row_stats, ext.row_stats_cap = stats.buf.data_ptr(), cap
    if ln is not None:
        assert ln.stats.buf.shape[1] == M and bias is not None
        ext.ln_stats, ext.ln_parts, ext.ln_eps = ln.stats.buf.data_ptr(), ln.stats.parts, float(ln.eps)
    cstats = None
    if chan_stats:
        assert not row_stats
        cstats = _chan_stats_buf(a, M, N, rows_per_image)
        if cstats is not None:
            ext.chan_stats = cstats.buf.data_ptr()
    rc = lib.pcdm_gemm(
        _l.ptr(a), C.c_longlong(a.stride(0)), _l.ptr(a2), C.c_longlong(a2.stride(0) if a2 is not None else 0),
        C.c_int(k1), _l.ptr(w), _l.ptr(out), C.c_longlong(out.stride(0)), _l.ptr(_f32(bias)), _l.ptr(_f32(rowvec)),
        C.c_longlong(rowvec.stride(0) if rowvec is not None else 0), C.c_int(rows_per_image), _l.ptr(residual), C.c_longlong(residual.stride(0) if residual is not None else 0),
        C.c_int(M), C.c_int(N), C.c_int(K), C.c_int(_dt(a)), C.c_int(flags), C.c_int(bn), C.byref(ext), _stream(a))
    _l.check(rc)
    if row_stats:
        stats.parts = int(ext.row_stats_parts)
        return out, stats
    if chan_stats:
        return out, cstats
    return out


def ln_gemm(x, gamma, beta, eps, w, out=None, *, bias=None, rowvec=None, rows_per_image=1, residual=None,
            out_f32=False, silu=False, gelu=False):
    """out[M, N] = act(LayerNorm(x[M, K]; gamma, beta, eps) @ w[N, K]^T (+bias) (+rowvec) (+residual)) — one launch for
    M <= 32 rows (the skinny kernel normalises the rows itself), otherwise pcdm_layernorm + pcdm_gemm inside the call."""
    lib = _l.load()
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K and x.stride(1) == 1 and w.is_contiguous()
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    scratch = None if M <= 32 and K <= 2048 and M * (K + 8) * 2 <= 100 * 1024 else torch.empty((M, K), device=x.device, dtype=x.dtype)
    flags = ((_l.FLAG_OUT_F32 if out_f32 else 0) | (_l.FLAG_SILU if silu else 0) | (_l.FLAG_GELU if gelu else 0) |
             _l.FLAG_W_STATIC)
    ext = _ext(x)
    rc = lib.pcdm_ln_gemm(
        _l.ptr(x), C.c_longlong(x.stride(0)), _l.ptr(_f32(gamma)), _l.ptr(_f32(beta)), C.c_float(eps), _l.ptr(scratch),
        _l.ptr(w), _l.ptr(out), C.c_longlong(out.stride(0)), _l.ptr(_f32(bias)), _l.ptr(_f32(rowvec)),
        C.c_longlong(rowvec.stride(0) if rowvec is not None else 0), C.c_int(rows_per_image), _l.ptr(residual),
        C.c_longlong(residual.stride(0) if residual is not None else 0), C.c_int(M), C.c_int(N), C.c_int(K),
        C.c_int(_dt(x)), C.c_int(flags), C.byref(ext), _stream(x))
    _l.check(rc, kernels=1 if scratch is None else 2)
    return out


def conv_tile_geometry(H, W):
    """The output sizes the implicit-GEMM tile map covers directly (an M-tile is 128 consecutive NHWC pixels = whole rows
    of a width that divides 128, or a 128-pixel segment of a row whose width is a multiple of 128), and the smallest
    such (Hp, Wp) >= (H, W) for everything else.  -> (ok, Hp, Wp)"""
    if W > 128:
        Wp = (W + 127) // 128 * 128
        return Wp == W, H, Wp
    Wp = 1 << max(W - 1, 0).bit_length()            # next power of two
    rows = 128 // Wp                                 # rows per tile
    if H * Wp >= 128 or H > rows:
        Hp = (H + rows - 1) // rows * rows
    else:
        Hp = 1 << max(H - 1, 0).bit_length()        # H * W < 128: the image must divide a tile
    return (Wp == W and Hp == H), Hp, Wp


def _pad_hw(t, Hp, Wp):
    """Zero-pad an NHWC tensor at the bottom / right (layout plumbing for the odd-canvas fallback below)."""
    B, H, W, C_ = t.shape
    o = torch.zeros((B, Hp, Wp, C_), device=t.device, dtype=t.dtype)
    o[:, :H, :W] = t
    return o


def conv3x3(x, w_packed, out=None, *, bias=None, rowvec=None, residual=None, stride=1, out_f32=False, silu=False,
            pad_br=False, bn=0, cta_group=0, chan_stats=False):
    """x: [B, Hin, Win, Cin] NHWC; w_packed: [Cout, 9*Cin]; returns [B, Hin/stride, Win/stride, Cout].
    pad_br (stride 2 only): zero padding on the bottom/right instead of all round (the VAE encoder's downsampler).
    chan_stats=True: -> `(out, ChanStats | None)`, the GroupNorm statistics of the output from the epilogue.

    Output sizes outside the tile map (a latent width that neither divides 128 nor is a multiple of it — no BASELINE
    configuration, but the reference accepts any canvas divisible by 8, stage2_batchtest_inpaint_model.py:258-260) run
    on the next covered size: zeros appended at the bottom / right of the input are exactly the convolution's own zero
    padding, so the outputs inside the real image are unchanged and the rest is cut off again.  Exact; costs the
    padded work plus two copies; no epilogue statistics (the consumer GroupNorm then takes its own)."""
    lib = _l.load()
    B, Hin, Win, Cin = x.shape
    Cout = w_packed.shape[0]
    assert w_packed.shape[1] == 9 * Cin and x.is_contiguous() and w_packed.is_contiguous()
    H, W = Hin // stride, Win // stride
    ok, Hp, Wp = conv_tile_geometry(H, W)
    if not ok:
        assert Hin == H * stride and Win == W * stride
        res_p = _pad_hw(residual, Hp, Wp) if residual is not None else None
        full = conv3x3(_pad_hw(x, Hp * stride, Wp * stride), w_packed, bias=bias, rowvec=rowvec, residual=res_p,
                       stride=stride, out_f32=out_f32, silu=silu, pad_br=pad_br, bn=bn, cta_group=cta_group)
        cut = full[:, :H, :W]
        if out is None:
            out = cut.contiguous()
        else:
            out.copy_(cut)
        return (out, None) if chan_stats else out
    if out is None:
        out = torch.empty((B, H, W, Cout), device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    assert out.is_contiguous() and tuple(out.shape) == (B, H, W, Cout)
    if residual is not None:
        assert residual.is_contiguous() and residual.shape == out.shape
    flags = (_l.FLAG_OUT_F32 if out_f32 else 0) | (_l.FLAG_SILU if silu else 0) | (_l.FLAG_PAD_BR if pad_br else 0)
    ext = _ext(x, cta_group)
    cstats = None
    if chan_stats and not out_f32:
        cstats = _chan_stats_buf(x, B * H * W, Cout, H * W)
        if cstats is not None:
            ext.chan_stats = cstats.buf.data_ptr()
    rc = lib.pcdm_conv3x3(_l.ptr(x), _l.ptr(w_packed), _l.ptr(out), _l.ptr(_f32(bias)), _l.ptr(_f32(rowvec)),
                          C.c_longlong(rowvec.stride(0) if rowvec is not None else 0), _l.ptr(residual), C.c_int(B), C.c_int(H), C.c_int(W), C.c_int(Cin), C.c_int(Cout),
                          C.c_int(stride), C.c_int(_dt(x)), C.c_int(flags), C.c_int(bn), C.byref(ext), _stream(x))
    _l.check(rc)
    return (out, cstats) if chan_stats else out


def conv3x3_up2x(x, w_up, out=None, *, bias=None, silu=False, bn=0, cta_group=0, chan_stats=False):
    """Upsample2D: nearest-2x + conv3x3 in one launch.  x: [B, H, W, Cin]; w_up: pack_upsample_conv_weight(...)
    [4, Cout, 4*Cin]; returns [B, 2H, 2W, Cout]."""
    lib = _l.load()
    B, H, W, Cin = x.shape
    Cout = w_up.shape[1]
    assert w_up.shape == (4, Cout, 4 * Cin) and x.is_contiguous() and w_up.is_contiguous()
    ok, Hp, Wp = conv_tile_geometry(H, W)
    if not ok or W > 128 or (W < 32 and (32 % W or (H * W) % 32) and H * W >= 32):   # odd canvases: see conv3x3
        if W > 128:
            raise NotImplementedError("conv3x3_up2x: input rows wider than 128 pixels (no caller has them)")
        if ok:                                       # tile map fine, the 32-pixel store boxes are not: pad to 32 | H*W
            Hp, Wp = (H + (32 // W) - 1) // (32 // W) * (32 // W), W
        cut = conv3x3_up2x(_pad_hw(x, Hp, Wp), w_up, bias=bias, silu=silu, bn=bn, cta_group=cta_group)[:, :2 * H, :2 * W]
        if out is None:
            out = cut.contiguous()
        else:
            out.copy_(cut)
        return (out, None) if chan_stats else out
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, Cout), device=x.device, dtype=x.dtype)
    assert out.is_contiguous() and tuple(out.shape) == (B, 2 * H, 2 * W, Cout)
    ext = _ext(x, cta_group)
    cstats = None
    if chan_stats:   # slabs: [image][parity plane][32 low-resolution pixels] — contiguous per image, as the consumer needs
        if (H * W) % 32 == 0:
            cstats = ChanStats(torch.empty((4 * B * H * W // 32, Cout, 2), device=x.device, dtype=torch.float32), 4 * H * W)
            ext.chan_stats = cstats.buf.data_ptr()
    rc = lib.pcdm_conv3x3_up2x(_l.ptr(x), _l.ptr(w_up), _l.ptr(out), _l.ptr(_f32(bias)), C.c_int(B), C.c_int(H),
                               C.c_int(W), C.c_int(Cin), C.c_int(Cout), C.c_int(_dt(x)),
                               C.c_int(_l.FLAG_SILU if silu else 0), C.c_int(bn), C.byref(ext), _stream(x))
    _l.check(rc)
    return (out, cstats) if chan_stats else out


# ---------------------------------------------------------------------------------------------------------------
# K4 / K5 norms
# ---------------------------------------------------------------------------------------------------------------
_gn_ws = {}


def _gn_workspace(device, B, groups):
    key = (device, B, groups)
    ws = _gn_ws.get(key)
    if ws is None:
        lib = _l.load()
        lib.pcdm_groupnorm_workspace_bytes.restype = C.c_longlong
        n = lib.pcdm_groupnorm_workspace_bytes(C.c_int(B), C.c_int(groups))
        ws = torch.zeros(int(n), dtype=torch.uint8, device=device)  # zero ONCE: the kernels keep the counters at 0
        _gn_ws[key] = ws
    return ws


def groupnorm(x1, gamma, beta, eps, *, x2=None, groups=32, silu=False, out=None, workspace=None, path=None,
              stats=None):
    """x1: [B, ..., C1] NHWC (x2 optional second channel segment); returns [B, ..., C1+C2].
    path: None (automatic), "two_pass" or "one_pass" (tests).
    stats=(ChanStats of x1, ChanStats of x2 | None): the statistics came out of the producers' epilogues — no pass over
    the activation for them (pcdm_groupnorm_apply); ignored (stand-alone kernels) when either is None."""
    lib = _l.load()
    lib.pcdm_groupnorm_workspace_bytes.restype = C.c_longlong
    B = x1.shape[0]
    C1 = x1.shape[-1]
    Ct = C1 + (x2.shape[-1] if x2 is not None else 0)
    HW = x1.numel() // (B * C1)
    assert x1.is_contiguous() and (x2 is None or x2.is_contiguous())
    if out is None:
        out = torch.empty((*x1.shape[:-1], Ct), device=x1.device, dtype=x1.dtype)
    if workspace is None:
        workspace = _gn_workspace(x1.device, B, groups)
    if stats is not None and stats[0] is not None and (x2 is None or stats[1] is not None) and HW % 32 == 0:
        s1, s2 = stats[0], (stats[1] if x2 is not None else None)
        assert s1.hw == HW and s1.buf.shape == (B * HW // 32, C1, 2)
        assert s2 is None or (s2.hw == HW and s2.buf.shape == (B * HW // 32, Ct - C1, 2))
        rc = lib.pcdm_groupnorm_apply(_l.ptr(x1), _l.ptr(s1.buf), _l.ptr(x2), _l.ptr(s2.buf if s2 is not None else None),
                                      C.c_int(C1), _l.ptr(out), _l.ptr(_f32(gamma)), _l.ptr(_f32(beta)), C.c_float(eps),
                                      C.c_int(B), C.c_int(HW), C.c_int(Ct), C.c_int(groups), C.c_int(_dt(x1)),
                                      C.c_int(_l.FLAG_SILU if silu else 0), _l.ptr(workspace), _stream(x1))
        _l.check(rc, kernels=2)
        return out
    rc = lib.pcdm_groupnorm(_l.ptr(x1), _l.ptr(x2), C.c_int(C1), _l.ptr(out), _l.ptr(_f32(gamma)), _l.ptr(_f32(beta)),
                            C.c_float(eps), C.c_int(B), C.c_int(HW), C.c_int(Ct), C.c_int(groups), C.c_int(_dt(x1)),
                            C.c_int((_l.FLAG_SILU if silu else 0) | {None: 0, "two_pass": _l.FLAG_GN_TWO_PASS,
                                                                     "one_pass": _l.FLAG_GN_ONE_PASS}[path]),
                            _l.ptr(workspace), _stream(x1))
    _l.check(rc, kernels=2)
    return out


def row_stats(x) -> RowStats:
    """RowStats of x [M, C] computed by a stand-alone kernel (rows that did not come out of a gemm(row_stats=True))."""
    lib = _l.load()
    M, Cc = x.shape
    assert x.stride(1) == 1
    buf = torch.empty((1, M, 2), device=x.device, dtype=torch.float32)
    rc = lib.pcdm_row_stats(_l.ptr(x), C.c_longlong(x.stride(0)), _l.ptr(buf), C.c_int(M), C.c_int(Cc), C.c_int(_dt(x)),
                            _stream(x))
    _l.check(rc)
    return RowStats(buf, 1)


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    """x: [M, C] rows (row stride free, unit column stride)."""
    lib = _l.load()
    M, Cc = x.shape
    assert x.stride(1) == 1
    if out is None:
        out = torch.empty((M, Cc), device=x.device, dtype=x.dtype)
    rc = lib.pcdm_layernorm(_l.ptr(x), C.c_longlong(x.stride(0)), _l.ptr(out), C.c_longlong(out.stride(0)),
                            _l.ptr(_f32(gamma)), _l.ptr(_f32(beta)), C.c_float(eps), C.c_int(M), C.c_int(Cc),
                            C.c_int(_dt(x)), _stream(x))
    _l.check(rc)
    return out


# ---------------------------------------------------------------------------------------------------------------
# K3 attention
# ---------------------------------------------------------------------------------------------------------------
def attention(q, k, v, B, heads, out=None, scale=0.125, head_dim=64):
    """q: [B*Sq, >=heads*head_dim] view, k/v: [B*Skv, ...] views (unit column stride; may be slices of a fused buffer).
    Head h of token row r lives at columns [h*head_dim, (h+1)*head_dim); head_dim is 64 or 128.
    Returns [B*Sq, heads*head_dim]."""
    lib = _l.load()
    assert q.stride(1) == 1 and k.stride(1) == 1 and v.stride(1) == 1
    Sq = q.shape[0] // B
    Skv = k.shape[0] // B
    if out is None:
        out = torch.empty((B * Sq, heads * head_dim), device=q.device, dtype=q.dtype)
    rc = lib.pcdm_attention_hd(_l.ptr(q), C.c_longlong(q.stride(0)), _l.ptr(k), C.c_longlong(k.stride(0)), _l.ptr(v),
                               C.c_longlong(v.stride(0)), _l.ptr(out), C.c_longlong(out.stride(0)), C.c_int(B),
                               C.c_int(heads), C.c_int(Sq), C.c_int(Skv), C.c_int(head_dim), C.c_float(scale),
                               C.c_int(_dt(q)), _stream(q))
    _l.check(rc)
    return out


# ---------------------------------------------------------------------------------------------------------------
# K6 and boundary helpers
# ---------------------------------------------------------------------------------------------------------------
def _any_dt(t: torch.Tensor) -> int:
    return {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}[t.dtype]


def nchw_to_nhwc_pad(x, cpad, dtype, out=None):
    lib = _l.load()
    B, Cc, H, W = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty((B, H, W, cpad), device=x.device, dtype=dtype)
    rc = lib.pcdm_nchw_to_nhwc_pad(_l.ptr(x), C.c_int(_any_dt(x)), _l.ptr(out), C.c_int(_any_dt(out)), C.c_int(B),
                                   C.c_int(Cc), C.c_int(H * W), C.c_int(cpad), _stream(x))
    _l.check(rc)
    return out


def nhwc_to_nchw(x, channels, out_dtype, out=None):
    """x: [B, H, W, ld] -> [B, channels, H, W]"""
    lib = _l.load()
    B, H, W, ld = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty((B, channels, H, W), device=x.device, dtype=out_dtype)
    rc = lib.pcdm_nhwc_to_nchw(_l.ptr(x), C.c_int(_any_dt(x)), C.c_longlong(ld), _l.ptr(out), C.c_int(_any_dt(out)),
                               C.c_int(B), C.c_int(channels), C.c_int(H * W), _stream(x))
    _l.check(rc)
    return out


def timestep_embedding(t, B, dim, dtype, out=None):
    """t: fp32 device tensor with 1 or B entries."""
    lib = _l.load()
    assert t.dtype == torch.float32 and t.is_cuda
    if out is None:
        out = torch.empty((B, dim), device=t.device, dtype=dtype)
    rc = lib.pcdm_timestep_embedding(_l.ptr(t), C.c_int(t.numel()), _l.ptr(out), C.c_int(_any_dt(out)), C.c_int(B),
                                     C.c_int(dim), _stream(t))
    _l.check(rc)
    return out


def upsample_nearest2x(x, out=None):
    lib = _l.load()
    B, H, W, Cc = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, Cc), device=x.device, dtype=x.dtype)
    rc = lib.pcdm_upsample_nearest2x(_l.ptr(x), _l.ptr(out), C.c_int(B), C.c_int(H), C.c_int(W), C.c_int(Cc),
                                     _stream(x))
    _l.check(rc)
    return out


def cfg_rescale_ratio(eps, guidance_scale, out=None, *, nhwc_channels=None):
    """Per-sample std(eps_cond) / std(cfg) of the reference's rescale_noise_cfg.  eps: the 2n-sample epsilon batch, either
    NCHW [2n, C, H, W] contiguous or (nhwc_channels=C) NHWC rows [2n, H, W, ld] of which the first C channels count.
    Returns [n] fp32."""
    lib = _l.load()
    assert eps.is_contiguous() and eps.shape[0] % 2 == 0
    n = eps.shape[0] // 2
    if nhwc_channels is None:
        Cc, HW = eps.shape[1], eps[0, 0].numel()
        sb, sc, sp = Cc * HW, HW, 1
    else:
        Cc, ld = nhwc_channels, eps.shape[-1]
        HW = eps[0].numel() // ld
        sb, sc, sp = HW * ld, 1, ld
    if out is None:
        out = torch.empty(n, device=eps.device, dtype=torch.float32)
    rc = lib.pcdm_cfg_rescale_ratio(_l.ptr(eps), C.c_int(_any_dt(eps)), C.c_longlong(sb), C.c_longlong(sc),
                                    C.c_longlong(sp), C.c_int(n), C.c_int(Cc), C.c_int(HW), C.c_float(guidance_scale),
                                    _l.ptr(out), _stream(eps))
    _l.check(rc)
    return out


def cfg_combine(eps, guidance_scale, guidance_rescale=0.0, out=None):
    """eps: [2n, ...] contiguous (rows [0, n) unconditional) -> [n, ...]: e_u + g (e_c - e_u), then the reference's
    rescale_noise_cfg when guidance_rescale > 0 (stage2_inpaint_pipeline.py:510-516)."""
    lib = _l.load()
    assert eps.is_contiguous() and eps.shape[0] % 2 == 0
    n = eps.shape[0] // 2
    if out is None:
        out = torch.empty((n, *eps.shape[1:]), device=eps.device, dtype=eps.dtype)
    ratio = cfg_rescale_ratio(eps.view(2 * n, eps.shape[1], -1) if eps.dim() > 2 else eps.view(2 * n, 1, -1),
                              guidance_scale) if guidance_rescale > 0.0 else None
    rc = lib.pcdm_cfg_combine(_l.ptr(eps), C.c_int(_any_dt(eps)), _l.ptr(out), C.c_int(_any_dt(out)), C.c_int(n),
                              C.c_longlong(eps[0].numel()), C.c_float(guidance_scale), _l.ptr(ratio),
                              C.c_float(guidance_rescale), _stream(eps))
    _l.check(rc)
    return out


def cfg_ddim_step(eps_rows, latents, x9, coef_table, step_counter, guidance_scale, t_table=None, t_cur=None,
                  ratio=None, guidance_rescale=0.0):
    """eps_rows: [2n, H, W, ld] UNet output rows; latents: [n, 4, H, W] fp32 (in place); x9: [2n, H, W, ld9].
    ratio: [n] fp32 from cfg_rescale_ratio (guidance_rescale > 0) or None."""
    lib = _l.load()
    n = latents.shape[0]
    HW = latents.shape[2] * latents.shape[3]
    assert latents.dtype == torch.float32 and latents.is_contiguous() and eps_rows.is_contiguous()
    assert coef_table.dtype == torch.float32 and step_counter.dtype == torch.int32 and step_counter.numel() == 2
    rc = lib.pcdm_cfg_ddim_step(_l.ptr(eps_rows), C.c_int(_any_dt(eps_rows)), C.c_longlong(eps_rows.shape[-1]),
                                _l.ptr(latents), _l.ptr(x9), C.c_int(_any_dt(x9)), C.c_longlong(x9.shape[-1]),
                                _l.ptr(coef_table), _l.ptr(step_counter), C.c_float(guidance_scale), C.c_int(n),
                                C.c_int(HW), _l.ptr(t_table), _l.ptr(t_cur), _l.ptr(ratio), C.c_float(guidance_rescale),
                                _stream(latents))
    _l.check(rc)


def add_noise(x0, noise, alphas_cumprod, timesteps, out=None):
    lib = _l.load()
    assert x0.is_contiguous() and noise.is_contiguous() and x0.dtype == noise.dtype
    assert alphas_cumprod.dtype == torch.float32 and timesteps.dtype == torch.int64
    if out is None:
        out = torch.empty_like(x0)
    B = x0.shape[0]
    rc = lib.pcdm_add_noise(_l.ptr(x0), _l.ptr(noise), _l.ptr(out), C.c_int(_any_dt(x0)), _l.ptr(alphas_cumprod),
                            _l.ptr(timesteps), C.c_int(B), C.c_longlong(x0.numel() // B), _stream(x0))
    _l.check(rc)
    return out


def ddim_step(model_output, sample, coefs, out=None):
    lib = _l.load()
    assert model_output.is_contiguous() and sample.is_contiguous() and model_output.shape == sample.shape
    if out is None:
        out = torch.empty_like(sample)
    rc = lib.pcdm_ddim_step(_l.ptr(model_output), C.c_int(_any_dt(model_output)), _l.ptr(sample), _l.ptr(out),
                            C.c_int(_any_dt(sample)), C.c_float(coefs[0]), C.c_float(coefs[1]), C.c_float(coefs[2]),
                            C.c_float(coefs[3]), C.c_longlong(sample.numel()), _stream(sample))
    _l.check(rc)
    return out


def ddim_step_eta(model_output, sample, noise, coefs, dir_coef, sigma, out=None):
    """The stochastic DDIM step (eta != 0): coefs as for ddim_step (its 4th entry unused), `noise` ~ N(0, 1) of the sample's
    shape and dtype."""
    lib = _l.load()
    assert model_output.is_contiguous() and sample.is_contiguous() and noise.is_contiguous()
    assert model_output.shape == sample.shape == noise.shape and noise.dtype == sample.dtype
    if out is None:
        out = torch.empty_like(sample)
    rc = lib.pcdm_ddim_step_eta(_l.ptr(model_output), C.c_int(_any_dt(model_output)), _l.ptr(sample), _l.ptr(noise),
                                _l.ptr(out), C.c_int(_any_dt(sample)), C.c_float(coefs[0]), C.c_float(coefs[1]),
                                C.c_float(coefs[2]), C.c_float(dir_coef), C.c_float(sigma), C.c_longlong(sample.numel()),
                                _stream(sample))
    _l.check(rc)
    return out


UNIPC_ROW = 16  # floats per step in the UniPC coefficient table (include/pcdm_b200.h)


def cfg_unipc_step(eps_rows, state, x9, coef_table, step_counter, guidance_scale, t_table=None, t_cur=None,
                   ratio=None, guidance_rescale=0.0):
    """eps_rows: [2n, H, W, ld] UNet output rows; state: [4, n, 4, H, W] fp32 {sample, last_sample, m0, m1} (in
    place); x9: [2n, H, W, ld9]; coef_table: [steps, 16] fp32 device."""
    lib = _l.load()
    n = state.shape[1]
    HW = state.shape[3] * state.shape[4]
    assert state.dtype == torch.float32 and state.is_contiguous() and state.shape[0] == 4 and eps_rows.is_contiguous()
    assert coef_table.dtype == torch.float32 and coef_table.shape[-1] == UNIPC_ROW and coef_table.is_contiguous()
    assert step_counter.dtype == torch.int32 and step_counter.numel() == 2
    rc = lib.pcdm_cfg_unipc_step(_l.ptr(eps_rows), C.c_int(_any_dt(eps_rows)), C.c_longlong(eps_rows.shape[-1]),
                                 _l.ptr(state), _l.ptr(x9), C.c_int(_any_dt(x9)), C.c_longlong(x9.shape[-1]),
                                 _l.ptr(coef_table), _l.ptr(step_counter), C.c_float(guidance_scale), C.c_int(n),
                                 C.c_int(HW), _l.ptr(t_table), _l.ptr(t_cur), _l.ptr(ratio), C.c_float(guidance_rescale),
                                 _stream(state))
    _l.check(rc)


def unipc_step(model_output, sample, last_sample, m0, m1, coef_row, out=None):
    """One UniPCMultistepScheduler.step on same-shape tensors; last_sample/m0/m1: fp32 history (updated in place);
    coef_row: 16 python floats."""
    lib = _l.load()
    assert model_output.is_contiguous() and sample.is_contiguous() and model_output.shape == sample.shape
    for t in (last_sample, m0, m1):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == sample.numel()
    if out is None:
        out = torch.empty_like(sample)
    row = (C.c_float * UNIPC_ROW)(*[float(v) for v in coef_row])
    rc = lib.pcdm_unipc_step(_l.ptr(model_output), C.c_int(_any_dt(model_output)), _l.ptr(sample), _l.ptr(out),
                             C.c_int(_any_dt(sample)), _l.ptr(last_sample), _l.ptr(m0), _l.ptr(m1), row,
                             C.c_longlong(sample.numel()), _stream(sample))
    _l.check(rc)
    return out


UNCLIP_ROW = 8  # floats per step in the UnCLIP coefficient table (include/pcdm_b200.h)


def cfg_unclip_step(pred, latents, xin, coef_table, noise_table, step_counter, guidance_scale, use_cfg, t_table=None,
                    t_cur=None):
    """pred: [n or 2n, E] fp32 prior outputs (rows [0, n) unconditional when use_cfg); latents: [n, E] fp32 (in place);
    xin: [n or 2n, E] model-input rows of the next step; coef_table: [steps, 8] fp32; noise_table: [steps, n, E] fp32."""
    lib = _l.load()
    n, E = latents.shape
    assert pred.dtype == torch.float32 and pred.stride(1) == 1 and pred.shape == ((2 if use_cfg else 1) * n, E)
    assert latents.dtype == torch.float32 and latents.is_contiguous() and xin.stride(1) == 1 and xin.shape == pred.shape
    assert coef_table.dtype == torch.float32 and coef_table.shape[-1] == UNCLIP_ROW and coef_table.is_contiguous()
    assert noise_table.dtype == torch.float32 and noise_table.is_contiguous() and noise_table.shape[1:] == (n, E)
    assert noise_table.shape[0] == coef_table.shape[0] and step_counter.dtype == torch.int32 and step_counter.numel() == 2
    rc = lib.pcdm_cfg_unclip_step(_l.ptr(pred), C.c_longlong(pred.stride(0)), _l.ptr(latents), _l.ptr(xin),
                                  C.c_int(_any_dt(xin)), C.c_longlong(xin.stride(0)), _l.ptr(coef_table),
                                  _l.ptr(noise_table), _l.ptr(step_counter), C.c_float(guidance_scale),
                                  C.c_int(1 if use_cfg else 0), C.c_int(n), C.c_int(E), _l.ptr(t_table), _l.ptr(t_cur),
                                  _stream(latents))
    _l.check(rc)


def unclip_step(model_output, sample, noise, coef_row, out=None):
    """One UnCLIPScheduler.step on same-shape tensors; noise: like sample, or None when the row's std is 0."""
    lib = _l.load()
    assert model_output.is_contiguous() and sample.is_contiguous() and model_output.shape == sample.shape
    assert noise is None or (noise.is_contiguous() and noise.shape == sample.shape and noise.dtype == sample.dtype)
    if out is None:
        out = torch.empty_like(sample)
    row = (C.c_float * UNCLIP_ROW)(*[float(v) for v in coef_row])
    rc = lib.pcdm_unclip_step(_l.ptr(model_output), C.c_int(_any_dt(model_output)), _l.ptr(sample), _l.ptr(noise),
                              _l.ptr(out), C.c_int(_any_dt(sample)), row, C.c_longlong(sample.numel()), _stream(sample))
    _l.check(rc)
    return out


def softmax_rows(x, scale, dtype, out=None):
    """x: [M, N] fp32 scores (unit column stride) -> softmax(scale * x) rows as `dtype` (fp16 / bf16)."""
    lib = _l.load()
    M, N = x.shape
    assert x.dtype == torch.float32 and x.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=dtype)
    assert out.stride(1) == 1 and out.shape == (M, N)
    rc = lib.pcdm_softmax_rows(_l.ptr(x), C.c_longlong(x.stride(0)), _l.ptr(out), C.c_longlong(out.stride(0)),
                               C.c_int(M), C.c_int(N), C.c_float(scale), C.c_int(_dt(out)), _stream(x))
    _l.check(rc)
    return out


def gaussian_sample(moments, B, channels, HW, noise=None, scale=1.0, out=None):
    """moments: [B*HW, ld] fp32 rows (mean | logvar); noise: [B, channels, HW] fp32 or None (mode) -> [B, channels, HW]
    fp32."""
    lib = _l.load()
    assert moments.dtype == torch.float32 and moments.stride(1) == 1 and moments.shape[0] == B * HW
    if noise is not None:
        assert noise.dtype == torch.float32 and noise.is_contiguous() and noise.numel() == B * channels * HW
    if out is None:
        out = torch.empty((B, channels, HW), device=moments.device, dtype=torch.float32)
    rc = lib.pcdm_gaussian_sample(_l.ptr(moments), C.c_longlong(moments.stride(0)), _l.ptr(noise), _l.ptr(out),
                                  C.c_int(B), C.c_int(channels), C.c_int(HW), C.c_float(scale), _stream(moments))
    _l.check(rc)
    return out
