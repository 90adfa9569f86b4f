"""Thin torch-tensor wrappers over the C ABI (one Python function per `pcdm_*` entry point).

Torch is plumbing here (device memory + streams); all arithmetic happens inside libpcdm_b200.so.  Activations are
NHWC: a tensor of logical shape [B, H, W, C] or [rows, C], contiguous in C.
"""
from __future__ import annotations

import ctypes as C

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


def _f32(t):
    if t is not None and t.dtype != torch.float32:
        raise TypeError("bias / rowvec vectors must be fp32")
    return t


# ---------------------------------------------------------------------------------------------------------------
# weight packing (host side, done once at load time)
# ---------------------------------------------------------------------------------------------------------------
def pack_conv3x3_weight(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """[Cout, Cin, 3, 3] (torch Conv2d layout) -> [Cout, 3, 3, Cin] flattened to [Cout, 9*Cin] (tap-major K)."""
    assert w.dim() == 4 and w.shape[2:] == (3, 3)
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(dtype).contiguous()


def geglu_row_permutation(n_out: int) -> torch.Tensor:
    """Row order for a GEGLU projection weight [2*n_out, K]: groups of [32 value rows | 32 gate rows]."""
    assert n_out % 32 == 0
    idx = torch.arange(n_out).view(-1, 32)
    return torch.cat([idx, idx + n_out], dim=1).reshape(-1)


# ---------------------------------------------------------------------------------------------------------------
# K1/K2
# ---------------------------------------------------------------------------------------------------------------
def gemm(a, w, out=None, *, a2=None, bias=None, rowvec=None, rows_per_image=1, residual=None, geglu=False,
         out_f32=False, bn=0):
    """out[M, N] = [a | a2][M, K] @ w[N, K]^T (+bias) (+rowvec[m // rows_per_image]) (+residual)."""
    lib = _l.load()
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
    flags = (_l.FLAG_GEGLU if geglu else 0) | (_l.FLAG_OUT_F32 if out_f32 else 0)
    rc = lib.pcdm_gemm(
        _l.ptr(a), C.c_longlong(a.stride(0)), _l.ptr(a2), C.c_longlong(a2.stride(0) if a2 is not None else 0),
        C.c_int(k1), _l.ptr(w), _l.ptr(out), C.c_longlong(out.stride(0)), _l.ptr(_f32(bias)), _l.ptr(_f32(rowvec)),
        C.c_int(rows_per_image), _l.ptr(residual), C.c_longlong(residual.stride(0) if residual is not None else 0),
        C.c_int(M), C.c_int(N), C.c_int(K), C.c_int(_dt(a)), C.c_int(flags), C.c_int(bn), _stream(a))
    _l.check(rc)
    return out


def conv3x3(x, w_packed, out=None, *, bias=None, rowvec=None, residual=None, stride=1, out_f32=False, bn=0):
    """x: [B, Hin, Win, Cin] NHWC; w_packed: [Cout, 9*Cin]; returns [B, Hin/stride, Win/stride, Cout]."""
    lib = _l.load()
    B, Hin, Win, Cin = x.shape
    Cout = w_packed.shape[0]
    assert w_packed.shape[1] == 9 * Cin and x.is_contiguous() and w_packed.is_contiguous()
    H, W = Hin // stride, Win // stride
    if out is None:
        out = torch.empty((B, H, W, Cout), device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    assert out.is_contiguous() and tuple(out.shape) == (B, H, W, Cout)
    if residual is not None:
        assert residual.is_contiguous() and residual.shape == out.shape
    flags = _l.FLAG_OUT_F32 if out_f32 else 0
    rc = lib.pcdm_conv3x3(_l.ptr(x), _l.ptr(w_packed), _l.ptr(out), _l.ptr(_f32(bias)), _l.ptr(_f32(rowvec)),
                          _l.ptr(residual), C.c_int(B), C.c_int(H), C.c_int(W), C.c_int(Cin), C.c_int(Cout),
                          C.c_int(stride), C.c_int(_dt(x)), C.c_int(flags), C.c_int(bn), _stream(x))
    _l.check(rc)
    return out
