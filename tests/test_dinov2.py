"""DINOv2 image encoder (reference call sites /root/reference/stage2_batchtest_inpaint_model.py:98,168).

The oracle here is the real thing: `transformers.Dinov2Model` (the library class the reference instantiates), random
weights, run on CPU in fp32.  (This image has transformers 5.5; the reference pins 4.32.1, whose only arithmetic
difference on this path is how the bicubic position-embedding interpolation derives its scale.)
CPU: host logic (weight folding / padding / interleaving, position interpolation, token layout) through
tests/mock_ops.py.  GPU: the kernels, at a tiny config and at dinov2-giant width (1536, 24 heads, SwiGLU 4096) with the
224 x 224 input the reference feeds a model trained at 518 x 518.
"""
import pytest
import torch

from pcdms_b200.dinov2 import B200Dinov2Model
from tests import mock_ops

transformers = pytest.importorskip("transformers")


def _hf(seed=0, **kw):
    from transformers import Dinov2Config, Dinov2Model
    torch.manual_seed(seed)
    cfg = Dinov2Config(use_swiglu_ffn=True, **kw)
    m = Dinov2Model(cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():   # make LayerScale / norms / biases non-trivial
        for n, p in m.named_parameters():
            if n.endswith("lambda1"):
                p.copy_(0.5 + 0.1 * torch.randn(p.shape, generator=g))
            elif "norm" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith(".bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif "position_embeddings" in n or "cls_token" in n:
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
    return cfg, m


TINY = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, mlp_ratio=4, image_size=56, patch_size=14)


@pytest.mark.parametrize("size", [56, 84])   # native grid, and a grid that needs position interpolation
def test_host_logic_matches_transformers(size):
    cfg, hf = _hf(**TINY)
    p = B200Dinov2Model(cfg, dtype=torch.float32, device="cpu")
    assert set(p.state_dict_shapes()) == set(hf.state_dict())
    assert p.ffn_hidden == 344 and p.ffn_hidden_padded == 384      # (int(128*4*2/3)+7)//8*8, padded to 64
    p.load_state_dict(hf.state_dict())
    p._guard = lambda x: None
    x = torch.randn(2, 3, size, size, generator=torch.Generator().manual_seed(3))
    with torch.no_grad(), mock_ops.patched():
        got = p(x)
        want = hf(x)
    torch.testing.assert_close(got.last_hidden_state, want.last_hidden_state, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(got.pooler_output, want.pooler_output, rtol=2e-4, atol=2e-5)


def test_surface_and_errors():
    p = B200Dinov2Model(device="cpu")     # dinov2-giant defaults
    assert p.config.hidden_size == 1536 and p.config.num_hidden_layers == 40 and p.ffn_hidden == 4096
    n = sum(torch.Size(s).numel() for k, s in p.state_dict_shapes().items())
    assert n == 1_136_480_768             # published size of facebook/dinov2-giant
    with pytest.raises(RuntimeError):
        p(torch.zeros(1, 3, 224, 224))    # not loaded
    with pytest.raises(NotImplementedError):
        B200Dinov2Model(device="cpu", use_swiglu_ffn=False)
    cfg, hf = _hf(**TINY)
    q = B200Dinov2Model(cfg, dtype=torch.float32, device="cpu")
    q.load_state_dict(hf.state_dict())
    with pytest.raises(RuntimeError):     # no CPU compute path
        q(torch.zeros(1, 3, 56, 56))
    with pytest.raises(ValueError):
        q._guard = lambda x: None
        q(torch.zeros(1, 3, 50, 56))


def _rel(got, want):
    return float((got.float().cpu() - want).abs().max() / want.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("dt,tol", [(torch.float16, 5e-3), (torch.bfloat16, 4e-2)])
def test_dinov2_tiny_gpu(dt, tol):
    cfg, hf = _hf(**TINY)
    p = B200Dinov2Model(cfg, dtype=dt)
    p.load_state_dict(hf.state_dict())
    for size in (56, 84):
        x = torch.randn(2, 3, size, size, generator=torch.Generator().manual_seed(3))
        with torch.no_grad():
            want = hf(x).last_hidden_state
        got = p(x.cuda()).last_hidden_state
        assert got.shape == want.shape and got.dtype == dt
        assert _rel(got, want) < tol


@pytest.mark.gpu
def test_dinov2_giant_width_gpu():
    """dinov2-giant's layer shape (1536 wide, 24 heads, SwiGLU hidden 4096, trained grid 37 x 37) at reduced depth,
    on the 224 x 224 crop the reference's CLIPImageProcessor produces: 257 tokens out."""
    cfg, hf = _hf(hidden_size=1536, num_hidden_layers=4, num_attention_heads=24, mlp_ratio=4, image_size=518,
                  patch_size=14)
    p = B200Dinov2Model(cfg, dtype=torch.float16)
    p.load_state_dict(hf.state_dict())
    x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want = hf(x).last_hidden_state
    got = p(x.cuda()).last_hidden_state
    assert got.shape == (1, 257, 1536)
    assert _rel(got, want) < 5e-3
    assert torch.equal(got, p(x.cuda()).last_hidden_state)
