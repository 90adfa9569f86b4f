"""CLIP ViT-H/14 image encoder (reference call sites
/root/reference/stage1_batchtest_prior_model.py:61,100-101, src/pipelines/stage1_prior_pipeline.py:282-289,
stage2_batchtest_inpaint_model.py:97,181-183).

The oracle is the library class the reference instantiates: `transformers.CLIPVisionModelWithProjection`, random
weights, CPU fp32.  CPU: host logic (head padding 80 -> 128 of q/k/v rows and out_proj columns, patch unfold, position
table) through tests/mock_ops.py.  GPU: the kernels (incl. the head_dim-128 attention variant) at a tiny config, at
ViT-H/14's layer shape, and the attention kernel alone against explicit softmax.
"""
import pytest
import torch

from pcdms_b200.clip import B200CLIPVisionModelWithProjection
from tests import mock_ops

transformers = pytest.importorskip("transformers")


def _hf(seed=0, **kw):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    torch.manual_seed(seed)
    cfg = CLIPVisionConfig(hidden_act="gelu", **kw)
    m = CLIPVisionModelWithProjection(cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():   # make norms / biases / embeddings non-trivial
        for n, p in m.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith(".bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif "position_embedding" in n or "class_embedding" in n:
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
            elif n.endswith(".weight"):
                p.copy_(torch.randn(p.shape, generator=g) * p[0].numel() ** -0.5)
    return cfg, m


TINY = dict(hidden_size=320, intermediate_size=1280, num_hidden_layers=2, num_attention_heads=4, image_size=56,
            patch_size=14, projection_dim=64)     # head_dim 80, as ViT-H/14


def test_host_logic_matches_transformers():
    cfg, hf = _hf(**TINY)
    p = B200CLIPVisionModelWithProjection(cfg, dtype=torch.float32, device="cpu")
    assert set(p.state_dict_shapes()) == set(hf.state_dict())
    assert p.head_dim == 80 and p.head_dim_padded == 128
    p.load_state_dict(hf.state_dict())
    p._guard = lambda x: None
    x = torch.randn(2, 3, 56, 56, generator=torch.Generator().manual_seed(3))
    with torch.no_grad(), mock_ops.patched():
        got = p(x)
        want = hf(x)
    torch.testing.assert_close(got.image_embeds, want.image_embeds, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(got["last_hidden_state"], want.last_hidden_state, rtol=2e-4, atol=2e-5)
    with torch.no_grad(), mock_ops.patched():    # a grid the position table was not trained on
        x2 = torch.randn(1, 3, 84, 84, generator=torch.Generator().manual_seed(4))
        with pytest.raises(ValueError):
            p(x2)
        got2, want2 = p(x2, interpolate_pos_encoding=True), hf(x2, interpolate_pos_encoding=True)
    torch.testing.assert_close(got2.image_embeds, want2.image_embeds, rtol=2e-4, atol=2e-5)


def test_surface_and_errors():
    p = B200CLIPVisionModelWithProjection(device="cpu")     # ViT-H/14 defaults
    assert p.config.hidden_size == 1280 and p.config.num_hidden_layers == 32 and p.head_dim == 80
    n = sum(torch.Size(s).numel() for k, s in p.state_dict_shapes().items())
    assert n == 632_076_800               # CLIP ViT-H/14 vision tower + projection (the published 632 M)
    with pytest.raises(RuntimeError):
        p(torch.zeros(1, 3, 224, 224))    # not loaded
    with pytest.raises(NotImplementedError):
        B200CLIPVisionModelWithProjection(device="cpu", hidden_act="quick_gelu")
    cfg, hf = _hf(**TINY)
    q = B200CLIPVisionModelWithProjection(cfg, dtype=torch.float32, device="cpu")
    sd = dict(hf.state_dict())
    sd["vision_model.embeddings.position_ids"] = torch.arange(17)[None]   # transformers < 4.31 persisted this buffer
    q.load_state_dict(sd)
    with pytest.raises(RuntimeError):     # no CPU compute path
        q(torch.zeros(1, 3, 56, 56))
    with pytest.raises(RuntimeError):
        q.load_state_dict({k: v for k, v in sd.items() if "fc1" not in k})


def _rel(got, want):
    return float((got.float().cpu() - want).abs().max() / want.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("dt,tol", [(torch.float16, 1e-3), (torch.bfloat16, 8e-3)])
@pytest.mark.parametrize("B,heads,Sq,Skv", [(2, 3, 257, 257), (1, 2, 128, 6), (1, 16, 300, 515)])
def test_attention_head_dim_128_gpu(dt, tol, B, heads, Sq, Skv):
    """pcdm_attention_hd at head_dim 128 against explicit softmax on the same 16-bit inputs; q/k/v are column slices
    of one fused buffer, as the encoder passes them."""
    from pcdms_b200 import ops
    g = torch.Generator().manual_seed(7)
    C = heads * 128
    qkv_q = torch.randn(B * Sq, C, generator=g).to(dt)
    kv = torch.randn(B * Skv, 2 * C, generator=g).to(dt)
    scale = 80 ** -0.5
    q4 = qkv_q.float().view(B, Sq, heads, 128).transpose(1, 2)
    k4 = kv[:, :C].float().view(B, Skv, heads, 128).transpose(1, 2)
    v4 = kv[:, C:].float().view(B, Skv, heads, 128).transpose(1, 2)
    want = (torch.softmax(q4 @ k4.transpose(-1, -2) * scale, -1) @ v4).transpose(1, 2).reshape(B * Sq, C)
    kvd = kv.cuda()
    got = ops.attention(qkv_q.cuda(), kvd[:, :C], kvd[:, C:], B, heads, scale=scale, head_dim=128)
    torch.testing.assert_close(got.float().cpu(), want, rtol=4 * tol, atol=4 * tol * 0.1)
    assert torch.equal(got, ops.attention(qkv_q.cuda(), kvd[:, :C], kvd[:, C:], B, heads, scale=scale, head_dim=128))


@pytest.mark.gpu
@pytest.mark.parametrize("dt,tol", [(torch.float16, 5e-3), (torch.bfloat16, 4e-2)])
def test_clip_tiny_gpu(dt, tol):
    cfg, hf = _hf(**TINY)
    p = B200CLIPVisionModelWithProjection(cfg, dtype=dt)
    p.load_state_dict(hf.state_dict())
    x = torch.randn(2, 3, 56, 56, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = hf(x)
    got = p(x.cuda())
    assert got.image_embeds.shape == want.image_embeds.shape and got.image_embeds.dtype == dt
    assert _rel(got.image_embeds, want.image_embeds) < tol
    assert _rel(got.last_hidden_state, want.last_hidden_state) < tol


@pytest.mark.gpu
def test_clip_vit_h_width_gpu():
    """ViT-H/14's layer shape (1280 wide, 16 heads of 80, MLP 5120, 224 x 224 -> 257 tokens, projection 1024) at
    reduced depth."""
    cfg, hf = _hf(hidden_size=1280, intermediate_size=5120, num_hidden_layers=3, num_attention_heads=16,
                  image_size=224, patch_size=14, projection_dim=1024)
    p = B200CLIPVisionModelWithProjection(cfg, dtype=torch.float16)
    p.load_state_dict(hf.state_dict())
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want = hf(x)
    got = p(x.cuda())
    assert got.image_embeds.shape == (2, 1024) and got.last_hidden_state.shape == (2, 257, 1280)
    assert _rel(got.image_embeds, want.image_embeds) < 5e-3
    assert _rel(got.last_hidden_state, want.last_hidden_state) < 5e-3
    assert torch.equal(got.image_embeds, p(x.cuda()).image_embeds)
