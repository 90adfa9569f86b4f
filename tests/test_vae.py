"""AutoencoderKL (reference call sites /root/reference/src/pipelines/stage2_inpaint_pipeline.py:443,528).

CPU: the oracle restatement (oracle/vae.py) is pinned by the published parameter count of the SD VAE and closed forms;
the product's host logic (key handling, weight packing / padding, block walk, attention decomposition) is checked
against it through the torch stand-ins of tests/mock_ops.py.  GPU: product vs oracle on identical weights and inputs.
"""
from dataclasses import asdict

import pytest
import torch
import torch.nn.functional as F

from oracle.vae import OracleAutoencoderKL, VAEAttention, VAEConfig, make_vae
from pcdms_b200.vae import B200AutoencoderKL
from tests import mock_ops


def _cfg_kw(cfg):
    return {k: v for k, v in asdict(cfg).items()}


def test_oracle_parameter_count_and_keys():
    m = OracleAutoencoderKL()
    assert sum(p.numel() for p in m.parameters()) == 83_653_863      # published size of the SD-1.x/2.x VAE
    assert len(m.state_dict()) == 248
    p = B200AutoencoderKL(device="cpu")
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == p.state_dict_shapes()


def test_oracle_attention_closed_form():
    torch.manual_seed(0)
    a = VAEAttention(64, 32).eval()
    x = torch.randn(2, 64, 4, 8)
    with torch.no_grad():
        got = a(x)
        hn = F.group_norm(x, 32, a.group_norm.weight, a.group_norm.bias, 1e-6).flatten(2).transpose(1, 2)
        q, k, v = a.to_q(hn), a.to_k(hn), a.to_v(hn)
        want = torch.stack([torch.softmax(q[b] @ k[b].t() / 8.0, -1) @ v[b] for b in range(2)])
        want = a.to_out[0](want).transpose(1, 2).reshape(2, 64, 4, 8) + x
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


def test_oracle_gaussian_and_downsample_padding():
    m = make_vae(VAEConfig.tiny())
    x = torch.randn(1, 3, 32, 64, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        d = m.encode(x).latent_dist
        noise = torch.randn(d.mean.shape, generator=torch.Generator().manual_seed(2))
        torch.testing.assert_close(d.sample(noise=noise), d.mean + torch.exp(0.5 * d.logvar) * noise)
        assert d.mean.shape == (1, 4, 4, 8)
        ds = m.encoder.down_blocks[0].downsamplers[0]
        h = torch.randn(1, 64, 6, 6)
        want = F.conv2d(F.pad(h, (0, 1, 0, 1)), ds.conv.weight, ds.conv.bias, stride=2)
        torch.testing.assert_close(ds(h), want)


@pytest.mark.parametrize("hw", [(64, 128), (32, 256)])
def test_host_logic_matches_oracle(hw):
    cfg = VAEConfig.tiny()
    o = make_vae(cfg, seed=3)
    p = B200AutoencoderKL(dtype=torch.float32, device="cpu", **_cfg_kw(cfg))
    p.load_state_dict(o.state_dict())
    x = torch.randn(2, 3, *hw, generator=torch.Generator().manual_seed(4))
    with torch.no_grad(), mock_ops.patched():
        p._check_input = lambda *a, **k: None   # the CUDA-only guard; the stand-ins run on CPU
        dist = p.encode(x).latent_dist
        want = o.encode(x).latent_dist
        torch.testing.assert_close(dist.mode(), want.mode(), rtol=1e-4, atol=1e-5)
        z = want.sample(generator=torch.Generator().manual_seed(5))
        got = p.decode(z, return_dict=False)[0]
        torch.testing.assert_close(got, o.decode(z).sample, rtol=1e-4, atol=1e-5)


def test_deprecated_attention_keys_and_errors():
    cfg = VAEConfig.tiny()
    sd = make_vae(cfg).state_dict()
    old = {}
    ren = {"to_q": "query", "to_k": "key", "to_v": "value", "to_out.0": "proj_attn"}
    for k, v in sd.items():
        for new, dep in ren.items():
            if f".attentions.0.{new}." in k:
                k = k.replace(f".{new}.", f".{dep}.")
                v = v[:, :, None, None] if v.dim() == 2 else v    # the deprecated block stored 1x1 convs
        old[k] = v
    assert any(".query." in k for k in old)
    p = B200AutoencoderKL(dtype=torch.float32, device="cpu", **_cfg_kw(cfg))
    p.load_state_dict(old)
    q = B200AutoencoderKL(dtype=torch.float32, device="cpu", **_cfg_kw(cfg))
    q.load_state_dict(sd)
    assert all(torch.equal(p._w[k], q._w[k]) for k in q._w)
    sd.pop("quant_conv.bias")
    with pytest.raises(RuntimeError):
        q.load_state_dict(sd)
    with pytest.raises(RuntimeError):   # no CPU compute path
        p.decode(torch.zeros(1, 4, 8, 16))
    with pytest.raises(NotImplementedError):
        B200AutoencoderKL(device="cpu", block_out_channels=(96, 128))


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
def _rel(got, want):
    return float((got.float().cpu() - want).abs().max() / want.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("hw", [(64, 128), (32, 256)])
def test_vae_tiny_gpu(dt, hw):
    cfg = VAEConfig.tiny()
    o = make_vae(cfg, seed=3)
    p = B200AutoencoderKL(dtype=dt, **_cfg_kw(cfg))
    p.load_state_dict(o.state_dict())
    x = torch.randn(2, 3, *hw, generator=torch.Generator().manual_seed(4))
    tol = 4e-2 if dt == torch.bfloat16 else 6e-3
    with torch.no_grad():
        want = o.encode(x).latent_dist
        dist = p.encode(x.cuda()).latent_dist
        assert _rel(dist.mode(), want.mode()) < tol
        g = torch.Generator().manual_seed(5)
        noise = torch.randn(want.mean.shape, generator=g)
        got = p.encode(x.cuda()).latent_dist.sample(generator=torch.Generator().manual_seed(5))
        assert _rel(got, want.sample(noise=noise)) < tol
        z = want.sample(noise=noise)
        img = p.decode(z.cuda().to(dt), return_dict=False)[0]
        assert img.shape == (2, 3, *hw) and img.dtype == dt
        assert _rel(img, o.decode(z.to(dt).float()).sample) < tol


@pytest.mark.gpu
def test_vae_full_size_gpu():
    """The real SD-2.1 VAE architecture (83.65 M parameters) on a 256 x 512 canvas (BASELINE config 2 images):
    encode -> latents 32 x 64, decode -> image, against the fp32 CPU oracle."""
    cfg = VAEConfig()
    o = make_vae(cfg, seed=1)
    p = B200AutoencoderKL(dtype=torch.float16)
    p.load_state_dict(o.state_dict())
    x = torch.randn(1, 3, 256, 512, generator=torch.Generator().manual_seed(2)).clamp(-1, 1)
    with torch.no_grad():
        want = o.encode(x).latent_dist
        got = p.encode(x.cuda()).latent_dist.mode()
        assert got.shape == (1, 4, 32, 64)
        assert _rel(got, want.mode()) < 1e-2
        z = want.mode()
        img = p.decode(z.cuda().half(), return_dict=False)[0]
        ref = o.decode(z.half().float()).sample
        assert img.shape == (1, 3, 256, 512)
        assert _rel(img, ref) < 1e-2
        assert torch.equal(img, p.decode(z.cuda().half(), return_dict=False)[0])   # deterministic
