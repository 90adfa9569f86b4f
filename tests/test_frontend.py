"""Conditioning front-end: `ImageProjModel_p` and `ControlNetConditioningEmbedding` (pose_proj).

CPU: the oracle's ImageProjModel_p must be BIT-EQUAL to the reference's own class, executed from
/root/reference/stage2_batchtest_inpaint_model.py:48-66 (skipped when the reference tree is absent: the GPU box);
the product's host logic (channel padding, key handling) against the oracle through tests/mock_ops.py.
GPU: product vs oracle on identical weights and inputs at the real sizes.
"""
import ast
import os

import pytest
import torch

from oracle.frontend import ControlNetConditioningEmbedding, ImageProjModel_p, make_frontend
from pcdms_b200.frontend import B200ControlNetConditioningEmbedding, B200ImageProjModel_p
from tests import mock_ops

REF = "/root/reference/stage2_batchtest_inpaint_model.py"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("path", [REF, "/root/reference/stage3_batchtest_refined_model.py"])
def test_oracle_image_proj_is_the_references_class(path):
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ImageProjModel_p")
    ns = {"torch": torch, "nn": torch.nn}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    ref = ns["ImageProjModel_p"](in_dim=1536, hidden_dim=768, out_dim=1024).eval()
    ours, _ = make_frontend(seed=0)
    assert set(ref.state_dict()) == set(ours.state_dict())
    ref.load_state_dict(ours.state_dict())
    x = torch.randn(1, 257, 1536, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        assert torch.equal(ref(x), ours(x))


def test_oracle_pose_embedding_structure():
    _, pose = make_frontend()
    convs = [(3, 16), (16, 16), (16, 32), (32, 32), (32, 96), (96, 96), (96, 256), (256, 320)]
    assert sum(p.numel() for p in pose.parameters()) == sum(9 * a * b + b for a, b in convs) == 1_086_480
    assert list(pose.state_dict()) == [f"{n}.{p}" for n in ["conv_in"] + [f"blocks.{i}" for i in range(6)] + ["conv_out"]
                                       for p in ("weight", "bias")]
    z = ControlNetConditioningEmbedding(320)
    with torch.no_grad():
        assert float(z(torch.randn(1, 3, 32, 64)).abs().max()) == 0.0   # zero_module(conv_out) at init
        y = pose(torch.randn(1, 3, 64, 128))
    assert y.shape == (1, 320, 8, 16)


def test_host_logic_matches_oracle():
    proj, pose = make_frontend(seed=2, pose_channels=(16, 32, 96, 128), out_channels=64, in_dim=128, hidden_dim=64,
                               out_dim=96)
    p1 = B200ImageProjModel_p(128, 64, 96, dtype=torch.float32, device="cpu")
    p1.load_state_dict(proj.state_dict())
    p2 = B200ControlNetConditioningEmbedding(64, 3, (16, 32, 96, 128), dtype=torch.float32, device="cpu")
    p2.load_state_dict(pose.state_dict())
    g = torch.Generator().manual_seed(3)
    x, c = torch.randn(1, 9, 128, generator=g), torch.randn(1, 3, 32, 64, generator=g)
    with torch.no_grad(), mock_ops.patched():
        p1._check = p2._check = lambda *a: None   # the CUDA-only guard; the stand-ins run on CPU
        torch.testing.assert_close(p1(x), proj(x), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(p2(c), pose(c), rtol=1e-4, atol=1e-5)
    with pytest.raises(RuntimeError):
        B200ImageProjModel_p(128, 64, 96, device="cpu").load_state_dict({})
    q = B200ControlNetConditioningEmbedding(64, 3, (16, 32, 96, 128), dtype=torch.float32, device="cpu")
    q.load_state_dict(pose.state_dict())
    with pytest.raises(RuntimeError):   # no CPU compute path
        q(c)


@pytest.mark.gpu
@pytest.mark.parametrize("dt,tol", [(torch.float16, 4e-3), (torch.bfloat16, 3e-2)])
def test_frontend_gpu_real_sizes(dt, tol):
    proj, pose = make_frontend(seed=1)
    p1 = B200ImageProjModel_p(dtype=dt)
    p1.load_state_dict(proj.state_dict())
    p2 = B200ControlNetConditioningEmbedding(dtype=dt)
    p2.load_state_dict(pose.state_dict())
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 257, 1536, generator=g).to(dt)
    c = torch.rand(1, 3, 256, 512, generator=g) * 2 - 1          # the 256 x (2 x 256) source|target pose canvas
    with torch.no_grad():
        want1, want2 = proj(x.float()), pose(c)
        got1, got2 = p1(x.cuda()), p2(c.cuda())
    assert got1.shape == (1, 257, 1024) and got2.shape == (1, 320, 32, 64) and got2.dtype == dt
    for got, want in ((got1, want1), (got2, want2)):
        err = (got.float().cpu() - want).abs().max() / want.abs().max()
        assert float(err) < tol, float(err)


@pytest.mark.gpu
def test_image_proj_against_reference_golden_gpu():
    """tests/golden/ref_image_proj.pt holds the output of the reference's own ImageProjModel_p class."""
    from pathlib import Path
    g = torch.load(Path(__file__).resolve().parent / "golden" / "ref_image_proj.pt")
    p = B200ImageProjModel_p(128, 64, 96, dtype=torch.float16)
    p.load_state_dict(g["state_dict"])
    got = p(g["x"].cuda())
    err = (got.float().cpu() - g["y"]).abs().max() / g["y"].abs().max()
    assert float(err) < 4e-3
