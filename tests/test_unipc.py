"""UniPCMultistepScheduler — the scheduler the reference's batch-test drivers install
(/root/reference/stage2_batchtest_inpaint_model.py:132).

CPU part: pins the oracle restatement (oracle/unipc.py; diffusers itself is not installable here) through properties of
the published algorithm — UniP-p converges with order p and UniP-p + UniC-p with order p + 1 on an analytic diffusion
whose probability-flow solution is known in closed form; the order-1 predictor is the DDIM step; timestep / sigma
tables — and checks the product's host logic (coefficient table + pipeline orchestration) against it through the torch
stand-ins of tests/mock_ops.py.  GPU part: the CUDA step kernels against the oracle, BIT-EXACT in fp32 (the kernels
evaluate the reference's expressions operation by operation with IEEE round-to-nearest arithmetic).
"""
from dataclasses import asdict

import numpy as np
import pytest
import torch

from oracle.schedulers import OracleDDIMScheduler
from oracle.unipc import UniPCMultistepScheduler as OracleUniPC
from pcdms_b200.scheduler import B200UniPCMultistepScheduler
from tests import mock_ops

# what `UniPCMultistepScheduler.from_config(pipe.scheduler.config)` receives for stable-diffusion-2-1-base (a PNDM
# config): keys UniPC does not know are ignored
SD21_SCHED = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                  steps_offset=1, timestep_spacing="leading", skip_prk_steps=True, set_alpha_to_one=False,
                  clip_sample=False, prediction_type="epsilon", trained_betas=None)


def _gauss_run(N, solver_order, corrector, s2=0.25):
    """Sample the probability-flow ODE of x0 ~ N(0, s2 I) on a uniform log-SNR grid with the exact epsilon model and
    return the max error against the closed-form solution x(t) = x_T sqrt(a_t^2 s2 + s_t^2) / sqrt(a_T^2 s2 + s_T^2)."""
    sch = OracleUniPC.from_config(SD21_SCHED, lower_order_final=False, solver_order=solver_order,
                                  disable_corrector=() if corrector else tuple(range(N)))
    sch.set_timesteps(N)
    lam = np.linspace(-3.0, 2.0, N + 1)
    sch.sigmas = torch.tensor(np.exp(-lam), dtype=torch.float64)
    sch.timesteps = torch.arange(N, 0, -1)

    def a_s(s):
        a = 1 / np.sqrt(s * s + 1)
        return a, s * a

    x = torch.randn(1, 4, 4, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    xT = x.clone()
    for i, t in enumerate(sch.timesteps):
        a, sg = a_s(float(sch.sigmas[i]))
        x = sch.step(sg * x / (a * a * s2 + sg * sg), t, x, return_dict=False)[0]
    a0, s0 = a_s(float(sch.sigmas[-1]))
    aT, sT = a_s(float(sch.sigmas[0]))
    exact = xT * np.sqrt(a0 * a0 * s2 + s0 * s0) / np.sqrt(aT * aT * s2 + sT * sT)
    return float((x - exact).abs().max())


@pytest.mark.parametrize("solver_order,corrector,order", [(1, False, 1), (2, False, 2), (1, True, 2), (2, True, 3)])
def test_oracle_convergence_order(solver_order, corrector, order):
    e40, e80, e160 = (_gauss_run(N, solver_order, corrector) for N in (40, 80, 160))
    for ratio in (e40 / e80, e80 / e160):
        assert 2 ** order * 0.85 < ratio < 2 ** order * 1.15, (e40, e80, e160)


def test_oracle_first_step_is_ddim():
    """UniP-1 in x0-prediction form: x_t = (sigma_t/sigma_s) x + alpha_t (1 - e^-h) x0 == the DDIM (eta 0) step."""
    u = OracleUniPC.from_config(SD21_SCHED)
    u.set_timesteps(20)
    ac = u.alphas_cumprod.double()
    t0, t1 = int(u.timesteps[0]), int(u.timesteps[1])
    g = torch.Generator().manual_seed(1)
    x, e = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    got = u.step(e, u.timesteps[0], x, return_dict=False)[0]
    x0 = (x.double() - (1 - ac[t0]).sqrt() * e.double()) / ac[t0].sqrt()
    want = ac[t1].sqrt() * x0 + (1 - ac[t1]).sqrt() * e.double()
    torch.testing.assert_close(got.double(), want, rtol=2e-5, atol=2e-6)
    d = OracleDDIMScheduler()   # same formula through the DDIM oracle (its grid: 981 -> 931 for 20 steps)
    d.set_timesteps(20)
    assert d.timesteps[0] == 951


def test_tables():
    u, p = OracleUniPC.from_config(SD21_SCHED), B200UniPCMultistepScheduler.from_config(SD21_SCHED)
    for n in (20, 50):
        u.set_timesteps(n)
        p.set_timesteps(n)
        assert u.timesteps.tolist() == p.timesteps.tolist()
        assert torch.equal(u.sigmas, p.sigmas) and len(p.sigmas) == n + 1
    u.set_timesteps(20)   # leading spacing over 21 intervals of 47, offset 1
    p.set_timesteps(20)
    assert u.timesteps.tolist() == [1 + 47 * k for k in range(20, 0, -1)]
    ac = u.alphas_cumprod
    torch.testing.assert_close(u.sigmas[0], ((1 - ac[941]) / ac[941]).sqrt())
    torch.testing.assert_close(u.sigmas[-1], ((1 - ac[0]) / ac[0]).sqrt())
    assert p.init_noise_sigma == 1.0 and p.order == 1 and p.config.solver_order == 2 and p.config.solver_type == "bh2"
    import inspect
    assert not {"eta", "generator"} & set(inspect.signature(p.step).parameters)   # reference :313-321 introspects
    rows = p.coefficient_rows()
    assert len(rows) == 20 and all(len(r) == 16 for r in rows)
    assert [r[15] for r in rows] == [1.0] + [2.0] * 18 + [1.0]      # predictor: warm-up, order 2, lower_order_final
    assert [r[2] for r in rows] == [0.0] + [1.0] * 19               # corrector on every step but the first
    assert [r[9] for r in rows[1:]] == [1.0] + [2.0] * 18           # corrector order = previous predictor order
    with pytest.raises(NotImplementedError):
        B200UniPCMultistepScheduler(solver_order=3)
    with pytest.raises(RuntimeError):
        p.step(torch.zeros(1, 4, 2, 2), 941, torch.zeros(1, 4, 2, 2))   # CPU tensors: no fallback


def _trajectory(sched, steps, x, eps_list, step_fn=None):
    sched.set_timesteps(steps)
    out = []
    for t, e in zip(sched.timesteps, eps_list):
        x = (step_fn or sched.step)(e, t, x, return_dict=False)[0]
        out.append(x)
    return out


@pytest.mark.parametrize("steps", [2, 3, 20])
def test_host_coefficients_reproduce_oracle_bit_exact(steps):
    """Product coefficient table + the documented kernel arithmetic (mock) == oracle trajectory, bit for bit."""
    g = torch.Generator().manual_seed(steps)
    x = torch.randn(2, 4, 8, 8, generator=g)
    eps = [torch.randn(2, 4, 8, 8, generator=g) for _ in range(steps)]
    want = _trajectory(OracleUniPC.from_config(SD21_SCHED), steps, x, eps)
    p = B200UniPCMultistepScheduler.from_config(SD21_SCHED)
    p.set_timesteps(steps)
    hist = torch.zeros(3, x.numel())
    cur = x
    for i, e in enumerate(eps):
        cur = mock_ops.unipc_step(e, cur, hist[0], hist[1], hist[2], p.coefficient_rows()[i])
        assert torch.equal(cur, want[i]), i


def test_pipeline_orchestration_with_unipc():
    from oracle.factory import make_inputs, make_unet
    from oracle.pipeline import denoise_loop, prepare_conditioning
    from oracle.unet import UNetConfig
    from pcdms_b200.pipeline import B200Stage2InpaintPipeline
    from pcdms_b200.unet import B200UNet2DConditionModel
    cfg = UNetConfig.tiny()
    o = make_unet(cfg, seed=0)
    m = B200UNet2DConditionModel(dtype=torch.float32, device="cpu", **asdict(cfg))
    m.load_state_dict(o.state_dict())
    pin = make_inputs(cfg, n=2, h=8, w=16, s_kv=6)
    cond = prepare_conditioning(s_img_proj_f=pin["s_img_proj_f"], pred_t_img_embed=pin["pred_t_img_embed"],
                                st_pose_f=pin["st_pose_f"], masked_latents=pin["masked_latents"], height=pin["height"],
                                width=pin["width"], num_images_per_prompt=2, guidance_scale=2.0)
    want = denoise_loop(o, OracleUniPC.from_config(SD21_SCHED), latents=pin["latents"], cond=cond,
                        num_inference_steps=6, guidance_scale=2.0)
    pipe = B200Stage2InpaintPipeline(vae=None, unet=m, scheduler=B200UniPCMultistepScheduler.from_config(SD21_SCHED))
    pipe.use_cuda_graph = False
    with mock_ops.patched():
        for _ in range(2):   # the second call must start from a clean multistep history
            got = pipe(height=pin["height"], width=pin["width"], num_inference_steps=6, guidance_scale=2.0,
                       num_images_per_prompt=2, latents=pin["latents"], output_type="latent",
                       s_img_proj_f=pin["s_img_proj_f"], st_pose_f=pin["st_pose_f"],
                       pred_t_img_embed=pin["pred_t_img_embed"], masked_latents=pin["masked_latents"]).images
            torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-5)


# ---------------------------------------------------------------------------------------------------------------
# GPU: the CUDA kernels
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("steps", [1, 2, 20, 50])
def test_unipc_step_kernel_bit_exact(steps):
    g = torch.Generator().manual_seed(100 + steps)
    x = torch.randn(3, 4, 16, 32, generator=g)
    eps = [torch.randn(3, 4, 16, 32, generator=g) for _ in range(steps)]
    want = _trajectory(OracleUniPC.from_config(SD21_SCHED), steps, x, eps)
    p = B200UniPCMultistepScheduler.from_config(SD21_SCHED)
    got = _trajectory(p, steps, x.cuda(), [e.cuda() for e in eps])
    for i, (a, b) in enumerate(zip(got, want)):
        assert torch.equal(a.cpu(), b), (i, float((a.cpu() - b).abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_unipc_step_kernel_16bit_io(dtype):
    """16-bit model output / sample tensors (what the reference's GPU loop hands the scheduler); fp32 history."""
    steps = 10
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 4, 8, 16, generator=g).to(dtype)
    eps = [torch.randn(2, 4, 8, 16, generator=g).to(dtype) for _ in range(steps)]
    o = OracleUniPC.from_config(SD21_SCHED)
    o.set_timesteps(steps)
    p = B200UniPCMultistepScheduler.from_config(SD21_SCHED)
    p.set_timesteps(steps)
    xo, xp = x.float(), x.cuda()
    for t, e in zip(o.timesteps, eps):
        # oracle in fp32 on the same 16-bit-rounded inputs; the product rounds only the returned sample
        xo = o.step(e.float(), t, xp.float().cpu(), return_dict=False)[0]
        xp = p.step(e.cuda(), t, xp, return_dict=False)[0]
        assert xp.dtype == dtype
        torch.testing.assert_close(xp.float().cpu(), xo, rtol=2 ** -7 if dtype == torch.bfloat16 else 2 ** -10,
                                   atol=1e-3)


@pytest.mark.gpu
def test_cfg_unipc_fused_kernel_matches_oracle():
    """The fused per-step kernel (CFG combine + UniPC + next-input rewrite, device step counter) over a whole
    trajectory against oracle CFG + oracle scheduler: bit-exact in fp32."""
    from pcdms_b200 import ops
    n, h, w, steps, guidance = 2, 8, 16, 12, 2.0
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, 4, h, w, generator=g)
    eps_rows = [torch.randn(2 * n, h, w, 32, generator=g) for _ in range(steps)]
    o = OracleUniPC.from_config(SD21_SCHED)
    o.set_timesteps(steps)
    p = B200UniPCMultistepScheduler.from_config(SD21_SCHED)
    p.set_timesteps(steps)
    state = torch.zeros(4, n, 4, h, w, device="cuda")
    state[0].copy_(x)
    x9 = torch.zeros(2 * n, h, w, 64, device="cuda", dtype=torch.bfloat16)
    coef = p.coefficient_table("cuda")
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    t_table = torch.cat([p.timesteps.float(), torch.zeros(1)]).cuda()
    t_cur = torch.zeros(1, device="cuda")
    xo = x
    for i, t in enumerate(o.timesteps):
        e = eps_rows[i][..., :4].permute(0, 3, 1, 2)
        e = e[:n] + guidance * (e[n:] - e[:n])
        xo = o.step(e, t, xo, return_dict=False)[0]
        ops.cfg_unipc_step(eps_rows[i].cuda(), state, x9, coef, counter, guidance, t_table, t_cur)
        assert torch.equal(state[0].cpu(), xo), i
        assert int(counter[0]) == i + 1 and int(counter[1]) == 0
        assert float(t_cur) == float(t_table[i + 1])
        want9 = torch.cat([xo, xo]).permute(0, 2, 3, 1).to(torch.bfloat16)
        assert torch.equal(x9[..., :4].cpu(), want9) and float(x9[..., 4:].abs().max()) == 0.0
