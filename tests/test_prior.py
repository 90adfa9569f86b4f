"""Stage-1 prior: /root/reference/src/models/stage1_prior_transformer.py,
src/pipelines/stage1_prior_pipeline.py, stage1_batchtest_prior_model.py.

CPU  * oracle/prior.py against the reference's OWN classes run unmodified over oracle/diffusers_shim: bit-equal (live
       where /root/reference exists; tests/golden/ref_prior_tiny.pt carries the reference's outputs everywhere else);
     * closed forms that pin the restated UnCLIP scheduler (diffusers itself is not installable: parity unpinned);
     * host logic of the B200 classes (weight packing, token template, step tables, pipeline loop) through
       tests/mock_ops.py, against the oracle.
GPU  * pcdm_unclip_step / pcdm_cfg_unclip_step bit-exact in fp32 against the oracle scheduler;
     * B200Stage1PriorTransformer.forward and the fused pipeline against the oracle and the reference's golden outputs,
       at the tiny config and at the real width (32 heads x 64 = 2048, reduced depth).
"""
import math
from pathlib import Path

import pytest
import torch

from oracle import reference_shim as rs
from oracle.prior import TINY, PriorTransformer, UnCLIPScheduler, make_prior, make_prior_inputs, prior_loop
from pcdms_b200.prior import B200Stage1PriorPipeline, B200Stage1PriorTransformer
from pcdms_b200.scheduler import B200UnCLIPScheduler
from tests import mock_ops

GOLD = Path(__file__).resolve().parent / "golden" / "ref_prior_tiny.pt"
needs_ref = pytest.mark.skipif(not rs.reference_available(), reason="/root/reference not present")


def _sched(cls=UnCLIPScheduler):
    return cls(prediction_type="sample", clip_sample=True, clip_sample_range=10.0)


# ------------------------------------------------------------------------------------------------------------------
# oracle vs the reference's own classes
# ------------------------------------------------------------------------------------------------------------------
@needs_ref
def test_reference_prior_forward_equals_oracle():
    o = make_prior(seed=3, **TINY)
    r = rs.build_reference_prior(**TINY)
    assert set(r.state_dict()) == set(o.state_dict()) and len(o.state_dict()) == 66
    r.load_state_dict(o.state_dict(), strict=True)
    i = make_prior_inputs(n=2, seed=5)
    x = i["latents"][:, None]
    with torch.no_grad():
        for t in (500, torch.tensor(37), torch.tensor([999, 0])):
            want = r(x, t, i["s_embed"], i["s_pose"], i["t_pose"], return_dict=False)[0]
            assert torch.equal(o(x, t, i["s_embed"], i["s_pose"], i["t_pose"]), want)
        x2, e2 = torch.cat([x, x]), torch.cat([torch.zeros_like(i["s_embed"]), i["s_embed"]])
        want = r(x2, 7, e2, i["s_pose"], i["t_pose"], test_flag=True).predicted_image_embedding
        assert torch.equal(o(x2, 7, e2, i["s_pose"], i["t_pose"], test_flag=True), want)
    assert len(r.attn_processors) == TINY["num_layers"]
    assert torch.equal(r.post_process_latents(x), o.post_process_latents(x))


@needs_ref
def test_reference_prior_pipeline_equals_oracle_loop():
    o = make_prior(seed=4, **TINY)
    r = rs.build_reference_prior(**TINY)
    r.load_state_dict(o.state_dict(), strict=True)
    i = make_prior_inputs(n=1, seed=6)
    kw = dict(s_embed=i["s_embed"], s_pose=i["s_pose"], t_pose=i["t_pose"], latents=i["latents"],
              num_inference_steps=5, guidance_scale=0)
    torch.manual_seed(11)      # the reference draws the variance noise from the global generator
    want, zero = rs.run_reference_prior_pipeline(r, zero_embed=torch.full((1, 1024), 0.25), **kw)
    torch.manual_seed(11)
    with torch.no_grad():
        got = prior_loop(o, _sched(), **kw)
    assert torch.equal(got, want) and torch.equal(zero, torch.full((1, 1024), 0.25))
    with pytest.raises(Exception):   # the reference's own CFG branch fails on its batch mismatch (documented)
        rs.run_reference_prior_pipeline(r, zero_embed=torch.zeros(1, 1024), **{**kw, "guidance_scale": 2.0})


def test_oracle_matches_reference_golden():
    gold = torch.load(GOLD)
    o = make_prior(seed=gold["seed"], **TINY)
    i = make_prior_inputs(n=1, seed=gold["input_seed"], steps=gold["steps"])
    x = i["latents"][:, None]
    with torch.no_grad():
        assert torch.equal(o(x, 500, i["s_embed"], i["s_pose"], i["t_pose"]), gold["forward"])
        x2, e2 = torch.cat([x, x]), torch.cat([torch.zeros_like(i["s_embed"]), i["s_embed"]])
        assert torch.equal(o(x2, torch.tensor(37), e2, i["s_pose"], i["t_pose"], test_flag=True),
                           gold["forward_test_flag"])
        got = prior_loop(o, _sched(), s_embed=i["s_embed"], s_pose=i["s_pose"], t_pose=i["t_pose"],
                         latents=i["latents"], num_inference_steps=gold["steps"], guidance_scale=0,
                         noises=gold["variance_noise"])
    assert torch.equal(got, gold["image_embeds"])


# ------------------------------------------------------------------------------------------------------------------
# UnCLIP scheduler: closed forms
# ------------------------------------------------------------------------------------------------------------------
def test_unclip_scheduler_closed_forms():
    s = _sched()
    T = 1000

    def abar(i):   # cumulative product telescopes to abar(i+1)/abar(0) while no beta is capped
        return math.cos(((i + 1) / T + 0.008) / 1.008 * math.pi / 2) ** 2 / math.cos(0.008 / 1.008 * math.pi / 2) ** 2
    for i in (0, 10, 500, 900):
        assert abs(float(s.alphas_cumprod[i]) - abar(i)) < 2e-6
    assert float(s.betas[-1]) == pytest.approx(0.999)               # the cap
    s.set_timesteps(25)
    ts = s.timesteps.tolist()
    assert ts[0] == 999 and ts[-1] == 0 and len(ts) == 25 and ts[1] == round(23 * 999 / 24)
    s.set_timesteps(5)
    assert s.timesteps.tolist() == [999, 749, 500, 250, 0]
    # posterior mean of q(x_prev | x_t, x0): with x_t = sqrt(a_t) x0 + sqrt(1-a_t) e the mean is
    # sqrt(a_prev) x0 + sqrt(1-a_prev) * sqrt(1 - var/(1-a_prev)) e ... checked through its two coefficients:
    for t, p in ((999, 749), (500, 250), (250, 0), (37, 36)):
        c0, c1, std, sa, sb = (float(v) for v in s.step_scalars(t, p))
        a_t, a_p = float(s.alphas_cumprod[t]), float(s.alphas_cumprod[p])
        beta = 1 - a_t / a_p
        assert c0 == pytest.approx(math.sqrt(a_p) * beta / (1 - a_t), rel=1e-3)   # fp32 `1 - ratio` cancellation
        # (abs: at t = 999 the fp32 `1 - a_t / a_prev` rounds to exactly 1, so the x_t coefficient is 0, not 1.2e-4)
        assert c1 == pytest.approx(math.sqrt(1 - beta) * (1 - a_p) / (1 - a_t), rel=1e-3, abs=2e-4)
        assert std == pytest.approx(math.sqrt((1 - a_p) / (1 - a_t) * beta), rel=1e-3)
        # consistency of the posterior: c0 + c1 * sqrt(a_t) = sqrt(a_prev)  (the x0 coefficient of the mean)
        assert c0 + c1 * math.sqrt(a_t) == pytest.approx(math.sqrt(a_p), rel=1e-3)
        assert c0 >= 0 and c1 >= 0 and std > 0
    # last step (t = 0, prev = None -> -1): alpha_prev = 1, no noise, returns the clipped x0 prediction (times
    # beta_0 / (1 - fp32(1 - beta_0)) = 0.99947: the rounding of 1 - beta_0 at beta_0 = 4e-5)
    c0, c1, std, _, _ = (float(v) for v in s.step_scalars(0, None))
    assert std == 0.0 and c1 == 0.0 and c0 == pytest.approx(1.0, rel=2e-3)
    x0 = torch.tensor([[-20.0, 0.5, 20.0]])
    assert torch.allclose(s.step(x0, 0, torch.randn(1, 3)), torch.tensor([[-10.0, 0.5, 10.0]]), rtol=2e-3)
    g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(1)
    mo, x = torch.randn(2, 8), torch.randn(2, 8)
    a = s.step(mo, 500, x, prev_timestep=250, generator=g1)
    b = s.step(mo, 500, x, prev_timestep=250, variance_noise=torch.randn(2, 8, generator=g2))
    assert torch.equal(a, b)
    e = UnCLIPScheduler(prediction_type="epsilon", clip_sample=False)     # epsilon form agrees with the sample form
    c = e.step_scalars(500, 250)
    x0p = (x - c[4] * mo) / c[3]
    nz = torch.randn(2, 8)
    assert torch.allclose(e.step(mo, 500, x, 250, variance_noise=nz),
                          UnCLIPScheduler(clip_sample=False).step(x0p, 500, x, 250, variance_noise=nz))


def test_b200_unclip_scheduler_tables_match_oracle():
    o, p = _sched(), _sched(B200UnCLIPScheduler)
    assert torch.equal(o.alphas_cumprod, p.alphas_cumprod)
    for n in (4, 20, 25):
        o.set_timesteps(n)
        p.set_timesteps(n)
        assert torch.equal(o.timesteps, p.timesteps)
        tab = p.coefficient_table("cpu")
        ts = o.timesteps.tolist()
        for i, t in enumerate(ts):
            want = o.step_scalars(t, ts[i + 1] if i + 1 < n else None)
            assert tab[i, :3].tolist() == [float(v) for v in want[:3]] and float(tab[i, 3]) == 10.0
    d = B200UnCLIPScheduler()       # diffusers' own defaults
    assert d.config.prediction_type == "epsilon" and d.config.clip_sample_range == 1.0
    assert B200UnCLIPScheduler.from_config(dict(prediction_type="sample", clip_sample_range=10.0, foo=1)
                                           ).config.clip_sample_range == 10.0
    with pytest.raises(RuntimeError):
        p.step(torch.zeros(1, 4), 500, torch.zeros(1, 4))       # no CPU path


def test_stage1_training_noise_schedule():
    """stage1_train_prior_model.py:155,287: DDPMScheduler(beta_schedule='squaredcos_cap_v2', prediction_type='sample')
    .add_noise — the B200 class carries the same cumulative-alpha table as the sampler's schedule (the per-element
    arithmetic is the schedule-independent pcdm_add_noise kernel, tests/test_kernels_gpu.py); no CPU path."""
    from pcdms_b200.scheduler import B200DDPMScheduler
    s = B200DDPMScheduler(beta_schedule="squaredcos_cap_v2", prediction_type="sample")
    assert torch.equal(s.alphas_cumprod, _sched().alphas_cumprod) and s.config.prediction_type == "sample"
    assert torch.equal(s.alphas_cumprod, _sched(B200UnCLIPScheduler).alphas_cumprod)
    with pytest.raises(RuntimeError):
        s.add_noise(torch.zeros(2, 8), torch.zeros(2, 8), torch.tensor([0, 999]))
    with pytest.raises(NotImplementedError):
        B200DDPMScheduler(beta_schedule="sigmoid")


# ------------------------------------------------------------------------------------------------------------------
# host logic of the product classes (mock kernels, CPU)
# ------------------------------------------------------------------------------------------------------------------
def _b200_prior(o, dtype, device):
    p = B200Stage1PriorTransformer(dtype=dtype, device=device, **TINY)
    assert set(p.state_dict_shapes()) == set(o.state_dict())
    for k, v in o.state_dict().items():
        assert tuple(v.shape) == p.state_dict_shapes()[k], k
    p.load_state_dict(o.state_dict())
    return p


def test_prior_host_logic_matches_oracle():
    o = make_prior(seed=3, **TINY)
    p = _b200_prior(o, torch.float32, "cpu")
    p._guard = lambda x: None
    i = make_prior_inputs(n=2, seed=5)
    x = i["latents"][:, None]
    with torch.no_grad(), mock_ops.patched():
        for t in (500, torch.tensor(37), torch.tensor([999, 0])):
            got = p(x, t, i["s_embed"], i["s_pose"], i["t_pose"]).predicted_image_embedding
            torch.testing.assert_close(got, o(x, t, i["s_embed"], i["s_pose"], i["t_pose"]), rtol=2e-4, atol=2e-5)
        x2, e2 = torch.cat([x, x]), torch.cat([torch.zeros_like(i["s_embed"]), i["s_embed"]])
        got = p(x2, 7, e2, i["s_pose"], i["t_pose"], test_flag=True, return_dict=False)[0]
        torch.testing.assert_close(got, o(x2, 7, e2, i["s_pose"], i["t_pose"], test_flag=True), rtol=2e-4, atol=2e-5)
        with pytest.raises(ValueError):
            p(x2, 7, e2, i["s_pose"][:1], i["t_pose"], test_flag=True)
        with pytest.raises(NotImplementedError):
            p(x, 7, i["s_embed"], i["s_pose"], i["t_pose"], attention_mask=torch.ones(2, 2))


@pytest.mark.parametrize("guidance,n", [(0, 1), (0, 3), (2.5, 2)])
def test_prior_pipeline_host_logic_matches_oracle(guidance, n):
    o = make_prior(seed=4, **TINY)
    p = _b200_prior(o, torch.float32, "cpu")
    p._guard = lambda x: None
    steps = 5
    i = make_prior_inputs(n=n, seed=6, steps=steps)

    class _Enc:
        config = type("c", (), {"image_size": 8})
        dtype = torch.float32

        def __call__(self, x):
            return {"image_embeds": torch.full((1, 1024), 0.5)}

    pipe = B200Stage1PriorPipeline(prior=p, image_encoder=_Enc(), scheduler=_sched(B200UnCLIPScheduler))
    pipe.use_cuda_graph = False
    kw = dict(s_embed=i["s_embed"], s_pose=i["s_pose"], t_pose=i["t_pose"], latents=i["latents"],
              num_inference_steps=steps, guidance_scale=guidance)
    with torch.no_grad(), mock_ops.patched():
        out = pipe(variance_noise=i["noises"], **kw)
        want = prior_loop(o, _sched(), noises=i["noises"], **kw)
    torch.testing.assert_close(torch.as_tensor(out[0]), want, rtol=2e-4, atol=2e-5)
    assert out.negative_image_embeds.shape == (n, 1024) and torch.equal(out["image_embeds"], out[0])


def test_prior_surface_and_errors():
    p = B200Stage1PriorTransformer(device="cpu", num_embeddings=2, embedding_dim=1024)   # the driver's overrides (:57)
    assert p.inner_dim == 2048 and p.config.num_layers == 20 and p.config.embedding_dim == 1024
    n = sum(torch.Size(s).numel() for s in p.state_dict_shapes().values())
    with torch.device("meta"):
        full = PriorTransformer(num_embeddings=2, embedding_dim=1024)
    assert n == sum(v.numel() for v in full.state_dict().values()) == 1_027_166_208
    assert {k: tuple(v.shape) for k, v in full.state_dict().items()} == p.state_dict_shapes()
    assert len(p.attn_processors) == 20
    p.set_default_attn_processor()
    with pytest.raises(ValueError):
        p.set_attn_processor({"x": None})
    with pytest.raises(RuntimeError):
        p(torch.zeros(1, 1, 1024), 1, torch.zeros(1, 1, 1024), torch.zeros(1, 1, 36), torch.zeros(1, 1, 36))
    with pytest.raises(NotImplementedError):
        B200Stage1PriorTransformer(device="cpu")                     # embedding_dim 768: the pose MLP emits 1024
    with pytest.raises(NotImplementedError):
        B200Stage1PriorTransformer(device="cpu", embedding_dim=1024, num_embeddings=77)
    pipe = B200Stage1PriorPipeline.from_pretrained("/nonexistent")
    assert pipe.scheduler.config.prediction_type == "sample" and pipe.scheduler.config.clip_sample_range == 10.0
    with pytest.raises(ValueError):
        pipe(torch.zeros(1, 1, 1024), torch.zeros(1, 1, 36), torch.zeros(1, 1, 36))     # no prior assigned
    pipe.enable_xformers_memory_efficient_attention()


# ------------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------------
def _rel(got, want):
    return float((got.float().cpu() - want).abs().max() / want.abs().max())


@pytest.mark.gpu
def test_unclip_step_kernels_bit_exact_gpu():
    from pcdms_b200 import ops
    o, p = _sched(), _sched(B200UnCLIPScheduler)
    steps, n, E = 7, 3, 1024
    o.set_timesteps(steps)
    p.set_timesteps(steps, device="cuda")
    g = torch.Generator().manual_seed(2)
    x = torch.randn(n, E, generator=g)
    xs = x.clone()
    noises = torch.randn(steps, n, E, generator=g)
    preds = 4.0 * torch.randn(steps, 2 * n, E, generator=g)             # some values beyond the clip range of 10
    preds[:, :, :5] *= 10
    # (a) protocol step
    ts = o.timesteps.tolist()
    xd = x.cuda()
    for i, t in enumerate(ts):
        prev = ts[i + 1] if i + 1 < steps else None
        want = o.step(preds[i, :n], t, xs, prev_timestep=prev, variance_noise=noises[i])
        row = p.step_coefficients(t, prev)
        got = ops.unclip_step(preds[i, :n].cuda().contiguous(), xd, noises[i].cuda() if row[2] else None, row)
        assert torch.equal(got.cpu(), want), i
        xs, xd = want, got
    # (b) fused CFG step replayed over the whole trajectory
    lat = x.cuda().clone()
    xin = torch.zeros(2 * n, E, device="cuda", dtype=torch.float16)
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    t_table = torch.cat([p.timesteps.float().cuda(), torch.zeros(1, device="cuda")])
    t_cur = torch.zeros(1, device="cuda")
    coef = p.coefficient_table("cuda")
    xs = x.clone()
    for i, t in enumerate(ts):
        u, c = preds[i, :n], preds[i, n:]
        mo = u + 2.5 * (c - u)
        xs = o.step(mo, t, xs, prev_timestep=ts[i + 1] if i + 1 < steps else None, variance_noise=noises[i])
        ops.cfg_unclip_step(preds[i].cuda().contiguous(), lat, xin, coef, noises.cuda(), counter, 2.5, True, t_table,
                            t_cur)
        assert torch.equal(lat.cpu(), xs), i
        assert torch.equal(xin.cpu(), torch.cat([xs, xs]).half())
        assert int(counter[0]) == i + 1 and float(t_cur) == float(t_table[i + 1])


@pytest.mark.gpu
@pytest.mark.parametrize("dt,tol", [(torch.float16, 5e-3), (torch.bfloat16, 4e-2)])
def test_prior_forward_gpu(dt, tol):
    gold = torch.load(GOLD)
    o = make_prior(seed=gold["seed"], **TINY)
    p = _b200_prior(o, dt, "cuda")
    i = make_prior_inputs(n=1, seed=gold["input_seed"], steps=gold["steps"])
    x = i["latents"][:, None]
    got = p(x.cuda(), 500, i["s_embed"].cuda(), i["s_pose"].cuda(), i["t_pose"].cuda()).predicted_image_embedding
    assert got.dtype == dt and _rel(got, gold["forward"]) < tol          # the reference's own output
    x2, e2 = torch.cat([x, x]), torch.cat([torch.zeros_like(i["s_embed"]), i["s_embed"]])
    got = p(x2.cuda(), torch.tensor(37), e2.cuda(), i["s_pose"].cuda(), i["t_pose"].cuda(), test_flag=True,
            return_dict=False)[0]
    assert _rel(got, gold["forward_test_flag"]) < tol


class _ZeroEncoder:
    config = type("c", (), {"image_size": 8})
    dtype = torch.float16

    def __call__(self, x):
        return {"image_embeds": torch.zeros(1, 1024, device=x.device)}


@pytest.mark.gpu
def test_prior_pipeline_gpu_against_reference_golden():
    gold = torch.load(GOLD)
    o = make_prior(seed=gold["seed"], **TINY)
    p = _b200_prior(o, torch.float16, "cuda")
    i = make_prior_inputs(n=1, seed=gold["input_seed"], steps=gold["steps"])
    pipe = B200Stage1PriorPipeline(prior=p, image_encoder=_ZeroEncoder(), scheduler=_sched(B200UnCLIPScheduler))
    kw = dict(s_embed=i["s_embed"].cuda(), s_pose=i["s_pose"].cuda(), t_pose=i["t_pose"].cuda(),
              latents=i["latents"].cuda(), num_inference_steps=gold["steps"], guidance_scale=0,
              variance_noise=gold["variance_noise"].cuda())
    out = pipe(**kw)
    assert _rel(out[0], gold["image_embeds"]) < 5e-3                     # Stage1_PriorPipeline.__call__'s own output
    pipe.use_cuda_graph = False
    eager = pipe(**kw)
    assert torch.equal(out[0], eager[0])                                 # graph replay == eager launches
    assert torch.equal(out[0], pipe(**kw)[0])


@pytest.mark.gpu
@pytest.mark.parametrize("guidance,n", [(0, 2), (2.5, 2)])
def test_prior_pipeline_full_width_gpu(guidance, n):
    """The real layer shape (32 heads x 64 = 2048 wide, FF 8192) at reduced depth, 25 steps (the pipeline default)."""
    cfg = dict(TINY, num_attention_heads=32, num_layers=3)
    o = make_prior(seed=8, **cfg)
    p = B200Stage1PriorTransformer(dtype=torch.float16, **cfg)
    p.load_state_dict(o.state_dict())
    steps = 25
    i = make_prior_inputs(n=n, seed=9, steps=steps)
    with torch.no_grad():
        want = prior_loop(o, _sched(), s_embed=i["s_embed"], s_pose=i["s_pose"], t_pose=i["t_pose"],
                          latents=i["latents"], num_inference_steps=steps, guidance_scale=guidance, noises=i["noises"])
    pipe = B200Stage1PriorPipeline(prior=p, image_encoder=_ZeroEncoder(), scheduler=_sched(B200UnCLIPScheduler))
    out = pipe(s_embed=i["s_embed"].cuda(), s_pose=i["s_pose"].cuda(), t_pose=i["t_pose"].cuda(),
               latents=i["latents"].cuda(), num_inference_steps=steps, guidance_scale=guidance,
               variance_noise=i["noises"].cuda())
    assert out.image_embeds.shape == (n, 1024)
    assert _rel(out.image_embeds, want) < 1e-2
    gen = torch.Generator(device="cuda").manual_seed(3)                  # generator-driven noise: reproducible
    a = pipe(s_embed=i["s_embed"].cuda(), s_pose=i["s_pose"].cuda(), t_pose=i["t_pose"].cuda(),
             num_inference_steps=steps, guidance_scale=guidance, generator=gen)
    gen.manual_seed(3)
    b = pipe(s_embed=i["s_embed"].cuda(), s_pose=i["s_pose"].cuda(), t_pose=i["t_pose"].cuda(),
             num_inference_steps=steps, guidance_scale=guidance, generator=gen)
    assert torch.equal(a[0], b[0])
