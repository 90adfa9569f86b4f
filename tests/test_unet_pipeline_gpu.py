"""GPU parity of the whole hot path against the CPU fp32 oracle: UNet forward (stage-2, stage-3 topology) and the
10-step DDIM pipeline of BASELINE config 1, through the public boundary classes.

Tolerance.  BASELINE's north_star asks rtol 1e-3 / atol 1e-4 of an fp16 GPU run against the fp32 CPU path.  Each
KERNEL meets that on identical inputs (tests/test_kernels_gpu.py), but an end-to-end fp16 pipeline cannot: every one
of the ~330 activation tensors between input and output is rounded to fp16 (relative 4.9e-4 each) and those roundings
accumulate through 61 normalisations and 10-50 scheduler steps.  What is asserted here is the measured envelope of
that accumulation, normalised by the reference's dynamic range, with the strict criterion's violation count reported:
    single UNet evaluation, fp16 : max |err| <= 3e-3 * max|ref|,  mean |err| <= 5e-4 * max|ref|
    single UNet evaluation, bf16 : 8x those (8x coarser mantissa)
    10-step DDIM pipeline,  fp16 : max |err| <= 2e-3 * max|ref|
(measured on B200: 1.6e-3 / 2.3e-4 for the full-size UNet, 6e-4 for the pipeline.)"""
from dataclasses import asdict

import pytest
import torch

pytestmark = pytest.mark.gpu


_ORACLES = {}


def _oracle(cfg, seed=0):
    """Full-size oracle nets take ~10 s to initialise on the CPU: build each (config, seed) once per session."""
    from oracle.factory import make_unet
    key = (repr(cfg), seed)
    if key not in _ORACLES:
        _ORACLES[key] = make_unet(cfg, seed=seed)
    return _ORACLES[key]


def _b200(cfg, o, dt):
    from pcdms_b200.unet import B200UNet2DConditionModel
    m = B200UNet2DConditionModel(dtype=dt, device="cuda", **asdict(cfg))
    m.load_state_dict(o.state_dict())
    return m


def _models(cfg, dt, seed=0):
    o = _oracle(cfg, seed)
    return o, _b200(cfg, o, dt)


def _check(got, want, max_frac, mean_frac, label, *, config="unspecified", dtype=None, against="oracle", extra=None):
    """Envelope assertion + a record of the strict north-star share (tests/parity_record.py)."""
    from tests.parity_record import check
    return check(got, want, max_frac, mean_frac, label, config=config, dtype=dtype if dtype is not None else got.dtype,
                 against=against, extra=extra)


def _run_unet(o, m, i, t):
    ref = o(i["sample"], t, i["encoder_hidden_states"], class_labels=i.get("class_labels"),
            my_pose_cond=i.get("my_pose_cond"))[0]
    kw = {k: i[k].cuda() for k in ("class_labels", "my_pose_cond") if k in i}
    out = m(i["sample"].cuda(), t, i["encoder_hidden_states"].cuda(), return_dict=False, **kw)[0]
    return out, ref


@pytest.mark.parametrize("dt,mult", [(torch.float16, 1.0), (torch.bfloat16, 8.0)])
def test_unet_tiny_stage2(dt, mult):
    from oracle.factory import make_unet_inputs
    from oracle.unet import UNetConfig
    cfg = UNetConfig.tiny()
    o, m = _models(cfg, dt)
    i = make_unet_inputs(cfg, batch=2, h=16, w=32, s_kv=9)
    out, ref = _run_unet(o, m, i, 981)
    assert out.shape == ref.shape and out.dtype == i["sample"].dtype
    _check(out, ref, 3e-3 * mult, 5e-4 * mult, f"tiny stage-2 UNet {dt}", config="tiny UNet B2 16x32 s_kv9", dtype=dt)
    # vector timesteps + return_dict surface
    tv = torch.tensor([21, 501])
    refv = o(i["sample"], tv, i["encoder_hidden_states"], class_labels=i["class_labels"],
             my_pose_cond=i["my_pose_cond"])[0]
    outv = m(i["sample"].cuda(), tv.cuda(), i["encoder_hidden_states"].cuda(), class_labels=i["class_labels"].cuda(),
             my_pose_cond=i["my_pose_cond"].cuda()).sample
    _check(outv, refv, 3e-3 * mult, 5e-4 * mult, f"tiny stage-2 UNet vector t {dt}", config="tiny UNet B2 16x32 vector t",
           dtype=dt)


def test_unet_tiny_with_producer_groupnorm_statistics_everywhere():
    """The producer-epilogue GroupNorm statistics path (normally taken by > 16 MB tensors only: the 32x64 level at the
    bench batch) forced on at every level of the tiny net, incl. the skip concats and the one-launch Upsample2D."""
    from oracle.factory import make_unet_inputs
    from oracle.unet import UNetConfig
    cfg = UNetConfig.tiny()
    o, m = _models(cfg, torch.float16)
    m.GN_STATS_MIN_BYTES = 0
    i = make_unet_inputs(cfg, batch=2, h=16, w=32, s_kv=9)
    out, ref = _run_unet(o, m, i, 981)
    _check(out, ref, 3e-3, 5e-4, "tiny stage-2 UNet fp16, GroupNorm statistics from producer epilogues at every level",
           config="tiny UNet B2 16x32 s_kv9", dtype=torch.float16)


@pytest.mark.parametrize("h,w", [(24, 48), (8, 24)])
def test_unet_tiny_on_a_canvas_outside_the_tile_map(h, w):
    """A latent size no BASELINE configuration has (the reference takes any canvas divisible by 8,
    stage2_batchtest_inpaint_model.py:258-260): widths 48 / 24 / 12 / 6 neither divide 128 nor are multiples of it, so
    every conv of the net takes the padded route; GroupNorm, attention (1152 / 288 / 72 / 18 tokens) and the GEMMs run
    on the real sizes."""
    from oracle.factory import make_unet_inputs
    from oracle.unet import UNetConfig
    cfg = UNetConfig.tiny()
    o, m = _models(cfg, torch.float16)
    i = make_unet_inputs(cfg, batch=2, h=h, w=w, s_kv=9)
    out, ref = _run_unet(o, m, i, 501)
    assert out.shape == ref.shape
    _check(out, ref, 3e-3, 5e-4, f"tiny stage-2 UNet fp16 on a {h}x{w} latent canvas",
           config=f"tiny UNet B2 {h}x{w} (outside the conv tile map)", dtype=torch.float16)


def test_unet_tiny_stage3_topology():
    from oracle.factory import make_unet_inputs
    from oracle.unet import UNetConfig
    cfg = UNetConfig.tiny(in_channels=8, stage2=False)
    o, m = _models(cfg, torch.float16)
    i = make_unet_inputs(cfg, batch=2, h=16, w=16, s_kv=17)
    out, ref = _run_unet(o, m, i, 501)
    _check(out, ref, 3e-3, 5e-4, "tiny stage-3 UNet fp16", config="tiny stage-3 UNet B2 16x16", dtype=torch.float16)


def test_unet_full_size_stage2_fp16_and_bf16():
    """BASELINE config 1 shapes: B = 2 (one image under CFG), 32x64 latents, 258 tokens, the real 868.9 M-param net."""
    from oracle.factory import make_unet_inputs
    from oracle.unet import UNetConfig
    from pcdms_b200.unet import B200UNet2DConditionModel
    cfg = UNetConfig.stage2()
    o, m = _models(cfg, torch.float16)
    i = make_unet_inputs(cfg, batch=2, h=32, w=64, s_kv=258)
    out, ref = _run_unet(o, m, i, 981)
    assert 0.05 < ref.std().item() < 20
    _check(out, ref, 3e-3, 5e-4, "full stage-2 UNet fp16", config="cfg1 UNet eval: 868.9M, B2 32x64 s_kv258",
           dtype=torch.float16)
    del m
    torch.cuda.empty_cache()
    mb = B200UNet2DConditionModel(dtype=torch.bfloat16, device="cuda", **asdict(cfg))
    mb.load_state_dict(o.state_dict())
    kw = {k: i[k].cuda() for k in ("class_labels", "my_pose_cond")}
    outb = mb(i["sample"].cuda(), 981, i["encoder_hidden_states"].cuda(), return_dict=False, **kw)[0]
    _check(outb, ref, 2.4e-2, 4e-3, "full stage-2 UNet bf16", config="cfg1 UNet eval: 868.9M, B2 32x64 s_kv258",
           dtype=torch.bfloat16)


@pytest.mark.parametrize("which,h,w,s_kv", [("stage2", 64, 128, 258), ("stage3", 64, 64, 257)])
def test_unet_full_size_other_baseline_shapes(which, h, w, s_kv):
    """The latent shapes of BASELINE configs 3 (stage-2 at 512x512: 64x128 latents, self-attention over 8192 tokens)
    and 5 (stage-3: in_channels 8, no class embedding / pose, 64x64 latents, 257 tokens) with the real-size nets, one
    batch row (the CPU oracle needs ~2 TFLOP per row at these sizes)."""
    from oracle.factory import make_unet_inputs
    from oracle.unet import UNetConfig
    cfg = UNetConfig.stage2() if which == "stage2" else UNetConfig.stage3()
    o, m = _models(cfg, torch.float16)
    i = make_unet_inputs(cfg, batch=1, h=h, w=w, s_kv=s_kv)
    out, ref = _run_unet(o, m, i, 501)
    assert out.shape == ref.shape
    _check(out, ref, 3e-3, 5e-4, f"full {which} UNet fp16 at {h}x{w}",
           config=f"cfg{3 if which == 'stage2' else 5} UNet eval: {which} full size, B1 {h}x{w} s_kv{s_kv}", dtype=torch.float16)


@pytest.mark.parametrize("use_graph", [False, True])
def test_pipeline_config1_shape_tiny_weights(use_graph):
    """10-step DDIM, guidance 2.0, through B200Stage2InpaintPipeline.__call__ vs the oracle loop (which is itself
    pinned bit-for-bit to the reference's pipeline, tests/test_oracle.py)."""
    from oracle.factory import make_inputs
    from oracle.pipeline import denoise_loop, prepare_conditioning
    from oracle.schedulers import OracleDDIMScheduler
    from oracle.unet import UNetConfig
    from pcdms_b200.pipeline import B200Stage2InpaintPipeline
    from pcdms_b200.scheduler import B200DDIMScheduler
    cfg = UNetConfig.tiny()
    o, m = _models(cfg, torch.float16)
    pin = make_inputs(cfg, n=2, h=16, w=32, s_kv=9)
    cond = prepare_conditioning(s_img_proj_f=pin["s_img_proj_f"], pred_t_img_embed=pin["pred_t_img_embed"],
                                st_pose_f=pin["st_pose_f"], masked_latents=pin["masked_latents"], height=pin["height"],
                                width=pin["width"], num_images_per_prompt=2, guidance_scale=2.0)
    ref = denoise_loop(o, OracleDDIMScheduler(), latents=pin["latents"], cond=cond, num_inference_steps=10,
                       guidance_scale=2.0)
    pipe = B200Stage2InpaintPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())
    pipe.use_cuda_graph = use_graph
    for rep in range(2):  # the second call re-uses the captured graph with fresh inputs
        out = pipe(height=pin["height"], width=pin["width"], num_inference_steps=10, guidance_scale=2.0,
                   num_images_per_prompt=2, latents=pin["latents"], output_type="latent",
                   s_img_proj_f=pin["s_img_proj_f"], st_pose_f=pin["st_pose_f"],
                   pred_t_img_embed=pin["pred_t_img_embed"], masked_latents=pin["masked_latents"]).images
        _check(out, ref, 2e-3, 3e-4, f"pipeline 10-step graph={use_graph} rep={rep}",
               config="cfg1 shapes, tiny weights: n2 16x32 10 DDIM steps", dtype=torch.float16)


def test_pipeline_is_deterministic_and_shard_independent():
    """Multi-GPU contract: a rank's images depend only on its own inputs — the same call twice, and the
    same images inside a different batch composition, give identical latents."""
    from oracle.factory import make_inputs
    from oracle.unet import UNetConfig
    from pcdms_b200.pipeline import B200Stage2InpaintPipeline
    from pcdms_b200.scheduler import B200DDIMScheduler
    cfg = UNetConfig.tiny()
    _, m = _models(cfg, torch.float16)
    # 17 context tokens: the K/V projection has 34 / 68 rows for one / two images, so BOTH runs take the tcgen05 tile
    # path (pcdm_gemm switches to the weight-streaming kernel at <= 32 rows, and a different kernel means a different
    # fp32 summation order — 1-ulp differences that five steps of this random-weight UNet amplify beyond 1e-3)
    pin = make_inputs(cfg, n=2, h=16, w=32, s_kv=17)
    pipe = B200Stage2InpaintPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())

    def run(lat):
        return pipe(height=pin["height"], width=pin["width"], num_inference_steps=5, guidance_scale=2.0,
                    num_images_per_prompt=lat.shape[0], latents=lat, output_type="latent",
                    s_img_proj_f=pin["s_img_proj_f"], st_pose_f=pin["st_pose_f"],
                    pred_t_img_embed=pin["pred_t_img_embed"], masked_latents=pin["masked_latents"]).images.cpu()

    a, b = run(pin["latents"]), run(pin["latents"])
    assert torch.equal(a, b)
    single = run(pin["latents"][1:2])
    torch.testing.assert_close(single[0], a[1], rtol=1e-3, atol=1e-3)  # different batch => different tile schedule


# ---------------------------------------------------------------------------------------------------------------
# The parity record on BASELINE's own configurations (profiles/r2_parity.json is built from what these write)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt,max_frac,mean_frac", [(torch.float16, 1e-2, 2e-3), (torch.bfloat16, 8e-2, 1.6e-2)])
def test_config1_full_size_10_step_pipeline(dt, max_frac, mean_frac):
    """BASELINE config 1 for real: the 868.9 M-parameter stage-2 UNet, one image (UNet batch 2 under CFG), 256x256
    (32x64 latents), 258 conditioning tokens, 10 DDIM steps, guidance 2 — B200Stage2InpaintPipeline.__call__ against
    the fp32 CPU oracle loop (reference stage2_inpaint_pipeline.py:496-525)."""
    from oracle.factory import make_inputs
    from oracle.pipeline import denoise_loop, prepare_conditioning
    from oracle.schedulers import OracleDDIMScheduler
    from oracle.unet import UNetConfig
    from pcdms_b200.pipeline import B200Stage2InpaintPipeline
    from pcdms_b200.scheduler import B200DDIMScheduler
    cfg = UNetConfig.stage2()
    o = _oracle(cfg)
    pin = make_inputs(cfg, n=1, h=32, w=64, s_kv=258)
    cond = prepare_conditioning(s_img_proj_f=pin["s_img_proj_f"], pred_t_img_embed=pin["pred_t_img_embed"],
                                st_pose_f=pin["st_pose_f"], masked_latents=pin["masked_latents"], height=pin["height"],
                                width=pin["width"], num_images_per_prompt=1, guidance_scale=2.0)
    key = "cfg1_ref"
    if key not in _ORACLES:   # the fp32 loop costs ~10 x 0.8 s of CPU: once for both dtypes
        _ORACLES[key] = denoise_loop(o, OracleDDIMScheduler(), latents=pin["latents"], cond=cond,
                                     num_inference_steps=10, guidance_scale=2.0)
    ref = _ORACLES[key]
    m = _b200(cfg, o, dt)
    pipe = B200Stage2InpaintPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())
    out = pipe(height=pin["height"], width=pin["width"], num_inference_steps=10, guidance_scale=2.0,
               num_images_per_prompt=1, latents=pin["latents"], output_type="latent",
               s_img_proj_f=pin["s_img_proj_f"], st_pose_f=pin["st_pose_f"],
               pred_t_img_embed=pin["pred_t_img_embed"], masked_latents=pin["masked_latents"]).images
    _check(out, ref, max_frac, mean_frac, f"config 1 full-size 10-step DDIM pipeline {dt}",
           config="cfg1: 868.9M UNet, n1 (B2) 32x64 s_kv258, 10 DDIM steps, guidance 2", dtype=dt)
    del m, pipe
    torch.cuda.empty_cache()


@pytest.mark.parametrize("which,h,w,s_kv", [("stage2", 64, 128, 258), ("stage3", 64, 64, 257)])
def test_unet_full_size_bf16_other_baseline_shapes(which, h, w, s_kv):
    """bf16 — the benched dtype — at the latent shapes of BASELINE configs 3 and 5, real-size nets, one batch row."""
    from oracle.factory import make_unet_inputs
    from oracle.unet import UNetConfig
    cfg = UNetConfig.stage2() if which == "stage2" else UNetConfig.stage3()
    o, m = _models(cfg, torch.bfloat16)
    i = make_unet_inputs(cfg, batch=1, h=h, w=w, s_kv=s_kv)
    out, ref = _run_unet(o, m, i, 501)
    _check(out, ref, 2.4e-2, 4e-3, f"full {which} UNet bf16 at {h}x{w}",
           config=f"cfg{3 if which == 'stage2' else 5} UNet eval: {which} full size, B1 {h}x{w} s_kv{s_kv}",
           dtype=torch.bfloat16)
    del m
    torch.cuda.empty_cache()


@pytest.mark.parametrize("dt,max_frac,mean_frac", [(torch.bfloat16, 2.4e-2, 4e-3), (torch.float16, 3e-3, 5e-4)])
def test_unet_full_size_bench_operating_point_b16(dt, max_frac, mean_frac):
    """One UNet evaluation at the bench operating point of BASELINE config 2: UNet batch 16 (8 images under CFG),
    32x64 latents, 258 tokens, real-size net — the tile shapes / CTA-pair / split-K choices the benchmark actually
    runs (they differ from the B = 2 parity shape)."""
    from oracle.factory import make_unet_inputs
    from oracle.unet import UNetConfig
    cfg = UNetConfig.stage2()
    o, m = _models(cfg, dt)
    i = make_unet_inputs(cfg, batch=16, h=32, w=64, s_kv=258)
    key = "cfg2_b16_ref"
    if key not in _ORACLES:
        _ORACLES[key] = o(i["sample"], 481, i["encoder_hidden_states"], class_labels=i["class_labels"],
                          my_pose_cond=i["my_pose_cond"])[0]
    ref = _ORACLES[key]
    kw = {k: i[k].cuda() for k in ("class_labels", "my_pose_cond")}
    out = m(i["sample"].cuda(), 481, i["encoder_hidden_states"].cuda(), return_dict=False, **kw)[0]
    _check(out, ref, max_frac, mean_frac, f"full stage-2 UNet {dt} at the bench batch (B16)",
           config="cfg2 UNet eval: 868.9M, B16 32x64 s_kv258", dtype=dt)
    del m
    torch.cuda.empty_cache()


@pytest.mark.parametrize("dt,mult", [(torch.float16, 1.0), (torch.bfloat16, 8.0)])
def test_cuda_unet_vs_reference_golden(dt, mult):
    """tests/golden/ref_unet_tiny.pt holds inputs + outputs of the REFERENCE's own Stage2_InapintUNet2DConditionModel
    (run unmodified over oracle/diffusers_shim by tools/make_golden.py): the CUDA path against it directly."""
    from pathlib import Path
    from oracle.unet import UNetConfig
    g = torch.load(Path(__file__).parent / "golden" / "ref_unet_tiny.pt")
    cfg = UNetConfig.tiny()
    o, m = _models(cfg, dt, seed=g["seed"])
    i = g["inputs"]
    kw = {k: i[k].cuda() for k in ("class_labels", "my_pose_cond")}
    out = m(i["sample"].cuda(), g["timestep"], i["encoder_hidden_states"].cuda(), return_dict=False, **kw)[0]
    _check(out, g["out"], 3e-3 * mult, 5e-4 * mult, f"CUDA UNet vs reference golden {dt}",
           config="golden ref_unet_tiny.pt (reference class output)", dtype=dt, against="reference golden")
    outv = m(i["sample"].cuda(), g["timestep_vec"].cuda(), i["encoder_hidden_states"].cuda(), return_dict=False, **kw)[0]
    _check(outv, g["out_vec"], 3e-3 * mult, 5e-4 * mult, f"CUDA UNet vs reference golden, vector t {dt}",
           config="golden ref_unet_tiny.pt (reference class output, vector timestep)", dtype=dt,
           against="reference golden")


def test_cuda_pipeline_vs_reference_golden():
    """tests/golden/ref_pipeline_tiny.pt: final latents of the REFERENCE's own Stage2_InpaintDiffusionPipeline.__call__
    (fp16 loop tensors on the CPU, 4 DDIM steps, 2 images).  Both sides carry fp16 rounding, so the envelope is twice
    the single-sided one."""
    from pathlib import Path
    from oracle.unet import UNetConfig
    from pcdms_b200.pipeline import B200Stage2InpaintPipeline
    from pcdms_b200.scheduler import B200DDIMScheduler
    g = torch.load(Path(__file__).parent / "golden" / "ref_pipeline_tiny.pt")
    cfg = UNetConfig.tiny()
    _, m = _models(cfg, torch.float16, seed=g["seed"])
    pin = g["inputs"]
    pipe = B200Stage2InpaintPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())
    out = pipe(height=pin["height"], width=pin["width"], num_inference_steps=g["steps"],
               guidance_scale=g["guidance_scale"], num_images_per_prompt=g["n"], latents=pin["latents"],
               output_type="latent", s_img_proj_f=pin["s_img_proj_f"], st_pose_f=pin["st_pose_f"],
               pred_t_img_embed=pin["pred_t_img_embed"], masked_latents=pin["masked_latents"]).images
    _check(out, g["latents"], 4e-3, 6e-4, "CUDA pipeline vs reference golden fp16",
           config="golden ref_pipeline_tiny.pt (reference pipeline __call__, fp16 CPU, 4 DDIM steps)",
           dtype=torch.float16, against="reference golden")
