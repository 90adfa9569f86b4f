"""CPU tests of the HOST side of the product: weight packing, topology walk, skip bookkeeping, conditioning layout and
step tables — with the CUDA kernels replaced by the torch stand-ins of tests/mock_ops.py (each follows the documented
semantics of its C-ABI entry point).  What these tests prove is that *given correct kernels* the orchestration in
pcdms_b200/{unet,pipeline,scheduler}.py reproduces the oracle; the kernels themselves are checked on the GPU
(tests/test_*_gpu.py)."""
from dataclasses import asdict

import pytest
import torch

from oracle.factory import make_inputs, make_unet, make_unet_inputs
from oracle.pipeline import denoise_loop, prepare_conditioning
from oracle.schedulers import OracleDDIMScheduler
from oracle.unet import UNetConfig
from pcdms_b200.pipeline import B200Stage2InpaintPipeline
from pcdms_b200.scheduler import B200DDIMScheduler
from pcdms_b200.unet import B200AttnProcessor, B200UNet2DConditionModel
from tests import mock_ops


def _models(cfg, dtype=torch.float32, seed=0):
    o = make_unet(cfg, seed=seed)
    m = B200UNet2DConditionModel(dtype=dtype, device="cpu", **asdict(cfg))
    m.load_state_dict(o.state_dict())
    return o, m


def _forward(m, cfg, i, t):
    x = mock_ops.nchw_to_nhwc_pad(i["sample"], 64, m.dtype)
    kv = m.context_kv(i["encoder_hidden_states"])
    pose = mock_ops.nchw_to_nhwc_pad(i["my_pose_cond"], i["my_pose_cond"].shape[1], m.dtype) if cfg.use_pose_cond else None
    rows = m.forward_nhwc(x, torch.tensor([float(t)]), kv, i.get("class_labels"), pose)
    return mock_ops.nhwc_to_nchw(rows, cfg.out_channels, torch.float32)


@pytest.mark.parametrize("stage2,gn_stats_everywhere", [(True, False), (False, False), (True, True), (False, True)])
def test_unet_orchestration_matches_oracle(stage2, gn_stats_everywhere):
    cfg = UNetConfig.tiny() if stage2 else UNetConfig.tiny(in_channels=8, stage2=False)
    o, m = _models(cfg)
    if gn_stats_everywhere:   # producer-epilogue GroupNorm statistics are normally reserved for > 16 MB tensors
        m.GN_STATS_MIN_BYTES = 0
    i = make_unet_inputs(cfg, batch=2, h=16, w=32, s_kv=9)
    want = o(i["sample"], 981, i["encoder_hidden_states"], class_labels=i.get("class_labels"),
             my_pose_cond=i.get("my_pose_cond"))[0]
    with mock_ops.patched():
        got = _forward(m, cfg, i, 981)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)


def test_unet_surface():
    cfg = UNetConfig.tiny()
    _, m = _models(cfg)
    assert m.config.in_channels == 9 and m.config.class_embed_type == "projection"
    assert m.config._diffusers_version and m.config.sample_size == 32
    assert len(m.attn_processors) == 32
    assert all(k.endswith(".processor") for k in m.attn_processors)
    m.set_default_attn_processor()
    m.set_attn_processor({k: B200AttnProcessor() for k in m.attn_processors})
    with pytest.raises(ValueError):
        m.set_attn_processor({"x": B200AttnProcessor()})
    m.enable_xformers_memory_efficient_attention()
    assert m.dtype == torch.float32 and m.device.type == "cpu"
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 9, 8, 16), 1, torch.zeros(1, 3, 128), attention_mask=torch.ones(1, 3))
    with pytest.raises(RuntimeError):  # no CPU compute path
        m(torch.zeros(1, 9, 8, 16), 1, torch.zeros(1, 3, 128), class_labels=torch.zeros(1, 1, 128),
          my_pose_cond=torch.zeros(1, 64, 8, 16))
    sd = make_unet(cfg).state_dict()
    sd.pop("conv_out.bias")
    with pytest.raises(RuntimeError):
        B200UNet2DConditionModel(dtype=torch.float16, device="cpu", **asdict(cfg)).load_state_dict(sd)
    with pytest.raises(NotImplementedError):
        B200UNet2DConditionModel(device="cpu", use_linear_projection=False)


def test_full_size_topology_keys():
    for cfg, n in ((UNetConfig.stage2(), 690), (UNetConfig.stage3(), 686)):
        m = B200UNet2DConditionModel(device="cpu", **asdict(cfg))
        assert len(m.expected_keys()) == n
        assert m._temb_total == 20160
        assert [c for _, _, c in m._resnets].count(320) == 5  # 2 down + 3 up at the 320 level


def test_custom_attention_processor_is_called():
    cfg = UNetConfig.tiny()
    o, m = _models(cfg)
    calls = []

    class Spy:
        def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
            calls.append(encoder_hidden_states is not None)
            return attn.fused_forward(hidden_states, encoder_hidden_states)

    m.set_attn_processor(Spy())
    i = make_unet_inputs(cfg, batch=1, h=8, w=16, s_kv=5)
    want = o(i["sample"], 5, i["encoder_hidden_states"], class_labels=i["class_labels"],
             my_pose_cond=i["my_pose_cond"])[0]
    with mock_ops.patched():
        got = _forward(m, cfg, i, 5)
    assert len(calls) == 32 and sum(calls) == 16
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)


def test_pipeline_orchestration_matches_oracle_loop():
    cfg = UNetConfig.tiny()
    o, m = _models(cfg)
    pin = make_inputs(cfg, n=2, h=8, w=16, s_kv=6)
    cond = prepare_conditioning(s_img_proj_f=pin["s_img_proj_f"], pred_t_img_embed=pin["pred_t_img_embed"],
                                st_pose_f=pin["st_pose_f"], masked_latents=pin["masked_latents"], height=pin["height"],
                                width=pin["width"], num_images_per_prompt=2, guidance_scale=2.0)
    want = denoise_loop(o, OracleDDIMScheduler(), latents=pin["latents"], cond=cond, num_inference_steps=5,
                        guidance_scale=2.0)
    pipe = B200Stage2InpaintPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())
    pipe.use_cuda_graph = False
    with mock_ops.patched():
        for _ in range(2):  # second call re-uses the cached static buffers
            got = pipe(height=pin["height"], width=pin["width"], num_inference_steps=5, guidance_scale=2.0,
                       num_images_per_prompt=2, latents=pin["latents"], output_type="latent",
                       s_img_proj_f=pin["s_img_proj_f"], st_pose_f=pin["st_pose_f"],
                       pred_t_img_embed=pin["pred_t_img_embed"], masked_latents=pin["masked_latents"]).images
            torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-5)
    with pytest.raises(ValueError):
        pipe(height=65, width=128, s_img_proj_f=pin["s_img_proj_f"], guidance_scale=2.0)
    with pytest.raises(NotImplementedError):
        pipe(height=64, width=128, s_img_proj_f=pin["s_img_proj_f"], guidance_scale=1.0)


def test_scheduler_protocol_and_tables():
    s, o = B200DDIMScheduler(), OracleDDIMScheduler()
    s.set_timesteps(50)
    o.set_timesteps(50)
    assert s.timesteps.tolist() == o.timesteps.tolist() == list(range(981, 0, -20))
    assert s.init_noise_sigma == 1.0 and s.order == 1 and s.config.steps_offset == 1 and not s.config.clip_sample
    import inspect
    assert {"eta", "generator"} <= set(inspect.signature(s.step).parameters)
    x, e = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    tab = s.coefficient_table("cpu")
    assert tab.shape == (50, 4)
    for row, t in ((0, 981), (49, 1)):
        c = tab[row]
        want = o.step(e, t, x, return_dict=False)[0]
        got = c[2] * (x - c[1] * e) * c[0] + c[3] * e
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    with pytest.raises(RuntimeError):
        s.step(e, 981, x)  # CPU tensors: no fallback
    assert torch.equal(s.scale_model_input(x, 3), x)


def test_from_pretrained_ignore_mismatched_sizes_like_the_reference_driver(tmp_path):
    """stage2_batchtest_inpaint_model.py:125-128: a stock SD-2.1 `unet/` folder (4-channel conv_in, no class
    embedding) loaded with in_channels=9 + projection class embedding and ignore_mismatched_sizes=True.  Tensors the
    checkpoint lacks or holds in another shape stay freshly initialised (diffusers semantics), everything else is
    loaded; then the driver's own `load_state_dict(unet_dict)` (:130) overwrites all of it."""
    import json
    from dataclasses import replace
    cfg = UNetConfig.tiny()
    stock_cfg = replace(cfg, in_channels=4, class_embed_type=None, projection_class_embeddings_input_dim=None,
                        use_pose_cond=False)
    stock = make_unet(stock_cfg, seed=3).state_dict()
    root = tmp_path / "sd21" / "unet"
    root.mkdir(parents=True)
    cj = {k: v for k, v in asdict(stock_cfg).items() if k != "use_pose_cond"}
    (root / "config.json").write_text(json.dumps(cj))
    torch.save(stock, root / "diffusion_pytorch_model.bin")
    m = B200UNet2DConditionModel.from_pretrained(str(tmp_path / "sd21"), subfolder="unet", in_channels=9,
                                                 class_embed_type="projection",
                                                 projection_class_embeddings_input_dim=cfg.projection_class_embeddings_input_dim,
                                                 torch_dtype=torch.float32, low_cpu_mem_usage=False,
                                                 ignore_mismatched_sizes=True, device="cpu")
    assert m.config.in_channels == 9 and m._loaded
    # loaded tensors come from the checkpoint ...
    k = "down_blocks.0.resnets.0.norm1.weight"
    assert torch.equal(m._w[k], stock[k])
    # ... the mismatched conv_in is NOT the zero-padded 4-channel weight but a fresh 9-channel one
    ci = m._w["conv_in.weight"].view(-1, 3, 3, 64)
    assert ci[..., 4:9].abs().sum() > 0 and ci[..., 9:].abs().sum() == 0
    assert "class_embedding.linear_1.weight" in m._w
    with pytest.raises(RuntimeError):   # strict (the default) still refuses
        B200UNet2DConditionModel(dtype=torch.float32, device="cpu", **asdict(cfg)).load_state_dict(stock)
    # the driver's second step: the real stage-2 checkpoint, strict
    v0 = m._weights_version
    res = m.load_state_dict(make_unet(cfg, seed=0).state_dict())
    assert not res.missing_keys and m._weights_version == v0 + 1


def test_weight_changes_invalidate_captured_graph_state():
    """Captured graphs hold raw weight pointers: load_state_dict / consolidate bump the model's weight generation and
    the pipeline's rebuild test includes it."""
    cfg = UNetConfig.tiny()
    o, m = _models(cfg)
    v = m._weights_version
    m.consolidate()
    assert m._weights_version == v + 1
    m.load_state_dict(o.state_dict())
    assert m._weights_version == v + 2 and m._arena is None


def test_pipeline_accepts_unipc_default_scheduler_and_reads_scheduler_config(tmp_path):
    """constructing with B200UniPCMultistepScheduler() (steps_offset 0, linspace spacing) must not raise —
    the reference only patches an outdated steps_offset with a deprecation warning (:86-98); from_pretrained reads
    scheduler/scheduler_config.json when present."""
    import json
    import warnings
    from pcdms_b200.scheduler import B200UniPCMultistepScheduler
    cfg = UNetConfig.tiny()
    _, m = _models(cfg)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        pipe = B200Stage2InpaintPipeline(vae=None, unet=m, scheduler=B200UniPCMultistepScheduler())
    assert pipe.scheduler.config.steps_offset == 1 and any("steps_offset" in str(x.message) for x in w)
    (tmp_path / "scheduler").mkdir()
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(
        {"_class_name": "DDIMScheduler", "beta_start": 0.001, "beta_end": 0.02, "steps_offset": 1,
         "clip_sample": False, "set_alpha_to_one": False}))
    p2 = B200Stage2InpaintPipeline.from_pretrained(str(tmp_path), unet=m)
    assert p2.scheduler.config.beta_start == 0.001 and p2.scheduler.config.beta_end == 0.02


def test_default_workspace_is_per_device_and_never_replaced():
    """'cuda' and 'cuda:0' must name the same scratch buffer and an existing one is never freed —
    captured graphs hold its address.  (Host logic only: no GPU here, so the allocation itself is patched.)"""
    from unittest import mock
    from pcdms_b200 import ops
    made = []
    real_zeros = torch.zeros

    def fake_zeros(n, dtype=None, device=None):
        made.append((n, str(device)))
        return real_zeros(8, dtype=torch.uint8)

    with mock.patch.object(ops, "_default_ws", {}), mock.patch("torch.cuda.current_device", return_value=0), \
            mock.patch("torch.zeros", fake_zeros):
        a = ops.ensure_workspace("cuda")
        b = ops.ensure_workspace("cuda:0")
        c = ops.ensure_workspace(torch.device("cuda", 0), nbytes=1)
        assert a is b is c and len(made) == 1 and made[0][0] == ops.WORKSPACE_BYTES
        assert ops.ensure_workspace("cpu") is None


@pytest.mark.parametrize("fast", [True, False])
def test_guidance_rescale_matches_reference_formula(fast):
    """rescale_noise_cfg (reference stage2_inpaint_pipeline.py:52-63, applied at :514-516): through the fused engine
    (per-sample std ratio kernel + the fused step) and through the generic protocol loop (pcdm_cfg_combine)."""
    cfg = UNetConfig.tiny()
    o, m = _models(cfg)
    pin = make_inputs(cfg, n=2, h=8, w=16, s_kv=6)
    cond = prepare_conditioning(s_img_proj_f=pin["s_img_proj_f"], pred_t_img_embed=pin["pred_t_img_embed"],
                                st_pose_f=pin["st_pose_f"], masked_latents=pin["masked_latents"], height=pin["height"],
                                width=pin["width"], num_images_per_prompt=2, guidance_scale=2.0)
    want = denoise_loop(o, OracleDDIMScheduler(), latents=pin["latents"], cond=cond, num_inference_steps=4,
                        guidance_scale=2.0, guidance_rescale=0.7)
    plain = denoise_loop(o, OracleDDIMScheduler(), latents=pin["latents"], cond=cond, num_inference_steps=4,
                         guidance_scale=2.0)
    assert (want - plain).abs().max() > 1e-2      # the rescale does something
    pipe = B200Stage2InpaintPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())
    pipe.use_cuda_graph = False
    calls = []
    with mock_ops.patched():
        got = pipe(height=pin["height"], width=pin["width"], num_inference_steps=4, guidance_scale=2.0,
                   guidance_rescale=0.7, num_images_per_prompt=2, latents=pin["latents"], output_type="latent",
                   s_img_proj_f=pin["s_img_proj_f"], st_pose_f=pin["st_pose_f"],
                   pred_t_img_embed=pin["pred_t_img_embed"], masked_latents=pin["masked_latents"],
                   callback=None if fast else (lambda i, t, x: calls.append(i))).images
    assert fast or calls == [0, 1, 2, 3]
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-5)
