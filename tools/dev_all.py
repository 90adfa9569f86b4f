"""Dev diagnostics (GPU): every kernel + the full UNet + the pipeline against the CPU oracle, with error statistics.
Usage: python tools/dev_all.py <section>   (sections: norm attn misc unet_tiny unet_full pipe_tiny perf)
       python tools/dev_all.py all         (runs each section in a subprocess with a timeout)
"""
import json
import subprocess
import sys
import time
from dataclasses import asdict

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")

SECTIONS = ["norm", "attn", "misc", "unet_tiny", "pipe_tiny", "unet_full", "perf"]
res = []
dev = "cuda"


def stats(name, got, want, **kw):
    got = got.detach().float().cpu()
    want = want.detach().float().cpu()
    err = (got - want).abs()
    tol = 1e-4 + 1e-3 * want.abs()
    r = dict(name=name, max_abs=err.max().item(), mean_abs=err.mean().item(), ref_absmax=want.abs().max().item(),
             ref_std=want.std().item(), frac_viol_1e3=(err > tol).float().mean().item(),
             max_err_over_tol=(err / tol).max().item(), nan=bool(torch.isnan(got).any()), **kw)
    res.append(r)
    print(json.dumps(r), flush=True)
    return r


def sec_norm():
    from pcdms_b200 import ops
    torch.manual_seed(0)
    for dt in (torch.float16, torch.bfloat16):
        for (B, H, W, C1, C2) in [(2, 32, 64, 320, 0), (2, 16, 32, 1280, 640), (3, 4, 8, 1280, 1280), (2, 8, 16, 640, 320),
                                  (1, 64, 128, 320, 0), (2, 2, 4, 64, 0)]:
            C = C1 + C2
            x1 = (torch.randn(B, C1, H, W) * 2 + 0.5).to(dt)
            x2 = (torch.randn(B, C2, H, W) - 0.3).to(dt) if C2 else None
            g, b = torch.randn(C), torch.randn(C)
            xc = torch.cat([x1, x2], 1) if C2 else x1
            for silu in (True, False):
                ref = F.group_norm(xc.float(), 32, g, b, 1e-5)
                if silu:
                    ref = F.silu(ref)
                out = ops.groupnorm(x1.permute(0, 2, 3, 1).contiguous().to(dev), g.to(dev), b.to(dev), 1e-5,
                                    x2=x2.permute(0, 2, 3, 1).contiguous().to(dev) if C2 else None, silu=silu)
                torch.cuda.synchronize()
                stats("groupnorm", out.permute(0, 3, 1, 2), ref, dt=str(dt), shape=[B, H, W, C1, C2], silu=silu)
        for (M, C) in [(4096, 320), (1000, 640), (77, 1280), (33, 64)]:
            x = (torch.randn(M, C) * 3 + 1).to(dt)
            g, b = torch.randn(C), torch.randn(C)
            ref = F.layer_norm(x.float(), (C,), g, b, 1e-5)
            out = ops.layernorm(x.to(dev), g.to(dev), b.to(dev))
            torch.cuda.synchronize()
            stats("layernorm", out, ref, dt=str(dt), shape=[M, C])


def sec_attn():
    from pcdms_b200 import ops
    torch.manual_seed(0)
    for dt in (torch.float16, torch.bfloat16):
        for (B, heads, Sq, Skv, sc) in [(2, 5, 2048, 2048, 1.0), (2, 10, 512, 512, 1.0), (2, 20, 128, 128, 1.0),
                                        (3, 20, 32, 32, 1.0), (2, 5, 2048, 258, 1.0), (2, 10, 512, 95, 1.0),
                                        (1, 20, 128, 257, 1.0), (1, 2, 384, 300, 4.0), (1, 1, 200, 130, 8.0)]:
            C = heads * 64
            q = (torch.randn(B, Sq, C) * sc).to(dt)
            k = (torch.randn(B, Skv, C) * sc).to(dt)
            v = torch.randn(B, Skv, C).to(dt)
            qh = q.float().view(B, Sq, heads, 64).transpose(1, 2)
            kh = k.float().view(B, Skv, heads, 64).transpose(1, 2)
            vh = v.float().view(B, Skv, heads, 64).transpose(1, 2)
            ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, Sq, C)
            # fused-buffer style strided inputs when shapes allow
            if Sq == Skv:
                qkv = torch.cat([q, k, v], dim=-1).reshape(B * Sq, 3 * C).to(dev)
                out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads)
            else:
                out = ops.attention(q.reshape(B * Sq, C).to(dev), k.reshape(B * Skv, C).to(dev),
                                    v.reshape(B * Skv, C).to(dev), B, heads)
            torch.cuda.synchronize()
            stats("attention", out.view(B, Sq, C), ref, dt=str(dt), shape=[B, heads, Sq, Skv], scale=sc)


def sec_misc():
    from pcdms_b200 import ops
    from pcdms_b200.scheduler import B200DDIMScheduler, B200DDPMScheduler
    from oracle.schedulers import OracleDDIMScheduler, ddpm_add_noise
    from oracle.pipeline import cfg_combine
    from oracle import blocks as OB
    torch.manual_seed(0)
    # timestep embedding
    for t in ([981.0], [1.0, 21.0, 501.0, 999.0]):
        tt = torch.tensor(t)
        B = 4
        ref = OB.Timesteps(320, True, 0)(tt.expand(B) if len(t) == 1 else tt)
        out = ops.timestep_embedding(tt.to(dev), B, 320, torch.float32)
        stats("timestep_embedding_f32", out, ref, t=t)
        out = ops.timestep_embedding(tt.to(dev), B, 320, torch.float16)
        stats("timestep_embedding_f16", out, ref.half(), t=t)
    # layout
    x = torch.randn(3, 9, 16, 32)
    out = ops.nchw_to_nhwc_pad(x.to(dev), 64, torch.float16)
    ref = torch.zeros(3, 16, 32, 64)
    ref[..., :9] = x.permute(0, 2, 3, 1).half().float()
    stats("nchw_to_nhwc_pad", out, ref)
    y = torch.randn(3, 16, 32, 32)
    out = ops.nhwc_to_nchw(y.to(dev), 4, torch.float32)
    stats("nhwc_to_nchw", out, y[..., :4].permute(0, 3, 1, 2))
    # upsample
    x = torch.randn(2, 4, 8, 128).half()
    out = ops.upsample_nearest2x(x.to(dev))
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    stats("upsample", out, ref)
    # fused step vs oracle
    n, h, w = 2, 16, 32
    sch = B200DDIMScheduler(); sch.set_timesteps(10)
    osch = OracleDDIMScheduler(); osch.set_timesteps(10)
    lat = torch.randn(n, 4, h, w)
    eps_rows = torch.randn(2 * n, h, w, 32)
    x9 = torch.zeros(2 * n, h, w, 64, dtype=torch.float16)
    coef = sch.coefficient_table(dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    t_table = torch.cat([sch.timesteps.float(), torch.zeros(1)]).to(dev)
    t_cur = torch.zeros(1, device=dev)
    lat_d, x9_d = lat.clone().to(dev), x9.to(dev)
    eps_nchw = eps_rows[..., :4].permute(0, 3, 1, 2)
    ref = lat.clone()
    for i, t in enumerate(osch.timesteps[:3]):
        ops.cfg_ddim_step(eps_rows.to(dev), lat_d, x9_d, coef, counter, 2.0, t_table, t_cur)
        ref = osch.step(cfg_combine(eps_nchw, 2.0), t, ref, return_dict=False)[0]
        torch.cuda.synchronize()
        stats("cfg_ddim_step_latents", lat_d, ref, step=i)
        stats("cfg_ddim_step_x9", x9_d[..., :4].float(), torch.cat([ref, ref]).permute(0, 2, 3, 1).half().float(), step=i)
        print("counter", counter.tolist(), "t_cur", t_cur.tolist(), "expected next t", float(osch.timesteps[i + 1]))
    # stand-alone step + add_noise
    e = torch.randn(2, 4, 8, 8); s = torch.randn(2, 4, 8, 8)
    out = sch.step(e.to(dev), 901, s.to(dev), return_dict=False)[0]
    stats("ddim_step", out, osch.step(e, 901, s, return_dict=False)[0])
    ts = torch.tensor([0, 500]);
    out = B200DDPMScheduler().add_noise(s.to(dev), e.to(dev), ts.to(dev))
    stats("add_noise", out, ddpm_add_noise(s, e, ts))
    # gemm silu + rowvec stride
    a = torch.randn(8, 128).half(); w = (torch.randn(256, 128) / 11).half(); b = torch.randn(256)
    rv = torch.randn(2, 1024)
    out = ops.gemm(a.to(dev), w.to(dev), bias=b.to(dev), rowvec=rv.to(dev)[:, 256:512], rows_per_image=4, silu=True)
    ref = F.silu(a.float() @ w.float().t() + b + rv[:, 256:512].repeat_interleave(4, 0))
    stats("gemm_silu_rowvec", out, ref)


def _mk_models(cfg, dt, seed=0):
    from oracle.factory import make_unet
    from pcdms_b200.unet import B200UNet2DConditionModel
    o = make_unet(cfg, seed=seed)
    d = asdict(cfg)
    m = B200UNet2DConditionModel(dtype=dt, device=dev, **d)
    m.load_state_dict(o.state_dict())
    return o, m


def _unet_compare(cfg, dt, batch, h, w, s_kv, label, debug_layers=False):
    from oracle.factory import make_unet_inputs
    o, m = _mk_models(cfg, dt)
    i = make_unet_inputs(cfg, batch=batch, h=h, w=w, s_kv=s_kv)
    t0 = time.time()
    with torch.no_grad():
        ref = o(i["sample"], 981, i["encoder_hidden_states"], class_labels=i.get("class_labels"),
                my_pose_cond=i.get("my_pose_cond"))[0]
    print(f"oracle forward {time.time() - t0:.1f}s", flush=True)
    kw = {}
    if "class_labels" in i:
        kw["class_labels"] = i["class_labels"].to(dev)
    if "my_pose_cond" in i:
        kw["my_pose_cond"] = i["my_pose_cond"].to(dev)
    out = m(i["sample"].to(dev), 981, i["encoder_hidden_states"].to(dev), return_dict=False, **kw)[0]
    torch.cuda.synchronize()
    stats(label, out, ref, dt=str(dt))
    # oracle fed the same 16-bit rounded weights/inputs isolates kernel error from quantisation error
    return o, m, i, ref, out


def sec_unet_tiny():
    from oracle.unet import UNetConfig
    for dt in (torch.float16, torch.bfloat16):
        _unet_compare(UNetConfig.tiny(), dt, 2, 16, 32, 9, "unet_tiny_stage2")
    _unet_compare(UNetConfig.tiny(in_channels=8, stage2=False), torch.float16, 2, 16, 16, 17, "unet_tiny_stage3")
    _unet_compare(UNetConfig.tiny(), torch.float16, 3, 32, 64, 258, "unet_tiny_stage2_32x64")


def sec_unet_full():
    from oracle.unet import UNetConfig
    o, m, i, ref, out = _unet_compare(UNetConfig.stage2(), torch.float16, 2, 32, 64, 258, "unet_full_stage2_fp16")
    print("weight bytes", m.weight_bytes())
    # timing eager
    kw = dict(class_labels=i["class_labels"].to(dev), my_pose_cond=i["my_pose_cond"].to(dev))
    s, e = i["sample"].to(dev), i["encoder_hidden_states"].to(dev)
    for _ in range(2):
        m(s, 981, e, return_dict=False, **kw)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(5):
        m(s, 981, e, return_dict=False, **kw)
    torch.cuda.synchronize()
    print("eager unet fwd B=2 ms", (time.time() - t0) / 5 * 1e3, flush=True)
    del m
    torch.cuda.empty_cache()
    from pcdms_b200.unet import B200UNet2DConditionModel
    m = B200UNet2DConditionModel(dtype=torch.bfloat16, device=dev, **asdict(UNetConfig.stage2()))
    m.load_state_dict(o.state_dict())
    out = m(s, 981, e, return_dict=False, **kw)[0]
    stats("unet_full_stage2_bf16", out, ref)


def sec_pipe_tiny():
    from oracle.unet import UNetConfig
    from oracle.factory import make_inputs
    from oracle.pipeline import prepare_conditioning, denoise_loop
    from oracle.schedulers import OracleDDIMScheduler
    from pcdms_b200.pipeline import B200Stage2InpaintPipeline
    from pcdms_b200.scheduler import B200DDIMScheduler
    cfg = UNetConfig.tiny()
    for dt in (torch.float16,):
        o, m = _mk_models(cfg, dt)
        pin = make_inputs(cfg, n=2, h=16, w=32, s_kv=9)
        cond = prepare_conditioning(s_img_proj_f=pin["s_img_proj_f"], pred_t_img_embed=pin["pred_t_img_embed"],
                                    st_pose_f=pin["st_pose_f"], masked_latents=pin["masked_latents"],
                                    height=pin["height"], width=pin["width"], num_images_per_prompt=2,
                                    guidance_scale=2.0)
        ref, traj = denoise_loop(o, OracleDDIMScheduler(), latents=pin["latents"], cond=cond, num_inference_steps=10,
                                 guidance_scale=2.0, return_trajectory=True)
        pipe = B200Stage2InpaintPipeline(vae=None, unet=m, scheduler=B200DDIMScheduler())
        for use_graph in (False, True, True):
            pipe.use_cuda_graph = use_graph
            out = pipe(height=pin["height"], width=pin["width"], num_inference_steps=10, guidance_scale=2.0,
                       num_images_per_prompt=2, latents=pin["latents"], output_type="latent",
                       s_img_proj_f=pin["s_img_proj_f"], st_pose_f=pin["st_pose_f"],
                       pred_t_img_embed=pin["pred_t_img_embed"], masked_latents=pin["masked_latents"]).images
            torch.cuda.synchronize()
            stats("pipeline_tiny_10step", out, ref, dt=str(dt), graph=use_graph)


def sec_perf():
    """First whole-step timing at BASELINE config 2 (B=16, 32x64, S_kv 258, bf16) with random weights, CUDA graph."""
    from oracle.unet import UNetConfig, OracleUNet
    from pcdms_b200.unet import B200UNet2DConditionModel
    from pcdms_b200 import ops
    cfg = UNetConfig.stage2()
    torch.manual_seed(0)
    with torch.device("meta"):
        shapes = {k: v.shape for k, v in OracleUNet(cfg).state_dict().items()}
    g = torch.Generator().manual_seed(0)
    sd = {k: (torch.randn(s, generator=g) * 0.02 if len(s) > 1 else torch.ones(s)) for k, s in shapes.items()}
    dt = torch.bfloat16
    m = B200UNet2DConditionModel(dtype=dt, device=dev, **asdict(cfg))
    m.load_state_dict(sd)
    B, h, w = 16, 32, 64
    x9 = torch.randn(B, h, w, 64, device=dev).to(dt)
    t = torch.tensor([981.0], device=dev)
    ctx = torch.randn(B, 258, 1024, device=dev).to(dt)
    cls = torch.randn(B, 1024, device=dev).to(dt)
    pose = torch.randn(B, h, w, 320, device=dev).to(dt)
    kv = m.context_kv(ctx)
    for _ in range(2):
        m.forward_nhwc(x9, t, kv, cls, pose)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        m.forward_nhwc(x9, t, kv, cls, pose)
    e1.record(); torch.cuda.synchronize()
    print("eager fwd B=16 ms", e0.elapsed_time(e1) / 5, flush=True)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        m.forward_nhwc(x9, t, kv, cls, pose)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps(dict(name="graph_fwd_B16_bf16", ms=ms, tflops=6.192 / ms * 1e3)), flush=True)
    res.append(dict(name="graph_fwd_B16_bf16", ms=ms, tflops=6.192 / ms * 1e3))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        for s in SECTIONS:
            t0 = time.time()
            p = subprocess.run([sys.executable, __file__, s], capture_output=True, text=True, timeout=None if False else 600)
            open(f"gpurun_out/dev_all_{s}.log", "w").write(p.stdout + "\n=== STDERR ===\n" + p.stderr[-6000:])
            print(f"== {s} rc={p.returncode} {time.time() - t0:.0f}s")
            print(p.stdout[-3500:])
            if p.returncode != 0:
                print(p.stderr[-2500:])
    else:
        globals()[f"sec_{which}"]()
        json.dump(res, open(f"gpurun_out/dev_all_{which}.json", "w"), indent=1)
