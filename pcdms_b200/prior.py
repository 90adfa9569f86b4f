"""Stage-1 prior on the B200 kernels: `B200Stage1PriorTransformer` is the drop-in for the reference's
`Stage1_PriorTransformer` (/root/reference/src/models/stage1_prior_transformer.py:50-301) and
`B200Stage1PriorPipeline` for `Stage1_PriorPipeline` (/root/reference/src/pipelines/stage1_prior_pipeline.py:127-505),
as driven by /root/reference/stage1_batchtest_prior_model.py:57-113.

The prior denoises ONE CLIP image embedding (1024 values) with a 20-block transformer over six tokens —
[source pose, target pose, source-image embedding, timestep, noisy embedding, learned query] — so a step is ~1.0 G
weight parameters streamed against 6 (12 under CFG) activation rows: weight-bandwidth work.  Everything runs through
the C ABI: with 6-12 rows every linear layer is a weight stream — `pcdm_gemm` routes them to the skinny mma.sync kernel
(16 weight rows per CTA, weights requested before the programmatic-dependency wait so that they stream in under the
previous kernel); the q/k/v projections are one GEMM, bias / GELU / residual live in its epilogue, attention over the 6
tokens is a one-warp-per-head kernel (a LayerNorm-in-front-of-the-GEMM variant, `pcdm_ln_gemm`, exists but measured
slower and is off by default), the positional embedding added in the epilogue of whichever GEMM produces a token
(`rowvec`), the UnCLIP scheduler step fused with the CFG combine (`pcdm_cfg_unclip_step`).  Step-invariant tokens
(both poses, the source embedding, the query) are computed once per call; the loop is one CUDA graph replayed
`num_inference_steps` times (device-side step counter, as the stage-2 engine).  Only the last token reaches the output
head, so `norm_out` and `proj_to_clip_embeddings` run on that row alone.  No PyTorch / CPU compute fallback.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import ops
from .arena import WeightArenaMixin
from .scheduler import B200UnCLIPScheduler
from .unet import _Config

_DEFAULT_CONFIG = dict(num_attention_heads=32, attention_head_dim=64, num_layers=20, embedding_dim=768,
                       num_embeddings=77, additional_embeddings=4, dropout=0.0)
_POSE_DIM, _POSE_HIDDEN, _POSE_OUT = 36, 512, 1024   # stage1_prior_transformer.py:91-92 (literals in the reference)


class PriorTransformerOutput:
    def __init__(self, predicted_image_embedding):
        self.predicted_image_embedding = predicted_image_embedding

    def __getitem__(self, i):
        return (self.predicted_image_embedding,)[i]


class _NoopProcessor:
    """The fused attention kernel is always on; the registry exists so `set_attn_processor` /
    `enable_xformers_memory_efficient_attention` (stage1_batchtest_prior_model.py:59) are harmless."""


def _ln_then_gemm(x, gamma, beta, eps, w, **kw):
    """A/B path (`fuse_layernorm = False`): the LayerNorm as its own launch."""
    return ops.gemm(ops.layernorm(x, gamma, beta, eps), w, **kw)


class B200Stage1PriorTransformer(WeightArenaMixin):
    # LayerNorm fused in front of the consuming GEMM (pcdm_ln_gemm: 106 instead of 147 launches per step) measured
    # SLOWER on B200 (1.18 vs 0.94 ms per step, profiles/r1_s3_prior_ab.md): every CTA of the GEMM re-normalises the
    # rows on its critical path, behind the dependency wait.  Kept as an option; tools/bench_stage1.py flips it.
    fuse_layernorm = os.environ.get("PCDM_PRIOR_FUSE_LN", "0") != "0"

    def __init__(self, dtype: torch.dtype = torch.float16, device="cuda", **kw):
        cfg = dict(_DEFAULT_CONFIG)
        cfg.update({k: v for k, v in kw.items() if k in cfg})
        c = _Config(cfg)
        if c.attention_head_dim != 64:
            raise NotImplementedError("pcdm_b200 prior: attention_head_dim must be 64")
        self.inner_dim = c.num_attention_heads * c.attention_head_dim
        if self.inner_dim > 2048 or c.embedding_dim % 64 or c.embedding_dim != _POSE_OUT:
            # the reference hard-codes the pose MLP's output width to 1024 and feeds it to a Linear(embedding_dim, .)
            raise NotImplementedError("pcdm_b200 prior: inner_dim <= 2048 and embedding_dim == 1024 (the reference's "
                                      "pose encoder emits 1024 channels)")
        if c.num_embeddings + c.additional_embeddings != 6:
            raise NotImplementedError("pcdm_b200 prior: the token sequence is the reference's six tokens "
                                      "(num_embeddings=2, additional_embeddings=4)")
        self.config = c
        self.num_attention_heads, self.attention_head_dim = c.num_attention_heads, c.attention_head_dim
        self.additional_embeddings = c.additional_embeddings
        self._dtype, self._device = dtype, torch.device(device)
        self._w: Dict[str, torch.Tensor] = {}
        self._loaded = False
        self.clip_mean, self.clip_std = torch.tensor(-0.016), torch.tensor(0.415)      # :134-135
        self._processors = {f"transformer_blocks.{i}.attn1.processor": _NoopProcessor() for i in range(c.num_layers)}
        ops.ensure_workspace(self._device)

    # -- surface -------------------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, path, subfolder=None, torch_dtype=torch.float16, device="cuda", low_cpu_mem_usage=False,
                        ignore_mismatched_sizes=False, **kw):
        """Config from <path>/<subfolder>/config.json, overridden by keyword (the reference passes num_embeddings=2,
        embedding_dim=1024 over the Kandinsky checkpoint, stage1_batchtest_prior_model.py:57); checkpoint tensors whose
        shape differs are skipped under ignore_mismatched_sizes, as diffusers does."""
        root = os.path.join(path, subfolder) if subfolder else path
        cfg = {}
        if os.path.exists(os.path.join(root, "config.json")):
            with open(os.path.join(root, "config.json")) as f:
                cfg = {k: v for k, v in json.load(f).items() if k in _DEFAULT_CONFIG}
        cfg.update(kw)
        m = cls(dtype=torch_dtype, device=device, **cfg)
        sd = None
        for name in ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin"):
            fp = os.path.join(root, name)
            if os.path.exists(fp):
                if name.endswith(".safetensors"):
                    from safetensors.torch import load_file
                    sd = load_file(fp)
                else:
                    sd = torch.load(fp, map_location="cpu")
                break
        if sd is not None:
            shapes = m.state_dict_shapes()
            if ignore_mismatched_sizes:
                sd = {k: v for k, v in sd.items() if k in shapes and tuple(v.shape) == shapes[k]}
            m.load_state_dict(sd, strict=not ignore_mismatched_sizes)
        return m

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype) and a != self._dtype:
                raise NotImplementedError("pcdm_b200 prior: choose the dtype at construction (weights are pre-packed)")
            if isinstance(a, (str, torch.device)) and torch.device(a).type != "cuda":
                raise RuntimeError("pcdm_b200 prior runs on CUDA only (no CPU fallback)")
        return self

    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        return self

    @property
    def attn_processors(self):
        return dict(self._processors)

    def set_attn_processor(self, processor):
        if isinstance(processor, dict):
            if len(processor) != len(self._processors):
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does "
                                 f"not match the number of attention layers: {len(self._processors)}.")
            self._processors = dict(processor)
        else:
            self._processors = {k: processor for k in self._processors}

    def set_default_attn_processor(self):
        self.set_attn_processor(_NoopProcessor())

    def post_process_latents(self, prior_latents):
        return prior_latents * self.clip_std.to(prior_latents.device) + self.clip_mean.to(prior_latents.device)

    # -- weights -------------------------------------------------------------------------------------------------------
    def state_dict_shapes(self) -> Dict[str, tuple]:
        c, C, E = self.config, self.inner_dim, self.config.embedding_dim
        sh = {}

        def lin(name, n_out, n_in):
            sh[f"{name}.weight"], sh[f"{name}.bias"] = (n_out, n_in), (n_out,)

        def ln(name, n):
            sh[f"{name}.weight"], sh[f"{name}.bias"] = (n,), (n,)
        for p in ("pose_encoder", "pose_encoder1"):
            lin(f"{p}.net.0", _POSE_HIDDEN, _POSE_DIM)
            ln(f"{p}.net.3", _POSE_HIDDEN)
            lin(f"{p}.net.4", _POSE_OUT, _POSE_HIDDEN)
            ln(f"{p}.net.6", _POSE_OUT)
        lin("time_embedding.linear_1", C, C)
        lin("time_embedding.linear_2", C, C)
        lin("proj_in", C, E)
        lin("embedding_proj", C, E)
        lin("encoder_hidden_states_proj", C, E)
        lin("encoder_hidden_states_proj1", C, E)
        sh["positional_embedding"] = (1, c.num_embeddings + c.additional_embeddings, C)
        sh["prd_embedding"] = (1, 1, C)
        for i in range(c.num_layers):
            b = f"transformer_blocks.{i}"
            ln(f"{b}.norm1", C)
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(f"{b}.attn1.{n}", C, C)
            ln(f"{b}.norm3", C)
            lin(f"{b}.ff.net.0.proj", 4 * C, C)
            lin(f"{b}.ff.net.2", C, 4 * C)
        ln("norm_out", C)
        lin("proj_to_clip_embeddings", E, C)
        return sh

    def load_state_dict(self, state_dict, strict: bool = True):
        shapes = self.state_dict_shapes()
        missing = [k for k in shapes if k not in state_dict]
        unexpected = [k for k in state_dict if k not in shapes]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for B200Stage1PriorTransformer: missing {missing[:5]} "
                               f"unexpected {unexpected[:5]}")
        for k, shp in shapes.items():
            if k in state_dict and tuple(state_dict[k].shape) != shp:
                raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(state_dict[k].shape)} vs model {shp}")
        if missing:   # non-strict partial load: start the absent tensors from the synthetic initialisation
            state_dict = {**{k: v for k, v in self.synthetic_state_dict(device="cpu").items() if k in missing},
                          **state_dict}
        sd, w, dev, dt = state_dict, self._w, self._device, self._dtype

        def f(k):
            return sd[k].detach().float()

        def mat(t):
            return t.to(device=dev, dtype=dt).contiguous()

        def vec(t):
            return t.to(device=dev, dtype=torch.float32).contiguous()

        def lin(name, dst=None):
            w[f"{dst or name}.weight"], w[f"{dst or name}.bias"] = mat(f(f"{name}.weight")), vec(f(f"{name}.bias"))

        for p in ("pose_encoder", "pose_encoder1"):
            w0 = f(f"{p}.net.0.weight")                                  # [512, 36] -> K zero-padded to 64
            w[f"{p}.net.0.weight"] = mat(torch.cat([w0, w0.new_zeros(w0.shape[0], 64 - w0.shape[1])], dim=1))
            w[f"{p}.net.0.bias"] = vec(f(f"{p}.net.0.bias"))
            lin(f"{p}.net.4")
            for n in ("net.3", "net.6"):
                w[f"{p}.{n}.weight"], w[f"{p}.{n}.bias"] = vec(f(f"{p}.{n}.weight")), vec(f(f"{p}.{n}.bias"))
        for n in ("time_embedding.linear_1", "time_embedding.linear_2", "proj_in", "embedding_proj",
                  "encoder_hidden_states_proj", "encoder_hidden_states_proj1", "proj_to_clip_embeddings"):
            lin(n)
        pos = f("positional_embedding")[0]                               # [6, C]
        w["pos"] = vec(pos)                                              # fp32 rows, added in GEMM epilogues
        w["pos16"] = mat(pos)
        w["prd_pos"] = mat(f("prd_embedding")[0, 0] + pos[-1])           # the learned query token + its position
        for i in range(self.config.num_layers):
            b = f"transformer_blocks.{i}"
            a = f"{b}.attn1"
            w[f"{i}.qkv.weight"] = mat(torch.cat([f(f"{a}.to_q.weight"), f(f"{a}.to_k.weight"), f(f"{a}.to_v.weight")]))
            w[f"{i}.qkv.bias"] = vec(torch.cat([f(f"{a}.to_q.bias"), f(f"{a}.to_k.bias"), f(f"{a}.to_v.bias")]))
            lin(f"{a}.to_out.0", f"{i}.out")
            lin(f"{b}.ff.net.0.proj", f"{i}.ff1")
            lin(f"{b}.ff.net.2", f"{i}.ff2")
            for n in ("norm1", "norm3"):
                w[f"{i}.{n}.weight"], w[f"{i}.{n}.bias"] = vec(f(f"{b}.{n}.weight")), vec(f(f"{b}.{n}.bias"))
        w["norm_out.weight"], w["norm_out.bias"] = vec(f("norm_out.weight")), vec(f("norm_out.bias"))
        self._arena = None
        self._loaded = True
        self._weights_version += 1
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def synthetic_state_dict(self, seed: int = 0, device=None) -> Dict[str, torch.Tensor]:
        dev = torch.device(device) if device is not None else self._device
        g = torch.Generator(device=dev).manual_seed(seed)
        sd = {}
        for k, shp in self.state_dict_shapes().items():
            if k in ("positional_embedding", "prd_embedding"):
                sd[k] = 0.5 * torch.randn(shp, generator=g, device=dev)
            elif len(shp) == 2:
                sd[k] = torch.randn(shp, generator=g, device=dev) * shp[1] ** -0.5
            elif k.endswith(".weight"):     # LayerNorm scale
                sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g, device=dev)
            else:
                sd[k] = 0.05 * torch.randn(shp, generator=g, device=dev)
        return sd

    # -- engine --------------------------------------------------------------------------------------------------------
    def _guard(self, x):
        if not self._loaded:
            raise RuntimeError("B200Stage1PriorTransformer: load_state_dict() first")
        if not x.is_cuda:
            raise RuntimeError("pcdm_b200 prior runs on CUDA tensors only (no CPU fallback)")

    def _pose_token(self, pose, which, out_rows, pos_row):
        """pose [b, 36] -> MLP (Linear+GELU, LN, Linear, LN) -> Linear(1024, C) + position, written into out_rows."""
        w, dt = self._w, self._dtype
        p = "pose_encoder" if which == 0 else "pose_encoder1"
        proj = "encoder_hidden_states_proj" if which == 0 else "encoder_hidden_states_proj1"
        a = torch.zeros((pose.shape[0], 64), device=pose.device, dtype=dt)
        a[:, :_POSE_DIM] = pose
        h = ops.gemm(a, w[f"{p}.net.0.weight"], bias=w[f"{p}.net.0.bias"], gelu=True)
        h = ops.layernorm(h, w[f"{p}.net.3.weight"], w[f"{p}.net.3.bias"], 1e-5)
        h = ops.gemm(h, w[f"{p}.net.4.weight"], bias=w[f"{p}.net.4.bias"])
        h = ops.layernorm(h, w[f"{p}.net.6.weight"], w[f"{p}.net.6.bias"], 1e-5)
        ops.gemm(h, w[f"{proj}.weight"], out=out_rows, bias=w[f"{proj}.bias"], rowvec=pos_row,
                 rows_per_image=max(1, h.shape[0]))

    def encode_condition(self, proj_embedding, pose_s, pose_t, test_flag=False, out=None):
        """The step-invariant tokens of a call: returns the [B, 6, C] token template with rows 0 (source pose),
        1 (target pose), 2 (source-image embedding) and 5 (learned query) filled in, position embedding included
        (stage1_prior_transformer.py:243-281).  proj_embedding [B, 1, E]; pose_s / pose_t [b, 1, 36] with b = B, or
        b = B/2 under test_flag — the unconditional half then gets all-zero pose tokens (:255-258)."""
        self._guard(proj_embedding)
        w, dt, C = self._w, self._dtype, self.inner_dim
        B = proj_embedding.shape[0]
        b = pose_s.shape[0]
        if (2 * b if test_flag else b) != B or pose_t.shape[0] != b:
            raise ValueError(f"batch mismatch: proj_embedding {B}, poses {b} / {pose_t.shape[0]} (test_flag={test_flag})")
        tok = out if out is not None else torch.empty((B, 6, C), device=proj_embedding.device, dtype=dt)
        pos = w["pos"]
        lo = B - b
        if lo:
            tok[:lo, 0] = w["pos16"][0]
            tok[:lo, 1] = w["pos16"][1]
        self._pose_token(pose_s.reshape(b, -1).to(dt), 0, tok[lo:, 0], pos[0:1])
        self._pose_token(pose_t.reshape(b, -1).to(dt), 1, tok[lo:, 1], pos[1:2])
        ops.gemm(proj_embedding.reshape(B, -1).to(dt).contiguous(), w["embedding_proj.weight"], out=tok[:, 2],
                 bias=w["embedding_proj.bias"], rowvec=pos[2:3], rows_per_image=B)
        tok[:, 5] = w["prd_pos"]
        return tok

    def forward_tokens(self, x_rows, t, tok, out_f32=True):
        """One evaluation: x_rows [B, E] 16-bit (x_t), t fp32 device tensor (1 or B entries), tok the template of
        `encode_condition` (rows 3 and 4 are rewritten here).  Returns the predicted embedding [B, E]."""
        w, c, C = self._w, self.config, self.inner_dim
        B, heads = x_rows.shape[0], c.num_attention_heads
        pos = w["pos"]
        temb = ops.timestep_embedding(t, B, C, self._dtype)                                   # Timesteps(C, True, 0)
        temb = ops.gemm(temb, w["time_embedding.linear_1.weight"], bias=w["time_embedding.linear_1.bias"], silu=True)
        ops.gemm(temb, w["time_embedding.linear_2.weight"], out=tok[:, 3], bias=w["time_embedding.linear_2.bias"],
                 rowvec=pos[3:4], rows_per_image=B)
        ops.gemm(x_rows, w["proj_in.weight"], out=tok[:, 4], bias=w["proj_in.bias"], rowvec=pos[4:5], rows_per_image=B)
        x = tok.view(B * 6, C)
        ln_gemm = ops.ln_gemm if self.fuse_layernorm else _ln_then_gemm
        for i in range(c.num_layers):
            qkv = ln_gemm(x, w[f"{i}.norm1.weight"], w[f"{i}.norm1.bias"], 1e-5, w[f"{i}.qkv.weight"],
                              bias=w[f"{i}.qkv.bias"])
            att = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads)
            x = ops.gemm(att, w[f"{i}.out.weight"], bias=w[f"{i}.out.bias"], residual=x)
            h = ln_gemm(x, w[f"{i}.norm3.weight"], w[f"{i}.norm3.bias"], 1e-5, w[f"{i}.ff1.weight"],
                            bias=w[f"{i}.ff1.bias"], gelu=True)
            x = ops.gemm(h, w[f"{i}.ff2.weight"], bias=w[f"{i}.ff2.bias"], residual=x)
        return ln_gemm(x.view(B, 6, C)[:, 5], w["norm_out.weight"], w["norm_out.bias"], 1e-5,
                           w["proj_to_clip_embeddings.weight"], bias=w["proj_to_clip_embeddings.bias"], out_f32=out_f32)

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    @torch.no_grad()
    def forward(self, hidden_states, timestep, proj_embedding, encoder_hidden_states, encoder_hidden_states1,
                attention_mask=None, return_dict: bool = True, do_classifier_free_guidance: bool = False,
                test_flag: bool = False):
        """The reference's forward (stage1_prior_transformer.py:200-297): hidden_states [B, 1, E] (or [B, E])."""
        if attention_mask is not None:
            raise NotImplementedError("pcdm_b200 prior: attention_mask is never passed on the reference path "
                                      "(stage1_prior_pipeline.py:467)")
        self._guard(hidden_states)
        B = hidden_states.shape[0]
        dev = hidden_states.device
        if not torch.is_tensor(timestep):
            t = torch.tensor([float(timestep)], dtype=torch.float32, device=dev)
        else:
            t = timestep.reshape(-1).to(device=dev, dtype=torch.float32)
        tok = self.encode_condition(proj_embedding, encoder_hidden_states, encoder_hidden_states1, test_flag)
        x_rows = hidden_states.reshape(B, -1).to(self._dtype).contiguous()
        pred = self.forward_tokens(x_rows, t, tok, out_f32=False)
        if not return_dict:
            return (pred,)
        return PriorTransformerOutput(pred)


class PriorPipelineOutput:
    """`KandinskyPriorPipelineOutput` (stage1_prior_pipeline.py:112-124): attribute, key and index access (the driver
    reads `output[0]`, stage1_batchtest_prior_model.py:116)."""

    def __init__(self, image_embeds, negative_image_embeds):
        self.image_embeds, self.negative_image_embeds = image_embeds, negative_image_embeds

    def __getitem__(self, k):
        if isinstance(k, str):
            return getattr(self, k)
        return (self.image_embeds, self.negative_image_embeds)[k]


class B200Stage1PriorPipeline:
    def __init__(self, prior: Optional[B200Stage1PriorTransformer] = None, image_encoder=None, scheduler=None,
                 image_processor=None):
        self.prior, self.image_encoder, self.image_processor = prior, image_encoder, image_processor
        self.scheduler = scheduler if scheduler is not None else B200UnCLIPScheduler(
            prediction_type="sample", clip_sample=True, clip_sample_range=10.0)     # kandinsky-2-2-prior's config
        self._graphs = {}
        self.use_cuda_graph = True

    @classmethod
    def from_pretrained(cls, path, torch_dtype=torch.float16, **kw):
        """Scheduler from <path>/scheduler/scheduler_config.json when present; the prior (and image encoder) are
        assigned by the caller afterwards, as the reference driver does (stage1_batchtest_prior_model.py:56-61)."""
        sched = None
        fp = os.path.join(str(path), "scheduler", "scheduler_config.json")
        if os.path.exists(fp):
            with open(fp) as f:
                sched = B200UnCLIPScheduler.from_config(json.load(f))
        return cls(prior=kw.get("prior"), image_encoder=kw.get("image_encoder"), scheduler=sched,
                   image_processor=kw.get("image_processor"))

    def to(self, *a, **k):
        return self

    @property
    def device(self):
        return self.prior.device if self.prior is not None else torch.device("cuda")

    @property
    def _execution_device(self):
        return self.device

    def enable_xformers_memory_efficient_attention(self, *a, **k):   # stage1_batchtest_prior_model.py:59
        return None

    def set_progress_bar_config(self, **k):
        return None

    def get_zero_embed(self, batch_size=1, device=None):
        """CLIP embedding of an all-zero image (stage1_prior_pipeline.py:282-289)."""
        if self.image_encoder is None:
            raise ValueError("get_zero_embed needs the pipeline's image_encoder")
        device = device or self.device
        size = self.image_encoder.config.image_size
        zero_img = torch.zeros(1, 3, size, size, device=device, dtype=self.image_encoder.dtype)
        return self.image_encoder(zero_img)["image_embeds"].repeat(batch_size, 1)

    # ------------------------------------------------------------------------------------------------------------
    def prepare_fused(self, s_embed, s_pose, t_pose, latents, guidance_scale, steps, noise):
        """Per-call set-up: step-invariant tokens, step tables, (re)capture of the one-step graph."""
        prior, sch = self.prior, self.scheduler
        dev, dt = prior.device, prior.dtype
        n, E = latents.shape
        cfg = guidance_scale > 1.0
        B = 2 * n if cfg else n
        key = (n, E, cfg, dt)
        st = self._graphs.get(key)
        if st is None:
            st = SimpleNamespace(tok=torch.empty((B, 6, prior.inner_dim), device=dev, dtype=dt),
                                 xin=torch.empty((B, E), device=dev, dtype=dt),
                                 latents=torch.empty((n, E), device=dev, dtype=torch.float32),
                                 latents_init=torch.empty((n, E), device=dev, dtype=torch.float32),
                                 t_cur=torch.zeros(1, device=dev, dtype=torch.float32),
                                 counter=torch.zeros(2, device=dev, dtype=torch.int32), cfg=cfg, guidance=None,
                                 graph=None, coef=None, noise=None, t_table=None, steps=None, launches_per_step=0)
            self._graphs[key] = st
        prompt = s_embed.to(dev)
        if cfg:   # unconditional half first: zero source embedding (:337-343) next to zero pose tokens (test_flag)
            prompt = torch.cat([torch.zeros_like(prompt), prompt])
        prior.encode_condition(prompt, s_pose.to(dev), t_pose.to(dev), test_flag=cfg, out=st.tok)
        st.latents_init.copy_(latents)
        coef = sch.coefficient_table(dev)
        t_table = torch.cat([sch.timesteps.to(dev, torch.float32), torch.zeros(1, device=dev)]).contiguous()
        wver = getattr(prior, "_weights_version", 0)
        rebuild = st.graph is None or st.guidance != guidance_scale or st.steps != steps or getattr(st, "wver", None) != wver
        st.guidance, st.steps, st.wver = guidance_scale, steps, wver
        if st.coef is None or st.coef.shape != coef.shape:
            st.coef, st.t_table, st.noise = coef, t_table, noise.clone()
            rebuild = True
        else:
            st.coef.copy_(coef)
            st.t_table.copy_(t_table)
            st.noise.copy_(noise)
        if self.use_cuda_graph and rebuild:
            self._reset_state(st)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._one_step(st)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            from . import lib as _l
            before = _l.launch_count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._one_step(st)
            st.launches_per_step = _l.launch_count - before
            st.graph = g
        return st

    def _one_step(self, st):
        pred = self.prior.forward_tokens(st.xin, st.t_cur, st.tok)
        ops.cfg_unclip_step(pred, st.latents, st.xin, st.coef, st.noise, st.counter, st.guidance, st.cfg, st.t_table,
                            st.t_cur)

    def _reset_state(self, st):
        st.latents.copy_(st.latents_init)
        st.xin.copy_(torch.cat([st.latents_init, st.latents_init]) if st.cfg else st.latents_init)
        st.counter.zero_()
        st.t_cur.copy_(st.t_table[:1])

    def replay_fused(self, st):
        """The sampling loop proper (stage1_prior_pipeline.py:456-483): `steps` replays of the one-step graph."""
        self._reset_state(st)
        for _ in range(st.steps):
            if self.use_cuda_graph:
                st.graph.replay()
            else:
                self._one_step(st)
        return st.latents

    @torch.no_grad()
    def __call__(self, s_embed, s_pose, t_pose, negative_prompt=None, num_images_per_prompt: int = 1,
                 num_inference_steps: int = 25, generator=None, latents=None, guidance_scale: float = 4.0,
                 output_type: Optional[str] = "pt", return_dict: bool = True, variance_noise=None):
        """s_embed [b, 1, E] source-image CLIP embedding, s_pose / t_pose [b, 1, 36] normalised keypoints.  Returns the
        predicted target-image embedding [b*num_images_per_prompt, E] (post-processed: * clip_std + clip_mean) and the
        zero-image embedding.  `variance_noise` ([steps, n, E], optional, not a reference argument) injects the
        scheduler's per-step noise instead of drawing it from `generator`.

        guidance_scale <= 1 (the driver's default 0) is the reference loop as is.  guidance_scale > 1: the reference
        doubles latents and source embedding but not the pose tokens and fails inside its own transformer call; here
        the unconditional half is what its `test_flag` branch builds (zero pose tokens, zero source embedding)."""
        if isinstance(negative_prompt, str):
            negative_prompt = [negative_prompt]
        elif not isinstance(negative_prompt, list) and negative_prompt is not None:
            raise ValueError(f"`negative_prompt` has to be of type `str` or `list` but is {type(negative_prompt)}")
        if output_type not in ["pt", "np"]:
            raise ValueError(f"Only the output types `pt` and `np` are supported not output_type={output_type}")
        if self.prior is None:
            raise ValueError("assign pipe.prior first")
        dev = self.device
        self.prior._guard(s_embed)        # loaded weights, CUDA tensors (there is no CPU path)
        b, seq, E = s_embed.shape
        k = int(num_images_per_prompt)
        n = b * k
        s_embed = s_embed.repeat(1, k, 1).view(n, seq, E)                                     # :335-336
        s_pose = s_pose.repeat(1, k, 1).view(n, 1, -1)
        t_pose = t_pose.repeat(1, k, 1).view(n, 1, -1)
        sch = self.scheduler
        sch.set_timesteps(num_inference_steps, device=dev)
        if latents is None:
            latents = torch.randn((n, E), generator=generator, device=dev, dtype=torch.float32)
        elif tuple(latents.shape) != (n, E):
            raise ValueError(f"Unexpected latents shape, got {latents.shape}, expected {(n, E)}")
        latents = latents.to(dev, torch.float32) * sch.init_noise_sigma
        if variance_noise is None:
            variance_noise = torch.randn((num_inference_steps, n, E), generator=generator, device=dev,
                                         dtype=torch.float32)
        st = self.prepare_fused(s_embed, s_pose, t_pose, latents, float(guidance_scale), num_inference_steps,
                                variance_noise.to(dev, torch.float32).contiguous())
        out = self.prior.post_process_latents(self.replay_fused(st))
        if negative_prompt is None:
            zero = self.get_zero_embed(out.shape[0], device=out.device)
        else:
            out, zero = out.chunk(2)
        if output_type == "np":
            out, zero = out.cpu().numpy(), zero.cpu().numpy()
        if not return_dict:
            return (out, zero)
        return PriorPipelineOutput(out, zero)
