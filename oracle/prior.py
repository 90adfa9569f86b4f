"""ORACLE (test infrastructure only — never imported by pcdms_b200/): CPU restatement of the reference's stage-1 prior
(SURVEY.md §8f-4): the 6-token prior transformer, the UnCLIP scheduler it is sampled with, and the sampling loop.

  * `PriorTransformer`  — /root/reference/src/models/stage1_prior_transformer.py:50-301 (`Stage1_PriorTransformer`);
    same parameter names / state-dict keys; built from oracle/blocks.py (diffusers 0.24.0 `BasicTransformerBlock` with
    `activation_fn="gelu"`, `attention_bias=True`, no cross attention; `Timesteps`, `TimestepEmbedding`).
  * `UnCLIPScheduler`   — diffusers 0.24.0 `schedulers/scheduling_unclip.py` (un-vendored third-party dependency pinned
    at README.md:37; instantiated by `Stage1_PriorPipeline.from_pretrained` from kandinsky-2-2-prior's
    scheduler_config.json: prediction_type "sample", variance_type "fixed_small_log", clip_sample True, range 10),
    restated from its published algorithm.  Call sites: src/pipelines/stage1_prior_pipeline.py:445-446,478-483.
  * `prior_loop`        — `Stage1_PriorPipeline.__call__`, src/pipelines/stage1_prior_pipeline.py:430-490.

Pin status: `PriorTransformer` and `prior_loop` are checked BIT-EQUAL against the reference's own classes run
unmodified over oracle/diffusers_shim (tests/test_prior.py, live; tests/golden/ref_prior_tiny.pt for the GPU box).
`UnCLIPScheduler` is parity UNPINNED against diffusers itself (not installable here); it is pinned by closed forms:
the cosine alpha-bar table, the DDPM posterior-mean identity, the "fixed_small_log" standard deviation, and the
deterministic last step (tests/test_prior.py).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
from torch import nn

from .blocks import BasicTransformerBlock, TimestepEmbedding, Timesteps


class PoseMLP(nn.Module):
    """`MLP` of the reference (stage1_prior_transformer.py:18-35): Linear-GELU-(Dropout)-LN-Linear-(Dropout)-LN.
    Index positions inside `net` follow the reference's nn.Sequential so the keys agree (net.0, net.3, net.4, net.6)."""

    def __init__(self, in_dim, hidden_dim, out_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.GELU(), nn.Dropout(dropout), nn.LayerNorm(hidden_dim),
                                 nn.Linear(hidden_dim, out_dim), nn.Dropout(dropout), nn.LayerNorm(out_dim))

    def forward(self, x):
        return self.net(x)


class PriorTransformer(nn.Module):
    def __init__(self, num_attention_heads=32, attention_head_dim=64, num_layers=20, embedding_dim=768, num_embeddings=77,
                 additional_embeddings=4, dropout=0.0):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.config = dict(num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim,
                           num_layers=num_layers, embedding_dim=embedding_dim, num_embeddings=num_embeddings,
                           additional_embeddings=additional_embeddings, dropout=dropout)
        self.inner_dim = inner
        self.pose_encoder = PoseMLP(36, 512, 1024)       # :91-92 (18 keypoints x 2 coordinates; out_dim is a literal)
        self.pose_encoder1 = PoseMLP(36, 512, 1024)
        self.time_proj = Timesteps(inner, True, 0)
        self.time_embedding = TimestepEmbedding(inner, inner)
        self.proj_in = nn.Linear(embedding_dim, inner)
        self.embedding_proj = nn.Linear(embedding_dim, inner)
        self.encoder_hidden_states_proj = nn.Linear(embedding_dim, inner)
        self.encoder_hidden_states_proj1 = nn.Linear(embedding_dim, inner)
        self.positional_embedding = nn.Parameter(torch.zeros(1, num_embeddings + additional_embeddings, inner))
        self.prd_embedding = nn.Parameter(torch.zeros(1, 1, inner))
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, dropout=dropout, activation_fn="gelu",
                                  attention_bias=True) for _ in range(num_layers)])
        self.norm_out = nn.LayerNorm(inner)
        self.proj_to_clip_embeddings = nn.Linear(inner, embedding_dim)
        self.clip_mean = torch.tensor(-0.016)            # :134-135
        self.clip_std = torch.tensor(0.415)

    @property
    def dtype(self):
        return self.proj_in.weight.dtype

    def forward(self, hidden_states, timestep, proj_embedding, encoder_hidden_states, encoder_hidden_states1,
                test_flag: bool = False):
        """hidden_states [B, 1, E] (x_t), proj_embedding [B, 1, E] (source-image CLIP embedding), encoder_hidden_states
        / encoder_hidden_states1 [b, 1, 36] (source / target pose).  Returns the predicted embedding [B, E] (:200-297).
        test_flag: prepend an all-zero copy of the two pose tokens (the unconditional half) — :255-258."""
        B = hidden_states.shape[0]
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long)
        elif t.dim() == 0:
            t = t[None]
        t = t * torch.ones(B, dtype=t.dtype)
        temb = self.time_embedding(self.time_proj(t).to(self.dtype))
        src = self.embedding_proj(proj_embedding)
        pose_s = self.encoder_hidden_states_proj(self.pose_encoder(encoder_hidden_states))
        pose_t = self.encoder_hidden_states_proj1(self.pose_encoder1(encoder_hidden_states1))
        x = self.proj_in(hidden_states)
        if test_flag:
            zeros = torch.zeros_like(pose_s)
            pose_s, pose_t = torch.cat([zeros, pose_s]), torch.cat([zeros, pose_t])
        tokens = torch.cat([pose_s, pose_t, src, temb[:, None, :], x,
                            self.prd_embedding.to(x.dtype).expand(B, -1, -1)], dim=1)
        tokens = tokens + self.positional_embedding.to(x.dtype)
        for blk in self.transformer_blocks:
            tokens = blk(tokens, attention_mask=None)
        return self.proj_to_clip_embeddings(self.norm_out(tokens)[:, -1])

    def post_process_latents(self, z):
        return z * self.clip_std + self.clip_mean


# ---------------------------------------------------------------------------------------------------------------------
def cosine_betas(n: int, max_beta: float = 0.999) -> torch.Tensor:
    """diffusers `betas_for_alpha_bar` ("squaredcos_cap_v2"): abar(s) = cos^2((s + 0.008) / 1.008 * pi / 2),
    beta_i = min(1 - abar((i+1)/n) / abar(i/n), max_beta), python doubles -> fp32 tensor."""
    def abar(s):
        return math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2
    return torch.tensor([min(1 - abar((i + 1) / n) / abar(i / n), max_beta) for i in range(n)], dtype=torch.float32)


class UnCLIPScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, variance_type="fixed_small_log", clip_sample=True,
                 clip_sample_range=10.0, prediction_type="sample", beta_schedule="squaredcos_cap_v2"):
        if beta_schedule != "squaredcos_cap_v2" or variance_type != "fixed_small_log":
            raise NotImplementedError
        if prediction_type not in ("sample", "epsilon"):
            raise ValueError(prediction_type)
        from types import SimpleNamespace
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, variance_type=variance_type,
                                      clip_sample=clip_sample, clip_sample_range=clip_sample_range,
                                      prediction_type=prediction_type, beta_schedule=beta_schedule)
        self.betas = cosine_betas(num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        """Evenly spaced over [0, T-1] INCLUDING both ends (unlike DDIM's "leading" spacing)."""
        self.num_inference_steps = num_inference_steps
        ratio = (self.config.num_train_timesteps - 1) / (num_inference_steps - 1)
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device or "cpu")

    def step_scalars(self, t: int, prev_t: Optional[int]):
        """The schedule-only fp32 scalars of one step, each computed with the torch fp32 tensor ops of the published
        `step` / `_get_variance`: (coefficient of x0, coefficient of x_t, standard deviation of the added noise,
        sqrt(abar_t), sqrt(1 - abar_t))."""
        if prev_t is None:
            prev_t = t - 1
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t, b_prev = 1 - a_t, 1 - a_prev
        if prev_t == t - 1:
            beta, alpha = self.betas[t], self.alphas[t]
        else:
            beta = 1 - a_t / a_prev
            alpha = 1 - beta
        c_x0 = (a_prev ** 0.5 * beta) / b_t
        c_xt = alpha ** 0.5 * b_prev / b_t
        std = torch.tensor(0.0)
        if t > 0:
            var = b_prev / b_t * beta
            std = torch.exp(0.5 * torch.log(torch.clamp(var, min=1e-20)))       # "fixed_small_log"
        return c_x0, c_xt, std, a_t ** 0.5, b_t ** 0.5

    def step(self, model_output, timestep, sample, prev_timestep=None, generator=None, variance_noise=None):
        """x_{prev} = c_x0 * clamp(x0) + c_xt * x_t + std * noise (noise only for t > 0).  `variance_noise` lets a
        test inject the noise the reference would draw from the global generator."""
        t = int(timestep)
        prev_t = None if prev_timestep is None else int(prev_timestep)
        c_x0, c_xt, std, sa, sb = self.step_scalars(t, prev_t)
        if self.config.prediction_type == "epsilon":
            x0 = (sample - sb * model_output) / sa
        else:
            x0 = model_output
        if self.config.clip_sample:
            x0 = torch.clamp(x0, -self.config.clip_sample_range, self.config.clip_sample_range)
        prev = c_x0 * x0 + c_xt * sample
        if t > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            prev = prev + std * variance_noise
        return prev


def prior_loop(prior: PriorTransformer, scheduler: UnCLIPScheduler, *, s_embed, s_pose, t_pose, latents,
               num_inference_steps: int, guidance_scale: float = 0.0, noises=None, generator=None):
    """`Stage1_PriorPipeline.__call__` (stage1_prior_pipeline.py:430-490) for `num_images_per_prompt = 1`.
    s_embed [b, 1, E]; s_pose / t_pose [b, 1, 36]; latents [b, E].  noises: optional list of per-step variance noise.

    guidance_scale <= 1 (the batch-test driver's default 0, stage1_batchtest_prior_model.py:151) is the reference's
    loop verbatim.  With guidance_scale > 1 the reference doubles the latents and the source embedding (:337-343,457)
    but not the pose tokens, so its own `prior(...)` call fails on the batch mismatch; the loop here uses the
    transformer's `test_flag` branch (:255-258), which builds exactly the missing unconditional half (zero pose tokens
    next to the zero source embedding) — the evident intent."""
    cfg = guidance_scale > 1.0
    prompt = torch.cat([torch.zeros_like(s_embed), s_embed]) if cfg else s_embed
    scheduler.set_timesteps(num_inference_steps)
    ts = scheduler.timesteps
    latents = latents * scheduler.init_noise_sigma
    for i, t in enumerate(ts):
        x = (torch.cat([latents] * 2) if cfg else latents).unsqueeze(1)
        pred = prior(x, t, prompt, s_pose, t_pose, test_flag=cfg)
        if cfg:
            u, c = pred.chunk(2)
            pred = u + guidance_scale * (c - u)
        prev_t = None if i + 1 == ts.shape[0] else ts[i + 1]
        latents = scheduler.step(pred, t, latents, prev_timestep=prev_t, generator=generator,
                                 variance_noise=None if noises is None else noises[i])
    return prior.post_process_latents(latents)


TINY = dict(num_attention_heads=2, attention_head_dim=64, num_layers=2, embedding_dim=1024, num_embeddings=2,
            additional_embeddings=4)


def make_prior(seed: int = 0, **cfg) -> PriorTransformer:
    """Seeded synthetic weights (torch default Linear init; positional / query embeddings, LayerNorm affines and
    biases made non-trivial so that every term of the forward is exercised)."""
    torch.manual_seed(seed)
    m = PriorTransformer(**cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n in ("positional_embedding", "prd_embedding"):
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
            elif "norm" in n or ".net.3." in n or ".net.6." in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith(".bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def make_prior_inputs(n: int = 1, seed: int = 0, steps: int = 4, embedding_dim: int = 1024):
    """Seeded inputs of one pipeline call: CLIP-like source embedding, normalised 18-keypoint poses in [0, 1] (the
    batch-test driver reads them from text files, stage1_batchtest_prior_model.py:20-28,84-85), initial latents and the
    per-step variance noise."""
    g = torch.Generator().manual_seed(seed)
    return dict(s_embed=torch.randn(n, 1, embedding_dim, generator=g), s_pose=torch.rand(n, 1, 36, generator=g),
                t_pose=torch.rand(n, 1, 36, generator=g), latents=torch.randn(n, embedding_dim, generator=g),
                noises=torch.randn(steps, n, embedding_dim, generator=g))
