"""N > 1 host logic on CPU (gloo, world_size 2): the packed-weight arena broadcast that replaces the reference's
per-rank checkpoint loads (stage2_batchtest_inpaint_model.py:103-104; stage1_batchtest_prior_model.py:53-61) — for
the UNet and for every other model class of the widened rows (VAE, DINOv2, CLIP, prior, conditioning modules) — and
the per-rank input seeding of bench.py."""
import hashlib
import os
import socket
from dataclasses import asdict

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.unet import UNetConfig
        from pcdms_b200.unet import B200UNet2DConditionModel
        m = B200UNet2DConditionModel(dtype=torch.float16, device="cpu", **asdict(UNetConfig.tiny()))
        if rank == 0:
            m.load_state_dict(m.synthetic_state_dict(seed=0, device="cpu"))
            m.consolidate()
        assert (len(m._w) > 0) == (rank == 0)
        nbytes = m.broadcast_weights(src=0)
        digest = hashlib.sha256(m._arena.numpy().tobytes()).hexdigest()
        keys = sorted(m._w.keys())
        sample = m._w["conv_in.weight"].float().abs().sum().item()
        # every packed tensor must be a view into the one arena
        base = m._arena.data_ptr()
        inside = all(base <= t.data_ptr() < base + nbytes for t in m._w.values())
        gathered = [None] * world
        dist.all_gather_object(gathered, (digest, len(keys), sample, inside, m._loaded))
        if rank == 0:
            torch.save(gathered, out)
    finally:
        dist.destroy_process_group()


def test_weight_arena_broadcast_gloo(tmp_path):
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    g = torch.load(out)
    assert g[0] == g[1], "rank 1 must hold bit-identical packed weights after one broadcast"
    digest, nkeys, sample, inside, loaded = g[0]
    assert nkeys > 600 and sample > 0 and inside and loaded


def test_bench_rank_inputs_are_independent_and_reproducible():
    import bench
    a0, b0, a1 = bench.host_inputs(0, 128, 64), bench.host_inputs(0, 128, 64), bench.host_inputs(1, 128, 64)
    assert all(torch.equal(a0[k], b0[k]) for k in a0)                 # same rank -> same shard, run to run
    assert not torch.equal(a0["latents"], a1["latents"])              # different ranks -> different images
    assert a0["latents"].shape == (bench.N_IMAGES, 4, bench.LAT_H, bench.LAT_W)
    assert a0["masked_latents"][..., bench.LAT_W // 2:].abs().sum() == 0


def _tiny_models():
    """One tiny instance of every model class that carries packed weights (device 'cpu': packing is host logic)."""
    from oracle.prior import TINY
    from oracle.vae import VAEConfig
    from pcdms_b200.clip import B200CLIPVisionModelWithProjection
    from pcdms_b200.dinov2 import B200Dinov2Model
    from pcdms_b200.frontend import B200ControlNetConditioningEmbedding, B200ImageProjModel_p
    from pcdms_b200.prior import B200Stage1PriorTransformer
    from pcdms_b200.vae import B200AutoencoderKL
    dt = torch.float16
    return {
        "vae": B200AutoencoderKL(dtype=dt, device="cpu", **asdict(VAEConfig.tiny())),
        "dinov2": B200Dinov2Model(dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, mlp_ratio=4,
                                       image_size=56, patch_size=14), dtype=dt, device="cpu"),
        "clip": B200CLIPVisionModelWithProjection(dict(hidden_size=320, intermediate_size=1280, num_hidden_layers=2,
                                                       num_attention_heads=4, image_size=56, patch_size=14,
                                                       projection_dim=64), dtype=dt, device="cpu"),
        "prior": B200Stage1PriorTransformer(dtype=dt, device="cpu", **TINY),
        "image_proj": B200ImageProjModel_p(in_dim=128, hidden_dim=64, out_dim=96, dtype=dt, device="cpu"),
        "pose_proj": B200ControlNetConditioningEmbedding(64, 3, (16, 32), dtype=dt, device="cpu"),
    }


def _worker_all(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        for name, m in _tiny_models().items():
            if rank == 0:
                m.load_state_dict(m.synthetic_state_dict(seed=1, device="cpu"))
            assert m._loaded == (rank == 0)
            nbytes = m.broadcast_weights(src=0)
            base = m._arena.data_ptr()
            res[name] = (hashlib.sha256(m._arena.numpy().tobytes()).hexdigest(), sorted(m._w.keys()), m._loaded,
                         all(base <= t.data_ptr() < base + nbytes for t in m._w.values()))
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        if rank == 0:
            torch.save(gathered, out)
    finally:
        dist.destroy_process_group()


def test_every_model_class_ships_its_weights_with_one_broadcast_gloo(tmp_path):
    out = str(tmp_path / "gathered_all.pt")
    mp.spawn(_worker_all, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out)
    assert set(r0) == {"vae", "dinov2", "clip", "prior", "image_proj", "pose_proj"}
    for name in r0:
        assert r0[name] == r1[name], f"{name}: rank 1 must hold bit-identical packed weights"
        digest, keys, loaded, inside = r0[name]
        assert loaded and inside and len(keys) > 3
    assert "_pos_raw" in r0["clip"][1] and "_pos_raw" in r0["dinov2"][1]   # everything forward() needs is in the arena
