"""N > 1 host logic on CPU (gloo, world_size 2): the packed-weight arena broadcast that replaces the reference's
per-rank checkpoint load (stage2_batchtest_inpaint_model.py:103-104), and the per-rank input seeding of bench.py."""
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
