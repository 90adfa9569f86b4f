"""GPU parity of every sm_100a kernel against the CPU oracle ops (torch fp32 on the SAME 16-bit-rounded inputs), called
through the C ABI (pcdms_b200.ops -> libpcdm_b200.so).  Tolerance: rtol 1e-3 / atol 1e-4 (BASELINE north_star) for
fp16 kernels whose only rounding is the final 16-bit store; attention additionally rounds the probabilities P to
16 bits before the PV product (as every tensor-core flash attention does), so it gets 4x that; bf16 has 8x coarser
mantissa than fp16, so bf16 runs use 8x the fp16 tolerance."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 1e-4
DTS = [torch.float16, torch.bfloat16]


def tol(dt, mult=1.0):
    k = mult * (8.0 if dt == torch.bfloat16 else 1.0)
    return dict(rtol=RTOL * k, atol=ATOL * k)


def close(got, want, dt, mult=1.0):
    torch.testing.assert_close(got.float().cpu(), want.float(), **tol(dt, mult))


@pytest.fixture(scope="module")
def ops():
    from pcdms_b200 import ops as _ops
    _ops.ensure_workspace("cuda")   # enables split-K for the tile-starved shapes below
    return _ops


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("M,N,K,bn,kw", [
    (256, 256, 128, 128, {}), (1000, 320, 320, 0, {"resid": True}), (4096, 1280, 1024, 256, {}),
    (516, 640, 1024, 64, {"bias": False}), (300, 2560, 320, 0, {"geglu": True}), (512, 320, 960, 160, {"split": 640}),
    (2, 1280, 320, 0, {"silu": True}), (4128, 640, 1024, 0, {"bias": False}),
    # K <= 320 with >= 2 tiles per CTA pair: the activation-stationary instances (contiguous tile runs, resident A slab,
    # ragged last m-tile; GEGLU additionally on the 16-epilogue-warp instance)
    (20000, 640, 320, 0, {"resid": True}), (33000, 2560, 320, 0, {"geglu": True}), (16448, 960, 256, 0, {}),
    (40000, 320, 64, 0, {"bias": False}), (32768, 320, 320, 0, {"resid": True, "silu": True}),
])
def test_gemm(ops, dt, M, N, K, bn, kw):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dt)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dt)
    b = torch.randn(N, generator=g) if kw.get("bias", True) else None
    r = torch.randn(M, N, generator=g).to(dt) if kw.get("resid") else None
    ref = a.float() @ w.float().t()
    if b is not None:
        ref = ref + b
    if r is not None:
        ref = ref + r.float()
    if kw.get("silu"):
        ref = F.silu(ref)
    wd, bd = w, b
    if kw.get("geglu"):
        h, gate = ref.chunk(2, dim=1)
        ref = h * F.gelu(gate)
        perm = ops.geglu_row_permutation(N // 2)
        wd, bd = w[perm].contiguous(), b[perm].contiguous()
    ad = a.cuda()
    extra = {}
    if kw.get("split"):
        extra["a2"] = ad[:, kw["split"]:]
        ad = ad[:, :kw["split"]]
    out = ops.gemm(ad, wd.cuda(), bias=bd.cuda() if bd is not None else None,
                   residual=r.cuda() if r is not None else None, geglu=bool(kw.get("geglu")),
                   silu=bool(kw.get("silu")), bn=bn, **extra)
    close(out, ref, dt)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,extra", [
    (2, 32, 64, 320, 320, 1, True), (2, 16, 32, 640, 640, 1, True), (2, 4, 8, 1280, 1280, 1, True),
    (6, 4, 8, 128, 64, 1, True), (1, 64, 128, 64, 64, 1, False), (2, 16, 32, 320, 320, 2, False),
    (3, 4, 8, 128, 128, 2, False), (2, 8, 16, 1920, 1280, 1, True), (2, 2, 4, 64, 32, 1, False),
])
def test_conv3x3(ops, dt, B, H, W, Cin, Cout, stride, extra):
    g = torch.Generator().manual_seed(B * H + Cin)
    x = torch.randn(B, Cin, H * stride, W * stride, generator=g).to(dt)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dt)
    b = torch.randn(Cout, generator=g)
    t = torch.randn(B, 3 * Cout, generator=g) if extra else None
    r = torch.randn(B, Cout, H, W, generator=g).to(dt) if extra else None
    ref = F.conv2d(x.float(), w.float(), b, stride=stride, padding=1)
    if extra:
        ref = ref + t[:, Cout:2 * Cout, None, None] + r.float()
    out = ops.conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), ops.pack_conv3x3_weight(w, dt).cuda(), bias=b.cuda(),
                      rowvec=t.cuda()[:, Cout:2 * Cout] if extra else None,
                      residual=r.permute(0, 2, 3, 1).contiguous().cuda() if extra else None, stride=stride)
    close(out.permute(0, 3, 1, 2), ref, dt)


@pytest.mark.parametrize("dt", DTS)
def test_pair_tiles_match_single_cta_tiles_bit_for_bit(ops, dt):
    """The N = 320 convs / GEMMs of the 32x64 level at batch 16: 256 CTA-pair tiles of 256 x 160 on 74 pairs (the automatic
    choice) against 512 single-CTA tiles of 128 x 160 (forced).  Same arithmetic per output element, so outputs AND the
    GroupNorm statistics of the epilogue must agree bit for bit, and with the oracle.  (Written for the remainder-split
    instance of round 2 — the last, partly filled wave's tiles cut into 96 + 64 columns — which passed it and was
    measured not faster: profiles/r2_s3_epilogue.md.)"""
    g = torch.Generator().manual_seed(3)
    B, H, W, Cin, Cout = 16, 32, 64, 64, 320
    x = torch.randn(B, H, W, Cin, generator=g).to(dt)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dt)
    b, t = torch.randn(Cout, generator=g), torch.randn(B, Cout, generator=g)
    r = torch.randn(B, H, W, Cout, generator=g).to(dt)
    wp = ops.pack_conv3x3_weight(w, dt).cuda()
    args = dict(bias=b.cuda(), rowvec=t.cuda(), residual=r.cuda())
    y, st = ops.conv3x3(x.cuda(), wp, chan_stats=True, **args)                 # automatic: pair tiles
    y1, st1 = ops.conv3x3(x.cuda(), wp, chan_stats=True, bn=160, cta_group=1, **args)   # 512 single-CTA tiles
    assert torch.equal(y, y1) and torch.equal(st.buf, st1.buf)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1) + t[:, :, None, None] + r.float().permute(0, 3, 1, 2)
    close(y.permute(0, 3, 1, 2), ref, dt)
    torch.testing.assert_close(st.buf.double().cpu(), _slab_sums(y.view(-1, Cout)), rtol=2e-5, atol=2e-4)
    # K = 5760 at N = 320 (conv1 of the 32x64 up-block resnets: 640 -> 320): the automatic choice is the 320-wide pair tile
    x6 = torch.randn(B, H, W, 640, generator=g).to(dt).cuda()
    w6 = (torch.randn(Cout, 9 * 640, generator=g) / (9 * 640) ** 0.5).to(dt).cuda()
    y6, st6 = ops.conv3x3(x6, w6, bias=b.cuda(), rowvec=t.cuda(), chan_stats=True)
    y7, st7 = ops.conv3x3(x6, w6, bias=b.cuda(), rowvec=t.cuda(), chan_stats=True, bn=160, cta_group=1)
    assert torch.equal(y6, y7) and torch.equal(st6.buf, st7.buf)
    # plain GEMM with residual (ff.net.2 of the 32x64 transformer blocks: M 32768, N 320, K 1280), ragged M
    M, N, K = 32768 - 96, 320, 256
    a, wg, rg = torch.randn(M, K, generator=g).to(dt), (torch.randn(N, K, generator=g) / K ** 0.5).to(dt), torch.randn(M, N, generator=g).to(dt)
    o = ops.gemm(a.cuda(), wg.cuda(), bias=b.cuda(), residual=rg.cuda())
    assert torch.equal(o, ops.gemm(a.cuda(), wg.cuda(), bias=b.cuda(), residual=rg.cuda(), bn=160, cta_group=1))
    close(o, a.float() @ wg.float().t() + b + rg.float(), dt)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(16, 4, 8, 1280, 1280), (2, 8, 16, 2560, 1280), (2, 4, 8, 1920, 640)])
def test_conv3x3_split_k_matches_single_pass(ops, B, H, W, Cin, Cout):
    """Tile-starved shapes take the split-K route (fp32 partials + finishing kernel); forcing bn disables it.  Both
    must agree with the oracle, and with each other to fp32-accumulation-order noise."""
    dt = torch.float16
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(B, Cin, H, W, generator=g).to(dt)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dt)
    b = torch.randn(Cout, generator=g)
    t = torch.randn(B, Cout, generator=g)
    r = torch.randn(B, Cout, H, W, generator=g).to(dt)
    ref = F.conv2d(x.float(), w.float(), b, padding=1) + t[:, :, None, None] + r.float()
    xd, wd = x.permute(0, 2, 3, 1).contiguous().cuda(), ops.pack_conv3x3_weight(w, dt).cuda()
    rd = r.permute(0, 2, 3, 1).contiguous().cuda()
    auto = ops.conv3x3(xd, wd, bias=b.cuda(), rowvec=t.cuda(), residual=rd)
    single = ops.conv3x3(xd, wd, bias=b.cuda(), rowvec=t.cuda(), residual=rd, bn=128)
    close(auto.permute(0, 3, 1, 2), ref, dt)
    close(single.permute(0, 3, 1, 2), ref, dt)
    torch.testing.assert_close(auto.float(), single.float(), rtol=2e-3, atol=2e-3)
    assert torch.equal(auto, ops.conv3x3(xd, wd, bias=b.cuda(), rowvec=t.cuda(), residual=rd))  # deterministic


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(16, 4, 8, 1280, 1280), (2, 8, 16, 1280, 1280), (2, 16, 32, 640, 640),
                                           (3, 2, 4, 64, 64), (2, 32, 32, 64, 128), (5, 4, 8, 128, 192)])
def test_conv3x3_up2x(ops, dt, B, H, W, Cin, Cout):
    """diffusers Upsample2D = F.interpolate(nearest, x2) + Conv2d(3, padding 1) as ONE launch of four per-parity 2x2
    convolutions over the low-resolution input (no 4x tensor, 16/36 of the MACs).
    (1) kernel arithmetic: against the same decomposition in fp32 on the kernel's own packed 16-bit weights, at the
        single-op tolerance;
    (2) the layer: against interpolate + conv2d with the fp32 MASTER weights.  Any 16-bit conv rounds its weights once
        (here: the pre-summed taps; conventionally: the nine taps), so the yardstick is the conventional 16-bit path —
        pcdm_upsample_nearest2x + pcdm_conv3x3 on the rounded 3x3 weights — whose error this one must not exceed."""
    g = torch.Generator().manual_seed(B * H + Cin + Cout)
    x = torch.randn(B, Cin, H, W, generator=g).to(dt)
    w32 = torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, generator=g)
    wp = ops.pack_upsample_conv_weight(w32, dt)
    assert wp.shape == (4, Cout, 4 * Cin)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = ops.conv3x3_up2x(xd, wp.cuda(), bias=b.cuda())
    assert out.shape == (B, 2 * H, 2 * W, Cout)
    xin = F.pad(x.float(), (1, 1, 1, 1))
    ref_same = torch.zeros(B, Cout, 2 * H, 2 * W)
    for py in (0, 1):
        for px in (0, 1):
            wk = wp[2 * py + px].float().view(Cout, 2, 2, Cin).permute(0, 3, 1, 2)
            ref_same[:, :, py::2, px::2] = F.conv2d(xin[:, :, py:py + H + 1, px:px + W + 1], wk, b)
    close(out.permute(0, 3, 1, 2), ref_same, dt)
    ref_true = F.conv2d(F.interpolate(x.float(), scale_factor=2, mode="nearest"), w32, b, padding=1)
    conventional = ops.conv3x3(ops.upsample_nearest2x(xd), ops.pack_conv3x3_weight(w32, dt).cuda(), bias=b.cuda())
    err = (out.permute(0, 3, 1, 2).float().cpu() - ref_true).abs()
    err_conv = (conventional.permute(0, 3, 1, 2).float().cpu() - ref_true).abs()
    assert err.max() <= 1.5 * err_conv.max() + tol(dt)["atol"] and err.mean() <= 1.25 * err_conv.mean() + 1e-6
    assert torch.equal(out, ops.conv3x3_up2x(xd, wp.cuda(), bias=b.cuda()))


def _slab_sums(y_rows):
    """[M, N] stored rows -> [M / 32, N, 2] (sum, sum of squares) per 32-row slab, fp64 on the CPU."""
    M, N = y_rows.shape
    yf = y_rows.double().cpu().view(M // 32, 32, N)
    return torch.stack([yf.sum(1), (yf ** 2).sum(1)], dim=-1)


@pytest.mark.parametrize("dt", DTS)
def test_groupnorm_statistics_from_producer_epilogues(ops, dt):
    """NS-1: the GroupNorm statistics are emitted by the epilogue of the conv / GEMM that PRODUCES the tensor
    (pcdm_ext.chan_stats), for every kind of producer on the UNet path — conv3x3 (+bias+temb+residual), stride-2 conv,
    the split-K route, the one-launch Upsample2D, the residual GEMM (proj_out) — and pcdm_groupnorm_apply normalises from
    them (incl. the skip concat with groups straddling the two sources).  Statistics are of the values AS STORED."""
    g = torch.Generator().manual_seed(11)

    def rnd(*shape, scale=1.0):
        return (scale * torch.randn(*shape, generator=g)).to(dt)

    def check_stats(st, y_rows, order=None):
        assert st is not None
        rows = y_rows if order is None else y_rows[order]
        want = _slab_sums(rows)
        assert st.buf.shape == want.shape
        torch.testing.assert_close(st.buf.double().cpu(), want, rtol=2e-5, atol=2e-4)

    # conv3x3 + bias + per-image vector + residual (a resnet's conv2), 2 images of 16 x 32, rows not a tile multiple
    B, H, W, Cin, Cout = 3, 8, 16, 128, 320
    x, wt = rnd(B, H, W, Cin), rnd(Cout, 9 * Cin, scale=(9 * Cin) ** -0.5)
    bias, tv, r = torch.randn(Cout, generator=g), torch.randn(B, Cout, generator=g), rnd(B, H, W, Cout)
    y1, s1 = ops.conv3x3(x.cuda(), wt.cuda(), bias=bias.cuda(), rowvec=tv.cuda(), residual=r.cuda(), chan_stats=True)
    assert torch.equal(y1, ops.conv3x3(x.cuda(), wt.cuda(), bias=bias.cuda(), rowvec=tv.cuda(), residual=r.cuda()))
    check_stats(s1, y1.view(-1, Cout))
    # stride 2
    y2, s2 = ops.conv3x3(rnd(2, 16, 32, 64).cuda(), rnd(64, 9 * 64, scale=1 / 24).cuda(), stride=2, chan_stats=True)
    check_stats(s2, y2.view(-1, 64))
    # a tile-starved shape (16 images of 4 x 8, K = 9 * 1280): with statistics the launch stays off the split-K route
    xs, ws, bs = _splitk_problem("cuda", 5)
    y3, s3 = ops.conv3x3(xs, ws, bias=bs, chan_stats=True)
    close(y3, ops.conv3x3(xs, ws, bias=bs).cpu(), dt if dt == torch.float16 else torch.float16, mult=2.0)
    check_stats(s3, y3.view(-1, 1280))
    # one-launch Upsample2D: slabs ordered [image][parity plane][32 low-resolution pixels]
    xu = rnd(2, 8, 16, 64)
    wu = ops.pack_upsample_conv_weight(torch.randn(128, 64, 3, 3, generator=g) / 24, dt)
    y4, s4 = ops.conv3x3_up2x(xu.cuda(), wu.cuda(), bias=torch.randn(128, generator=g).cuda(), chan_stats=True)
    planes = torch.stack([y4[:, py::2, px::2] for py in (0, 1) for px in (0, 1)], dim=1)
    check_stats(s4, planes.reshape(-1, 128))
    assert s4.hw == 4 * 8 * 16
    # residual GEMM with rows_per_image (a transformer's proj_out)
    a, wg, rg = rnd(B * H * W, 192), rnd(Cout, 192, scale=192 ** -0.5), rnd(B * H * W, Cout)
    y5, s5 = ops.gemm(a.cuda(), wg.cuda(), bias=bias.cuda(), residual=rg.cuda(), rows_per_image=H * W, chan_stats=True)
    check_stats(s5, y5)
    # shapes that cannot carry slab statistics hand back None (the consumer then runs the stand-alone kernels)
    assert ops.conv3x3(rnd(2, 2, 4, 64).cuda(), rnd(64, 576, scale=1 / 24).cuda(), chan_stats=True)[1] is None
    # consumer: GroupNorm(32) (+SiLU) over [y1 | y5 reshaped]: 640 channels, 20 per group; and over a concat whose groups
    # straddle the two sources (320 + 128 = 448 channels, 14 per group: group 22 takes 12 from y1 and 2 from y4')
    gamma, beta = 1 + 0.1 * torch.randn(640, generator=g), 0.1 * torch.randn(640, generator=g)
    xa, xb = y1, y5.view(B, H, W, Cout)
    got = ops.groupnorm(xa, gamma.cuda(), beta.cuda(), 1e-5, x2=xb, silu=True, stats=(s1, s5))
    ref = F.silu(F.group_norm(torch.cat([xa, xb], -1).float().cpu().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5))
    close(got.permute(0, 3, 1, 2), ref, dt)
    close(got, ops.groupnorm(xa, gamma.cuda(), beta.cuda(), 1e-5, x2=xb, silu=True).cpu(), dt)   # the stand-alone path
    y6, s6 = ops.conv3x3(x.cuda(), rnd(128, 9 * Cin, scale=(9 * Cin) ** -0.5).cuda(), chan_stats=True)
    g2, b2 = 1 + 0.1 * torch.randn(448, generator=g), 0.1 * torch.randn(448, generator=g)
    got = ops.groupnorm(y1, g2.cuda(), b2.cuda(), 1e-6, x2=y6, stats=(s1, s6))
    ref = F.group_norm(torch.cat([y1, y6], -1).float().cpu().permute(0, 3, 1, 2), 32, g2, b2, 1e-6)
    close(got.permute(0, 3, 1, 2), ref, dt)
    assert torch.equal(got, ops.groupnorm(y1, g2.cuda(), b2.cuda(), 1e-6, x2=y6, stats=(s1, s6)))   # reproducible
    # single source, the up-sampled tensor (its slab order differs, the per-image fold must not care)
    g3, b3 = 1 + 0.1 * torch.randn(128, generator=g), 0.1 * torch.randn(128, generator=g)
    got = ops.groupnorm(y4, g3.cuda(), b3.cuda(), 1e-5, silu=True, stats=(s4, None))
    close(got.permute(0, 3, 1, 2), F.silu(F.group_norm(y4.float().cpu().permute(0, 3, 1, 2), 32, g3, b3, 1e-5)), dt)


def _splitk_problem(dev, seed):
    dt = torch.float16
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(16, 4, 8, 1280, generator=g).to(dt).to(dev)
    w = (torch.randn(1280, 9 * 1280, generator=g) / (9 * 1280) ** 0.5).to(dt).to(dev)
    b = torch.randn(1280, generator=g).to(dev)
    return x, w, b


def test_two_streams_with_their_own_workspaces_are_independent(ops):
    """The C ABI carries no process-wide state: split-K scratch is a per-call argument.  Two
    streams running the split-K conv concurrently, each with its own workspace, give exactly the serial results."""
    dev = torch.device("cuda", torch.cuda.current_device())
    probs = [_splitk_problem(dev, s) for s in (1, 2)]
    serial = [ops.conv3x3(x, w, bias=b) for x, w, b in probs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    wss = [torch.zeros(ops.WORKSPACE_BYTES, dtype=torch.uint8, device=dev) for _ in streams]
    outs = [[], []]
    torch.cuda.synchronize()
    for it in range(20):                      # interleaved enqueue: the two streams overlap on the device
        for k, (s, ws) in enumerate(zip(streams, wss)):
            with torch.cuda.stream(s), ops.use_workspace(ws):
                x, w, b = probs[k]
                outs[k].append(ops.conv3x3(x, w, bias=b))
    torch.cuda.synchronize()
    for k in range(2):
        for o in outs[k]:
            assert torch.equal(o, serial[k])


def test_two_devices_from_one_process(ops):
    """Per-device kernel attributes / SM counts / workspaces: the same problem on cuda:0 and cuda:1 from one process
    is bit-identical (needs a >= 2-GPU lease; the round-end 8-GPU box runs it)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    outs = []
    for d in (0, 1):
        with torch.cuda.device(d):
            x, w, b = _splitk_problem(torch.device("cuda", d), 3)
            q = torch.randn(2 * 256, 3 * 320, generator=torch.Generator().manual_seed(4)).half().to(f"cuda:{d}")
            o1 = ops.conv3x3(x, w, bias=b)
            o2 = ops.attention(q[:, :320], q[:, 320:640], q[:, 640:], 2, 5)
            torch.cuda.synchronize()
            outs.append((o1.cpu(), o2.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_conv3x3_rejects_unsupported_shapes(ops):
    """The C entry point itself still refuses output sizes its tile map does not cover (the host wrapper pads them)."""
    import ctypes as C
    from pcdms_b200 import lib as _l
    L = _l.load()
    x = torch.zeros(1, 8, 24, 64, device="cuda", dtype=torch.float16)  # W = 24 does not divide 128
    w = torch.zeros(64, 9 * 64, device="cuda", dtype=torch.float16)
    out = torch.empty(1, 8, 24, 64, device="cuda", dtype=torch.float16)
    rc = L.pcdm_conv3x3(_l.ptr(x), _l.ptr(w), _l.ptr(out), None, None, C.c_longlong(0), None, C.c_int(1), C.c_int(8),
                        C.c_int(24), C.c_int(64), C.c_int(64), C.c_int(1), C.c_int(0), C.c_int(0), C.c_int(0), None, None)
    assert rc == _l.ERR_UNSUPPORTED


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,H,W,Cin,Cout,stride", [(2, 12, 24, 64, 64, 1), (2, 6, 12, 128, 64, 1), (1, 3, 6, 64, 128, 1),
                                                   (2, 48, 96, 64, 32, 1), (2, 12, 24, 64, 64, 2), (1, 5, 3, 64, 64, 1),
                                                   (1, 11, 22, 64, 64, 2), (1, 40, 200, 64, 32, 1)])
def test_conv3x3_any_canvas(ops, dt, B, H, W, Cin, Cout, stride):
    """Latent sizes outside the tile map (widths that neither divide 128 nor are multiples of it, row counts that do not
    fill whole tiles) — the reference accepts any canvas divisible by 8 (stage2_batchtest_inpaint_model.py:258-260).
    The wrapper runs them on the next covered size: appended zeros ARE the convolution's zero padding."""
    g = torch.Generator().manual_seed(H * W + Cin)
    x = torch.randn(B, Cin, H * stride, W * stride, generator=g).to(dt)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dt)
    b = torch.randn(Cout, generator=g)
    t = torch.randn(B, Cout, generator=g)
    r = torch.randn(B, Cout, H, W, generator=g).to(dt)
    ref = F.conv2d(x.float(), w.float(), b, stride=stride, padding=1) + t[:, :, None, None] + r.float()
    out, st = ops.conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), ops.pack_conv3x3_weight(w, dt).cuda(), bias=b.cuda(),
                          rowvec=t.cuda(), residual=r.permute(0, 2, 3, 1).contiguous().cuda(), stride=stride,
                          chan_stats=True)
    assert st is None and out.shape == (B, H, W, Cout)
    close(out.permute(0, 3, 1, 2), ref, dt)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 6, 12, 64, 64), (1, 12, 24, 128, 64), (1, 3, 6, 64, 64)])
def test_conv3x3_up2x_any_canvas(ops, B, H, W, Cin, Cout):
    dt = torch.float16
    g = torch.Generator().manual_seed(H * W + Cout)
    x = torch.randn(B, Cin, H, W, generator=g).to(dt)
    w32 = torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, generator=g)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = ops.conv3x3_up2x(xd, ops.pack_upsample_conv_weight(w32, dt).cuda(), bias=b.cuda())
    assert out.shape == (B, 2 * H, 2 * W, Cout)
    ref = F.conv2d(F.interpolate(x.float(), scale_factor=2, mode="nearest"), w32, b, padding=1)
    torch.testing.assert_close(out.permute(0, 3, 1, 2).float().cpu(), ref, rtol=4e-3, atol=4e-3)   # taps pre-summed in 16 bits


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,H,W,C1,C2", [(2, 32, 64, 320, 0), (2, 16, 32, 1280, 640), (3, 4, 8, 1280, 1280),
                                          (2, 8, 16, 640, 320), (1, 64, 128, 320, 0), (2, 2, 4, 64, 0),
                                          (16, 32, 64, 640, 320), (16, 16, 32, 640, 0), (16, 8, 16, 1280, 0),
                                          (5, 4, 8, 1280, 0), (1, 1, 8, 512, 0), (3, 12, 20, 96, 160)])
@pytest.mark.parametrize("silu", [True, False])
def test_groupnorm(ops, dt, B, H, W, C1, C2, silu):
    g = torch.Generator().manual_seed(C1 + C2 + H)
    C = C1 + C2
    x1 = (torch.randn(B, C1, H, W, generator=g) * 2 + 0.5).to(dt)
    x2 = (torch.randn(B, C2, H, W, generator=g) - 0.3).to(dt) if C2 else None
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    xc = torch.cat([x1, x2], 1) if C2 else x1
    ref = F.group_norm(xc.float(), 32, gamma, beta, 1e-5)   # straddling groups: 1920/32 = 60 ch, 1280/60 not integral
    if silu:
        ref = F.silu(ref)
    out = ops.groupnorm(x1.permute(0, 2, 3, 1).contiguous().cuda(), gamma.cuda(), beta.cuda(), 1e-5,
                        x2=x2.permute(0, 2, 3, 1).contiguous().cuda() if C2 else None, silu=silu)
    close(out.permute(0, 3, 1, 2), ref, dt)
    # the single-pass (register-resident, cluster-reduced) path and the statistics + apply path agree to rounding,
    # and each is bit-reproducible
    args = (x1.permute(0, 2, 3, 1).contiguous().cuda(), gamma.cuda(), beta.cuda(), 1e-5)
    kw = dict(x2=x2.permute(0, 2, 3, 1).contiguous().cuda() if C2 else None, silu=silu)
    assert torch.equal(out, ops.groupnorm(*args, **kw))
    two = ops.groupnorm(*args, path="two_pass", **kw)
    one = ops.groupnorm(*args, path="one_pass", **kw)
    assert torch.equal(one, ops.groupnorm(*args, path="one_pass", **kw))
    close(two.permute(0, 3, 1, 2), ref, dt)
    close(one.permute(0, 3, 1, 2), ref, dt)
    torch.testing.assert_close(one.float(), two.float(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("M,C", [(4096, 320), (1000, 640), (77, 1280), (33, 64), (32768, 320), (5000, 1280), (9, 2048)])
def test_layernorm(ops, dt, M, C):
    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 3 + 1).to(dt)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    close(ops.layernorm(x.cuda(), gamma.cuda(), beta.cuda()), ref, dt)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,heads,Sq,Skv,sc", [
    (2, 5, 2048, 2048, 1.0), (2, 10, 512, 512, 1.0), (2, 20, 128, 128, 1.0), (3, 20, 32, 32, 1.0),  # self (a9)
    (2, 5, 2048, 258, 1.0), (2, 10, 512, 95, 1.0), (1, 20, 128, 257, 1.0),                          # cross
    (1, 2, 384, 300, 4.0), (1, 1, 200, 130, 8.0), (1, 5, 8192, 8192, 1.0),                          # ragged / sharp / cfg 3
])
def test_attention(ops, dt, B, heads, Sq, Skv, sc):
    g = torch.Generator().manual_seed(Sq + Skv + heads)
    C = heads * 64
    q = (torch.randn(B, Sq, C, generator=g) * sc).to(dt)
    k = (torch.randn(B, Skv, C, generator=g) * sc).to(dt)
    v = torch.randn(B, Skv, C, generator=g).to(dt)
    qh, kh, vh = (t.float().view(B, -1, heads, 64).transpose(1, 2) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, Sq, C)
    if Sq == Skv:  # q/k/v as column slices of one fused projection buffer, like the UNet's self-attention
        qkv = torch.cat([q, k, v], dim=-1).reshape(B * Sq, 3 * C).cuda()
        out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads)
    else:
        out = ops.attention(q.reshape(B * Sq, C).cuda(), k.reshape(B * Skv, C).cuda(), v.reshape(B * Skv, C).cuda(),
                            B, heads)
    close(out.view(B, Sq, C), ref, dt, mult=4.0)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,heads,Sq,Skv", [(2, 5, 1024, 1024), (1, 2, 384, 300), (2, 3, 200, 2048), (1, 3, 257, 257)])
def test_attention_growing_row_maxima(ops, dt, B, heads, Sq, Skv):
    """Key norms ramp up along the sequence, so every later K/V block brings much larger scores: the two-tile kernel's
    speculative exponentials (taken against the row's previous reference maximum) must detect the growth, redo the
    block from the scores still in TMEM and rescale O / l — also across ragged last blocks and odd tile counts."""
    g = torch.Generator().manual_seed(7 * Sq + Skv)
    C = heads * 64
    q = torch.randn(B, Sq, C, generator=g).to(dt)
    k = (torch.randn(B, Skv, C, generator=g) * torch.linspace(0.2, 6.0, Skv)[None, :, None]).to(dt)
    v = torch.randn(B, Skv, C, generator=g).to(dt)
    qh, kh, vh = (t.float().view(B, -1, heads, 64).transpose(1, 2) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, Sq, C)
    out = ops.attention(q.reshape(B * Sq, C).cuda(), k.reshape(B * Skv, C).cuda(), v.reshape(B * Skv, C).cuda(), B, heads)
    # sharp rows (|scores| up to ~40): the error is the 16-bit rounding of P and of the output, measured 2e-3 (fp16) /
    # 2e-2 (bf16) for both attention kernels
    bound = 6e-3 if dt == torch.float16 else 5e-2
    assert (out.view(B, Sq, C).float().cpu() - ref).abs().max().item() < bound
    # decreasing norms: the first block sets a reference maximum that is never exceeded again
    k2 = k.flip(1).contiguous()
    ref2 = F.scaled_dot_product_attention(qh, k2.float().view(B, -1, heads, 64).transpose(1, 2), vh).transpose(1, 2).reshape(B, Sq, C)
    out2 = ops.attention(q.reshape(B * Sq, C).cuda(), k2.reshape(B * Skv, C).cuda(), v.reshape(B * Skv, C).cuda(), B, heads)
    assert (out2.view(B, Sq, C).float().cpu() - ref2).abs().max().item() < bound


def test_attention_more_items_than_sms_is_deterministic(ops):
    """The persistent kernel at a work list longer than one round (B 16, 5 heads, 2048 tokens: 640 tile pairs on 148
    CTAs, the last partial round issued as single tiles) against the fp32 reference on a sample of rows, and bit-equal
    run to run."""
    dt = torch.bfloat16
    B, heads, S = 16, 5, 2048
    g = torch.Generator().manual_seed(11)
    qkv = torch.randn(B * S, 3 * heads * 64, generator=g).to(dt).cuda()
    C = heads * 64
    a = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads)
    b = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads)
    assert torch.equal(a, b)
    for bi in (0, 7, 15):
        rows = slice(bi * S, (bi + 1) * S)
        qh, kh, vh = (qkv[rows, i * C:(i + 1) * C].float().view(1, S, heads, 64).transpose(1, 2) for i in range(3))
        ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(S, C)
        close(a[rows], ref.cpu(), dt, mult=4.0)


def test_boundary_and_misc_kernels(ops):
    from oracle import blocks as OB
    for t in ([981.0], [1.0, 21.0, 501.0, 999.0]):
        tt = torch.tensor(t)
        ref = OB.Timesteps(320, True, 0)(tt.expand(4) if len(t) == 1 else tt)
        out = ops.timestep_embedding(tt.cuda(), 4, 320, torch.float32)
        torch.testing.assert_close(out.cpu(), ref, rtol=RTOL, atol=ATOL)
        out = ops.timestep_embedding(tt.cuda(), 4, 320, torch.float16)
        torch.testing.assert_close(out.float().cpu(), ref, rtol=RTOL, atol=5e-4)  # fp16 store near |x| = 1
    x = torch.randn(3, 9, 16, 32)
    out = ops.nchw_to_nhwc_pad(x.cuda(), 64, torch.float16).cpu()
    assert torch.equal(out[..., :9], x.permute(0, 2, 3, 1).half()) and out[..., 9:].abs().sum() == 0
    y = torch.randn(3, 16, 32, 32)
    assert torch.equal(ops.nhwc_to_nchw(y.cuda(), 4, torch.float32).cpu(), y[..., :4].permute(0, 3, 1, 2))
    x = torch.randn(2, 4, 8, 128).half()
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(ops.upsample_nearest2x(x.cuda()).float().cpu(), ref)


def test_fused_step_and_schedulers_match_oracle(ops):
    from oracle.pipeline import cfg_combine
    from oracle.schedulers import OracleDDIMScheduler, ddpm_add_noise
    from pcdms_b200.scheduler import B200DDIMScheduler, B200DDPMScheduler
    n, h, w = 2, 16, 32
    sch, osch = B200DDIMScheduler(), OracleDDIMScheduler()
    sch.set_timesteps(10)
    osch.set_timesteps(10)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(n, 4, h, w, generator=g)
    eps_rows = torch.randn(2 * n, h, w, 32, generator=g)
    x9 = torch.zeros(2 * n, h, w, 64, dtype=torch.float16, device="cuda")
    coef = sch.coefficient_table("cuda")
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    t_table = torch.cat([sch.timesteps.float(), torch.zeros(1)]).cuda()
    t_cur = torch.zeros(1, device="cuda")
    lat_d = lat.clone().cuda()
    ref = lat.clone()
    eps_nchw = eps_rows[..., :4].permute(0, 3, 1, 2)
    for i, t in enumerate(osch.timesteps[:4]):
        ops.cfg_ddim_step(eps_rows.cuda(), lat_d, x9, coef, counter, 2.0, t_table, t_cur)
        ref = osch.step(cfg_combine(eps_nchw, 2.0), t, ref, return_dict=False)[0]
        torch.testing.assert_close(lat_d.cpu(), ref, rtol=1e-5, atol=1e-5)
        want9 = torch.cat([ref, ref]).permute(0, 2, 3, 1).half()
        torch.testing.assert_close(x9[..., :4].cpu(), want9, rtol=RTOL, atol=ATOL)
        assert counter.tolist() == [i + 1, 0] and t_cur.item() == float(osch.timesteps[i + 1])
    e, s = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    out = sch.step(e.cuda(), 901, s.cuda(), return_dict=False)[0]
    torch.testing.assert_close(out.cpu(), osch.step(e, 901, s, return_dict=False)[0], rtol=1e-5, atol=1e-5)
    ts = torch.tensor([0, 500])
    out = B200DDPMScheduler().add_noise(s.cuda(), e.cuda(), ts.cuda())
    torch.testing.assert_close(out.cpu(), ddpm_add_noise(s, e, ts), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("M,C,N,geglu,res", [(1000, 320, 960, False, True), (4096, 640, 640, False, False),
                                            (300, 1280, 2560, True, True), (2048, 320, 2560, True, False),
                                            (16, 128, 256, False, True), (32768, 320, 320, False, True)])
def test_layernorm_folded_into_gemms(ops, dt, M, C, N, geglu, res):
    """LayerNorm without a pass of its own (BasicTransformerBlock norm1/2/3): the PRODUCING GEMM emits
    per-row (sum, sum of squares) from its epilogue (pcdm_ext.row_stats), the CONSUMING GEMM takes the raw rows, a
    gamma-scaled weight and finishes the normalisation in its epilogue (pcdm_ext.ln_*).  Checked against
    torch LayerNorm -> Linear (-> GEGLU) in fp32 on the same 16-bit tensors, and against the stand-alone statistics
    kernel."""
    g = torch.Generator().manual_seed(M + C + N)
    a = torch.randn(M, C, generator=g).to(dt)
    w0 = (torch.randn(C, C, generator=g) / C ** 0.5).to(dt)
    b0 = torch.randn(C, generator=g)
    r = (2.0 * torch.randn(M, C, generator=g) + 1.5).to(dt) if res else None     # rows with a non-zero mean
    gamma, beta = 1.0 + 0.2 * torch.randn(C, generator=g), 0.3 * torch.randn(C, generator=g)
    W = torch.randn(N, C, generator=g) / C ** 0.5
    bias = torch.randn(N, generator=g)
    # producer: h = a @ w0^T + b0 (+ r), with row statistics
    h, st = ops.gemm(a.cuda(), w0.cuda(), bias=b0.cuda(), residual=r.cuda() if res else None, row_stats=True)
    h_ref = a.float() @ w0.float().t() + b0 + (r.float() if res else 0)
    close(h, h_ref, dt)
    assert 1 <= st.parts <= st.buf.shape[0]
    hf = h.float().cpu()

    def mean_rstd(stats):
        sums = stats.buf[: stats.parts].sum(0).cpu()
        mean = sums[:, 0] / C
        return mean, torch.rsqrt((sums[:, 1] / C - mean * mean).clamp_min(0) + 1e-5)

    # the epilogue's statistics are of the fp32 values it is about to round: equal to the statistics of the stored rows up
    # to that rounding; the stand-alone kernel reads the stored rows
    want_mean, want_rstd = hf.mean(1), torch.rsqrt(hf.var(1, unbiased=False) + 1e-5)
    alone = ops.row_stats(h)
    assert alone.parts == 1
    for stats, k in ((st, 1.0), (alone, 0.1)):
        mean, rstd = mean_rstd(stats)
        torch.testing.assert_close(mean, want_mean, rtol=0, atol=k * tol(dt)["rtol"] * float(hf.abs().max()))
        torch.testing.assert_close(rstd, want_rstd, rtol=k * tol(dt)["rtol"], atol=0)
    # consumer: LN(h) @ W^T + bias (-> GEGLU), W folded once (gamma-scaled, rows centred)
    Wf, cb = ops.fold_layernorm_weight(W, gamma, beta, bias, dt)
    assert float(Wf.float().sum(1).abs().max()) < 0.05        # centred rows (up to the 16-bit rounding of each entry)
    # (2) the layer it replaces: torch LayerNorm(gamma, beta) -> Linear(W, bias) with the fp32 MASTER weight.  Any 16-bit
    #     path rounds the weight once (here: the centred W * gamma; conventionally: W, plus the normalised rows), so the
    #     yardstick is the conventional path — pcdm_layernorm + pcdm_gemm — whose error this one must not exceed
    ref_true = F.layer_norm(hf, (C,), gamma, beta, 1e-5) @ W.t() + bias

    def act(y):
        if not geglu:
            return y
        hh, gate = y.chunk(2, dim=1)
        return hh * F.gelu(gate)

    Wd, cbd, Wu, bu = Wf, cb, W.to(dt), bias
    if geglu:
        perm = ops.geglu_row_permutation(N // 2)
        Wd, cbd, Wu, bu = Wf[perm].contiguous(), cb[perm].contiguous(), Wu[perm].contiguous(), bias[perm].contiguous()
    conventional = ops.gemm(ops.layernorm(h, gamma.cuda(), beta.cuda()), Wu.cuda(), bias=bu.cuda(), geglu=geglu)
    err_conv = (conventional.float().cpu() - act(ref_true)).abs()
    for stats in (st, alone):
        out = ops.gemm(h, Wd.cuda(), bias=cbd.cuda(), geglu=geglu, ln=ops.FoldedLN(stats, 1e-5))
        # (1) the kernel's arithmetic on ITS inputs: rstd (from the statistics it was given) * (h . W'^T) + bias'
        ref_same = act(mean_rstd(stats)[1][:, None] * (hf @ Wf.float().t()) + cb)
        close(out, ref_same, dt)
        err = (out.float().cpu() - act(ref_true)).abs()
        assert err.max() <= 1.5 * err_conv.max() + tol(dt)["atol"] and err.mean() <= 1.25 * err_conv.mean() + 1e-6
    assert torch.equal(out, ops.gemm(h, Wd.cuda(), bias=cbd.cuda(), geglu=geglu, ln=ops.FoldedLN(alone, 1e-5)))


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
def test_cfg_combine_and_rescale(ops, dt):
    """a11: e_u + g (e_c - e_u) and the reference's rescale_noise_cfg (stage2_inpaint_pipeline.py:52-63) as kernels —
    stand-alone (the generic protocol loop) and inside the fused DDIM step."""
    from oracle.pipeline import cfg_combine, rescale_noise_cfg
    from oracle.schedulers import OracleDDIMScheduler
    from pcdms_b200.scheduler import B200DDIMScheduler
    n, h, w = 3, 16, 32
    g = torch.Generator().manual_seed(5)
    eps = (torch.randn(2 * n, 4, h, w, generator=g) * torch.tensor([0.5, 1.0, 2.0, 1.0, 0.7, 3.0]).view(-1, 1, 1, 1)).to(dt)
    ref = cfg_combine(eps.float(), 2.0)
    ref_r = rescale_noise_cfg(ref, eps.float()[n:], guidance_rescale=0.7)
    tol_ = dict(rtol=1e-5, atol=1e-5) if dt == torch.float32 else tol(dt)
    torch.testing.assert_close(ops.cfg_combine(eps.cuda(), 2.0).float().cpu(), ref.to(dt).float(), **tol_)
    out = ops.cfg_combine(eps.cuda(), 2.0, 0.7)
    assert out.dtype == dt and out.shape == (n, 4, h, w)
    torch.testing.assert_close(out.float().cpu(), ref_r, **tol_)
    want_ratio = eps.float()[n:].std(dim=(1, 2, 3)) / ref.std(dim=(1, 2, 3))
    torch.testing.assert_close(ops.cfg_rescale_ratio(eps.cuda(), 2.0).cpu(), want_ratio, rtol=1e-5, atol=1e-6)
    assert torch.equal(ops.cfg_rescale_ratio(eps.cuda(), 2.0), ops.cfg_rescale_ratio(eps.cuda(), 2.0))   # reproducible
    if dt != torch.float32:
        return
    # fused: NHWC fp32 rows with a leading dimension, ratio read per sample inside the step kernel
    rows = torch.zeros(2 * n, h, w, 32)
    rows[..., :4] = eps.permute(0, 2, 3, 1)
    rows[..., 4:] = 99.0                      # padding columns must not enter the statistics
    ratio = ops.cfg_rescale_ratio(rows.cuda(), 2.0, nhwc_channels=4)
    torch.testing.assert_close(ratio.cpu(), want_ratio, rtol=1e-5, atol=1e-6)
    sch, osch = B200DDIMScheduler(), OracleDDIMScheduler()
    sch.set_timesteps(10)
    osch.set_timesteps(10)
    lat = torch.randn(n, 4, h, w, generator=g)
    lat_d = lat.clone().cuda()
    x9 = torch.zeros(2 * n, h, w, 64, dtype=torch.float16, device="cuda")
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    ops.cfg_ddim_step(rows.cuda(), lat_d, x9, sch.coefficient_table("cuda"), counter, 2.0, ratio=ratio,
                      guidance_rescale=0.7)
    want = osch.step(ref_r, osch.timesteps[0], lat, return_dict=False)[0]
    torch.testing.assert_close(lat_d.cpu(), want, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------------------------------
# VAE / conditioning front-end shapes: rows wider than a tile, bottom/right-padded stride 2,
# SiLU / GELU epilogues, fp32 score softmax
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,pad_br", [
    (1, 8, 256, 64, 64, 1, False), (2, 6, 512, 128, 128, 1, False), (1, 3, 384, 64, 32, 1, False),
    (2, 4, 256, 64, 128, 2, False), (2, 16, 32, 128, 128, 2, True), (1, 4, 128, 64, 64, 2, True),
    (2, 5, 256, 64, 64, 2, True),
])
def test_conv3x3_wide_and_asymmetric(ops, dt, B, H, W, Cin, Cout, stride, pad_br):
    g = torch.Generator().manual_seed(B * H + W + Cin)
    x = torch.randn(B, Cin, H * stride, W * stride, generator=g).to(dt)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dt)
    b = torch.randn(Cout, generator=g)
    if pad_br:
        ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), w.float(), b, stride=2, padding=0)
    else:
        ref = F.conv2d(x.float(), w.float(), b, stride=stride, padding=1)
    r = torch.randn(B, Cout, H, W, generator=g).to(dt)
    ref = F.silu(ref + r.float())
    out = ops.conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), ops.pack_conv3x3_weight(w, dt).cuda(), bias=b.cuda(),
                      residual=r.permute(0, 2, 3, 1).contiguous().cuda(), stride=stride, pad_br=pad_br, silu=True)
    close(out.permute(0, 3, 1, 2), ref, dt)


@pytest.mark.parametrize("dt", DTS)
def test_gemm_gelu_epilogue(ops, dt):
    g = torch.Generator().manual_seed(5)
    a = torch.randn(514, 1536, generator=g).to(dt)
    w = (torch.randn(768, 1536, generator=g) / 1536 ** 0.5).to(dt)
    b = torch.randn(768, generator=g)
    ref = F.gelu(a.float() @ w.float().t() + b)
    close(ops.gemm(a.cuda(), w.cuda(), bias=b.cuda(), gelu=True), ref, dt)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("M,N", [(64, 2048), (300, 8192), (5, 36), (17, 16384)])
def test_softmax_rows(ops, dt, M, N):
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, N, generator=g) * 30.0
    scale = 512 ** -0.5
    got = ops.softmax_rows(x.cuda(), scale, dt)
    want = torch.softmax(x * scale, dim=-1)
    torch.testing.assert_close(got.float().cpu(), want, rtol=(2 ** -7 if dt == torch.bfloat16 else 2 ** -10), atol=1e-6)
    assert torch.equal(got, ops.softmax_rows(x.cuda(), scale, dt))


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("M,N,K,kw", [
    (6, 2048, 2048, {"resid": True}), (12, 6144, 2048, {}), (6, 8192, 2048, {"gelu": True}),
    (6, 2048, 8192, {"resid": True}), (1, 2048, 1024, {"rowvec": True}), (8, 512, 64, {"gelu": True}),
    (16, 1280, 320, {"silu": True}), (16, 20160, 1280, {"f32": True}), (24, 1024, 2048, {"f32": True, "bias": False}),
    (32, 2048, 1024, {"resid": True, "rowvec": True}), (9, 64, 128, {}),
])
def test_skinny_gemm(ops, dt, M, N, K, kw):
    """pcdm_gemm with M <= 32 rows: the weight-streaming mma.sync kernel (skinny.cu) against torch fp32 on the same 16-bit
    inputs, against the tcgen05 tile path it replaces (PCDM_FLAG_NO_SKINNY), through strided views, and run to run."""
    import ctypes as C
    from pcdms_b200 import lib
    g = torch.Generator().manual_seed(M + N + K)
    a_full = torch.randn(M, 3, K, generator=g).to(dt)              # rows taken as a strided view (token 1 of 3)
    a = a_full[:, 1]
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dt)
    b = torch.randn(N, generator=g) if kw.get("bias", True) else None
    r = torch.randn(M, N, generator=g).to(dt) if kw.get("resid") else None
    rv = torch.randn(2, N, generator=g) if kw.get("rowvec") else None
    rpi = (M + 1) // 2
    ref = a.float() @ w.float().t()
    if b is not None:
        ref = ref + b
    if rv is not None:
        ref = ref + rv.repeat_interleave(rpi, 0)[:M]
    if r is not None:
        ref = ref + r.float()
    if kw.get("silu"):
        ref = F.silu(ref)
    if kw.get("gelu"):
        ref = F.gelu(ref)
    ad = a_full.cuda()[:, 1]
    assert ad.stride(0) == 3 * K
    args = dict(bias=b.cuda() if b is not None else None, residual=r.cuda() if r is not None else None,
                rowvec=rv.cuda() if rv is not None else None, rows_per_image=rpi, silu=bool(kw.get("silu")),
                gelu=bool(kw.get("gelu")), out_f32=bool(kw.get("f32")))
    out_buf = torch.zeros(M, 2 * N, device="cuda", dtype=torch.float32 if kw.get("f32") else dt)
    before = lib.launch_count
    out = ops.gemm(ad, w.cuda(), out=out_buf[:, :N], **args)
    assert lib.launch_count == before + 1
    assert out.dtype == (torch.float32 if kw.get("f32") else dt) and not out_buf[:, N:].any()
    close(out, ref, dt)
    assert torch.equal(out, ops.gemm(ad, w.cuda(), **args))            # bit-reproducible
    if M > 1 and rv is None:   # a row's result does not depend on how many other rows ride along (shard independence)
        k = M // 2
        sub = dict(args, residual=args["residual"][:k] if args["residual"] is not None else None)
        assert torch.equal(ops.gemm(ad[:k], w.cuda(), **sub), out[:k])
    if N % 32 == 0:                                                     # the tile path needs N % 32 == 0
        tiles = ops.gemm(ad, w.cuda(), skinny=False, **args)
        close(out, tiles.float().cpu(), dt, mult=2.0)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,heads,Sq,Skv", [(1, 32, 6, 6), (2, 32, 6, 6), (3, 5, 32, 32), (2, 3, 1, 17), (1, 2, 20, 3)])
def test_attention_short_sequences(ops, dt, B, heads, Sq, Skv):
    """Sq, Skv <= 32 at head_dim 64: the one-warp-per-head CUDA-core kernel (the stage-1 prior's six tokens); q / k / v
    are column slices of fused buffers."""
    g = torch.Generator().manual_seed(B * 100 + Sq + Skv)
    C_ = heads * 64
    q = torch.randn(B * Sq, C_ + 64, generator=g).to(dt)
    kv = torch.randn(B * Skv, 2 * C_, generator=g).to(dt)
    q4 = q[:, :C_].float().view(B, Sq, heads, 64).transpose(1, 2)
    k4 = kv[:, :C_].float().view(B, Skv, heads, 64).transpose(1, 2)
    v4 = kv[:, C_:].float().view(B, Skv, heads, 64).transpose(1, 2)
    want = (torch.softmax(q4 @ k4.transpose(-1, -2) * 0.125, -1) @ v4).transpose(1, 2).reshape(B * Sq, C_)
    qd, kvd = q.cuda(), kv.cuda()
    got = ops.attention(qd[:, :C_], kvd[:, :C_], kvd[:, C_:], B, heads)
    close(got, want, dt)
    assert torch.equal(got, ops.attention(qd[:, :C_], kvd[:, :C_], kvd[:, C_:], B, heads))


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("M,N,K,kw", [(6, 6144, 2048, {}), (12, 8192, 2048, {"gelu": True}), (2, 1024, 2048, {"f32": True}),
                                      (24, 640, 320, {"resid": True}), (40, 640, 320, {"gelu": True}),
                                      (32, 512, 2048, {})])
def test_ln_gemm(ops, dt, M, N, K, kw):
    """pcdm_ln_gemm: LayerNorm fused in front of the skinny GEMM (M <= 32 rows that fit) or layernorm + gemm inside the
    call (M = 40; M = 32 x K = 2048 exceeds the shared-memory budget) — against torch fp32 with the normalised rows
    rounded to 16 bits, as both paths do."""
    from pcdms_b200 import lib
    g = torch.Generator().manual_seed(M + N + K)
    x_full = (2.0 * torch.randn(M, 2, K, generator=g) + 0.5).to(dt)
    x = x_full[:, 1]
    gamma, beta = 1.0 + 0.1 * torch.randn(K, generator=g), 0.1 * torch.randn(K, generator=g)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dt)
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g).to(dt) if kw.get("resid") else None
    n16 = F.layer_norm(x.float(), (K,), gamma, beta, 1e-5).to(dt).float()
    ref = n16 @ w.float().t() + b
    if r is not None:
        ref = ref + r.float()
    if kw.get("gelu"):
        ref = F.gelu(ref)
    xd = x_full.cuda()[:, 1]
    before = lib.launch_count
    out = ops.ln_gemm(xd, gamma.cuda(), beta.cuda(), 1e-5, w.cuda(), bias=b.cuda(),
                      residual=r.cuda() if r is not None else None, gelu=bool(kw.get("gelu")), out_f32=bool(kw.get("f32")))
    fused = M <= 32 and M * (K + 8) * 2 <= 100 * 1024
    assert lib.launch_count - before == (1 if fused else 2)
    # a rounding flip of a normalised value (fp32 mean / rstd evaluated in a different order) moves an output by one
    # 16-bit ulp of that value times a weight: allow 2x the single-rounding tolerance
    close(out, ref, dt, mult=2.0)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
def test_ddim_stochastic_step(ops, dt):
    """DDIMScheduler.step with eta != 0 (the `eta` / `generator` the reference pipelines forward to step(),
    stage2_inpaint_pipeline.py:307-322,519): sigma_t noise added, direction term shrunk — against the oracle's
    restatement of diffusers' formula with the same variance noise; the generator path draws the same numbers as the
    oracle from an identically seeded CPU generator (fp32)."""
    from oracle.schedulers import OracleDDIMScheduler
    from pcdms_b200.scheduler import B200DDIMScheduler
    g = torch.Generator().manual_seed(3)
    e, s = torch.randn(2, 4, 16, 32, generator=g).to(dt), torch.randn(2, 4, 16, 32, generator=g).to(dt)
    noise = torch.randn(2, 4, 16, 32, generator=g).to(dt)
    sch, osch = B200DDIMScheduler(), OracleDDIMScheduler()
    sch.set_timesteps(10)
    osch.set_timesteps(10)
    for t, eta in ((901, 1.0), (401, 0.5), (1, 0.3)):
        want = osch.step(e.float(), t, s.float(), eta=eta, variance_noise=noise.float(), return_dict=False)[0]
        got = sch.step(e.cuda(), t, s.cuda(), eta=eta, variance_noise=noise.cuda(), return_dict=False)[0]
        assert got.dtype == dt
        close(got, want, torch.float16 if dt == torch.float32 else dt)
        assert (want - osch.step(e.float(), t, s.float(), return_dict=False)[0]).abs().max() > 1e-3   # eta matters
    if dt == torch.float32:
        want = osch.step(e, 501, s, eta=0.8, generator=torch.Generator().manual_seed(9), return_dict=False)[0]
        got = sch.step(e.cuda(), 501, s.cuda(), eta=0.8, generator=torch.Generator().manual_seed(9), return_dict=False)[0]
        torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        sch.step(e.cuda(), 501, s.cuda(), eta=0.5, generator=torch.Generator(), variance_noise=noise.cuda())


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("shape", [(2, 4, 16, 32), (3, 4, 3, 3)])    # 16-byte vector path / scalar path (numel % 8 != 0)
def test_scheduler_protocol_16bit_vector_paths(ops, dt, shape):
    """DDIMScheduler.step and DDPMScheduler.add_noise on same-dtype 16-bit tensors (the reference's fp16 loop tensors,
    stage2_inpaint_pipeline.py:431-501; training add_noise, stage2_train_inpaint_model.py:361): the vectorised kernels
    against the fp32 formulas on the same 16-bit inputs."""
    from oracle.schedulers import OracleDDIMScheduler, ddpm_add_noise
    from pcdms_b200.scheduler import B200DDIMScheduler, B200DDPMScheduler
    g = torch.Generator().manual_seed(sum(shape))
    e, s = torch.randn(shape, generator=g).to(dt), torch.randn(shape, generator=g).to(dt)
    sch, osch = B200DDIMScheduler(), OracleDDIMScheduler()
    sch.set_timesteps(10)
    osch.set_timesteps(10)
    out = sch.step(e.cuda(), 501, s.cuda(), return_dict=False)[0]
    assert out.dtype == dt
    close(out, osch.step(e.float(), 501, s.float(), return_dict=False)[0], dt)
    ts = torch.tensor([0, 500, 999][: shape[0]])
    out = B200DDPMScheduler().add_noise(s.cuda(), e.cuda(), ts.cuda())
    assert out.dtype == dt
    close(out, ddpm_add_noise(s.float(), e.float(), ts), dt)
