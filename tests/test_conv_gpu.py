"""GPU parity of the tcgen05 implicit-GEMM convolution kernels (forward / dgrad / wgrad) against a plain
PyTorch fp32 convolution on the same bf16-rounded operands (reference op: hourglass.py:100 nn.Conv2d).

Tolerance: inputs are exactly representable in bf16 on both sides, products accumulate in fp32 on both sides,
so the only differences are summation order and the final bf16 rounding of the NHWC output
(2^-9 relative): |err| <= 1e-2 * max|ref| elementwise and relative L2 error <= 3e-3.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf16_round(t):
    return t.to(torch.bfloat16).float()


def _rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


CASES = [
    # N, H, W, cin, cout, k
    (2, 32, 32, 128, 128, 3),
    (2, 16, 16, 128, 128, 3),
    (3, 8, 8, 128, 128, 3),
    (2, 4, 4, 128, 128, 3),
    (5, 2, 2, 128, 128, 3),
    (1, 64, 64, 64, 128, 3),
    (1, 64, 64, 64, 128, 1),
    (2, 32, 32, 128, 128, 1),
    (1, 40, 24, 128, 128, 3),   # tiles that do not divide the image (masked rows)
    (2, 5, 5, 128, 128, 3),
    (1, 128, 128, 128, 128, 3),
]


@pytest.mark.parametrize("N,H,W,cin,cout,k", CASES)
def test_conv_forward(cuda_device, N, H, W, cin, cout, k):
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(1234 + H * 7 + cin)
    x = _bf16_round(torch.randn(N, cin, H, W, generator=g))
    w = _bf16_round(torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k) ** 0.5))
    bias = torch.randn(cout, generator=g)
    ref = F.conv2d(x, w, bias, padding=(k - 1) // 2)

    xd = ops.to_nhwc(x.to(cuda_device))
    wp = ops.pack_weight(w.to(cuda_device), mode=0)
    stats = torch.zeros(2, cout, device=cuda_device)
    y = ops.conv2d_igemm(xd, wp, cout, k, bias=bias.to(cuda_device), stats=stats)
    out = ops.to_nchw(y).cpu()
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 1e-2 * scale
    assert _rel_l2(out, ref) <= 3e-3
    # BN statistics are accumulated from the fp32 accumulators, before the bf16 rounding of the output
    s1 = ref.sum(dim=(0, 2, 3))
    s2 = (ref * ref).sum(dim=(0, 2, 3))
    assert torch.allclose(stats[0].cpu(), s1, rtol=1e-3, atol=1e-3 * s2.sqrt().max().item())
    assert torch.allclose(stats[1].cpu(), s2, rtol=1e-3)


def test_conv_forward_addend(cuda_device):
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(7)
    x = _bf16_round(torch.randn(2, 128, 16, 16, generator=g))
    r = _bf16_round(torch.randn(2, 128, 16, 16, generator=g))
    w = _bf16_round(torch.randn(128, 128, 1, 1, generator=g) * 0.1)
    ref = F.conv2d(x, w) + r
    y = ops.conv2d_igemm(ops.to_nhwc(x.to(cuda_device)), ops.pack_weight(w.to(cuda_device)), 128, 1,
                         addend=ops.to_nhwc(r.to(cuda_device)))
    out = ops.to_nchw(y).cpu()
    assert _rel_l2(out, ref) <= 3e-3


@pytest.mark.parametrize("N,H,W,cin,cout,k", CASES)
def test_conv_dgrad(cuda_device, N, H, W, cin, cout, k):
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(99 + H + cin)
    dy = _bf16_round(torch.randn(N, cout, H, W, generator=g))
    w = _bf16_round(torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cout * k * k) ** 0.5))
    ref = torch.nn.grad.conv2d_input((N, cin, H, W), w, dy, padding=(k - 1) // 2)
    wd = ops.pack_weight(w.to(cuda_device), mode=1)
    dx = ops.conv2d_igemm(ops.to_nhwc(dy.to(cuda_device)), wd, cin, k)
    out = ops.to_nchw(dx).cpu()
    assert (out - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert _rel_l2(out, ref) <= 3e-3


@pytest.mark.parametrize("N,H,W,cin,cout,k", CASES)
def test_conv_wgrad(cuda_device, N, H, W, cin, cout, k):
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(5 + H + cin)
    x = _bf16_round(torch.randn(N, cin, H, W, generator=g))
    dy = _bf16_round(torch.randn(N, cout, H, W, generator=g))
    ref = torch.nn.grad.conv2d_weight(x, (cout, cin, k, k), dy, padding=(k - 1) // 2)
    gw = ops.conv2d_wgrad(ops.to_nhwc(x.to(cuda_device)), ops.to_nhwc(dy.to(cuda_device)), cin, k).cpu()
    # fp32 output, fp32 accumulation: only summation order differs
    assert _rel_l2(gw, ref) <= 1e-4
    assert (gw - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


def test_head_conv(cuda_device):
    """1x1 128->6 head writing the fp32 NCHW logits slice (hourglass.py:189-195, :237 torch.stack)."""
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = _bf16_round(torch.randn(2, 128, 32, 32, generator=g))
    w = _bf16_round(torch.randn(6, 128, 1, 1, generator=g) * 0.1)
    b = torch.randn(6, generator=g)
    ref = F.conv2d(x, w, b)
    logits = torch.zeros(2, 2, 6, 32, 32, device=cuda_device)
    pad = torch.full((2, 32, 32, 64), 7.0, dtype=torch.bfloat16, device=cuda_device)
    ops.conv2d_igemm(ops.to_nhwc(x.to(cuda_device)), ops.pack_weight(w.to(cuda_device)), 6, 1,
                     bias=b.to(cuda_device), head_out=logits, stack_idx=1, out2=pad)
    out = logits.cpu()
    assert torch.count_nonzero(out[:, 0]) == 0
    assert torch.allclose(out[:, 1], ref, rtol=1e-4, atol=1e-4)
    padc = ops.to_nchw(pad).cpu()
    assert _rel_l2(padc[:, :6], ref) <= 3e-3
    assert torch.count_nonzero(padc[:, 6:16]) == 0


HALO_CASES = [(2, 32, 32, 128, 128, 3), (1, 64, 48, 128, 128, 3), (1, 40, 24, 128, 128, 3), (2, 16, 16, 64, 128, 3),
              (3, 17, 33, 128, 128, 3), (2, 32, 32, 128, 128, 1), (2, 16, 48, 64, 128, 1), (1, 40, 24, 128, 128, 1)]


@pytest.mark.parametrize("variant", [2])
@pytest.mark.parametrize("N,H,W,cin,cout,k", HALO_CASES)
def test_conv_halo_variant(cuda_device, N, H, W, cin, cout, k, variant):
    """The halo kernel (transposed product: M = 128 channels, N = 256 pixels, TMA-store epilogue), forced onto shapes
    the dispatcher would give to the generic kernel too (ragged 17x33 / 40x24 maps, cin = 64), must agree
    with PyTorch exactly like the generic kernel: forward with bias, residual addend and BN statistics, and dgrad."""
    from real_time_helmet_detection_b200 import ops, _lib
    g = torch.Generator().manual_seed(77 + H + W)
    x = _bf16_round(torch.randn(N, cin, H, W, generator=g))
    r = _bf16_round(torch.randn(N, cout, H, W, generator=g))
    w = _bf16_round(torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k) ** 0.5))
    bias = torch.randn(cout, generator=g)
    conv = F.conv2d(x, w, bias, padding=(k - 1) // 2)
    ref = conv + r
    _lib.lib().hd_set_conv_variant(variant)
    try:
        stats = torch.zeros(2, cout, device=cuda_device)
        y = ops.conv2d_igemm(ops.to_nhwc(x.to(cuda_device)), ops.pack_weight(w.to(cuda_device)), cout, k,
                             bias=bias.to(cuda_device), addend=ops.to_nhwc(r.to(cuda_device)), stats=stats)
        out = ops.to_nchw(y).cpu()
        if cin == cout:
            dy = _bf16_round(torch.randn(N, cout, H, W, generator=g))
            dref = torch.nn.grad.conv2d_input((N, cin, H, W), w, dy, padding=(k - 1) // 2)
            dx = ops.to_nchw(ops.conv2d_igemm(ops.to_nhwc(dy.to(cuda_device)), ops.pack_weight(w.to(cuda_device), mode=1),
                                              cin, k)).cpu()
    finally:
        _lib.lib().hd_set_conv_variant(0)
    assert (out - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert _rel_l2(out, ref) <= 3e-3
    s1, s2 = ref.sum(dim=(0, 2, 3)), (ref * ref).sum(dim=(0, 2, 3))
    assert torch.allclose(stats[0].cpu(), s1, rtol=1e-3, atol=1e-3 * s2.sqrt().max().item())
    assert torch.allclose(stats[1].cpu(), s2, rtol=1e-3)
    if cin == cout:
        assert _rel_l2(dx, dref) <= 3e-3


@pytest.mark.parametrize("N,H,W,cin,cout,k,relu,with_add", [
    (2, 32, 32, 128, 128, 3, True, True),      # halo kernel (forced below) and generic kernel
    (1, 40, 24, 128, 128, 3, False, True),
    (2, 16, 16, 64, 128, 1, False, False),     # skip conv 1x1 64->128
    (3, 8, 8, 128, 128, 3, True, False),
    (1, 32, 32, 192, 64, 1, True, False),      # stem GEMM shape (K = 192, 64 channels)
])
def test_conv_affine_epilogue(cuda_device, N, H, W, cin, cout, k, relu, with_add):
    """hd_conv2d_igemm_affine: eval-mode Convolution (BN folded to scale/shift) + residual addend + ReLU in the conv
    epilogue == relu((conv + bias) * scale + shift + addend) of PyTorch, on the generic and the halo kernel."""
    from real_time_helmet_detection_b200 import ops, _lib
    g = torch.Generator().manual_seed(5 + H + cin)
    x = _bf16_round(torch.randn(N, cin, H, W, generator=g))
    w = _bf16_round(torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k) ** 0.5))
    bias = torch.randn(cout, generator=g)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    r = _bf16_round(torch.randn(N, cout, H, W, generator=g)) if with_add else None
    ref = (F.conv2d(x, w, bias, padding=(k - 1) // 2)) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if with_add:
        ref = ref + r
    if relu:
        ref = torch.relu(ref)
    rows = 128 if cout == 128 else 64
    wp = ops.pack_weight(w.to(cuda_device), rows_pad=rows) if cout != 128 else ops.pack_weight(w.to(cuda_device))
    for variant in ((1, 2) if (k == 3 and H >= 16 and cout == 128) else (0,)):
        _lib.lib().hd_set_conv_variant(variant)
        try:
            y = ops.conv2d_igemm_affine(ops.to_nhwc(x.to(cuda_device)), wp, cout, k, scale.to(cuda_device),
                                        shift.to(cuda_device), relu, bias=bias.to(cuda_device),
                                        addend=ops.to_nhwc(r.to(cuda_device)) if with_add else None)
        finally:
            _lib.lib().hd_set_conv_variant(0)
        out = ops.to_nchw(y).cpu()
        assert _rel_l2(out, ref) <= 3e-3, (variant, _rel_l2(out, ref))
        assert (out - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-2


def test_bn_fold_all(cuda_device):
    import ctypes
    from real_time_helmet_detection_b200 import _lib
    class Job(ctypes.Structure):
        _fields_ = [("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("mean", ctypes.c_void_p),
                    ("var", ctypes.c_void_p), ("out", ctypes.c_void_p), ("channels", ctypes.c_int), ("eps", ctypes.c_float)]
    g = torch.Generator().manual_seed(1)
    chans = [64, 128, 128, 6]
    ts = [[torch.randn(c, generator=g).to(cuda_device) for _ in range(3)] + [(torch.rand(c, generator=g) + 0.1).to(cuda_device)]
          for c in chans]
    outs = [torch.zeros(2 * c, device=cuda_device) for c in chans]
    jobs = (Job * len(chans))(*[Job(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), o.data_ptr(), c, 1e-5)
                                for t, o, c in zip(ts, outs, chans)])
    dev_jobs = torch.zeros(len(chans) * ctypes.sizeof(Job), dtype=torch.uint8, device=cuda_device)
    _lib.check(_lib.lib().hd_bn_fold_all(ctypes.cast(jobs, ctypes.c_void_p), len(chans), _lib.ptr(dev_jobs), _lib.stream()))
    for t, o, c in zip(ts, outs, chans):
        sc = t[0] * torch.rsqrt(t[3] + 1e-5)
        assert torch.allclose(o[:c], sc, rtol=1e-6, atol=1e-7) and torch.allclose(o[c:], t[1] - t[2] * sc, rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------
# 64-output-channel halo kernel (conv_igemm_n64_kernel): the 256x256 level of PreLayer - forced here on small maps with
# hd_set_conv_variant(2), incl. image sizes that are not multiples of the 16x16 tile (zero-filled loads, clipped stores,
# masked statistics), bias / addend / BN-statistics epilogue, and the fused second 1x1 input (the skip-branch dgrad).
@pytest.fixture
def force_halo():
    from real_time_helmet_detection_b200 import _lib
    _lib.lib().hd_set_conv_variant(2)
    yield
    _lib.lib().hd_set_conv_variant(0)


@pytest.mark.parametrize("N,H,W,cin,k", [(2, 32, 32, 128, 3), (1, 48, 40, 128, 3), (3, 16, 16, 64, 3), (2, 32, 48, 128, 1),
                                         (1, 20, 28, 64, 3), (1, 256, 256, 128, 3)])
def test_conv_n64_forward_and_stats(cuda_device, force_halo, N, H, W, cin, k):
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(77 + H + cin + k)
    cout = 64
    x = _bf16_round(torch.randn(N, cin, H, W, generator=g))
    w = _bf16_round(torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k) ** 0.5))
    bias = torch.randn(cout, generator=g)
    r = _bf16_round(torch.randn(N, cout, H, W, generator=g))
    ref = F.conv2d(x, w, bias, padding=(k - 1) // 2)
    xd = ops.to_nhwc(x.to(cuda_device))
    wp = ops.pack_weight(w.to(cuda_device), mode=0)
    assert wp.shape[1] == 64
    stats = torch.zeros(2, cout, device=cuda_device)
    y = ops.to_nchw(ops.conv2d_igemm(xd, wp, cout, k, bias=bias.to(cuda_device), stats=stats)).cpu()
    scale = ref.abs().max().item()
    assert (y - ref).abs().max().item() <= 1e-2 * scale and _rel_l2(y, ref) <= 3e-3
    s1, s2 = ref.sum(dim=(0, 2, 3)), (ref * ref).sum(dim=(0, 2, 3))
    assert torch.allclose(stats[0].cpu(), s1, rtol=1e-3, atol=1e-3 * s2.sqrt().max().item())
    assert torch.allclose(stats[1].cpu(), s2, rtol=1e-3)
    # the generic kernel (variant 1) gives the same tensor up to the bf16 rounding of the output
    from real_time_helmet_detection_b200 import _lib
    _lib.lib().hd_set_conv_variant(1)
    y1 = ops.to_nchw(ops.conv2d_igemm(xd, wp, cout, k, bias=bias.to(cuda_device))).cpu()
    _lib.lib().hd_set_conv_variant(2)
    assert _rel_l2(y, y1) <= 2e-3
    # addend epilogue
    ya = ops.to_nchw(ops.conv2d_igemm(xd, wp, cout, k, addend=ops.to_nhwc(r.to(cuda_device)))).cpu()
    assert _rel_l2(ya, F.conv2d(x, w, None, padding=(k - 1) // 2) + r) <= 3e-3


@pytest.mark.parametrize("N,H,W", [(2, 32, 32), (1, 40, 24), (1, 256, 256)])
def test_conv_n64_dual_input_is_fused_residual_dgrad(cuda_device, force_halo, N, H, W):
    """dX = dgrad3x3(dY1; W1) + dgrad1x1(dYs; Ws) of `Residual(64, 128)` (hourglass.py:111-127) in one launch ==
    torch's conv_transpose pair in fp32 on the same bf16 operands."""
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(5 + H)
    dy1 = _bf16_round(torch.randn(N, 128, H, W, generator=g))
    dys = _bf16_round(torch.randn(N, 128, H, W, generator=g))
    w1 = _bf16_round(torch.randn(128, 64, 3, 3, generator=g) * 0.03)
    ws = _bf16_round(torch.randn(128, 64, 1, 1, generator=g) * 0.1)
    ref = (torch.nn.grad.conv2d_input((N, 64, H, W), w1, dy1, padding=1) +
           torch.nn.grad.conv2d_input((N, 64, H, W), ws, dys, padding=0))
    w1d = ops.pack_weight(w1.to(cuda_device), mode=1)
    wsd = ops.pack_weight(ws.to(cuda_device), mode=1)
    assert w1d.shape == (9, 64, 128) and wsd.shape == (1, 64, 128)
    out = ops.to_nchw(ops.conv2d_igemm_dual(ops.to_nhwc(dy1.to(cuda_device)), w1d, ops.to_nhwc(dys.to(cuda_device)), wsd,
                                            64, 3)).cpu()
    assert (out - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() and _rel_l2(out, ref) <= 3e-3


@pytest.mark.parametrize("N,H,W,k", [(2, 32, 32, 3), (1, 40, 24, 3), (3, 17, 33, 3), (2, 16, 48, 1), (32, 64, 64, 3)])
def test_conv_dgrad_with_bn_backward_statistics(cuda_device, force_halo, N, H, W, k):
    """hd_conv2d_igemm_bwdstat: the halo kernel's output is bit-identical to the plain launch, and the BN-backward
    coefficients / dgamma / dbeta its last CTA leaves equal hd_bn_bwd_reduce_fin's on that output (ragged maps: pixels
    outside the image must not count); the scratch (sums, ticket) is left zeroed, twice in a row."""
    import ctypes
    from real_time_helmet_detection_b200 import ops, _lib

    class Fuse(ctypes.Structure):
        _fields_ = [(n_, ctypes.c_void_p) for n_ in ("gamma", "mean", "rstd", "coef", "dgamma", "dbeta", "gamma_s", "mean_s",
                                                     "rstd_s", "coef_s", "dgamma_s", "dbeta_s")] + \
                   [("count", ctypes.c_float), ("counter", ctypes.c_void_p)]

    C = 128
    d = cuda_device
    L = _lib.lib()
    g = torch.Generator().manual_seed(5 + H + W + k)
    dy2 = ops.to_nhwc(_bf16_round(torch.randn(N, C, H, W, generator=g)).to(d))
    y1 = ops.to_nhwc(_bf16_round(torch.randn(N, C, H, W, generator=g)).to(d))
    w = _bf16_round(torch.randn(C, C, k, k, generator=g) * (1.0 / (C * k * k) ** 0.5)).to(d)
    wpd = ops.pack_weight(w, mode=1)
    bnp = torch.stack([torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.5, torch.randn(C, generator=g) * 0.1,
                       torch.rand(C, generator=g) + 0.5]).to(d)                     # scale | shift | mean | rstd
    gamma = (torch.rand(C, generator=g) + 0.5).to(d)
    npix = N * H * W
    assert L.hd_conv2d_igemm_halo_eligible(N, H, W, C, k) == 1
    plain = ops.conv2d_igemm(dy2, wpd, C, k)

    def scratch():
        s = torch.zeros(16 * 256, device=d)
        outs = [torch.empty(C, device=d) for _ in range(2)]
        f = Fuse()
        f.gamma, f.mean, f.rstd = gamma.data_ptr(), bnp[2].data_ptr(), bnp[3].data_ptr()
        f.coef, f.dgamma, f.dbeta = s[768:].data_ptr(), outs[0].data_ptr(), outs[1].data_ptr()
        f.count, f.counter = float(npix), s[9 * 256:].data_ptr()
        return s, outs, f

    sA, outsA, fA = scratch()
    _lib.check(L.hd_bn_bwd_reduce_fin(_lib.ptr(plain), None, _lib.ptr(bnp[0]), _lib.ptr(bnp[1]), None, None, _lib.ptr(y1), None,
                                      _lib.ptr(sA), npix, C, ctypes.byref(fA), _lib.stream()))
    sB, outsB, fB = scratch()
    for rep in range(2):
        out = torch.empty_like(plain)
        _lib.check(L.hd_conv2d_igemm_bwdstat(_lib.ptr(dy2), _lib.ptr(wpd), _lib.ptr(out), N, H, W, C, C, k, _lib.ptr(y1),
                                             _lib.ptr(bnp[0]), _lib.ptr(bnp[1]), _lib.ptr(sB), ctypes.byref(fB), _lib.stream()))
        assert torch.equal(out, plain)
        for a, b in zip(outsA, outsB):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-4 * float(a.abs().max()) + 1e-5), float((a - b).abs().max())
        assert torch.allclose(sA[768:768 + 3 * C], sB[768:768 + 3 * C], rtol=1e-4, atol=1e-6)
        torch.cuda.synchronize()
        assert float(sB[:768].abs().max()) == 0.0 and int(sB[9 * 256:].view(torch.int32)[0]) == 0
