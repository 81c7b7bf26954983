"""GPU parity of the elementwise / reduction kernels (csrc/elementwise.cu, stem.cu, head.cu) against plain PyTorch
fp32 (CPU, autograd) on the same bf16-rounded inputs. Outputs are bf16: |err| <= 2^-8 relative per element plus
fp32 summation-order noise; reductions (fp32) to 1e-4 relative."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def bf(t):
    return t.to(torch.bfloat16).float()


def nhwc(t, dev):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)


def nchw(t):
    return t.float().cpu().permute(0, 3, 1, 2).contiguous()


def close_bf16(a, b, extra=0.0):
    tol = 2 ** -7 * b.abs() + 1e-2 * b.abs().max() * 2 ** -7 + extra
    assert ((a - b).abs() <= tol).all(), ((a - b).abs().max().item(), b.abs().max().item())


@pytest.mark.parametrize("C,H,W,N", [(128, 16, 16, 2), (64, 32, 32, 1), (128, 6, 10, 3)])
def test_bn_forward_and_residual_tail(cuda_device, C, H, W, N):
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    y = bf(torch.randn(N, C, H, W, generator=g) * 2 + 0.5)
    ys = bf(torch.randn(N, C, H, W, generator=g))
    x = bf(torch.randn(N, C, H, W, generator=g))
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    rm, rv = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5
    d = cuda_device
    for training in (True, False):
        stats = torch.stack([y.sum((0, 2, 3)), (y * y).sum((0, 2, 3))]).to(d)
        rmd, rvd, nbt = rm.clone().to(d), rv.clone().to(d), torch.zeros((), dtype=torch.int64, device=d)
        bnp = ops.bn_finalize(stats, N * H * W, gamma.to(d), beta.to(d), rmd, rvd, nbt, training=training)
        rm2, rv2 = rm.clone(), rv.clone()
        ref = F.batch_norm(y, rm2, rv2, gamma, beta, training=training, momentum=0.1, eps=1e-5)
        z = nchw(ops.bn_act(nhwc(y, d), bnp, relu=True))
        close_bf16(z, F.relu(ref))
        if training:
            assert torch.allclose(rmd.cpu(), rm2, rtol=1e-4, atol=1e-5) and torch.allclose(rvd.cpu(), rv2, rtol=1e-4)
            assert int(nbt.item()) == 1
        else:
            assert torch.equal(rmd.cpu(), rm) and int(nbt.item()) == 0
        out = nchw(ops.bn_add_relu(nhwc(y, d), bnp, nhwc(x, d)))
        close_bf16(out, F.relu(ref + x))
        stats_s = torch.stack([ys.sum((0, 2, 3)), (ys * ys).sum((0, 2, 3))]).to(d)
        bnp_s = ops.bn_finalize(stats_s, N * H * W, gamma.to(d), beta.to(d), rm.clone().to(d), rv.clone().to(d),
                                None, training=training)
        ref_s = F.batch_norm(ys, rm.clone(), rv.clone(), gamma, beta, training=training, eps=1e-5)
        out = nchw(ops.bn_add_relu(nhwc(y, d), bnp, nhwc(ys, d), bnp_s))
        close_bf16(out, F.relu(ref + ref_s))


def test_pool_upsample_add(cuda_device):
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(0)
    d = cuda_device
    x = bf(torch.randn(2, 128, 12, 20, generator=g))
    low = bf(torch.randn(2, 128, 6, 10, generator=g))
    assert torch.equal(nchw(ops.maxpool2(nhwc(x, d))), F.max_pool2d(x, 2, 2))
    close_bf16(nchw(ops.upsample2_add(nhwc(x, d), nhwc(low, d))), x + F.interpolate(low, scale_factor=2))
    a, b, c = (bf(torch.randn(2, 128, 4, 4, generator=g)) for _ in range(3))
    close_bf16(nchw(ops.add(nhwc(a, d), nhwc(b, d), nhwc(c, d))), a + b + c)
    assert torch.allclose(ops.colsum(nhwc(x, d)).cpu(), x.sum((0, 2, 3)), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("skip_bn", [False, True])
def test_bn_backward(cuda_device, skip_bn):
    """d/dy and d/dys of out = relu(bn(y) + (bn_s(ys) | x)), plus dgamma/dbeta, vs autograd."""
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(21)
    d = cuda_device
    N, C, H, W = 2, 128, 8, 12
    y = bf(torch.randn(N, C, H, W, generator=g)).requires_grad_(True)
    ys = bf(torch.randn(N, C, H, W, generator=g)).requires_grad_(True)
    gamma = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
    beta = torch.randn(C, generator=g).requires_grad_(True)
    gamma_s = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
    beta_s = torch.randn(C, generator=g).requires_grad_(True)
    dout = bf(torch.randn(N, C, H, W, generator=g))
    branch = F.batch_norm(ys, None, None, gamma_s, beta_s, training=True) if skip_bn else ys
    out = F.relu(F.batch_norm(y, None, None, gamma, beta, training=True) + branch)
    out.backward(dout)

    def stats(t):
        return torch.stack([t.sum((0, 2, 3)), (t * t).sum((0, 2, 3))]).to(d)
    bnp = ops.bn_finalize(stats(y.detach()), N * H * W, gamma.detach().to(d), beta.detach().to(d))
    bnp_s = ops.bn_finalize(stats(ys.detach()), N * H * W, gamma_s.detach().to(d), beta_s.detach().to(d))
    out_d = nhwc(bf(out.detach()), d)
    dy, dys, gout, (dg, db), (dgs, dbs) = ops.bn_bwd(
        nhwc(dout, d), out_d, nhwc(y.detach(), d), bnp, gamma.detach().to(d),
        ys=nhwc(ys.detach(), d) if skip_bn else None, bnp_s=bnp_s if skip_bn else None,
        gamma_s=gamma_s.detach().to(d) if skip_bn else None, want_g=True)
    scale = y.grad.abs().max().item()
    close_bf16(nchw(dy), y.grad, extra=2e-3 * scale)
    assert torch.allclose(dg.cpu(), gamma.grad, rtol=2e-3, atol=2e-3 * gamma.grad.abs().max().item())
    assert torch.allclose(db.cpu(), beta.grad, rtol=2e-3, atol=2e-3 * beta.grad.abs().max().item())
    gref = dout * (out.detach() > 0)
    close_bf16(nchw(gout), gref)
    if not skip_bn:
        # plain conv+BN+ReLU unit: same result with the ReLU mask recomputed from y instead of read from `out`
        yb = y.detach()
        z = F.relu(F.batch_norm(yb, None, None, gamma.detach(), beta.detach(), training=True)).requires_grad_(False)
        yr = yb.clone().requires_grad_(True)
        F.relu(F.batch_norm(yr, None, None, gamma.detach(), beta.detach(), training=True)).backward(dout)
        dy2 = ops.bn_bwd(nhwc(dout, d), None, nhwc(yb, d), bnp, gamma.detach().to(d), remask=True)[0]
        close_bf16(nchw(dy2), yr.grad, extra=2e-3 * yr.grad.abs().max().item())
    if skip_bn:
        # two-branch tail: identical results with the mask rebuilt from y and ys instead of read from `out`
        r = ops.bn_bwd(nhwc(dout, d), None, nhwc(y.detach(), d), bnp, gamma.detach().to(d), ys=nhwc(ys.detach(), d),
                       bnp_s=bnp_s, gamma_s=gamma_s.detach().to(d), want_g=True, remask=True)
        # (elements whose pre-activation is within bf16 rounding of zero may flip: compare with tolerance)
        close_bf16(nchw(r[0]), y.grad, extra=2e-3 * scale)
        close_bf16(nchw(r[1]), ys.grad, extra=2e-3 * scale)
        assert (nchw(r[2]) != nchw(gout)).float().mean() < 2e-3
        assert torch.allclose(r[3][0].cpu(), gamma.grad, rtol=2e-3, atol=2e-3 * gamma.grad.abs().max().item())
        close_bf16(nchw(dys), ys.grad, extra=2e-3 * scale)
        assert torch.allclose(dgs.cpu(), gamma_s.grad, rtol=2e-3, atol=2e-3 * gamma_s.grad.abs().max().item())
    else:
        close_bf16(nchw(gout), ys.grad)


def test_pool_upsample_backward(cuda_device):
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(2)
    d = cuda_device
    x = bf(torch.randn(2, 128, 8, 8, generator=g))
    x[:, :, :2, :2] = 0.0                                 # ties: the first element of the window gets the gradient
    x = x.requires_grad_(True)
    dp = bf(torch.randn(2, 128, 4, 4, generator=g))
    a1 = bf(torch.randn(2, 128, 8, 8, generator=g))
    F.max_pool2d(x, 2, 2).backward(dp)
    dx = nchw(ops.maxpool2_bwd(nhwc(x.detach(), d), nhwc(dp, d), nhwc(a1, d)))
    close_bf16(dx, x.grad + a1)
    dx0 = nchw(ops.maxpool2_bwd(nhwc(x.detach(), d), nhwc(dp, d)))
    assert torch.equal(dx0, x.grad)
    dout = bf(torch.randn(2, 128, 8, 8, generator=g))
    low = torch.zeros(2, 128, 4, 4, requires_grad=True)
    F.interpolate(low, scale_factor=2).backward(dout)
    close_bf16(nchw(ops.sum2x2(nhwc(dout, d))), low.grad)


@pytest.mark.parametrize("C,H,W,N", [(128, 16, 24, 2), (64, 8, 8, 3)])
def test_fused_residual_tail_pool(cuda_device, C, H, W, N):
    """bn_add_relu_pool2 == maxpool2(bn_add_relu(...)) bit for bit, its argmax points at the first maximum, and
    maxpool2_bwd_idx == maxpool2_bwd (which recomputes the argmax from the un-pooled tensor)."""
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    d = cuda_device
    y2, ys = (nhwc(bf(torch.randn(N, C, H, W, generator=g)), d) for _ in range(2))

    def bnp():
        return torch.stack([torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.5]).to(d)
    b2, bs = bnp(), bnp()
    out = ops.bn_add_relu(y2, b2, ys, bs)                 # relu -> many exact zeros -> ties inside windows
    want = ops.maxpool2(out)
    pooled, idx = ops.bn_add_relu_pool2(y2, b2, ys, bs)
    assert torch.equal(pooled, want)
    win = out.view(N, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 5, 2, 4).reshape(N, H // 2, W // 2, C, 4).float()
    assert torch.equal(idx.long(), win.argmax(dim=-1).long()) or torch.equal(
        win.gather(-1, idx.long().unsqueeze(-1)).squeeze(-1), pooled.float())
    first = (win == win.max(dim=-1, keepdim=True).values).float().argmax(dim=-1)      # first maximum in scan order
    assert torch.equal(idx.long(), first)
    dp = nhwc(bf(torch.randn(N, C, H // 2, W // 2, generator=g)), d)
    a1 = nhwc(bf(torch.randn(N, C, H, W, generator=g)), d)
    assert torch.equal(ops.maxpool2_bwd_idx(idx, dp), ops.maxpool2_bwd(out, dp))
    assert torch.equal(ops.maxpool2_bwd_idx(idx, dp, a1, a1), ops.maxpool2_bwd(out, dp, a1, a1))


def test_relu_mask_bits_path(cuda_device):
    """bn_add_relu_mask stores (out > 0) as bits; BN backward from the bits == BN backward from the activated tensor."""
    import ctypes
    from real_time_helmet_detection_b200 import ops, _lib
    g = torch.Generator().manual_seed(8)
    d = cuda_device
    N, C, H, W = 2, 128, 16, 12
    y2, x = (nhwc(bf(torch.randn(N, C, H, W, generator=g)), d) for _ in range(2))
    b2 = torch.stack([torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.5, torch.randn(C, generator=g) * 0.1,
                      torch.rand(C, generator=g) + 0.5]).to(d)                      # scale | shift | mean | rstd
    out = torch.empty_like(y2)
    mask = torch.zeros(N * H * W * C // 8, dtype=torch.uint8, device=d)
    L = _lib.lib()
    npix = N * H * W
    _lib.check(L.hd_bn_add_relu_mask(_lib.ptr(y2), _lib.ptr(b2[0]), _lib.ptr(b2[1]), _lib.ptr(x), None, None, _lib.ptr(out),
                                     _lib.ptr(mask), npix, C, _lib.stream()))
    assert torch.equal(out, ops.bn_add_relu(y2, b2, x))
    bits = ((mask.view(-1, 1).int() >> torch.arange(8, device=d).view(1, 8)) & 1).bool().view(N, H, W, C)
    assert torch.equal(bits, out.float() > 0)
    dout = nhwc(bf(torch.randn(N, C, H, W, generator=g)), d)
    gamma = (torch.rand(C, generator=g) + 0.5).to(d)
    ref = ops.bn_bwd(dout, out, y2, b2, gamma, want_g=True)
    sums = torch.zeros(3, C, device=d)
    _lib.check(L.hd_bn_bwd_reduce_fin_mask(_lib.ptr(dout), _lib.ptr(mask), _lib.ptr(y2), _lib.ptr(sums), npix, C, None,
                                           _lib.stream()))
    coef, dgm, dbt = torch.empty(3, C, device=d), torch.empty(C, device=d), torch.empty(C, device=d)
    _lib.check(L.hd_bn_bwd_finalize(_lib.ptr(sums[0]), _lib.ptr(sums[1]), float(npix), _lib.ptr(gamma), _lib.ptr(b2[2]),
                                    _lib.ptr(b2[3]), _lib.ptr(coef), _lib.ptr(dgm), _lib.ptr(dbt), 0, C, _lib.stream()))
    dy, gout = torch.empty_like(y2), torch.empty_like(y2)
    _lib.check(L.hd_bn_bwd_apply_mask(_lib.ptr(dout), _lib.ptr(mask), _lib.ptr(y2), _lib.ptr(coef), _lib.ptr(dy),
                                      _lib.ptr(gout), npix, C, _lib.stream()))
    assert torch.equal(gout, ref[2])
    # the per-channel sums come from fp32 atomics (order varies run to run): dy may differ in the last bf16 bit
    assert (dy.float() - ref[0].float()).abs().max() <= 2 ** -7 * ref[0].float().abs().max()
    assert torch.allclose(dgm, ref[3][0], rtol=1e-5, atol=1e-5) and torch.allclose(dbt, ref[3][1], rtol=1e-5, atol=1e-5)


def test_stem_forward_wgrad(cuda_device):
    """7x7 stride-2 stem (hourglass.py:163) = im2col (K=147->192) + 1x1 tcgen05 GEMM; wgrad via the 1x1 wgrad kernel."""
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(9)
    d = cuda_device
    x = torch.randn(2, 3, 64, 96, generator=g)
    w = bf(torch.randn(64, 3, 7, 7, generator=g) * 0.1)
    b = torch.randn(64, generator=g)
    ref = F.conv2d(bf(x), w, b, stride=2, padding=3)
    patches = ops.stem_im2col(x.to(d))
    stats = torch.zeros(2, 64, device=d)
    y = ops.conv2d_igemm(patches, ops.stem_pack_weight(w.to(d)), 64, 1, bias=b.to(d), stats=stats)
    close_bf16(nchw(y), ref, extra=1e-3 * ref.abs().max().item())
    assert torch.allclose(stats[0].cpu(), ref.sum((0, 2, 3)), rtol=1e-3, atol=1e-2)
    dy = bf(torch.randn(2, 64, 32, 48, generator=g))
    gref = torch.nn.grad.conv2d_weight(bf(x), (64, 3, 7, 7), dy, stride=2, padding=3)
    gw = ops.conv2d_wgrad(patches, nhwc(dy, d), 147, 1, stem_perm=True).cpu()
    assert gw.shape == (64, 3, 7, 7)
    assert ((gw - gref).norm() / gref.norm()).item() <= 1e-4


def test_stem_space_to_depth_path(cuda_device):
    """The default stem: x-unfolded space-to-depth tensor (hd_stem_unfold) + four vertical taps (hd_conv2d_igemm_vtaps,
    weights packed by pack mode 3) + wgrad with the space-to-depth gradient mapping (stem_perm 2) == F.conv2d(7x7, s2, p3)."""
    import ctypes
    from real_time_helmet_detection_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(19)
    d = cuda_device
    N, H, W = 2, 64, 96
    x = bf(torch.randn(N, 3, H, W, generator=g))
    w = bf(torch.randn(64, 3, 7, 7, generator=g) * 0.1)
    b = torch.randn(64, generator=g)
    ref = F.conv2d(x, w, b, stride=2, padding=3)
    Ho, Wo = H // 2, W // 2
    xd = x.to(d).contiguous()
    U = torch.full((N, Ho, Wo, 64), 7.0, dtype=torch.bfloat16, device=d)
    _lib.check(L.hd_stem_unfold(_lib.ptr(xd), _lib.ptr(U), N, H, W, _lib.stream()))
    # U[n, Y, X, dx*12 + (c*2+sy)*2 + sx] = x[n, c, 2Y+sy, 2(X+dx-2)+sx], zero outside / for k >= 48
    xp = F.pad(x, (4, 4, 0, 0))
    want = torch.zeros(N, Ho, Wo, 64)
    for dx in range(4):
        for c in range(3):
            for sy in range(2):
                for sx in range(2):
                    cols = torch.arange(Wo) * 2 + 2 * dx + sx              # 2(X+dx-2)+sx+4 in the padded image
                    want[..., dx * 12 + (c * 2 + sy) * 2 + sx] = xp[:, c, sy::2][:, :, cols]
    assert torch.equal(U.float().cpu(), want)

    class Job(ctypes.Structure):
        _fields_ = [("w", ctypes.c_void_p), ("out", ctypes.c_void_p), ("cout", ctypes.c_int), ("cin", ctypes.c_int),
                    ("taps", ctypes.c_int), ("rows_pad", ctypes.c_int), ("k_pad", ctypes.c_int), ("mode", ctypes.c_int),
                    ("start", ctypes.c_longlong)]
    wd = w.to(d).contiguous()
    wp = torch.empty(4, 64, 64, dtype=torch.bfloat16, device=d)
    job = Job(wd.data_ptr(), wp.data_ptr(), 64, 3, 4, 64, 64, 3, 0)
    jd = torch.frombuffer(bytearray(bytes(job)), dtype=torch.uint8).to(d)
    _lib.check(L.hd_pack_all_weights(_lib.ptr(jd), 1, 4 * 64 * 64, _lib.stream()))
    y = torch.empty(N, Ho, Wo, 64, dtype=torch.bfloat16, device=d)
    stats = torch.zeros(2, 64, device=d)
    bd = b.to(d)
    _lib.check(L.hd_conv2d_igemm_vtaps(_lib.ptr(U), _lib.ptr(wp), _lib.ptr(y), _lib.ptr(bd), _lib.ptr(stats[0]),
                                       _lib.ptr(stats[1]), N, Ho, Wo, 64, 64, 64, 4, 2, 64, None, None, None, 0, _lib.stream()))
    close_bf16(nchw(y), ref, extra=1e-3 * ref.abs().max().item())
    assert torch.allclose(stats[0].cpu(), ref.sum((0, 2, 3)), rtol=1e-3, atol=1e-2)
    dy = bf(torch.randn(N, 64, Ho, Wo, generator=g))
    gref = torch.nn.grad.conv2d_weight(x, (64, 3, 7, 7), dy, stride=2, padding=3)
    gw = torch.full((64, 3, 7, 7), 9.0, device=d)
    dyd = nhwc(dy, d)
    ws = torch.empty(L.hd_conv2d_wgrad_workspace_bytes(N, Ho, Wo, 64, 1), dtype=torch.uint8, device=d)
    _lib.check(L.hd_conv2d_wgrad(_lib.ptr(U), _lib.ptr(dyd), _lib.ptr(gw), _lib.ptr(ws), N, Ho, Wo, 64, 48, 64, 1, 0, 2,
                                 _lib.stream()))
    assert ((gw.cpu() - gref).norm() / gref.norm()).item() <= 1e-4


@pytest.mark.parametrize("N,H,W,with_extra", [(2, 16, 16, True), (3, 20, 12, False), (1, 8, 8, True), (5, 32, 32, False)])
def test_head_backward(cuda_device, N, H, W, with_extra):
    """Head 1x1 conv backward (dfeat, dW, dbias in one kernel), incl. pixel counts that are not a multiple of the 64-pixel
    tile and the optional extra gradient (merge_prediction's dgrad for stacks > 1)."""
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator().manual_seed(4 + H)
    d = cuda_device
    feat = bf(torch.randn(N, 128, H, W, generator=g)).requires_grad_(True)
    w = bf(torch.randn(6, 128, 1, 1, generator=g) * 0.1).requires_grad_(True)
    b = torch.zeros(6, requires_grad=True)
    dlog_full = torch.randn(N, 2, 6, H, W, generator=g)
    extra = bf(torch.randn(N, 6, H, W, generator=g)) if with_extra else None
    F.conv2d(feat, w, b).backward(dlog_full[:, 1] + (extra if with_extra else 0))
    extra_d = ops.to_nhwc(extra.to(d), c_pad=64) if with_extra else None
    dfeat, dw, db = ops.head_backward(dlog_full.to(d)[:, 1], nhwc(feat.detach(), d), ops.pack_weight(w.detach().to(d)),
                                      6, extra=extra_d)
    close_bf16(nchw(dfeat), feat.grad, extra=1e-3 * feat.grad.abs().max().item())
    assert torch.allclose(dw.cpu(), w.grad.view(6, 128), rtol=1e-3, atol=1e-3 * w.grad.abs().max().item())
    assert torch.allclose(db.cpu(), b.grad, rtol=1e-3, atol=1e-3 * b.grad.abs().max().item() + 1e-3)


@pytest.mark.parametrize("N,H,W", [(2, 16, 12), (32, 8, 8), (32, 32, 32), (1, 2, 2)])
def test_bn_backward_fused_small(cuda_device, N, H, W):
    """hd_bn_bwd_fused_small (reduce + coefficients + apply in ONE launch behind an epoch flag; the deep hourglass levels)
    == the two-launch path, for both mask sources (stored bits with the masked gradient output, recomputed mask), and
    back-to-back launches on the same scratch (sums / ticket / epoch are left ready for the next one)."""
    import ctypes
    from real_time_helmet_detection_b200 import ops, _lib

    class Fuse(ctypes.Structure):
        _fields_ = [(k, ctypes.c_void_p) for k in ("gamma", "mean", "rstd", "coef", "dgamma", "dbeta", "gamma_s", "mean_s",
                                                   "rstd_s", "coef_s", "dgamma_s", "dbeta_s")] + \
                   [("count", ctypes.c_float), ("counter", ctypes.c_void_p)]

    g = torch.Generator().manual_seed(80 + H)
    d = cuda_device
    C = 128
    L = _lib.lib()
    npix = N * H * W
    y2, x = (nhwc(bf(torch.randn(N, C, H, W, generator=g)), d) for _ in range(2))
    b2 = torch.stack([torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.5, torch.randn(C, generator=g) * 0.1,
                      torch.rand(C, generator=g) + 0.5]).to(d)                      # scale | shift | mean | rstd
    out = torch.empty_like(y2)
    mask = torch.zeros(npix * C // 8, dtype=torch.uint8, device=d)
    _lib.check(L.hd_bn_add_relu_mask(_lib.ptr(y2), _lib.ptr(b2[0]), _lib.ptr(b2[1]), _lib.ptr(x), None, None, _lib.ptr(out),
                                     _lib.ptr(mask), npix, C, _lib.stream()))
    dout = nhwc(bf(torch.randn(N, C, H, W, generator=g)), d)
    gamma = (torch.rand(C, generator=g) + 0.5).to(d)
    scratch = torch.zeros(16 * 256, device=d)                    # sums [3][256] | coef [3][256] | ... | counter, epoch
    sums, coef = scratch[:768], scratch[768:1536]
    words = scratch[9 * 256:].view(torch.int32)
    dgm, dbt = torch.empty(C, device=d), torch.empty(C, device=d)
    fin = Fuse()
    fin.gamma, fin.mean, fin.rstd = gamma.data_ptr(), b2[2].data_ptr(), b2[3].data_ptr()
    fin.coef, fin.dgamma, fin.dbeta = coef.data_ptr(), dgm.data_ptr(), dbt.data_ptr()
    fin.count, fin.counter = float(npix), words.data_ptr()
    epoch = ctypes.c_void_p(words.data_ptr() + 4)
    for rep in range(3):                                         # back-to-back on the same scratch
        # (1) stored mask bits + masked gradient output: the Residual tail
        ref = ops.bn_bwd(dout, out, y2, b2, gamma, want_g=True)
        dy, gout = torch.empty_like(y2), torch.empty_like(y2)
        _lib.check(L.hd_bn_bwd_fused_small(_lib.ptr(dout), _lib.ptr(mask), None, None, _lib.ptr(y2), _lib.ptr(sums),
                                           _lib.ptr(dy), _lib.ptr(gout), npix, C, ctypes.byref(fin), epoch, _lib.stream()))
        assert torch.equal(gout, ref[2])
        assert (dy.float() - ref[0].float()).abs().max() <= 2 ** -7 * ref[0].float().abs().max() + 1e-6
        assert torch.allclose(dgm, ref[3][0], rtol=1e-4, atol=1e-4) and torch.allclose(dbt, ref[3][1], rtol=1e-4, atol=1e-4)
        # (2) mask recomputed from y * scale + shift: conv1's BN + ReLU
        ref = ops.bn_bwd(dout, None, y2, b2, gamma, remask=True)
        dy2 = torch.empty_like(y2)
        _lib.check(L.hd_bn_bwd_fused_small(_lib.ptr(dout), None, _lib.ptr(b2[0]), _lib.ptr(b2[1]), _lib.ptr(y2),
                                           _lib.ptr(sums), _lib.ptr(dy2), None, npix, C, ctypes.byref(fin), epoch,
                                           _lib.stream()))
        assert (dy2.float() - ref[0].float()).abs().max() <= 2 ** -7 * ref[0].float().abs().max() + 1e-6
        assert torch.allclose(dgm, ref[3][0], rtol=1e-4, atol=1e-4) and torch.allclose(dbt, ref[3][1], rtol=1e-4, atol=1e-4)
    torch.cuda.synchronize()
    assert int(words[0]) == 0 and int(words[1]) == 6             # ticket left at zero, six launches raised the epoch
    assert float(sums.abs().max()) == 0.0                        # accumulators left zeroed


@pytest.mark.parametrize("C,N,H,W", [(128, 2, 16, 24), (128, 3, 64, 64), (64, 1, 8, 8), (128, 1, 2, 2)])
def test_bn_backward_through_fused_pool(cuda_device, C, N, H, W):
    """hd_bn_bwd_reduce_pool_fin / hd_bn_bwd_apply_pool (the pool's backward folded into the two-branch BN backward of
    PreLayer's Residual(64,128)) == hd_maxpool2_bwd_idx -> hd_bn_bwd_reduce_fin -> hd_bn_bwd_apply on the routed gradient:
    coefficients / dgamma / dbeta to fp32 summation-order noise, dy / dys BIT-exact when both applies get the same
    coefficients; the scratch (sums, ticket) is left ready for the next launch."""
    import ctypes
    from real_time_helmet_detection_b200 import ops, _lib

    class Fuse(ctypes.Structure):
        _fields_ = [(k, ctypes.c_void_p) for k in ("gamma", "mean", "rstd", "coef", "dgamma", "dbeta", "gamma_s", "mean_s",
                                                   "rstd_s", "coef_s", "dgamma_s", "dbeta_s")] + \
                   [("count", ctypes.c_float), ("counter", ctypes.c_void_p)]

    g = torch.Generator().manual_seed(C + H + N)
    d = cuda_device
    L = _lib.lib()
    npix = N * H * W
    y2, ys = (nhwc(bf(torch.randn(N, C, H, W, generator=g)), d) for _ in range(2))

    def bnp():      # scale | shift | mean | rstd
        return torch.stack([torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.5,
                            torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5]).to(d)
    b2, bs = bnp(), bnp()
    gamma, gamma_s = ((torch.rand(C, generator=g) + 0.5).to(d) for _ in range(2))
    pooled, idx = ops.bn_add_relu_pool2(y2, b2[:2].contiguous(), ys, bs[:2].contiguous())
    dpool = nhwc(bf(torch.randn(N, C, H // 2, W // 2, generator=g)), d)

    def scratch():
        s = torch.zeros(16 * 256, device=d)
        return s, s[:768], s[768:1536], s[1536:2304], s[9 * 256:].view(torch.int32)

    def fuse(coef, coef_s, words, outs):
        f = Fuse()
        f.gamma, f.mean, f.rstd = gamma.data_ptr(), b2[2].data_ptr(), b2[3].data_ptr()
        f.gamma_s, f.mean_s, f.rstd_s = gamma_s.data_ptr(), bs[2].data_ptr(), bs[3].data_ptr()
        f.coef, f.coef_s = coef.data_ptr(), coef_s.data_ptr()
        f.dgamma, f.dbeta, f.dgamma_s, f.dbeta_s = (o.data_ptr() for o in outs)
        f.count, f.counter = float(npix), words.data_ptr()
        return f

    # unfused: route, then the two-branch reduce / apply with the mask rebuilt from y2 / ys
    dout = ops.maxpool2_bwd_idx(idx, dpool)
    sA, sumsA, coefA, coefsA, wordsA = scratch()
    outsA = [torch.empty(C, device=d) for _ in range(4)]
    fA = fuse(coefA, coefsA, wordsA, outsA)
    _lib.check(L.hd_bn_bwd_reduce_fin(_lib.ptr(dout), None, _lib.ptr(b2[0]), _lib.ptr(b2[1]), _lib.ptr(bs[0]), _lib.ptr(bs[1]),
                                      _lib.ptr(y2), _lib.ptr(ys), _lib.ptr(sumsA), npix, C, ctypes.byref(fA), _lib.stream()))
    dyA, dysA = torch.empty_like(y2), torch.empty_like(y2)
    _lib.check(L.hd_bn_bwd_apply(_lib.ptr(dout), None, _lib.ptr(b2[0]), _lib.ptr(b2[1]), _lib.ptr(bs[0]), _lib.ptr(bs[1]),
                                 _lib.ptr(y2), _lib.ptr(coefA), _lib.ptr(dyA), _lib.ptr(ys), _lib.ptr(coefsA), _lib.ptr(dysA),
                                 None, npix, C, _lib.stream()))
    # folded
    sB, sumsB, coefB, coefsB, wordsB = scratch()
    outsB = [torch.empty(C, device=d) for _ in range(4)]
    fB = fuse(coefB, coefsB, wordsB, outsB)
    for rep in range(2):
        _lib.check(L.hd_bn_bwd_reduce_pool_fin(_lib.ptr(dpool), _lib.ptr(idx), _lib.ptr(b2[0]), _lib.ptr(b2[1]),
                                               _lib.ptr(bs[0]), _lib.ptr(bs[1]), _lib.ptr(y2), _lib.ptr(ys), _lib.ptr(sumsB),
                                               N, H, W, C, ctypes.byref(fB), _lib.stream()))
        for a, b in zip(outsA, outsB):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-3 * float(a.abs().max()) + 1e-6)
        assert torch.allclose(coefA[:3 * C], coefB[:3 * C], rtol=1e-4, atol=1e-5)
        assert torch.allclose(coefsA[:3 * C], coefsB[:3 * C], rtol=1e-4, atol=1e-5)
        torch.cuda.synchronize()
        assert int(wordsB[0]) == 0 and float(sumsB.abs().max()) == 0.0
    dyB, dysB = torch.empty_like(y2), torch.empty_like(y2)
    _lib.check(L.hd_bn_bwd_apply_pool(_lib.ptr(dpool), _lib.ptr(idx), _lib.ptr(b2[0]), _lib.ptr(b2[1]), _lib.ptr(bs[0]),
                                      _lib.ptr(bs[1]), _lib.ptr(y2), _lib.ptr(ys), _lib.ptr(coefA), _lib.ptr(coefsA),
                                      _lib.ptr(dyB), _lib.ptr(dysB), N, H, W, C, _lib.stream()))
    assert torch.equal(dyA, dyB) and torch.equal(dysA, dysB)
