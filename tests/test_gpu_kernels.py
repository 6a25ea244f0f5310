"""GPU parity tests of the individual C-ABI kernels against the CPU oracle (torch fp32 ATen ops
in oracle/pose_oracle.py and the reference-generated vectors in tests/golden/)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pose_oracle as po

pytestmark = pytest.mark.gpu

cuda = torch.device("cuda")


@pytest.fixture(scope="module")
def ops():
    from pose_estimation_amitai_b200 import ops as _ops
    return _ops


def _kat(golden_dir):
    return np.load(os.path.join(golden_dir, "kat.npz"))


# ------------------------------------------------------------------------------------- peaks
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_argmax_kat_nhwc_and_nchw(ops, golden_dir, dtype):
    fx = _kat(golden_dir)
    for key_in, key_out in (("argmax_in", "argmax_out"), ("argmax_big_in", "argmax_big_out")):
        hm = torch.from_numpy(fx[key_in].astype(np.float32))
        if dtype == torch.bfloat16:
            hm = hm.to(torch.bfloat16)
            want = po.find_peaks_argmax(hm.float())  # peaks are defined on the tensor as given
        else:
            want = fx[key_out]
        got_nhwc = ops.peaks_argmax(hm.to(cuda), layout="nhwc").cpu().numpy()
        np.testing.assert_array_equal(got_nhwc, want)
        nchw = hm.permute(0, 3, 1, 2).contiguous().to(cuda)
        got_nchw, vals = ops.peaks_argmax(nchw, layout="nchw", want_values=True)
        np.testing.assert_array_equal(got_nchw.cpu().numpy(), want)
        ref_vals = hm.float().reshape(hm.shape[0], -1, hm.shape[3]).max(dim=1).values
        np.testing.assert_array_equal(vals.cpu().numpy(), ref_vals.numpy())
        # channels_last view of an NCHW tensor (what the reference's trainer transposes to)
        got_view = ops.peaks_argmax(nchw.permute(0, 2, 3, 1), layout="nhwc").cpu().numpy()
        np.testing.assert_array_equal(got_view, want)


def test_argmax_full_size_properties(ops):
    g = torch.Generator().manual_seed(11)
    n, c = 64, 36
    hm = torch.rand(n, c, 192, 192, generator=g).to(cuda)
    ys = torch.randint(0, 192, (n, c), generator=g)
    xs = torch.randint(0, 192, (n, c), generator=g)
    idx_n = torch.arange(n).view(n, 1).expand(n, c)
    idx_c = torch.arange(c).view(1, c).expand(n, c)
    hm[idx_n, idx_c, ys, xs] = 2.0
    # a later duplicate of the maximum must not win
    hm[0, 0, 191, 191] = 2.0
    peaks = ops.peaks_argmax(hm).cpu()
    assert torch.equal(peaks[..., 0], xs.float()) and torch.equal(peaks[..., 1], ys.float())
    assert ops.peaks_argmax(hm[:0]).shape == (0, c, 2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_argmax_vector_scan_is_bit_exact(ops, dtype, monkeypatch):
    """the two-phase planar scan (per-16-byte-vector maxima, then the element rule inside the winner) returns the
    same peaks and maxima as the oracle (Augmentor.tf_find_peaks, pytorch/Augmentor.py:105-148) and as the element-wise
    kernel it replaces (POSEB200_ARGMAX_V1): heavy ties, NaNs, -0 / +0, -inf maps; contiguous maps, row-strided views
    and a width that is not a whole number of vectors (falls back to the element-wise kernel)."""
    g = torch.Generator().manual_seed(5)
    v = 16 // torch.empty((), dtype=dtype).element_size()
    for (n, c, h, w, wpad) in ((2, 3, 192, 192, 0), (3, 5, 24, 2 * v, 0), (2, 4, 20, 3 * v, v), (2, 3, 17, v + 3, 0),
                               (1, 2, 1, v, 0), (8, 36, 192, 192, 0)):
        full = torch.randint(-3, 4, (n, c, h, w + wpad), generator=g).float()       # 7 values: ties everywhere
        full[0, 0] = float("-inf")
        full[0, 1] = -0.0
        full[0, 1, h // 2, (w // 2):] = 0.0
        if n > 1:
            full[1, 0, h - 1, w - 1] = float("nan")
            full[1, 1, h // 2, 1:] = float("nan")                                     # first NaN is not a vector's first element
            full[1, 2, :, :] = 9.0
            full[1, 2, 0, min(w - 1, v + 1)] = float("inf")
        full = full.to(dtype).to(cuda)
        hm = full[..., :w] if wpad else full
        want = po.find_peaks_argmax(hm.float().cpu().permute(0, 2, 3, 1).contiguous())
        want_v = hm.float().cpu().reshape(n, c, -1).max(dim=2).values.numpy()
        got, got_v = ops.peaks_argmax(hm, want_values=True)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
        np.testing.assert_array_equal(got_v.cpu().numpy(), want_v + 0.0)
        monkeypatch.setenv("POSEB200_ARGMAX_V1", "1")
        old, old_v = ops.peaks_argmax(hm, want_values=True)
        monkeypatch.delenv("POSEB200_ARGMAX_V1")
        assert torch.equal(old, got)
        np.testing.assert_array_equal(old_v.cpu().numpy(), got_v.cpu().numpy())


def test_softargmax_kat(ops, golden_dir):
    fx = _kat(golden_dir)
    hm = torch.from_numpy(fx["soft_in"].astype(np.float32))
    got = ops.peaks_softargmax(hm.to(cuda), layout="nhwc").cpu().numpy()
    np.testing.assert_allclose(got, fx["soft_out"], rtol=1e-4, atol=2e-3)
    got2 = ops.peaks_softargmax(hm.permute(0, 3, 1, 2).contiguous().to(cuda)).cpu().numpy()
    np.testing.assert_allclose(got2, fx["soft_out"], rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_softargmax_table_form_matches_direct_form(ops, dtype, monkeypatch):
    """the planar soft arg-max with its linspace weights in shared-memory tables returns what the per-element form
    (POSEB200_SOFTARGMAX_V1) returns (same products in the same order; 1e-4 px allowed for a differently contracted
    weight), and both match the oracle (find_peaks_soft_argmax, pytorch/utils.py:47-83) within the 2e-3 px gate:
    contiguous maps, row-strided views, odd widths (per-element form)."""
    g = torch.Generator().manual_seed(9)
    v = 16 // torch.empty((), dtype=dtype).element_size()
    for (n, c, h, w, wpad) in ((2, 3, 192, 192, 0), (3, 5, 24, 2 * v, 0), (2, 4, 20, 3 * v, v), (2, 3, 17, v + 3, 0),
                               (1, 2, 2, v, 0), (4, 36, 192, 192, 0), (1, 1, 3, 40 * v, 0)):
        full = (torch.rand(n, c, h, w + wpad, generator=g) + 0.05).to(dtype).to(cuda)
        hm = full[..., :w] if wpad else full
        got = ops.peaks_softargmax(hm)
        monkeypatch.setenv("POSEB200_SOFTARGMAX_V1", "1")
        old = ops.peaks_softargmax(hm)
        monkeypatch.delenv("POSEB200_SOFTARGMAX_V1")
        np.testing.assert_allclose(got.cpu().numpy(), old.cpu().numpy(), rtol=1e-6, atol=1e-4, err_msg=str((n, c, h, w, wpad)))
        want = po.find_peaks_soft_argmax(hm.float().cpu().permute(0, 2, 3, 1).contiguous().numpy())
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=2e-3)


# ------------------------------------------------------------------------------------- targets / loss
def test_gaussian_kat(ops, golden_dir):
    fx = _kat(golden_dir)
    pts = torch.from_numpy(fx["gauss_means"].astype(np.float32)).view(1, -1, 2).to(cuda)
    got = ops.gaussian_heatmaps(pts)[0].cpu().numpy()
    # reference renders in float64 (simple_data_generator.py:119-125); fp32 expf tolerance
    np.testing.assert_allclose(got, fx["gauss_out"], rtol=1e-5, atol=1e-30)
    got6 = ops.gaussian_heatmaps(pts[:, :1], sigma=6.0)[0, 0].cpu().numpy()
    np.testing.assert_allclose(got6, fx["gauss_sigma6"], rtol=1e-5, atol=1e-30)
    peaks = ops.peaks_argmax(ops.gaussian_heatmaps(pts[:, :1])).cpu().numpy()
    np.testing.assert_array_equal(peaks[0, 0], [120.0, 50.0])  # render -> peak round trip


def test_gaussian_separable_render_vs_float64_reference(ops, monkeypatch):
    """the separable renderer (column / row factors in double, one product per pixel) against the reference's float64
    rendering (SimpleDataGenerator.get_gaussian, tensorflow/simple_data_generator.py:119-125) rounded to fp32:
    fractional, off-image and far-corner means, three sigmas, non-square maps, few maps (row-split grid) and many.
    2e-6 relative -- five times tighter than the per-pixel fp32 expf form (POSEB200_GAUSS_V1) is held to."""
    rs = np.random.RandomState(21)
    for (b, c, h, w, sigma) in ((2, 3, 192, 192, 3.0), (1, 1, 192, 192, 6.0), (3, 5, 48, 64, 1.5), (40, 36, 96, 96, 3.0),
                                (1, 2, 7, 8, 3.0)):
        pts = rs.uniform(-4.0, max(h, w) + 4.0, size=(b, c, 2)).astype(np.float32)
        pts[0, 0] = (0.0, 0.0)
        pts[-1, -1] = (w - 1.0, h - 1.0)
        want = np.stack([np.stack([po.gaussian_heatmap(pts[i, j].astype(np.float64), sigma, (w, h))
                                   for j in range(c)]) for i in range(b)]).astype(np.float32)
        got = ops.gaussian_heatmaps(torch.from_numpy(pts).to(cuda), sigma=sigma, size=(h, w)).cpu().numpy()
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-30)
        monkeypatch.setenv("POSEB200_GAUSS_V1", "1")
        old = ops.gaussian_heatmaps(torch.from_numpy(pts).to(cuda), sigma=sigma, size=(h, w)).cpu().numpy()
        monkeypatch.delenv("POSEB200_GAUSS_V1")
        np.testing.assert_allclose(old, want, rtol=1e-4, atol=1e-30)


@pytest.mark.parametrize("c,cpad", [(36, 36), (36, 48), (18, 32), (36, 64), (5, 8)])
def test_mse_loss_and_grad(ops, c, cpad):
    g = torch.Generator().manual_seed(3)
    b = 3
    out = (torch.rand(b, c, 192, 192, generator=g) - 0.3)
    pts = po.synthetic_points(b, c, seed=4)
    tgt = torch.from_numpy(po.gaussian_targets(pts))
    acc = 3
    want_loss = po.mse_loss(out, tgt, acc).item()
    want_grad = po.mse_loss_grad(out, tgt, acc)
    for fused in (False, True):
        loss_sum, g_nchw, g_nhwc = ops.mse_loss_fwd_bwd(
            out.to(cuda), None if fused else tgt.to(cuda), points=torch.from_numpy(pts).to(cuda) if fused else None,
            accumulation_steps=acc, want_grad_nchw=True, grad_nhwc_dtype=torch.float32, cpad=cpad)
        got_loss = loss_sum.item() / out.numel() / acc
        assert abs(got_loss - want_loss) <= 2e-6 * abs(want_loss)
        np.testing.assert_allclose(g_nchw.cpu().numpy(), want_grad.numpy(), rtol=1e-5, atol=1e-12)
        want_nhwc = (want_grad * torch.where(out > 0, 1.0, 0.1)).permute(0, 2, 3, 1)
        np.testing.assert_allclose(g_nhwc[..., :c].cpu().numpy(), want_nhwc.numpy(), rtol=1e-5, atol=1e-12)
        assert torch.count_nonzero(g_nhwc[..., c:]).item() == 0
        # training path: bf16 NHWC gradient only (the channel-pair kernel when cpad % 8 == 0)
        loss_b, none_nchw, g_b = ops.mse_loss_fwd_bwd(
            out.to(cuda), None if fused else tgt.to(cuda), points=torch.from_numpy(pts).to(cuda) if fused else None,
            accumulation_steps=acc, grad_nhwc_dtype=torch.bfloat16, cpad=cpad)
        assert none_nchw is None
        # the fused path renders the target with one ex2.approx per element: |target error| <= 4e-6 absolute
        tgt_tol = 4e-6 if fused else 0.0
        assert abs(loss_b.item() / out.numel() / acc - want_loss) <= (2e-6 + 2 * tgt_tol) * abs(want_loss)
        np.testing.assert_allclose(g_b[..., :c].float().cpu().numpy(), want_nhwc.numpy(), rtol=2 ** -8,
                                   atol=1e-12 + tgt_tol * 2.0 / (out.numel() * acc))
        assert torch.count_nonzero(g_b[..., c:]).item() == 0
    # ingest path == same thing from an upstream gradient
    gi = ops.grad_ingest(want_grad.to(cuda), out.to(cuda), torch.float32, cpad=cpad)
    np.testing.assert_allclose(gi[..., :c].cpu().numpy(), want_nhwc.numpy(), rtol=1e-6, atol=1e-12)


def test_adam_matches_oracle(ops):
    g = torch.Generator().manual_seed(5)
    n = 100003
    p, gr = torch.randn(n, generator=g), torch.randn(n, generator=g) * 1e-3
    m, v = torch.zeros(n), torch.zeros(n)
    pc, mc, vc = p.to(cuda), m.to(cuda), v.to(cuda)
    for step in (1, 2, 3):
        p, m, v = po.adam_step(p, gr, m, v, step)
        ops.adam_step(pc, gr.to(cuda), mc, vc, step)
    np.testing.assert_allclose(pc.cpu().numpy(), p.numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(vc.cpu().numpy(), v.numpy(), rtol=1e-5, atol=1e-12)
    opt_p = torch.nn.Parameter(torch.randn(n, generator=torch.Generator().manual_seed(5)))
    opt = torch.optim.Adam([opt_p], lr=1e-3)
    for _ in range(3):
        opt_p.grad = gr.clone()
        opt.step()
    np.testing.assert_allclose(pc.cpu().numpy(), opt_p.detach().numpy(), rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------- contractions
def _ref_layer(kind, x, w, b, dilation):
    if kind == "conv":
        return F.conv2d(x, w, b, padding=dilation, dilation=dilation)
    if kind == "convT1":
        return F.conv_transpose2d(x, w, b, stride=1, padding=1)
    if kind == "convT2":
        return F.conv_transpose2d(x, w, b, stride=2, padding=1, output_padding=1)
    raise ValueError(kind)


def _layer_case(ops, kind, cin, cout, h, w, dilation, impl, dtype, seed=0):
    """forward (bias + lrelu + residual + mask), dgrad and wgrad of one layer vs torch CPU autograd."""
    g = torch.Generator().manual_seed(seed)
    n = 2
    spec = ops.Contraction(kind, cin, cout, dilation=dilation)
    wshape = (cout, cin, 3, 3) if kind == "conv" else (cin, cout, 3, 3)
    wt = (torch.rand(wshape, generator=g) - 0.5) * (2.0 / (3 * cin ** 0.5))
    bias = torch.rand(cout, generator=g) - 0.5
    x = torch.rand(n, cin, h, w, generator=g) - 0.5
    if dtype == torch.bfloat16:  # identical operand values on both sides
        wt, x = wt.bfloat16().float(), x.bfloat16().float()
    oh, ow = spec.out_hw(h, w)
    res = (torch.rand(n, cout, oh, ow, generator=g) - 0.5)
    if dtype == torch.bfloat16:
        res = res.bfloat16().float()
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    pre = _ref_layer(kind, xr, wr, br, dilation)
    y = F.leaky_relu(pre, 0.1) + res
    gy = torch.rand(y.shape, generator=g) - 0.5
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    # dC = gy * lrelu'(pre): what our dgrad / wgrad consume
    dc = gy * torch.where(pre.detach() > 0, 1.0, 0.1)
    if dtype == torch.bfloat16:
        dc = dc.bfloat16().float()
    pre.backward(dc)

    tol = dict(rtol=2e-2, atol=2e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-5)
    to_nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(cuda, dtype)
    wc = wt.to(cuda)
    xg, resg, dcg = to_nhwc(x), to_nhwc(res), to_nhwc(dc)
    mask = torch.zeros((n * oh * ow, (cout + 31) // 32), device=cuda, dtype=torch.int32)
    if impl == "tc":
        from pose_estimation_amitai_b200 import tc_support
        wf = ops.pack_weights(wc, spec, "oi", torch.bfloat16, ipad=tc_support.pad_n(cout))
        wd = ops.pack_weights(wc, spec, "io", torch.bfloat16)
    else:
        wf = ops.pack_weights(wc, spec, "io", torch.float32)
        wd = ops.pack_weights(wc, spec, "oi", torch.float32)
    yg = ops.conv(impl, xg, wf, spec.fwd_taps(), n, h, w, cin, oh, ow, cout, bias=bias.to(cuda),
                  act=ops.PB_ACT_LRELU, add1=resg, mask_out=mask, act_dtype=dtype)
    np.testing.assert_allclose(yg.float().cpu().permute(0, 3, 1, 2).numpy(), y.detach().numpy(), **tol)
    # sign mask == (pre > 0) wherever |pre| is clear of rounding noise
    bits = ((mask.view(n, oh, ow, -1, 1) >> torch.arange(32, device=cuda, dtype=torch.int32)) & 1).reshape(
        n, oh, ow, -1)[..., :cout].bool().cpu().permute(0, 3, 1, 2)
    clear = pre.detach().abs() > (1e-2 if dtype == torch.bfloat16 else 1e-5)
    assert torch.equal(bits[clear], (pre.detach() > 0)[clear])
    # dgrad (+ skip add + mask-mul epilogue exercised in the network tests)
    gx = ops.conv(impl, dcg, wd, spec.dgrad_taps(), n, oh, ow, cout, h, w, cin, act_dtype=dtype)
    np.testing.assert_allclose(gx.float().cpu().permute(0, 3, 1, 2).numpy(), xr.grad.numpy(), **tol)
    # wgrad
    dw, db = torch.empty_like(wc), torch.empty(cout, device=cuda)
    ops.wgrad("simt" if impl == "simt" else "tc", spec, xg, dcg, n, h, w, dw, db, act_dtype=dtype)
    scale = wr.grad.abs().max().item()
    np.testing.assert_allclose(dw.cpu().numpy(), wr.grad.numpy(), rtol=tol["rtol"], atol=tol["rtol"] * scale)
    np.testing.assert_allclose(db.cpu().numpy(), br.grad.numpy(), rtol=tol["rtol"],
                               atol=tol["rtol"] * br.grad.abs().max().item())


@pytest.mark.parametrize("kind,cin,cout,h,w,dil", [
    ("conv", 4, 64, 24, 20, 2), ("conv", 64, 64, 16, 24, 2), ("conv", 32, 80, 10, 12, 1),
    ("convT1", 64, 64, 12, 12, 1), ("convT2", 64, 36, 12, 16, 1), ("convT2", 128, 64, 6, 6, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_simt_layer(ops, kind, cin, cout, h, w, dil, dtype):
    _layer_case(ops, kind, cin, cout, h, w, dil, "simt", dtype)


# ------------------------------------------------------------------------------------- tcgen05 building blocks
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("n,k", [(64, 64), (128, 192), (256, 128)])
def test_tcgen05_selftest_gemm(a_mn, b_mn, n, k):
    from pose_estimation_amitai_b200 import _lib
    g = torch.Generator().manual_seed(a_mn * 2 + b_mn + n + k)
    a = (torch.randint(-4, 5, (128, k), generator=g).float() / 4).bfloat16()
    b = (torch.randint(-4, 5, (n, k), generator=g).float() / 4).bfloat16()
    want = a.float() @ b.float().t()  # exact in fp32 for these small dyadic values
    ad = (a.t().contiguous() if a_mn else a).to(cuda)
    bd = (b.t().contiguous() if b_mn else b).to(cuda)
    d = torch.full((128, n), float("nan"), device=cuda)
    args = _lib.STRUCTS["pb_gemm_selftest_args"]()
    args.a, args.b, args.d = ad.data_ptr(), bd.data_ptr(), d.data_ptr()
    args.M, args.N, args.K, args.a_mn_major, args.b_mn_major = 128, n, k, a_mn, b_mn
    _lib.call("pb_gemm_selftest", args, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d.cpu().numpy(), want.numpy())


@pytest.mark.parametrize("a_f16,b_f16", [(1, 1)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (0, 1)])
def test_tcgen05_selftest_gemm_fp16(a_f16, b_f16, a_mn, b_mn):
    """fp16 x fp16 operands (the "fp16" precision's forward).  Operand values use fp16's three extra significand
    bits (multiples of 1/512 up to 1), which bf16 cannot hold, so a descriptor that decoded them as bf16 would not
    reproduce the exact fp32 product.  Mixed formats (fp16 x bf16) are NOT tested: measured on the B200 they raise
    cudaErrorIllegalInstruction (gpurun_out/r2a_tests.log), which is why the weight gradients read a bf16 twin."""
    from pose_estimation_amitai_b200 import _lib
    n, k = 128, 128
    g = torch.Generator().manual_seed(7 + 2 * a_f16 + b_f16)
    def draw(rows, f16):
        if f16:
            return (torch.randint(-512, 513, (rows, k), generator=g).float() / 512).half()
        return (torch.randint(-4, 5, (rows, k), generator=g).float() / 4).bfloat16()
    a, b = draw(128, a_f16), draw(n, b_f16)
    want = (a.double() @ b.double().t()).float()
    ad = (a.t().contiguous() if a_mn else a).to(cuda)
    bd = (b.t().contiguous() if b_mn else b).to(cuda)
    d = torch.full((128, n), float("nan"), device=cuda)
    args = _lib.STRUCTS["pb_gemm_selftest_args"]()
    args.a, args.b, args.d = ad.data_ptr(), bd.data_ptr(), d.data_ptr()
    args.M, args.N, args.K, args.a_mn_major, args.b_mn_major = 128, n, k, a_mn, b_mn
    args.a_f16, args.b_f16 = a_f16, b_f16
    _lib.call("pb_gemm_selftest", args, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_allclose(d.cpu().numpy(), want.numpy(), rtol=0, atol=2e-5)


@pytest.mark.parametrize("kind,cin,cout,h,w,dil", [
    ("conv", 64, 64, 16, 32, 2), ("conv", 128, 256, 24, 24, 2), ("conv", 256, 256, 48, 48, 2),
    ("convT1", 128, 128, 20, 12, 1), ("convT2", 256, 128, 12, 12, 1), ("convT2", 128, 36, 24, 24, 1),
    ("convT2", 256, 256, 24, 16, 1), ("conv", 64, 128, 40, 24, 2),
    ("convT2", 1280, 640, 8, 8, 1), ("convT1", 640, 640, 16, 16, 1)])   # four-camera decoder widths (CNNs.py:210-216)
def test_tc_layer_fwd_dgrad(ops, kind, cin, cout, h, w, dil):
    """tcgen05 forward and input-gradient contraction vs torch CPU (wgrad checked separately)."""
    _tc_fwd_dgrad_case(ops, kind, cin, cout, h, w, dil, 2)


@pytest.mark.parametrize("kind,cin,cout,h,w,dil,n", [
    ("conv", 64, 64, 16, 48, 2, 3),      # 9 pixel groups: the last cta pair runs a padding group
    ("conv", 128, 128, 16, 16, 2, 1),    # a single group: one CTA of the only pair is padding
    ("convT2", 256, 128, 16, 8, 1, 3),   # stride-2 transposed conv, 2 passes, odd group count
    ("conv", 128, 64, 32, 16, 2, 1)])    # 64 output channels with two K chunks: resident half-tiles per CTA
def test_tc_pair_ragged_group_counts(ops, kind, cin, cout, h, w, dil, n):
    """cta_group::2 pair mode when the number of pixel groups is odd (or 1): the peer CTA's padding group must
    contribute nothing and store nothing."""
    _tc_fwd_dgrad_case(ops, kind, cin, cout, h, w, dil, n)


@pytest.mark.parametrize("kind,cin,cout,h,w,dil", [
    ("conv", 64, 64, 16, 32, 2), ("conv", 64, 128, 40, 24, 2), ("conv", 256, 256, 48, 48, 2),
    ("convT1", 128, 128, 20, 12, 1), ("convT2", 256, 128, 12, 12, 1), ("convT2", 128, 36, 24, 24, 1),
    ("linear", 64, 64, 32, 32, 1), ("linear", 256, 128, 1, 512, 1)])   # the second takes the per-tap kernel (tc_conv.cu)
def test_tc_layer_fwd_fp16(ops, kind, cin, cout, h, w, dil):
    """the "fp16" precision's forward: IEEE-half activations / weights / residual / output, the bf16 twin the weight
    gradient reads, the NCHW fp32 head.  Operand values are exact in fp16 (NOT in bf16), so decoding any operand as
    bf16 -- or rounding the output to bf16 -- fails the fp16-sized tolerance."""
    from pose_estimation_amitai_b200 import tc_support
    g = torch.Generator().manual_seed(2)
    n = 2
    spec = ops.Contraction(kind, cin, cout, dilation=dil)
    k = 1 if kind == "linear" else 3
    wshape = (cout, cin) if kind == "linear" else ((cout, cin, 3, 3) if kind == "conv" else (cin, cout, 3, 3))
    wt = ((torch.rand(wshape, generator=g) - 0.5) * (2.0 / (k * cin ** 0.5))).half().float()
    bias = torch.rand(cout, generator=g) - 0.5
    x = (torch.rand(n, cin, h, w, generator=g) - 0.5).half().float()
    oh, ow = spec.out_hw(h, w)
    res = (torch.rand(n, cout, oh, ow, generator=g) - 0.5).half().float()
    pre = F.conv2d(x, wt.view(cout, cin, 1, 1), bias) if kind == "linear" else _ref_layer(kind, x, wt, bias, dil)
    y = F.leaky_relu(pre, 0.1) + res
    to_nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(cuda, torch.float16)
    wf = ops.pack_weights(wt.to(cuda), spec, "oi", torch.float16, ipad=tc_support.pad_n(cout))
    assert wf.dtype == torch.float16
    mask = torch.zeros((n * oh * ow, (cout + 31) // 32), device=cuda, dtype=torch.int32)
    twin = torch.zeros((n, oh, ow, cout), device=cuda, dtype=torch.bfloat16)
    yg = ops.conv("tc", to_nhwc(x), wf, spec.fwd_taps(), n, h, w, cin, oh, ow, cout, bias=bias.to(cuda),
                  act=ops.PB_ACT_LRELU, add1=to_nhwc(res), mask_out=mask, act_dtype=torch.float16, out2=twin)
    torch.cuda.synchronize()
    assert yg.dtype == torch.float16
    want = y.numpy()
    got = yg.float().cpu().permute(0, 3, 1, 2).numpy()
    np.testing.assert_allclose(got, want, rtol=1.5e-3, atol=1e-3)       # fp16 output rounding: 2^-11 relative
    assert np.abs(got - want).max() < 0.25 * np.abs(yg.bfloat16().float().cpu().permute(0, 3, 1, 2).numpy() - want).max()
    # the twin is the bf16 rounding of the same fp32 value (allow the double rounding through fp16 to differ by an ulp)
    np.testing.assert_allclose(twin.float().cpu().permute(0, 3, 1, 2).numpy(), want, rtol=8e-3, atol=4e-3)
    if kind != "linear":
        yn = ops.conv("tc", to_nhwc(x), wf, spec.fwd_taps(), n, h, w, cin, oh, ow, cout, bias=bias.to(cuda),
                      act=ops.PB_ACT_LRELU, act_dtype=torch.float16, out_nchw=True)
        np.testing.assert_allclose(yn.cpu().numpy(), F.leaky_relu(pre, 0.1).numpy(), rtol=1e-4, atol=1e-4)


def test_pool_fp16_and_twin(ops):
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(2, 12, 16, 64, generator=g) - 0.4).half()
    want = F.leaky_relu(F.max_pool2d(x.float().permute(0, 3, 1, 2), 2, 2), 0.1).permute(0, 2, 3, 1)
    y, y2 = ops.maxpool_lrelu_fwd(x.to(cuda), twin=True)
    assert y.dtype == torch.float16 and y2.dtype == torch.bfloat16
    assert torch.equal(y.cpu(), want.half()) and torch.equal(y2.cpu(), want.bfloat16())
    # backward: fp16 forward activations select the arg-max, gradients are bf16
    gy = (torch.rand(2, 6, 8, 64, generator=g) - 0.5).bfloat16()
    mask = torch.randint(-2 ** 31, 2 ** 31 - 1, (2 * 12 * 16, 2), generator=g, dtype=torch.int64).to(torch.int32)
    gx, gxm = ops.maxpool_lrelu_bwd(x.to(cuda), gy.to(cuda), mask.to(cuda))
    gx_ref, gxm_ref = ops.maxpool_lrelu_bwd(x.bfloat16().to(cuda), gy.to(cuda), mask.to(cuda))
    assert gx.dtype == torch.bfloat16
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.leaky_relu(F.max_pool2d(xr, 2, 2), 0.1).backward(gy.float().permute(0, 3, 1, 2))
    assert torch.equal(gx.float().cpu(), xr.grad.permute(0, 2, 3, 1).bfloat16().float())
    assert gxm_ref.shape == gxm.shape


def _tc_fwd_dgrad_case(ops, kind, cin, cout, h, w, dil, n):
    from pose_estimation_amitai_b200 import tc_support
    g = torch.Generator().manual_seed(1)
    spec = ops.Contraction(kind, cin, cout, dilation=dil)
    wshape = (cout, cin, 3, 3) if kind == "conv" else (cin, cout, 3, 3)
    wt = ((torch.rand(wshape, generator=g) - 0.5) * (2.0 / (3 * cin ** 0.5))).bfloat16().float()
    bias = torch.rand(cout, generator=g) - 0.5
    x = (torch.rand(n, cin, h, w, generator=g) - 0.5).bfloat16().float()
    oh, ow = spec.out_hw(h, w)
    res = (torch.rand(n, cout, oh, ow, generator=g) - 0.5).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    pre = _ref_layer(kind, xr, wt, bias, dil)
    y = F.leaky_relu(pre, 0.1) + res
    dc = (torch.rand(y.shape, generator=g) - 0.5).bfloat16().float()
    pre.backward(dc)
    to_nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(cuda, torch.bfloat16)
    wc = wt.to(cuda)
    wf = ops.pack_weights(wc, spec, "oi", torch.bfloat16, ipad=tc_support.pad_n(cout))
    mask = torch.zeros((n * oh * ow, (cout + 31) // 32), device=cuda, dtype=torch.int32)
    yg = ops.conv("tc", to_nhwc(x), wf, spec.fwd_taps(), n, h, w, cin, oh, ow, cout, bias=bias.to(cuda),
                  act=ops.PB_ACT_LRELU, add1=to_nhwc(res), mask_out=mask, act_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    np.testing.assert_allclose(yg.float().cpu().permute(0, 3, 1, 2).numpy(), y.detach().numpy(), rtol=2e-2, atol=2e-2)
    bits = ((mask.view(n, oh, ow, -1, 1) >> torch.arange(32, device=cuda, dtype=torch.int32)) & 1).reshape(
        n, oh, ow, -1)[..., :cout].bool().cpu().permute(0, 3, 1, 2)
    clear = pre.detach().abs() > 1e-2
    assert torch.equal(bits[clear], (pre.detach() > 0)[clear])
    # NCHW fp32 output variant (network head)
    yn = ops.conv("tc", to_nhwc(x), wf, spec.fwd_taps(), n, h, w, cin, oh, ow, cout, bias=bias.to(cuda),
                  act=ops.PB_ACT_LRELU, act_dtype=torch.bfloat16, out_nchw=True)
    np.testing.assert_allclose(yn.cpu().numpy(), F.leaky_relu(pre.detach(), 0.1).numpy(), rtol=1e-3, atol=1e-3)
    # input gradient with skip-add, G output and LeakyReLU' mask multiply
    cpad = (cout + 7) // 8 * 8
    dcg = torch.zeros((n, oh, ow, cpad), device=cuda, dtype=torch.bfloat16)
    dcg[..., :cout] = to_nhwc(dc)
    wd = ops.pack_weights(wc, spec, "io", torch.bfloat16, jpad=cpad)
    skip = (torch.rand(n, cin, h, w, generator=g) - 0.5).bfloat16().float()
    mprev = torch.randint(-2 ** 31, 2 ** 31 - 1, (n * h * w, (cin + 31) // 32), generator=g, dtype=torch.int64).to(
        torch.int32).to(cuda)
    gpre = torch.empty((n, h, w, cin), device=cuda, dtype=torch.bfloat16)
    gx = ops.conv("tc", dcg, wd, spec.dgrad_taps(), n, oh, ow, cpad, h, w, cin, add0=to_nhwc(skip), pre_out=gpre,
                  act=ops.PB_ACT_MASKMUL, mask_in=mprev, act_dtype=torch.bfloat16)
    want_g = xr.grad + skip
    np.testing.assert_allclose(gpre.float().cpu().permute(0, 3, 1, 2).numpy(), want_g.numpy(), rtol=2e-2, atol=2e-2)
    mbits = ((mprev.view(n, h, w, -1, 1) >> torch.arange(32, device=cuda, dtype=torch.int32)) & 1).reshape(
        n, h, w, -1)[..., :cin].bool().cpu().permute(0, 3, 1, 2)
    want_dc = want_g * torch.where(mbits, 1.0, 0.1)
    np.testing.assert_allclose(gx.float().cpu().permute(0, 3, 1, 2).numpy(), want_dc.numpy(), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("kind,cin,cout,h,w,dil,cpad", [
    ("conv", 64, 64, 16, 32, 2, 64), ("conv", 128, 256, 24, 24, 2, 256), ("conv", 256, 256, 12, 20, 2, 256),
    ("conv", 64, 128, 24, 24, 2, 128), ("convT1", 128, 128, 20, 12, 1, 128), ("convT2", 256, 128, 12, 12, 1, 128),
    ("convT2", 128, 36, 24, 24, 1, 48),
    # csrc/tc_wgrad_up.cu (all nine taps of the narrow stride-2 head per CTA): ragged tiles, 2 ci blocks, N = 16 / 32 / 48
    ("convT2", 128, 36, 33, 17, 1, 48), ("convT2", 256, 18, 16, 8, 1, 32), ("convT2", 128, 5, 20, 12, 1, 8),
    ("convT2", 1280, 640, 8, 8, 1, 640), ("convT1", 640, 640, 16, 16, 1, 640), ("convT1", 640, 640, 6, 6, 1, 640)])
def test_tc_wgrad(ops, kind, cin, cout, h, w, dil, cpad):
    """tcgen05 weight gradient (MN-major operands straight from NHWC) vs torch CPU autograd.
    Operands are exactly representable in bf16, products accumulate in fp32 on both sides."""
    g = torch.Generator().manual_seed(2)
    n = 3
    spec = ops.Contraction(kind, cin, cout, dilation=dil)
    wshape = (cout, cin, 3, 3) if kind == "conv" else (cin, cout, 3, 3)
    wt = torch.zeros(wshape, requires_grad=True)
    bias = torch.zeros(cout, requires_grad=True)
    x = (torch.randint(-8, 9, (n, cin, h, w), generator=g).float() / 8)
    oh, ow = spec.out_hw(h, w)
    dc = (torch.randint(-8, 9, (n, cout, oh, ow), generator=g).float() / 8)
    _ref_layer(kind, x, wt, bias, dil).backward(dc)
    xg = x.permute(0, 2, 3, 1).contiguous().to(cuda, torch.bfloat16)
    dcg = torch.zeros((n, oh, ow, cpad), device=cuda, dtype=torch.bfloat16)
    dcg[..., :cout] = dc.permute(0, 2, 3, 1).to(cuda, torch.bfloat16)
    dw = torch.full(wshape, float("nan"), device=cuda)
    db = torch.full((cout,), float("nan"), device=cuda)
    ops.wgrad("tc", spec, xg, dcg, n, h, w, dw, db, act_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    np.testing.assert_allclose(dw.cpu().numpy(), wt.grad.numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(db.cpu().numpy(), bias.grad.numpy(), rtol=1e-5, atol=1e-3)
    # accumulate mode (accumulation_steps > 1): dw = 1*dw + new
    ops.wgrad("tc", spec, xg, dcg, n, h, w, dw, db, act_dtype=torch.bfloat16, beta=1.0)
    np.testing.assert_allclose(dw.cpu().numpy(), 2 * wt.grad.numpy(), rtol=1e-5, atol=2e-3)


@pytest.mark.parametrize("shape", [(2, 144, 3, 256), (1, 64, 2, 128), (3, 128, 1, 256), (1, 80, 2, 64), (1, 256, 1, 64)])
@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2e-2), (torch.float32, 1e-4)])
def test_attention_fwd_bwd(ops, dtype, tol, shape, monkeypatch):
    """softmax(q k^T * d^-1/2) v per (sample, head) and its autograd (pytorch_vit_encoder.py:59-78); the bf16
    case runs the fused two-product tcgen05 kernels (csrc/tc_attn.cu: one or two 128-row tiles, 1..4 feature chunks;
    (1, 256, 1, 64) does not fit their three operand images and takes the single-product kernels)."""
    from pose_estimation_amitai_b200 import vit_ops
    b, s, h, d = shape
    if dtype == torch.float32 and shape != (2, 144, 3, 256):
        pytest.skip("fp32 mode runs the CUDA-core kernels: one shape is enough")
    g = torch.Generator().manual_seed(7)
    qkv = (torch.randn(b * s, 3 * h * d, generator=g) * 0.5)
    go = torch.randn(b * s, h * d, generator=g) * 0.5
    if dtype == torch.bfloat16:
        qkv, go = qkv.bfloat16().float(), go.bfloat16().float()
    scale = d ** -0.5
    ref_in = qkv.clone().requires_grad_(True)
    q, k, v = ref_in.view(b, s, 3, h, d).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)
    out = (att @ v).permute(0, 2, 1, 3).reshape(b * s, h * d)
    out.backward(go)
    o, probs = vit_ops.attention_fwd(qkv.to(cuda, dtype), b, s, h, d, scale)
    torch.cuda.synchronize()
    sc = out.detach().abs().max().item()
    np.testing.assert_allclose(o.float().cpu().numpy(), out.detach().numpy(), rtol=tol, atol=tol * sc)
    gq = vit_ops.attention_bwd(qkv.to(cuda, dtype), probs, go.to(cuda, dtype), b, s, h, d, scale)
    torch.cuda.synchronize()
    sg = ref_in.grad.abs().max().item()
    np.testing.assert_allclose(gq.float().cpu().numpy(), ref_in.grad.numpy(), rtol=tol, atol=tol * sg)
    if dtype == torch.bfloat16:
        # fused and single-product kernels compute the same bf16 products: results agree to rounding of the bf16 P / dS
        monkeypatch.setenv("POSEB200_ATTN_UNFUSED", "1")
        o2, probs2 = vit_ops.attention_fwd(qkv.to(cuda, dtype), b, s, h, d, scale)
        gq2 = vit_ops.attention_bwd(qkv.to(cuda, dtype), probs2, go.to(cuda, dtype), b, s, h, d, scale)
        np.testing.assert_allclose(o2.float().cpu().numpy(), o.float().cpu().numpy(), rtol=1e-2, atol=1e-2 * sc)
        np.testing.assert_allclose(gq2.float().cpu().numpy(), gq.float().cpu().numpy(), rtol=1e-2, atol=1e-2 * sg)


@pytest.mark.parametrize("cin,cout,rows", [(256, 1024, 1152), (256, 768, 640), (1024, 256, 1152)])
def test_tc_wgrad_linear(ops, cin, cout, rows):
    """weight gradient of nn.Linear on the tensor cores, incl. outputs wider than one 256-column accumulator
    (to_qkv 256->9216, MLP 256->1024; pytorch_vit_encoder.py:20-23,52)."""
    g = torch.Generator().manual_seed(3)
    spec = ops.Contraction("linear", cin, cout)
    x = torch.randint(-8, 9, (rows, cin), generator=g).float() / 8
    dc = torch.randint(-8, 9, (rows, cout), generator=g).float() / 8
    want_w, want_b = dc.t() @ x, dc.sum(0)
    dw = torch.full((cout, cin), float("nan"), device=cuda)
    db = torch.full((cout,), float("nan"), device=cuda)
    ops.wgrad("tc", spec, x.to(cuda, torch.bfloat16).view(1, 1, rows, cin), dc.to(cuda, torch.bfloat16).view(1, 1, rows, cout),
              1, 1, rows, dw, db, act_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    np.testing.assert_allclose(dw.cpu().numpy(), want_w.numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(db.cpu().numpy(), want_b.numpy(), rtol=1e-5, atol=1e-3)


def test_find_peaks_reference_surfaces(golden_dir):
    """utils.py wrappers with the reference's three call surfaces: Augmentor.tf_find_peaks -> (N,C,2),
    Preprocessor.tf_find_peaks -> (N,3,C) with the peak value, utils.find_peaks_soft_argmax -> (N,C,2)."""
    from pose_estimation_amitai_b200 import utils
    fx = _kat(golden_dir)
    hm = fx["argmax_in"].astype(np.float32)                      # (N,H,W,C)
    got = utils.tf_find_peaks(hm).cpu().numpy()
    np.testing.assert_array_equal(got, fx["argmax_out"])
    got3 = utils.tf_find_peaks_with_values(hm)
    assert got3.shape == (hm.shape[0], 3, hm.shape[3])
    # pytorch/preprocessor.py:648-666 restated
    t = torch.from_numpy(hm)
    flat = t.reshape(t.shape[0], t.shape[1] * t.shape[2], t.shape[3])
    vals, idx = torch.max(flat, dim=1)
    want3 = torch.stack([(idx % t.shape[2]).float(), (idx // t.shape[2]).float(), vals], dim=1).numpy()
    np.testing.assert_array_equal(got3, want3)
    soft = utils.find_peaks_soft_argmax(fx["soft_in"].astype(np.float32))
    np.testing.assert_allclose(soft, fx["soft_out"], rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("c,cpad,hw", [(36, 48, 192), (5, 8, 64), (18, 32, 96)])
@pytest.mark.parametrize("with_target", [False, True])
def test_minmax_mse_fused_tail_equals_the_four_op_chain(ops, c, cpad, hw, with_target):
    """pb_minmax_mse_fwd_bwd (normalize_between_0_and_1 + MSELoss + both backwards + LeakyReLU', VITs.py:44-58 and
    train_pytorch.py:134-137) against the separate entry points it replaces and against torch autograd on the CPU."""
    from pose_estimation_amitai_b200 import vit_ops
    g = torch.Generator().manual_seed(c)
    b = 2
    pre = torch.randn(b, c, hw, hw, generator=g)
    x = F.leaky_relu(pre, 0.1)                       # what deconv4 hands to the normalisation
    pts = torch.randint(8, hw - 8, (b, c, 2), generator=g).float()
    tgt = torch.from_numpy(po.gaussian_targets(pts.numpy(), size=hw))
    xg, pg, tg = x.to(cuda), pts.to(cuda), (tgt.to(cuda) if with_target else None)
    loss_f, dc_f = ops.minmax_mse_fwd_bwd(xg, tg, points=None if with_target else pg, cpad=cpad)
    y, scratch = vit_ops.minmax_normalize_fwd(xg)
    loss_c, g_nchw, _ = ops.mse_loss_fwd_bwd(y, tg, points=None if with_target else pg, want_grad_nchw=True)
    g_pre = vit_ops.minmax_normalize_bwd(xg, g_nchw, scratch)
    dc_c = ops.grad_ingest(g_pre, xg, torch.bfloat16, cpad=cpad)
    torch.cuda.synchronize()
    assert abs(loss_f.item() - loss_c.item()) <= 1e-5 * abs(loss_c.item())
    a, r = dc_f.float().cpu(), dc_c.float().cpu()
    assert (a[..., c:] == 0).all()
    np.testing.assert_allclose(a.numpy(), r.numpy(), rtol=2e-2, atol=1e-3 * r.abs().max().item())
    # torch autograd on the CPU (oracle arithmetic): d loss / d pre through normalise and LeakyReLU
    pr = pre.clone().requires_grad_(True)
    xr = F.leaky_relu(pr, 0.1)
    yr = (xr - xr.min()) / (xr.max() - xr.min())
    loss = torch.nn.MSELoss()(yr, tgt)
    loss.backward()
    assert abs(loss_f.item() / x.numel() - loss.item()) <= 1e-4 * loss.item()
    want = pr.grad.permute(0, 2, 3, 1)
    got = a[..., :c]
    cos = (got.double() * want.double()).sum() / (got.double().norm() * want.double().norm())
    assert cos.item() >= 0.9999


@pytest.mark.parametrize("rows,dim", [(1152, 1024), (100, 256), (7, 2048)])
def test_gelu_bwd_with_fused_bias_column_sums(ops, rows, dim):
    """gx = gy * gelu'(pre) (erf GELU, pytorch_vit_encoder.py:21) and, from the same pass, the column sums of gx =
    the bias gradient of the nn.Linear in front of it."""
    from pose_estimation_amitai_b200 import vit_ops
    g = torch.Generator().manual_seed(rows)
    pre = torch.randn(rows, dim, generator=g).bfloat16()
    gy = torch.randn(rows, dim, generator=g).bfloat16()
    pr = pre.float().requires_grad_(True)
    F.gelu(pr).backward(gy.float())
    gx, (part, nblk) = vit_ops.gelu_bwd(pre.to(cuda), gy.to(cuda), want_colsum=True)
    np.testing.assert_allclose(gx.float().cpu().numpy(), pr.grad.numpy(), rtol=1e-2, atol=1e-2)
    plain = vit_ops.gelu_bwd(pre.to(cuda), gy.to(cuda))
    assert torch.equal(plain, gx)
    db = torch.full((dim,), 2.0, device=cuda)
    vit_ops.colsum(part, db, nblk, dim, beta=1.0)
    want = gx.float().sum(dim=0) + 2.0
    np.testing.assert_allclose(db.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("rows,with_add", [(9216, True), (1000, False), (37, True)])
def test_layernorm_bwd_with_fused_output_column_sums(ops, rows, with_add):
    """LayerNorm backward at the ViT's width (pytorch_vit_encoder.py:17,45,125) against torch autograd, and from the
    same pass the column sums of its OUTPUT gx (+ the residual-branch gradient) = the bias gradient of the nn.Linear
    that closes the previous block; gx itself is bit-identical to the plain call."""
    from pose_estimation_amitai_b200 import vit_ops
    dim = 256
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, dim, generator=g).bfloat16()
    gy = torch.randn(rows, dim, generator=g).bfloat16()
    add = torch.randn(rows, dim, generator=g).bfloat16() if with_add else None
    gamma = torch.rand(dim, generator=g) + 0.5
    beta_p = torch.rand(dim, generator=g) - 0.5
    xr = x.float().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = beta_p.clone().requires_grad_(True)
    F.layer_norm(xr, (dim,), gr, br, 1e-5).backward(gy.float())
    want_gx = xr.grad + (add.float() if with_add else 0.0)
    _, mean, rstd = vit_ops.layernorm_fwd(x.to(cuda), gamma.to(cuda), beta_p.to(cuda), save=True)
    dg = torch.zeros(dim, device=cuda)
    db = torch.zeros(dim, device=cuda)
    args = (x.to(cuda), gy.to(cuda), gamma.to(cuda), mean, rstd)
    gx, (part, nblk) = vit_ops.layernorm_bwd(*args, dg, db, gx_add=add.to(cuda) if with_add else None,
                                             want_colsum=True)
    np.testing.assert_allclose(gx.float().cpu().numpy(), want_gx.numpy(), rtol=2e-2, atol=2e-2)
    np.testing.assert_allclose(dg.cpu().numpy(), gr.grad.numpy(), rtol=2e-2, atol=2e-2 * rows ** 0.5)
    np.testing.assert_allclose(db.cpu().numpy(), br.grad.numpy(), rtol=1e-3, atol=1e-3 * rows ** 0.5)
    dg2 = torch.zeros(dim, device=cuda)
    db2 = torch.zeros(dim, device=cuda)
    plain = vit_ops.layernorm_bwd(*args, dg2, db2, gx_add=add.to(cuda) if with_add else None)
    assert torch.equal(plain, gx) and torch.equal(dg, dg2) and torch.equal(db, db2)
    bias_grad = torch.full((dim,), 2.0, device=cuda)
    vit_ops.colsum(part, bias_grad, nblk, dim, beta=1.0)
    want = gx.double().sum(dim=0).float() + 2.0
    np.testing.assert_allclose(bias_grad.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=1e-3)
    # outside the bf16 / dim 256 kernel the wrapper reports "no partials" and the caller keeps its own pass
    x32 = torch.randn(8, 64, generator=g).to(cuda)
    _, m32, r32 = vit_ops.layernorm_fwd(x32, torch.ones(64, device=cuda), torch.zeros(64, device=cuda), save=True)
    _, none_part = vit_ops.layernorm_bwd(x32, x32, torch.ones(64, device=cuda), m32, r32, torch.zeros(64, device=cuda),
                                         torch.zeros(64, device=cuda), want_colsum=True)
    assert none_part is None


# ------------------------------------------------------------------------------------- fused network head
def _head_case(ops, n, cin, cout, ih, iw, seed, dtype=torch.bfloat16):
    from pose_estimation_amitai_b200 import tc_support
    g = torch.Generator().manual_seed(seed)
    spec = ops.Contraction("convT2", cin, cout)
    wt = ((torch.rand(cin, cout, 3, 3, generator=g) - 0.5) * (2.0 / (3 * cin ** 0.5))).to(cuda)
    bias = (torch.rand(cout, generator=g) - 0.5).to(cuda)
    x = (torch.rand(n, ih, iw, cin, generator=g) - 0.5).to(cuda, dtype)
    wf = ops.pack_weights(wt, spec, "oi", dtype, ipad=tc_support.pad_n(cout))
    return spec, wf, bias, x


@pytest.mark.parametrize("n,cin,cout,ih,iw", [(3, 128, 36, 32, 24), (2, 128, 18, 16, 8), (5, 64, 36, 48, 48),
                                              (1, 640, 18, 16, 16)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_head_argmax_fused_is_bit_exact(ops, n, cin, cout, ih, iw, dtype):
    """pb_convT_argmax_fused == pb_peaks_argmax on the heatmaps the same layer materialises (peaks and maxima),
    including crafted ties (constant maps -> index 0; a duplicated maximum -> the first) and NaNs (NaN is the
    maximum, the first NaN wins), odd batch / group counts (pair-mode padding groups), C = 36 and C = 18."""
    spec, wf, bias, x = _head_case(ops, n, cin, cout, ih, iw, seed=n + cout, dtype=dtype)
    x[0] = 0                                   # image 0: every map is the constant lrelu(bias) -> ties everywhere
    if n > 1:
        x[1, ih // 2, iw // 3, :] = float("nan")   # image 1: a NaN patch in every map
        x[1, ih - 1, iw - 1, :] = float("nan")
    args = (x, wf, spec.fwd_taps(), n, ih, iw, cin, cout)
    hm = ops.conv("tc", *args[:3], n, ih, iw, cin, 2 * ih, 2 * iw, cout, bias=bias, act=ops.PB_ACT_LRELU,
                  act_dtype=dtype, out_nchw=True)
    want_pk, want_v = ops.peaks_argmax(hm, want_values=True)
    got_pk, got_v = ops.head_argmax_fused(*args, bias=bias, want_values=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(got_pk.cpu().numpy(), want_pk.cpu().numpy())
    np.testing.assert_array_equal(got_v.cpu().numpy(), want_v.cpu().numpy())
    assert (got_pk[0] == 0).all()              # constant maps: lowest flat index
    if n > 1:
        assert torch.isnan(got_v[1]).all()
    # and against the oracle's torch.max restatement on the same heatmaps
    np.testing.assert_array_equal(got_pk.cpu().numpy(),
                                  po.find_peaks_argmax(hm.cpu().permute(0, 2, 3, 1).contiguous()))


@pytest.mark.parametrize("n,cin,cout,ih,iw", [(3, 128, 36, 32, 24), (2, 128, 18, 16, 8), (1, 640, 18, 16, 16)])
@pytest.mark.parametrize("with_target", [False, True])
def test_head_mse_fused_equals_head_then_mse(ops, n, cin, cout, ih, iw, with_target):
    """pb_convT_mse_fused == pb_conv_tc (NCHW fp32 heatmaps) followed by pb_mse_loss_fwd_bwd: same gradient bits,
    same loss up to the summation order; Gaussian targets rendered on the fly or a target tensor read."""
    spec, wf, bias, x = _head_case(ops, n, cin, cout, ih, iw, seed=7 * n + cout)
    g = torch.Generator().manual_seed(5)
    pts = torch.randint(4, 2 * min(ih, iw) - 4, (n, cout, 2), generator=g).float().to(cuda)
    tgt = ops.gaussian_heatmaps(pts, size=(2 * ih, 2 * iw)) if with_target else None
    hm = ops.conv("tc", x, wf, spec.fwd_taps(), n, ih, iw, cin, 2 * ih, 2 * iw, cout, bias=bias,
                  act=ops.PB_ACT_LRELU, act_dtype=torch.bfloat16, out_nchw=True)
    cpad = (cout + 15) // 16 * 16
    want_loss, _, want_g = ops.mse_loss_fwd_bwd(hm, tgt, points=None if with_target else pts,
                                                grad_nhwc_dtype=torch.bfloat16, cpad=cpad, accumulation_steps=3)
    got_loss, got_g = ops.head_mse_fused(x, wf, spec.fwd_taps(), n, ih, iw, cin, cout, bias=bias, target=tgt,
                                         points=None if with_target else pts, accumulation_steps=3)
    torch.cuda.synchronize()
    assert got_g.shape == want_g.shape == (n, 2 * ih, 2 * iw, cpad)
    assert torch.equal(got_g, want_g)
    assert (got_g[..., cout:] == 0).all()
    assert abs(got_loss.item() - want_loss.item()) <= 1e-5 * abs(want_loss.item())


@pytest.mark.parametrize("n,cin,cout,ih,iw", [(3, 128, 36, 32, 24), (2, 128, 18, 16, 8), (5, 64, 36, 48, 48),
                                              (2, 64, 5, 20, 12), (1, 64, 64, 17, 9), (7, 128, 36, 24, 24)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_head_folded_parities_vs_torch_and_generic_kernel(ops, n, cin, cout, ih, iw, dtype, monkeypatch):
    """csrc/tc_head.cu (the four output parities of the stride-2 head folded into the MMA's N; tap tiles grouped by
    input shift) against torch's ConvTranspose2d(k3, s2, p1, op1) + LeakyReLU on the same 16-bit operand values, and
    against the generic halo kernel (POSEB200_HEAD_V2=0) -- ragged image sizes (zero-filled halo rows / columns),
    every N tile (16 / 32 / 48 / 64), tile ranges that cross image boundaries."""
    spec, wf, bias, x = _head_case(ops, n, cin, cout, ih, iw, seed=3 * n + cout, dtype=dtype)
    conv = lambda: ops.conv("tc", x, wf, spec.fwd_taps(), n, ih, iw, cin, 2 * ih, 2 * iw, cout, bias=bias,
                            act=ops.PB_ACT_LRELU, act_dtype=dtype, out_nchw=True)
    got = conv()
    monkeypatch.setenv("POSEB200_HEAD_V2", "0")
    old = conv()
    monkeypatch.delenv("POSEB200_HEAD_V2")
    torch.cuda.synchronize()
    wt = wf[:, :cout, :].float().cpu().permute(2, 1, 0).reshape(cin, cout, 3, 3)      # [t][co][ci] -> (ci, co, r, s)
    want = F.leaky_relu(F.conv_transpose2d(x.float().cpu().permute(0, 3, 1, 2), wt, bias.cpu(), stride=2, padding=1,
                                           output_padding=1), 0.1)
    assert got.shape == want.shape == (n, cout, 2 * ih, 2 * iw)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(got.cpu().numpy(), old.cpu().numpy(), rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("n,cin,cout,h,w,dil", [(2, 4, 64, 192, 192, 2), (3, 4, 64, 37, 50, 2), (1, 3, 32, 16, 33, 1),
                                                (2, 1, 128, 9, 70, 3)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_first_layer_direct_vs_im2col_form_and_torch(ops, n, cin, cout, h, w, dil, dtype):
    """csrc/tc_conv1.cu (conv1 + bias + LeakyReLU + sign mask straight from the NCHW fp32 crop, the operand rows built
    in shared memory) against the two-kernel form it replaces (pb_im2col_first + the 1-tap tcgen05 contraction: same
    operand values, same MMA sequence -> same bits) and against torch's Conv2d on the 16-bit-rounded operands; ragged
    image sizes (partial tiles, zero padding at every border)."""
    from pose_estimation_amitai_b200 import tc_support
    g = torch.Generator().manual_seed(n + cout)
    x = (torch.rand(n, cin, h, w, generator=g) - 0.5).to(cuda)
    wt = ((torch.rand(cout, cin, 3, 3, generator=g) - 0.5) * (2.0 / (3 * cin ** 0.5))).to(cuda)
    bias = (torch.rand(cout, generator=g) - 0.5).to(cuda)
    lin = ops.Contraction("linear", cin * 9, cout)
    wp = ops.pack_weights(wt, lin, "oi", dtype, ipad=tc_support.pad_n(cout), jpad=64)
    words = cout // 32
    m_direct = torch.zeros((n * h * w, words), device=cuda, dtype=torch.int32)
    m_two = torch.zeros_like(m_direct)
    twin = torch.full((n, h, w, cout), 9.0, device=cuda, dtype=torch.bfloat16) if dtype == torch.float16 else None
    got = ops.conv_first(x, wp, bias, cout, dil, dtype, mask_out=m_direct, out2=twin)
    cols = ops.im2col_first(x, 3, dil, 64, dtype)
    two = ops.conv("tc", cols, wp, lin.fwd_taps(), n, h, w, 64, h, w, cout, bias=bias, act=ops.PB_ACT_LRELU,
                   mask_out=m_two, act_dtype=dtype)
    torch.cuda.synchronize()
    assert got.shape == two.shape == (n, h, w, cout) and got.dtype == dtype
    assert torch.equal(got, two)
    assert torch.equal(m_direct, m_two)
    if twin is not None:      # the bf16 twin ("fp16" training) is the same fp32 value rounded to bf16 instead of fp16
        np.testing.assert_allclose(twin.float().cpu().numpy(), got.float().cpu().numpy(), rtol=2 ** -7, atol=1e-6)
    xr = x.to(dtype).float().cpu()
    wr = wt.to(dtype).float().cpu()
    want = F.leaky_relu(F.conv2d(xr, wr, bias.cpu(), padding=dil, dilation=dil), 0.1)
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-3      # output rounding of the 16-bit format
    np.testing.assert_allclose(got.float().cpu().permute(0, 3, 1, 2).numpy(), want.numpy(), rtol=tol, atol=tol)


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cin,cout,with_bias", [
    (1000, 256, 768, False),    # staged stores only (few M tiles -> 64-column tiles); ragged last M tile (104 of 128
                                # rows): the store clips the rows past the matrix
    (9216, 256, 3072, False),   # to_out's input gradient: resident 192-column weight tile (16 N tiles), ~8 tiles per
                                # CTA, most CTAs switch the weight tile once
    (5000, 128, 4096, True),    # resident 256-column tile, two K chunks, bias, ragged rows
    (38000, 256, 192, False),   # staged stores only, 297 M tiles on 148 CTAs; ragged rows
    (10800, 64, 192, True),     # staged stores only, one 192-column tile (three boxes) per M tile, one K chunk
    (2304, 3072, 256, True),    # K = 3072: stays on the register-path epilogue
    (128, 256, 64, True)])      # a single tile, a single box
def test_linear_tma_store_epilogue(ops, rows, cin, cout, with_bias):
    """nn.Linear on a token matrix (pytorch_vit_encoder.py:20-23,52) through the per-tap kernel's staged TMA-store
    epilogue and its resident-weight-tile form (csrc/tc_conv.cu, TE_TMA / TE_BRES): against torch on the bf16-rounded
    operands, result inside guard bands."""
    g = torch.Generator().manual_seed(rows + cout)
    x = (torch.rand(rows, cin, generator=g) - 0.5).bfloat16().to(cuda)
    wt = ((torch.rand(cout, cin, generator=g) - 0.5) * (2.0 / cin ** 0.5)).to(cuda)
    bias = (torch.rand(cout, generator=g) - 0.5).to(cuda) if with_bias else None
    lin = ops.Contraction("linear", cin, cout)
    wp = ops.pack_weights(wt, lin, "oi", torch.bfloat16)
    buf, out, pad = _guarded((1, 1, rows, cout), torch.bfloat16, 7.0)
    ops.conv("tc", x, wp, lin.fwd_taps(), 1, 1, rows, cin, 1, rows, cout, bias=bias, out=out, act_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    assert _guards_intact(buf, pad, 7.0)
    want = x.float().cpu() @ wt.bfloat16().float().cpu().t()
    if with_bias:
        want = want + bias.cpu()
    np.testing.assert_allclose(out.view(rows, cout).float().cpu().numpy(), want.numpy(), rtol=2 ** -7, atol=2e-3)


def _guarded(shape, dtype, fill):
    """a contiguous tensor of `shape` carved out of a larger allocation with 4 KB guard bands on both sides."""
    n = int(np.prod(shape))
    pad = 4096 // torch.empty((), dtype=dtype).element_size()
    buf = torch.full((n + 2 * pad,), fill, device=cuda, dtype=dtype)
    return buf, buf[pad:pad + n].view(shape), pad


def _guards_intact(buf, pad, fill):
    return bool((buf[:pad] == fill).all() and (buf[-pad:] == fill).all())


@pytest.mark.parametrize("n,cin,cout,ih,iw", [(3, 128, 36, 17, 9), (2, 64, 18, 33, 31), (1, 128, 5, 16, 8)])
def test_new_kernels_stay_inside_their_buffers(ops, n, cin, cout, ih, iw):
    """compute-sanitizer is not available on the GPU pool (profiles/r2_sanitize_summary.txt), so the round-2 kernels
    are run on ragged shapes -- partial tiles on every edge -- with guard bands around every output and the bands are
    checked afterwards: tc_head.cu (NCHW store, fused MSE gradient; the arg-max form only writes N*C keys) and
    tc_conv1.cu (NHWC store + sign mask)."""
    spec, wf, bias, x = _head_case(ops, n, cin, cout, ih, iw, seed=11)
    obuf, out, opad = _guarded((n, cout, 2 * ih, 2 * iw), torch.float32, 7.0)
    ops.conv("tc", x, wf, spec.fwd_taps(), n, ih, iw, cin, 2 * ih, 2 * iw, cout, bias=bias, act=ops.PB_ACT_LRELU,
             act_dtype=torch.bfloat16, out_nchw=True, out=out)
    cpad = (cout + 15) // 16 * 16
    gbuf, grad, gpad = _guarded((n, 2 * ih, 2 * iw, cpad), torch.bfloat16, 7.0)
    pts = torch.randint(2, 2 * min(ih, iw) - 2, (n, cout, 2), generator=torch.Generator().manual_seed(3)).float().to(cuda)
    ops.head_mse_fused(x, wf, spec.fwd_taps(), n, ih, iw, cin, cout, bias=bias, points=pts, grad_out=grad)
    pk = ops.head_argmax_fused(x, wf, spec.fwd_taps(), n, ih, iw, cin, cout, bias=bias)
    torch.cuda.synchronize()
    assert _guards_intact(obuf, opad, 7.0) and _guards_intact(gbuf, gpad, 7.0)
    assert torch.isfinite(out).all() and (out != 7.0).any() and torch.isfinite(grad.float()).all()
    assert torch.equal(pk, ops.peaks_argmax(out))
    # first layer: H, W not multiples of the 4 x 32 tile
    from pose_estimation_amitai_b200 import tc_support
    h, w = 2 * ih + 1, 2 * iw + 3
    xin = torch.rand(n, 4, h, w, device=cuda)
    lin = ops.Contraction("linear", 36, 64)
    wt = (torch.rand(64, 4, 3, 3, device=cuda) - 0.5) * 0.3
    wp = ops.pack_weights(wt, lin, "oi", torch.bfloat16, ipad=tc_support.pad_n(64), jpad=64)
    ybuf, y, ypad = _guarded((n, h, w, 64), torch.bfloat16, 7.0)
    mbuf, mask, mpad = _guarded((n * h * w, 2), torch.int32, 7)
    ops.conv_first(xin, wp, None, 64, 2, torch.bfloat16, mask_out=mask, out=y)
    torch.cuda.synchronize()
    assert _guards_intact(ybuf, ypad, 7.0) and _guards_intact(mbuf, mpad, 7)
    want = F.leaky_relu(F.conv2d(xin.bfloat16().float().cpu(), wt.bfloat16().float().cpu(), None, padding=2, dilation=2), 0.1)
    np.testing.assert_allclose(y.float().cpu().permute(0, 3, 1, 2).numpy(), want.numpy(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("kind,cin,cout,h,w,dil,n", [("conv", 64, 64, 32, 48, 2, 2), ("conv", 128, 128, 16, 16, 2, 3),
                                                     ("conv", 64, 64, 24, 20, 2, 1), ("conv", 64, 128, 48, 24, 2, 2),
                                                     ("conv", 256, 256, 18, 10, 2, 1)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_fused_maxpool_epilogue_equals_conv_then_pool(ops, kind, cin, cout, h, w, dil, n, dtype):
    """pb_conv_args.pool_out (csrc/tc_conv2.cu): lrelu(maxpool2x2(out)) emitted by the conv's own epilogue (CNNs.py:77,82)
    is bit-identical to the conv followed by pb_maxpool_lrelu_fwd, with and without the full-resolution store
    (pool_only), on ragged image sizes (partial tiles; cta pairs with a padding group) and with 1-4 64-channel blocks."""
    from pose_estimation_amitai_b200 import tc_support
    g = torch.Generator().manual_seed(cin + h)
    spec = ops.Contraction(kind, cin, cout, dilation=dil)
    assert tc_support.pool_fusable(spec, h, w)
    wt = ((torch.rand(cout, cin, 3, 3, generator=g) - 0.5) * (2.0 / (3 * cin ** 0.5))).to(cuda)
    bias = (torch.rand(cout, generator=g) - 0.5).to(cuda)
    x = (torch.rand(n, h, w, cin, generator=g) - 0.5).to(cuda, dtype)
    res = (torch.rand(n, h, w, cout, generator=g) - 0.5).to(cuda, dtype)
    wp = ops.pack_weights(wt, spec, "oi", dtype, ipad=tc_support.pad_n(cout))
    kw = dict(bias=bias, act=ops.PB_ACT_LRELU, add1=res, act_dtype=dtype)
    mask0 = torch.zeros((n * h * w, (cout + 31) // 32), device=cuda, dtype=torch.int32)
    mask1 = torch.zeros_like(mask0)
    y_ref = ops.conv("tc", x, wp, spec.fwd_taps(), n, h, w, cin, h, w, cout, mask_out=mask0, **kw)
    p_ref = ops.maxpool_lrelu_fwd(y_ref)
    pooled = torch.full((n, h // 2, w // 2, cout), 9.0, device=cuda, dtype=dtype)
    y = ops.conv("tc", x, wp, spec.fwd_taps(), n, h, w, cin, h, w, cout, mask_out=mask1, pool_out=pooled, **kw)
    pooled_only = torch.full_like(pooled, 9.0)
    ops.conv("tc", x, wp, spec.fwd_taps(), n, h, w, cin, h, w, cout, pool_out=pooled_only, pool_only=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(y, y_ref) and torch.equal(mask1, mask0)
    assert torch.equal(pooled, p_ref)
    assert torch.equal(pooled_only, p_ref)


@pytest.mark.parametrize("n,cin,cout,ih,iw", [(3, 128, 36, 32, 24), (2, 64, 18, 16, 8)])
def test_head_mse_fused_bias_gradient(ops, n, cin, cout, ih, iw):
    """pb_head_fused_args.dbias: the fused head's epilogue also sums its gradient over all pixels (= the layer's bias
    gradient, otherwise a separate pass over the 226 MB gradient tensor).  Equals the column sums of the bf16 gradient
    it stores up to that tensor's rounding; one row per CTA, summed in a fixed order (bit-reproducible); accumulates."""
    spec, wf, bias, x = _head_case(ops, n, cin, cout, ih, iw, seed=5 * n + cout)
    pts = torch.randint(4, 2 * min(ih, iw) - 4, (n, cout, 2), generator=torch.Generator().manual_seed(6)).float().to(cuda)
    db = ops.head_dbias_buffer(cout, cuda)
    _, grad = ops.head_mse_fused(x, wf, spec.fwd_taps(), n, ih, iw, cin, cout, bias=bias, points=pts, dbias_out=db)
    torch.cuda.synchronize()
    want = grad.float().sum(dim=(0, 1, 2))[:cout].cpu().numpy()
    scale = np.abs(want).max()
    first = db.sum(dim=0).cpu().numpy()
    np.testing.assert_allclose(first, want, rtol=5e-3, atol=5e-3 * scale)
    db2 = ops.head_dbias_buffer(cout, cuda)
    ops.head_mse_fused(x, wf, spec.fwd_taps(), n, ih, iw, cin, cout, bias=bias, points=pts, dbias_out=db2)
    ops.head_mse_fused(x, wf, spec.fwd_taps(), n, ih, iw, cin, cout, bias=bias, points=pts, dbias_out=db)
    torch.cuda.synchronize()
    assert torch.equal(db2.sum(dim=0), torch.from_numpy(first).to(cuda))                      # fixed summation order: same bits every run
    np.testing.assert_allclose(db.sum(dim=0).cpu().numpy(), 2 * want, rtol=5e-3, atol=1e-2 * scale)


@pytest.mark.parametrize("n,cin,h,w,dil", [(2, 4, 192, 192, 2), (3, 4, 37, 50, 2), (1, 3, 16, 33, 1), (5, 1, 9, 70, 3)])
def test_first_layer_wgrad_direct_vs_im2col_form_and_torch(ops, n, cin, h, w, dil):
    """csrc/tc_wgrad1.cu (conv1's weight + bias gradient straight from the NCHW fp32 crop: the im2col rows are built in
    shared memory, the bias gradient comes from a ones column of the operand) against the form it replaces
    (pb_im2col_first + the 1-tap tcgen05 weight gradient) and against torch's Conv2d autograd on bf16-rounded operands;
    ragged sizes, fewer tiles than SMs, accumulate mode."""
    cout = 64
    g = torch.Generator().manual_seed(n + h)
    x = (torch.randint(-8, 9, (n, cin, h, w), generator=g).float() / 8).to(cuda)          # exact in bf16
    dc = (torch.randint(-8, 9, (n, h, w, cout), generator=g).float() / 8).to(cuda, torch.bfloat16)
    dw = torch.full((cout, cin, 3, 3), float("nan"), device=cuda)
    db = torch.full((cout,), float("nan"), device=cuda)
    ops.wgrad_first(x, dc, dw, db, dil)
    lin = ops.Contraction("linear", cin * 9, cout)
    cols = ops.im2col_first(x, 3, dil, 64, torch.bfloat16)
    dw2 = torch.full((cout, cin * 9), float("nan"), device=cuda)
    db2 = torch.full((cout,), float("nan"), device=cuda)
    ops.wgrad("tc", lin, cols, dc, n, h, w, dw2, db2, act_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    wt = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    bias = torch.zeros(cout, requires_grad=True)
    F.conv2d(x.cpu(), wt, bias, padding=dil, dilation=dil).backward(dc.float().cpu().permute(0, 3, 1, 2))
    np.testing.assert_allclose(dw.cpu().numpy(), wt.grad.numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(db.cpu().numpy(), bias.grad.numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(dw.reshape(cout, -1).cpu().numpy(), dw2.cpu().numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(db.cpu().numpy(), db2.cpu().numpy(), rtol=1e-5, atol=1e-3)
    ops.wgrad_first(x, dc, dw, db, dil, beta=1.0)           # accumulation_steps > 1
    np.testing.assert_allclose(dw.cpu().numpy(), 2 * wt.grad.numpy(), rtol=1e-5, atol=2e-3)
    np.testing.assert_allclose(db.cpu().numpy(), 2 * bias.grad.numpy(), rtol=1e-5, atol=2e-3)
