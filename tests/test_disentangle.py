"""FourCamerasDisentanglement / FTL / InvFTL / train-mode BatchNorm (pytorch/CNNs.py:240-352):
structure on the CPU; on the GPU the kernels against torch / the oracle and the model against vectors produced by the
real reference module (tests/golden/multicam_next.npz) and against autograd of the oracle restatement."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pose_oracle as po

CFG = {"model type": "ALL_CAMS_DISENTANGLED_PER_WING_CNN", "number of base filters": 64, "convolution kernel size": 3,
       "dilation rate": 2, "dropout ratio": 0.5}


def _fx(golden_dir):
    return np.load(os.path.join(golden_dir, "multicam_next.npz"), allow_pickle=False)


def _cos(a, b):
    a, b = a.flatten().double().cpu(), b.flatten().double().cpu()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-300)).item()


def _build(precision, seed=5, joints=72):
    from pose_estimation_amitai_b200 import CNNs, Network
    torch.manual_seed(seed)
    model = Network.Network(dict(CFG, precision=precision), (192, 192, 16), joints).model
    assert isinstance(model, CNNs.FourCamerasDisentanglement)
    return model


def test_disentanglement_parameters_and_seeded_init(golden_dir):
    fx = _fx(golden_dir)
    model = _build("bf16")
    sd = model.state_dict()
    keys = [str(k) for k in fx["dis_param_keys"]]
    assert [k for k, v in sd.items() if v.is_floating_point()] == keys
    for k, s in zip(keys, fx["dis_param_sum"]):
        assert np.isclose(sd[k].double().sum().item(), s, rtol=0, atol=1e-9 + 1e-12 * abs(s)), k
    assert model.fusion_layer_1.weight.shape == (400, 1600, 1, 1) and model.batch_norm3.num_features == 300
    assert len(model._live_params()) == 40
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 16, 192, 192), torch.zeros(1, 4, 3, 4), torch.zeros(1, 4, 4, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ftl_kernels_match_the_raw_reinterpretation(dtype):
    """pb_ftl == the reference's reshape / matmul / reshape (oracle.ftl / inv_ftl), incl. batch strides wider than the
    tensor (channel padding), the input shared by several views (in_batch_mod) and accumulation; its transpose is the
    backward."""
    from pose_estimation_amitai_b200 import CNNs, ops
    g = torch.Generator().manual_seed(0)
    x4 = torch.rand(3, 400, 48, 48, generator=g) - 0.5
    x3 = torch.rand(3, 300, 48, 48, generator=g) - 0.5
    P = torch.randn(3, 3, 4, generator=g)
    Pinv = torch.randn(3, 4, 3, generator=g)
    if dtype == torch.float32:
        got = CNNs.FTL()(x4.cuda(), P.cuda()).cpu()
        np.testing.assert_allclose(got.numpy(), po.ftl(x4, P).numpy(), rtol=1e-5, atol=1e-6)
        got = CNNs.InvFTL()(x3.cuda(), Pinv.cuda()).cpu()
        np.testing.assert_allclose(got.numpy(), po.inv_ftl(x3, Pinv).numpy(), rtol=1e-5, atol=1e-6)
    tol = dict(rtol=1e-5, atol=1e-6) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    npix = 48 * 48
    # padded batch strides, one input serving two "views" of a 1-sample batch, accumulation
    xin = torch.zeros(1, 448, npix)
    xin[:, :400] = (x4[:1].reshape(1, 400, npix)).to(dtype).float()
    mats = torch.stack([P[0], P[1]])
    out = torch.zeros(2, 320, npix).to("cuda", dtype)
    ops.ftl(xin.to("cuda", dtype), mats.cuda(), out, kin=4, kout=3, groups=400 * npix // 4, in_batch_stride=448 * npix,
            out_batch_stride=320 * npix, in_batch_mod=1)
    for v in range(2):
        want = po.ftl(xin[:, :400].reshape(1, 400, 48, 48), mats[v:v + 1]).reshape(300, npix)
        np.testing.assert_allclose(out[v, :300].float().cpu().numpy(), want.numpy(), **tol)
    assert (out[:, 300:] == 0).all()
    # backward = the transposed matrices; check <FTL(x), g> == <x, FTL^T(g)>
    gy = (torch.rand(1, 300, npix, generator=g) - 0.5).to(dtype).float()
    gx = torch.zeros(1, 400, npix).to("cuda", dtype)
    ops.ftl(gy.to("cuda", dtype), P[:1].transpose(1, 2).contiguous().cuda(), gx, kin=3, kout=4, groups=300 * npix // 3,
            in_batch_stride=300 * npix, out_batch_stride=400 * npix)
    lhs = (po.ftl(xin[:, :400].reshape(1, 400, 48, 48), P[:1]).reshape(-1).double() * gy.reshape(-1).double()).sum()
    rhs = (xin[:, :400].reshape(-1).double() * gx.float().cpu().reshape(-1).double()).sum()
    assert abs(lhs - rhs) <= (1e-4 if dtype == torch.float32 else 2e-2) * abs(lhs)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,groups,c,cs", [(torch.float32, 1, 400, 400), (torch.bfloat16, 1, 400, 448),
                                               (torch.bfloat16, 4, 300, 320), (torch.float32, 4, 300, 300)])
def test_batchnorm_relu_kernels_vs_torch(dtype, groups, c, cs):
    """training-mode BatchNorm + ReLU forward / backward per group, running statistics updated group after group
    (what four calls of one nn.BatchNorm2d do), eval mode, zero channel padding kept zero."""
    from pose_estimation_amitai_b200 import ops
    g = torch.Generator().manual_seed(1)
    rpg = 2 * 48 * 48
    x = torch.zeros(groups * rpg, cs)
    x[:, :c] = torch.randn(groups * rpg, c, generator=g) * 1.5 + 0.3
    x = x.to(dtype).float()
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.rand(c, generator=g) - 0.5
    gy = torch.zeros(groups * rpg, cs)
    gy[:, :c] = torch.randn(groups * rpg, c, generator=g)
    gy = gy.to(dtype).float()
    bn = torch.nn.BatchNorm2d(c)
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    bn.train()
    want_y, want_gx = [], []
    for gi in range(groups):
        xi = x[gi * rpg:(gi + 1) * rpg, :c].t().reshape(1, c, rpg, 1).clone().requires_grad_(True)
        yi = F.relu(bn(xi))
        yi.backward(gy[gi * rpg:(gi + 1) * rpg, :c].t().reshape(1, c, rpg, 1))
        want_y.append(yi.detach().reshape(c, rpg).t())
        want_gx.append(xi.grad.reshape(c, rpg).t())
    want_y, want_gx = torch.cat(want_y), torch.cat(want_gx)
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    y, mean, rstd = ops.batchnorm_fwd(x.to("cuda", dtype), gamma.cuda(), beta.cuda(), rm, rv, groups=groups, channels=c,
                                      training=True)
    tol = dict(rtol=1e-4, atol=1e-4) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    np.testing.assert_allclose(y[:, :c].float().cpu().numpy(), want_y.numpy(), **tol)
    assert (y[:, c:] == 0).all()
    np.testing.assert_allclose(rm.cpu().numpy(), bn.running_mean.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rv.cpu().numpy(), bn.running_var.numpy(), rtol=1e-4, atol=1e-5)
    dgm, dbt = torch.ones(c, device="cuda"), torch.ones(c, device="cuda")
    gx = ops.batchnorm_bwd(x.to("cuda", dtype), y, gy.to("cuda", dtype), gamma.cuda(), mean, rstd, dgm, dbt,
                           groups=groups, channels=c, beta_acc=1.0)
    scale = want_gx.abs().max().item()
    np.testing.assert_allclose(gx[:, :c].float().cpu().numpy(), want_gx.numpy(), rtol=tol["rtol"], atol=tol["atol"] * scale)
    np.testing.assert_allclose(dgm.cpu().numpy() - 1.0, bn.weight.grad.numpy(), rtol=5 * tol["rtol"],
                               atol=5 * tol["atol"] * bn.weight.grad.abs().max().item())
    np.testing.assert_allclose(dbt.cpu().numpy() - 1.0, bn.bias.grad.numpy(), rtol=5 * tol["rtol"],
                               atol=5 * tol["atol"] * bn.bias.grad.abs().max().item())
    # eval mode: running statistics
    bn.eval()
    with torch.no_grad():
        want_e = F.relu(bn(x[:rpg, :c].t().reshape(1, c, rpg, 1))).reshape(c, rpg).t()
    ye, _, _ = ops.batchnorm_fwd(x[:rpg].to("cuda", dtype), gamma.cuda(), beta.cuda(), rm, rv, groups=1, channels=c,
                                 training=False)
    np.testing.assert_allclose(ye[:, :c].float().cpu().numpy(), want_e.numpy(), **tol)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,gate", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_disentanglement_vs_reference_golden(golden_dir, precision, gate):
    """train-mode forward (batch statistics, running statistics updated) against the REAL reference module's output."""
    fx = _fx(golden_dir)
    model = _build(precision).cuda().train()
    x = torch.rand(2, 16, 192, 192, generator=torch.Generator().manual_seed(11)).cuda()
    cams, cams_inv = torch.from_numpy(fx["cams"]).cuda(), torch.from_numpy(fx["cams_inv"]).cuda()
    with torch.no_grad():
        out = model(x, cams, cams_inv)
    assert out.shape == (2, 72, 192, 192) and out.dtype == torch.float32
    ref = torch.from_numpy(fx["dis_out_sub"])
    m = po.heatmap_parity(out.cpu()[:, ::24, ::3, ::3], ref)
    print(f"[parity FourCamerasDisentanglement golden b2 {precision}] " + "  ".join(f"{k} {v:.3e}" for k, v in m.items()))
    if precision == "fp32":
        assert m["floor10"] <= gate, m
        stats = np.array([out.mean().item(), out.std().item(), out.min().item(), out.max().item()])
        np.testing.assert_allclose(stats, fx["dis_out_stats"], rtol=1e-3, atol=1e-5)
    else:       # bf16 operands: heatmap scale (see tests/test_gpu_network.py for the element-wise floor of the format)
        assert m["worst"] <= gate and m["rms"] <= gate, m
    assert int(model.batch_norm3.num_batches_tracked) == 4 and int(model.batch_norm1.num_batches_tracked) == 1
    assert not torch.equal(model.batch_norm1.running_mean, torch.zeros_like(model.batch_norm1.running_mean))


@pytest.mark.gpu
def test_disentanglement_gradients_vs_oracle_autograd_and_fused_step(golden_dir):
    torch.set_num_threads(os.cpu_count() or 1)
    fx = _fx(golden_dir)
    joints, b = 72, 2
    x = torch.rand(b, 16, 192, 192, generator=torch.Generator().manual_seed(3))
    cams, cams_inv = torch.from_numpy(fx["cams"]), torch.from_numpy(fx["cams_inv"])
    pts = po.synthetic_points(b, joints, seed=4)
    tgt = torch.from_numpy(po.gaussian_targets(pts))
    model = _build("fp32", seed=2).cuda().train()
    live = dict(model._live_params())
    ref_params = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    for k in live:
        ref_params[k].requires_grad_(True)
    ref_out = po.four_cameras_disentanglement_forward(ref_params, x, cams, cams_inv, training=True)
    ref_loss = po.mse_loss(ref_out, tgt)
    ref_loss.backward()
    out = model(x.cuda(), cams.cuda(), cams_inv.cuda())
    loss = torch.nn.MSELoss()(out, tgt.cuda())
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item())
    assert po.heatmap_parity(out.detach().cpu(), ref_out.detach())["floor10"] <= 1e-4
    for k, p in live.items():
        assert p.grad is not None, k
        if k in ("fusion_layer_1.bias", "fusion_layer_2.bias"):
            # a bias in front of a training-mode BatchNorm has NO gradient (the batch mean removes it): both sides
            # hold rounding noise only
            wn = live[k.replace(".bias", ".weight")].grad.double().norm().item()
            assert p.grad.double().norm().item() <= 1e-4 * wn and ref_params[k].grad.double().norm().item() <= 1e-4 * wn, k
            continue
        assert _cos(p.grad, ref_params[k].grad) >= 0.9999, k
        n_ref = ref_params[k].grad.double().norm().item()
        assert abs(p.grad.double().norm().item() - n_ref) <= 5e-3 * n_ref + 1e-12, k
    for k, p in model.named_parameters():
        if k not in live:
            assert p.grad is None, k
    # fused step (bf16, loss inside the head's epilogue) vs the module's autograd path
    model = _build("bf16", seed=2).cuda().train()
    out = model(x.cuda(), cams.cuda(), cams_inv.cuda())
    loss = torch.nn.MSELoss()(out, tgt.cuda())
    loss.backward()
    named = dict(model._live_params())
    want = {k: p.grad.clone() for k, p in named.items()}
    for p in model.parameters():
        p.grad = None
    loss2 = model.train_step(x.cuda(), tgt.cuda(), camera_matrices=cams.cuda(), camera_matrices_inv=cams_inv.cuda())
    assert abs(loss2.item() - loss.item()) <= 1e-4 * abs(loss.item())
    for k, g in want.items():
        if k in ("fusion_layer_1.bias", "fusion_layer_2.bias"):
            continue
        assert _cos(named[k].grad, g) >= 0.999, k
    # data-parallel step plumbing passes the camera matrices through
    from pose_estimation_amitai_b200 import parallel
    dp = parallel.DataParallelStep(model, lr=1e-3)
    l3 = dp.step(x.cuda(), points=torch.from_numpy(pts).cuda(), camera_matrices=cams.cuda(),
                 camera_matrices_inv=cams_inv.cuda())
    assert np.isfinite(l3.item())
    pk = model.predict_peaks(x.cuda(), cams.cuda(), cams_inv.cuda())
    assert pk.shape == (b, joints, 2)
