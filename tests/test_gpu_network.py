"""Whole-network GPU parity: the drop-in BasicNet against vectors produced by the real reference
modules (tests/golden/basicnet_c36.npz) and against the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import pose_oracle as po

pytestmark = pytest.mark.gpu
cuda = torch.device("cuda")

CFG = {"model type": "MODEL_18_POINTS_PER_WING", "number of base filters": 64, "convolution kernel size": 3,
       "dilation rate": 2, "dropout ratio": 0.5}


# ---- heatmap parity gates.  Metric: oracle.heatmap_parity; THE GATE is `floor10` =
#      max |err| / (|ref| + 0.1 max|ref|) against the fp32 reference -- element-wise relative with a 10 % floor, because
#      LeakyReLU heatmaps cross zero.
#        fp32 mode : floor10 <= 1e-4
#        fp16 mode : floor10 <= 2e-2      (north_star's 16-bit tolerance, met with a 3x margin)
#        bf16 mode : a bf16-OPERAND forward cannot meet 2e-2 on this metric whatever the kernels do -- rounding only
#                    the weights to bf16, everything else fp32, already gives 2.4e-2, all operands 6.2e-2
#                    (oracle.basicnet_forward_operand_rounded, tests/test_oracle_golden.py pins both numbers).  So the
#                    gate is (a) 2e-2 on the heatmap scale (`worst`, `rms`) and (b) the CUDA path sits AT the format's
#                    floor: its floor10 / rms are not larger than the operand-rounded oracle's (x1.5 / x1.15 slack for
#                    the different summation order).
GATE = {"fp32": 1e-4, "fp16": 2e-2, "bf16": 2e-2}


def _check_parity(precision: str, got: torch.Tensor, ref: torch.Tensor, floor_ref: torch.Tensor = None, what: str = ""):
    m = po.heatmap_parity(got, ref)
    line = f"[parity {what} {precision}] " + "  ".join(f"{k} {v:.3e}" for k, v in m.items())
    if precision == "bf16" and floor_ref is not None:
        f = po.heatmap_parity(floor_ref, ref)
        line += "  | bf16-operand floor: " + "  ".join(f"{k} {v:.3e}" for k, v in f.items())
        assert m["worst"] <= GATE["bf16"] and m["rms"] <= GATE["bf16"], line
        assert m["floor10"] <= 1.5 * f["floor10"] and m["rms"] <= 1.15 * f["rms"], line
    else:
        assert m["floor10"] <= GATE[precision], line
    print(line)
    return m


def _cos(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.flatten().double(), b.flatten().double()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-300)).item()


def _build(precision, joints=36):
    from pose_estimation_amitai_b200 import CNNs
    torch.manual_seed(0)
    return CNNs.BasicNet(dict(CFG, precision=precision), np.array((192, 192, 4)), joints).to(cuda)


@pytest.mark.parametrize("precision,tol_loss,min_cos", [("fp32", 1e-5, 0.99999), ("fp16", 1e-3, 0.999),
                                                        ("bf16", 1e-3, 0.999)])
def test_basicnet_vs_reference_golden(golden_dir, precision, tol_loss, min_cos):
    fx = np.load(os.path.join(golden_dir, "basicnet_c36.npz"))
    joints, batch = int(fx["joints"]), int(fx["batch"])
    model = _build(precision, joints)
    x = po.synthetic_crops(batch, seed=1).to(cuda)
    tgt = torch.from_numpy(po.gaussian_targets(fx["points"])).to(cuda)
    # autograd path exactly as train_pytorch.py:132-137 (without AMP)
    model.train()
    out = model(x)
    assert out.shape == (batch, joints, 192, 192) and out.dtype == torch.float32
    loss = torch.nn.MSELoss()(out, tgt)
    loss.backward()
    # reference heatmaps: the golden file holds every 6th map of the REAL reference module's output; the oracle
    # reproduces those (checked here) and supplies the other maps, so the gate runs over all 36
    sd_ref = po.basicnet_state_dict(joints, seed=0)
    ref = po.basicnet_forward(sd_ref, x.cpu())
    np.testing.assert_allclose(ref[:, ::6].numpy(), fx["out_sub"], rtol=1e-5, atol=1e-7)
    floor = po.basicnet_forward_operand_rounded(sd_ref, x.cpu(), "bf16") if precision == "bf16" else None
    _check_parity(precision, out.detach().cpu(), ref, floor, "BasicNet golden b2")
    assert abs(loss.item() - float(fx["loss"])) <= tol_loss * float(fx["loss"])
    named = dict(model.named_parameters())
    for k, n in zip([str(s) for s in fx["grad_keys"]], fx["grad_norm"]):
        g = named[k].grad
        assert g is not None, k
        assert abs(g.double().norm().item() - n) <= (1e-3 if precision == "fp32" else 3e-2) * n, k
        if "grad::" + k in fx.files:
            assert _cos(g.cpu(), torch.from_numpy(fx["grad::" + k])) >= min_cos, k
    for k in (str(s) for s in fx["grad_none_keys"]):
        assert named[k].grad is None, k  # inert BatchNorm parameters (CNNs.py:25-43)
    # fused train step == autograd path
    grads_autograd = {k: p.grad.clone() for k, p in named.items() if p.grad is not None}
    for p in model.parameters():
        p.grad = None
    loss2 = model.train_step(x, tgt)
    assert abs(loss2.item() - loss.item()) <= 1e-5 * abs(loss.item())
    for k, g in grads_autograd.items():
        assert _cos(named[k].grad, g) >= 0.99999, k
    # fused Gaussian target == materialised target
    loss3 = model.train_step(x, points=torch.from_numpy(fx["points"]).to(cuda), accumulate=True)
    assert abs(loss3.item() - loss.item()) <= 1e-4 * abs(loss.item())
    for k, g in grads_autograd.items():
        assert _cos(named[k].grad, g) >= 0.9999, k
        assert abs(named[k].grad.norm().item() - 2 * g.norm().item()) <= 1e-2 * g.norm().item(), k  # accumulated


def test_peaks_bit_exact_on_network_output(golden_dir):
    from pose_estimation_amitai_b200 import ops
    model = _build("bf16", 36).eval()
    x = po.synthetic_crops(4, seed=9).to(cuda)
    with torch.no_grad():
        out = model(x)
        got = model.predict_peaks(x).cpu().numpy()
    want = po.find_peaks_argmax(out.cpu().permute(0, 2, 3, 1).contiguous())
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(ops.peaks_argmax(out.to(torch.bfloat16)).cpu().numpy(),
                                  po.find_peaks_argmax(out.to(torch.bfloat16).float().cpu().permute(0, 2, 3, 1).contiguous()))


def test_state_dict_round_trip_with_oracle_weights():
    """weights drawn by the oracle's reference-order constructor load strict=False-free of misses."""
    model = _build("fp32", 18)
    sd = po.basicnet_state_dict(18, seed=3)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(".bn" in k for k in missing)
    x = po.synthetic_crops(1, seed=5)
    with torch.no_grad():
        got = model.to(cuda)(x.to(cuda)).cpu()
        want = po.basicnet_forward(sd, x)
    _check_parity("fp32", got, want, what="BasicNet C18 oracle weights")


VIT_CFG = dict(CFG, **{"model type": "MODEL_18_POINTS_PER_WING_VIT", "optimizer": "adam", "patch size": 16,
                       "projection dim": 256, "num heads": 12, "dim head": -1, "transformer layers": 8})


@pytest.mark.parametrize("precision,tol_loss,min_cos", [("fp32", 1e-4, 0.9999), ("bf16", 2e-2, 0.99)])
def test_vit_vs_reference_golden(golden_dir, precision, tol_loss, min_cos):
    from pose_estimation_amitai_b200 import VITs
    fx = np.load(os.path.join(golden_dir, "vit_c36.npz"))
    joints, batch = int(fx["joints"]), int(fx["batch"])
    torch.manual_seed(0)
    model = VITs.VIT_encoder_CNN_decoder(dict(VIT_CFG, precision=precision), np.array((192, 192, 4)), joints)
    sd = model.state_dict()
    assert len(sd) == int(fx["state_dict_len"]) == 104
    for k, s in zip([str(s) for s in fx["param_keys"]], fx["param_sum"]):
        assert abs(sd[k].double().sum().item() - s) <= 1e-9 + 1e-12 * abs(s), k   # same seeded init as the reference
    model = model.to(cuda).train()
    x = po.synthetic_crops(batch, seed=1).to(cuda)
    tgt = torch.from_numpy(po.gaussian_targets(fx["points"])).to(cuda)
    out = model(x)
    loss = torch.nn.MSELoss()(out, tgt)
    loss.backward()
    assert out.shape == (batch, joints, 192, 192)
    # the ViT's heatmaps are min-max normalised to [0, 1]: bf16 meets the strict element-wise gate directly
    m = po.heatmap_parity(out.detach().cpu()[:, ::6], torch.from_numpy(fx["out_sub"]))
    print(f"[parity ViT golden b2 {precision}] " + "  ".join(f"{k} {v:.3e}" for k, v in m.items()))
    assert m["floor10"] <= GATE[precision], m
    assert abs(loss.item() - float(fx["loss"])) <= tol_loss * float(fx["loss"])
    named = dict(model.named_parameters())
    worst = 1.0
    for k, n in zip([str(s) for s in fx["grad_keys"]], fx["grad_norm"]):
        g = named[k].grad
        assert g is not None, k
        assert abs(g.double().norm().item() - n) <= (2e-3 if precision == "fp32" else 8e-2) * n + 1e-12, k
        if "grad::" + k in fx.files:
            worst = min(worst, _cos(g.cpu(), torch.from_numpy(fx["grad::" + k])))
    assert worst >= min_cos
    assert named["vit_encoder.cls_token"].grad is None
    # fused train step == autograd path
    grads_autograd = {k: p.grad.clone() for k, p in named.items() if p.grad is not None}
    for p in model.parameters():
        p.grad = None
    loss2 = model.train_step(x, tgt)
    assert abs(loss2.item() - loss.item()) <= 1e-5 * abs(loss.item())
    for k, g in grads_autograd.items():
        assert _cos(named[k].grad, g) >= 0.9999, k


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_full_size_forward_vs_oracle(precision):
    """BASELINE.json configs[1] size: batch 64, C = 36.  The fp32 CPU oracle forward of the whole batch (a few
    seconds on the host) against the CUDA forward, every element of all 64 x 36 heatmaps; peaks bit-exact on the
    CUDA heatmaps, and agreeing with the oracle's peaks wherever the oracle's own maximum is unambiguous."""
    B, C = 64, 36
    model = _build(precision, C).eval()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    x = po.synthetic_crops(B, seed=11)
    with torch.no_grad():
        want = torch.cat([po.basicnet_forward(sd, x[i:i + 16]) for i in range(0, B, 16)])
        got = model(x.to(cuda)).cpu()
    floor = None
    if precision == "bf16":
        floor = torch.cat([po.basicnet_forward_operand_rounded(sd, x[i:i + 16], "bf16") for i in range(0, B, 16)])
    _check_parity(precision, got, want, floor, "BasicNet b64")
    pk = model.predict_peaks(x.to(cuda)).cpu().numpy()
    np.testing.assert_array_equal(pk, po.find_peaks_argmax(got.permute(0, 2, 3, 1).contiguous()))
    # keypoint agreement with the fp32 reference: random-init maps are nearly flat around their maximum (the runner-up
    # is usually the neighbouring pixel), so the exact location may differ; the reference's heatmap AT the predicted
    # location must be its maximum to within the heatmap tolerance, and most locations coincide exactly
    ref_pk = po.find_peaks_argmax(want.permute(0, 2, 3, 1).contiguous())
    xs, ys = torch.from_numpy(pk[..., 0]).long(), torch.from_numpy(pk[..., 1]).long()
    at_pred = want.flatten(2).gather(2, (ys * want.shape[-1] + xs).unsqueeze(-1)).squeeze(-1)
    slack = (2e-3 if precision == "fp16" else 2e-2) * want.abs().max()
    assert (at_pred >= want.flatten(2).max(dim=2).values - slack).all()
    agree = float((pk == ref_pk).all(axis=-1).mean())
    print(f"[peaks BasicNet b64 {precision}] identical to the fp32 reference's: {100 * agree:.1f} %")
    assert agree > (0.9 if precision == "fp16" else 0.6)


def test_cpu_tensor_raises():
    model = _build("bf16", 18)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 4, 192, 192))


def test_batched_repack_equals_lazy_packing():
    """after fused-Adam steps the packed operands refreshed by ONE pb_pack_weights_multi launch are bit-identical to
    operands packed from scratch (pb_pack_weights per layer)."""
    from pose_estimation_amitai_b200 import parallel
    model = _build("bf16", 36)
    dp = parallel.DataParallelStep(model, lr=1e-3)
    x = po.synthetic_crops(4, seed=3).to(cuda)
    pts = torch.from_numpy(po.synthetic_points(4, 36, seed=4)).to(cuda)
    for _ in range(3):
        dp.step(x, points=pts)
    eng = model.encoder._engine()
    assert getattr(eng, "_pack_table", None) is not None and eng._pack_table[2] >= 9
    model.eval()
    with torch.no_grad():
        out_repacked = model(x).clone()
        model.invalidate_packed_weights()
        out_fresh = model(x)
    assert torch.equal(out_repacked, out_fresh)


def test_full_size_step_is_additive_over_batch_shards():
    """BASELINE.json configs[1] size (batch 64, C=36, bf16): the fused training step is per-sample independent, so
    the gradient of the whole batch equals the sum of the gradients of its two halves (each scaled by its share of
    the loss mean) -- the property batch-sharded data parallelism relies on.  Size-independent check
    at the full benchmark size, where the CPU oracle would take minutes."""
    model = _build("bf16", 36)
    B = 64
    x = po.synthetic_crops(B, seed=7).to(cuda)
    pts = torch.from_numpy(po.synthetic_points(B, 36, seed=8)).to(cuda)
    loss_full = model.train_step(x, points=pts).item()
    named = {k: p for k, p in model.named_parameters() if p.grad is not None}
    g_full = {k: p.grad.clone() for k, p in named.items()}
    assert len(g_full) == 26 and np.isfinite(loss_full)
    # two half batches accumulated: each half's mean is over B/2 samples, so accumulation_steps=2 restores 1/B
    l0 = model.train_step(x[:B // 2], points=pts[:B // 2], accumulation_steps=2).item()
    l1 = model.train_step(x[B // 2:], points=pts[B // 2:], accumulation_steps=2, accumulate=True).item()
    assert abs((l0 + l1) - loss_full) <= 1e-5 * abs(loss_full)
    for k, p in named.items():
        assert _cos(p.grad, g_full[k]) >= 0.9999, k
        assert abs(p.grad.norm().item() / g_full[k].norm().item() - 1.0) <= 2e-3, k
    # peaks of the full-size forward are reproducible and in range
    pk = model.predict_peaks(x)
    assert pk.shape == (B, 36, 2) and torch.equal(pk, model.predict_peaks(x))
    assert (pk >= 0).all() and (pk <= 191).all()


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_cuda_graph_step_equals_eager_step(precision):
    """parallel.DataParallelStep.enable_graph(): six optimisation steps replayed from the captured CUDA graph (two eager
    steps, the capture, three replays -- on CHANGING batches) leave the same parameters, Adam moments and losses as the
    same six steps launched kernel by kernel.  The captured Adam launch reads the step number and the learning rate
    from device memory, so the bias corrections advance at every replay and an lr change reaches the replays."""
    from pose_estimation_amitai_b200 import parallel
    B, J = 4, 18
    runs = []
    for graph in (False, True):
        model = _build(precision, J)
        dp = parallel.DataParallelStep(model, lr=1e-3)
        if graph:
            dp.enable_graph()
        losses = []
        for step in range(6):
            x = po.synthetic_crops(B, seed=20 + step).to(cuda)
            pts = torch.from_numpy(po.synthetic_points(B, J, seed=40 + step)).to(cuda)
            if step == 4:
                dp.opt.lr = 5e-4      # ReduceLROnPlateau-style change between steps
            losses.append(dp.step(x, points=pts).item())
        torch.cuda.synchronize()
        runs.append((losses, dp.buckets.flat_param.clone(), dp.opt.exp_avg.clone(), dp.opt.exp_avg_sq.clone(),
                     dp.opt.step_count, len(dp._graphs)))
        pk = model.predict_peaks(x)      # eager inference after the replays sees the re-packed weights
        runs[-1] += (pk,)
    (l0, p0, m0, v0, s0, g0, k0), (l1, p1, m1, v1, s1, g1, k1) = runs
    assert g0 == 0 and g1 == 1 and s0 == s1 == 6
    np.testing.assert_allclose(l1, l0, rtol=1e-5)     # the loss scalar is an atomically ordered fp32 sum over CTAs
    np.testing.assert_allclose(p1.cpu().numpy(), p0.cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(m1.cpu().numpy(), m0.cpu().numpy(), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(v1.cpu().numpy(), v0.cpu().numpy(), rtol=1e-5, atol=1e-12)
    assert torch.equal(k0, k1) or (k0 - k1).abs().max().item() <= 1.0
