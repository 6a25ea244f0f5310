"""VIT4CamerasBaseLine / CrossAttention (pytorch/VITs.py:235-306): structure on the CPU, parity on the
GPU against vectors produced by the real reference module (tests/golden/multicam_next.npz) and against autograd of the
oracle restatement; the column-block glue kernel against torch indexing."""
import os

import numpy as np
import pytest
import torch

from oracle import pose_oracle as po

CFG = {"model type": "ALL_CAMS_18_POINTS_VIT", "number of base filters": 64, "convolution kernel size": 3,
       "dilation rate": 2, "dropout ratio": 0.5, "optimizer": "adam", "patch size": 16, "projection dim": 256,
       "num heads": 12, "dim head": -1, "transformer layers": 8}


def _fx(golden_dir):
    return np.load(os.path.join(golden_dir, "multicam_next.npz"), allow_pickle=False)


def _cos(a, b):
    a, b = a.flatten().double().cpu(), b.flatten().double().cpu()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-300)).item()


def _build(precision, seed=6, joints=72):
    from pose_estimation_amitai_b200 import Network, VITs
    torch.manual_seed(seed)
    model = Network.Network(dict(CFG, precision=precision), (192, 192, 4), joints).model
    assert isinstance(model, VITs.VIT4CamerasBaseLine)
    return model


def test_vit4_parameters_and_seeded_init(golden_dir):
    fx = _fx(golden_dir)
    model = _build("bf16")
    sd = model.state_dict()
    keys = [str(k) for k in fx["vit4_param_keys"]]
    assert [k for k, v in sd.items() if v.is_floating_point()] == keys and len(keys) == 172
    for k, s in zip(keys, fx["vit4_param_sum"]):
        assert np.isclose(sd[k].double().sum().item(), s, rtol=0, atol=1e-9 + 1e-12 * abs(s)), k
    assert model.cross_attentions[0].layers[0].layers[0][0].to_qkv.weight.shape == (3072, 1280)
    assert model.cross_attentions[3].layers[2].weight.shape == (256, 1280)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 16, 192, 192))
    live = [n for n, _ in model._live_params()]
    assert "shared_vit_encoder.cls_token" not in live and len(live) == 171


@pytest.mark.gpu
def test_colblock_against_torch_indexing():
    """pb_colblock: concatenation / split along features, row-modulo broadcast, folded sums, accumulation; fp32 and
    bf16; 8-element vector path and scalar path."""
    from pose_estimation_amitai_b200 import vit_ops
    g = torch.Generator().manual_seed(0)
    for dtype, cols in ((torch.float32, 24), (torch.bfloat16, 64), (torch.bfloat16, 20)):
        src = torch.rand(4 * 6, 40 + cols, generator=g).to("cuda", dtype)
        dst = torch.rand(6, 3 * cols, generator=g).to("cuda", dtype)
        want = dst.clone().float()
        want[:, cols:2 * cols] = src.float().view(4, 6, -1)[:, :, 40:40 + cols].sum(0)
        vit_ops.colblock(src, dst, rows=6, ncols=cols, src_row_stride=40 + cols, dst_row_stride=3 * cols, src_col0=40,
                         dst_col0=cols, nfold=4, fold_stride=6 * (40 + cols))
        torch.testing.assert_close(dst.float(), want.to(dtype).float(), rtol=0, atol=0)
        # broadcast of 2 source rows over 6 destination rows, accumulated
        small = torch.rand(2, cols, generator=g).to("cuda", dtype)
        before = dst.clone().float()
        vit_ops.colblock(small, dst, rows=6, ncols=cols, src_row_stride=cols, dst_row_stride=3 * cols, dst_col0=2 * cols,
                         src_rows_mod=2, accumulate=True)
        before[:, 2 * cols:] += small.float().repeat(3, 1)
        torch.testing.assert_close(dst.float(), before.to(dtype).float(), rtol=0, atol=0)


@pytest.mark.gpu
def test_view_rearrangements_follow_split_and_cat_order():
    from pose_estimation_amitai_b200 import VITs
    x = torch.arange(3 * 16 * 8 * 2, device="cuda").float().reshape(3, 16, 8, 2)
    vb = VITs.VIT4CamerasBaseLine._views_to_batch(x)
    for v, part in enumerate(torch.split(x, 4, dim=1)):
        assert torch.equal(vb[3 * v:3 * v + 3], part)
    assert torch.equal(VITs.VIT4CamerasBaseLine._batch_to_views(vb), x)
    pts = torch.rand(3, 72, 2, device="cuda")
    assert torch.equal(VITs.VIT4CamerasBaseLine._views_to_batch(pts), torch.cat(torch.split(pts, 18, dim=1), dim=0))


@pytest.mark.gpu
@pytest.mark.parametrize("precision,gate", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_vit4_vs_reference_golden(golden_dir, precision, gate):
    """forward against the REAL reference module's output (every 24th map, every 3rd pixel) on the same seeded init
    and input; heatmap gate floor10 = max |err| / (|ref| + 0.1 max|ref|)."""
    fx = _fx(golden_dir)
    model = _build(precision).cuda().eval()
    x = torch.rand(2, 16, 192, 192, generator=torch.Generator().manual_seed(11))[:1].cuda()
    with torch.no_grad():
        out = model(x)
    assert out.shape == (1, 72, 192, 192) and out.dtype == torch.float32
    m = po.heatmap_parity(out.cpu()[:, ::24, ::3, ::3], torch.from_numpy(fx["vit4_out_sub"]))
    print(f"[parity VIT4Cameras golden b1 {precision}] " + "  ".join(f"{k} {v:.3e}" for k, v in m.items()))
    assert m["floor10"] <= gate, m
    # every view's maps are normalised by themselves: each 18-map group spans exactly [0, 1]
    for v in range(4):
        grp = out[:, 18 * v:18 * (v + 1)]
        assert grp.min().item() == 0.0 and abs(grp.max().item() - 1.0) < 1e-6


@pytest.mark.gpu
def test_vit4_gradients_vs_oracle_autograd_and_fused_step():
    """autograd through the drop-in module == torch autograd of the oracle restatement (fp32 mode, every live
    parameter); the fused train step (bf16, per-view normalise + MSE tail) == the module's own autograd path."""
    torch.set_num_threads(os.cpu_count() or 1)
    joints, b = 72, 2
    x = torch.rand(b, 16, 192, 192, generator=torch.Generator().manual_seed(3))
    pts = po.synthetic_points(b, joints, seed=4)
    tgt = torch.from_numpy(po.gaussian_targets(pts))
    model = _build("fp32", seed=1).cuda().train()
    ref_params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    ref_out = po.vit_four_cameras_forward(ref_params, x)
    ref_loss = po.mse_loss(ref_out, tgt)
    ref_loss.backward()
    out = model(x.cuda())
    loss = torch.nn.MSELoss()(out, tgt.cuda())
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item())
    m = po.heatmap_parity(out.detach().cpu(), ref_out.detach())
    assert m["floor10"] <= 1e-4, m
    worst = 1.0
    for k, p in model.named_parameters():
        if k.endswith("cls_token"):
            assert p.grad is None and ref_params[k].grad is None
            continue
        assert p.grad is not None, k
        worst = min(worst, _cos(p.grad, ref_params[k].grad))
        n_ref = ref_params[k].grad.double().norm().item()
        assert abs(p.grad.double().norm().item() - n_ref) <= 5e-3 * n_ref + 1e-12, k
    assert worst >= 0.9999, worst
    # fused step vs the autograd path, bf16
    model = _build("bf16", seed=1).cuda().train()
    out = model(x.cuda())
    loss = torch.nn.MSELoss()(out, tgt.cuda())
    loss.backward()
    named = {k: p for k, p in model.named_parameters() if p.grad is not None}
    want = {k: p.grad.clone() for k, p in named.items()}
    for p in model.parameters():
        p.grad = None
    loss2 = model.train_step(x.cuda(), tgt.cuda())
    assert abs(loss2.item() - loss.item()) <= 1e-4 * abs(loss.item())
    for k, g in want.items():
        assert _cos(named[k].grad, g) >= 0.999, k
    loss3 = model.train_step(x.cuda(), points=torch.from_numpy(pts).cuda(), accumulate=True)
    assert abs(loss3.item() - loss.item()) <= 1e-3 * abs(loss.item())
    # peaks come from the same forward
    pk = model.predict_peaks(x.cuda())
    assert pk.shape == (b, joints, 2)
    with torch.no_grad():
        assert np.array_equal(pk.cpu().numpy(),
                              po.find_peaks_argmax(model(x.cuda()).cpu().permute(0, 2, 3, 1).contiguous()))
