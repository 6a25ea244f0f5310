"""FourCamerasBaseLine (pytorch/CNNs.py:189-237): structure on the CPU, parity on the GPU against
vectors produced by the real reference module (tests/golden/fourcam_c72.npz) and against the oracle restatement."""
import os

import numpy as np
import pytest
import torch

from oracle import pose_oracle as po
from oracle import ref_shim

CFG = {"model type": "ALL_CAMS_18_POINTS", "number of base filters": 64, "convolution kernel size": 3,
       "dilation rate": 2, "dropout ratio": 0.5}


def _fx(golden_dir):
    return np.load(os.path.join(golden_dir, "fourcam_c72.npz"), allow_pickle=False)


def _cos(a, b):
    a, b = a.flatten().double().cpu(), b.flatten().double().cpu()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-300)).item()


def _rel(got, ref):
    return max(((got - ref).abs().max() / ref.abs().max()).item(),
               ((got - ref).double().norm() / ref.double().norm()).item())


def test_fourcam_parameters_and_seeded_init(golden_dir):
    from pose_estimation_amitai_b200 import CNNs, Network
    fx = _fx(golden_dir)
    joints, size = int(fx["joints"]), int(fx["size"])
    torch.manual_seed(0)
    net = Network.Network(dict(CFG), (size, size, 16), joints)
    model = net.model
    assert isinstance(model, CNNs.FourCamerasBaseLine)
    sd = model.state_dict()
    assert len(sd) == int(fx["state_dict_len"])
    keys = [str(k) for k in fx["param_keys"]]
    assert [k for k, v in sd.items() if v.is_floating_point()] == keys
    for k, shp, s in zip(keys, fx["param_shapes"], fx["param_sum"]):
        assert ",".join(str(d) for d in sd[k].shape) == str(shp), k
        assert np.isclose(sd[k].double().sum().item(), s, rtol=0, atol=1e-9 + 1e-12 * abs(s)), k
    assert model.shared_conv2d.weight.shape == (1024, 1024, 1, 1)
    assert model.shared_decoder.conv2dTranspose1.weight.shape == (1280, 640, 3, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 16, size, size))
    # view <-> batch re-arrangements are each other's inverse and follow torch.split / torch.cat order
    x = torch.arange(2 * 16 * 3).float().reshape(2, 16, 3, 1)
    vb = CNNs.FourCamerasBaseLine._views_to_batch(x)
    for v, part in enumerate(torch.split(x, 4, dim=1)):
        assert torch.equal(vb[2 * v:2 * v + 2], part)
    assert torch.equal(CNNs.FourCamerasBaseLine._batch_to_views(vb), x)


def test_fourcam_oracle_seeded_init_matches_reference(golden_dir):
    fx = _fx(golden_dir)
    sd = po.four_cameras_state_dict(int(fx["joints"]))
    for k, shp, s in zip([str(k) for k in fx["param_keys"]], fx["param_shapes"], fx["param_sum"]):
        if ".bn" in k:
            continue
        assert ",".join(str(d) for d in sd[k].shape) == str(shp), k
        assert np.isclose(sd[k].double().sum().item(), s, rtol=0, atol=1e-9 + 1e-12 * abs(s)), k


def test_fourcam_oracle_matches_reference_golden(golden_dir):
    """the oracle restatement against the real module's outputs / loss / gradient norms."""
    torch.set_num_threads(os.cpu_count() or 1)
    fx = _fx(golden_dir)
    joints, size = int(fx["joints"]), int(fx["size"])
    from pose_estimation_amitai_b200 import CNNs
    torch.manual_seed(0)
    sd = {k: v for k, v in CNNs.FourCamerasBaseLine(dict(CFG), np.array((size, size, 16)), joints).state_dict().items()
          if ".bn" not in k}
    x = po.synthetic_crops(1, seed=1, cin=16, size=size)
    tgt = torch.from_numpy(po.gaussian_targets(po.synthetic_points(1, joints, seed=2, size=size), size=size))
    out, loss, grads = po.train_step_reference(sd, x, tgt, model="cnn4")
    np.testing.assert_allclose(out[:, ::9].numpy(), fx["out_sub"], rtol=1e-4, atol=1e-6)
    assert np.isclose(loss.item(), float(fx["loss"]), rtol=1e-5)
    gkeys = [str(k) for k in fx["grad_keys"]]
    assert set(gkeys) == set(grads.keys())
    for k, n in zip(gkeys, fx["grad_norm"]):
        assert np.isclose(grads[k].double().norm().item(), n, rtol=2e-3, atol=1e-12), k


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol_out,tol_loss,min_cos,tol_norm",
                         [("fp32", 1e-4, 1e-5, 0.9999, 1e-3), ("bf16", 2e-2, 2e-3, 0.99, 5e-2)])
def test_fourcam_vs_reference_golden(golden_dir, precision, tol_out, tol_loss, min_cos, tol_norm):
    from pose_estimation_amitai_b200 import CNNs
    fx = _fx(golden_dir)
    joints, size = int(fx["joints"]), int(fx["size"])
    torch.manual_seed(0)
    model = CNNs.FourCamerasBaseLine(dict(CFG, precision=precision), np.array((size, size, 16)), joints).cuda()
    x = po.synthetic_crops(1, seed=1, cin=16, size=size).cuda()
    pts = po.synthetic_points(1, joints, seed=2, size=size)
    tgt = torch.from_numpy(po.gaussian_targets(pts, size=size)).cuda()
    # autograd path, as train_pytorch.py:132-137 drives any model
    model.train()
    out = model(x)
    assert out.shape == (1, joints, size, size) and out.dtype == torch.float32
    loss = torch.nn.MSELoss()(out, tgt)
    loss.backward()
    assert _rel(out.detach().cpu()[:, ::9], torch.from_numpy(fx["out_sub"])) <= tol_out
    assert abs(loss.item() - float(fx["loss"])) <= tol_loss * float(fx["loss"])
    named = dict(model.named_parameters())
    for k, n in zip([str(s) for s in fx["grad_keys"]], fx["grad_norm"]):
        g = named[k].grad
        assert g is not None, k
        assert abs(g.double().norm().item() - n) <= tol_norm * n, k
        if "grad::" + k in fx.files:
            assert _cos(g, torch.from_numpy(fx["grad::" + k])) >= min_cos, k
    # engine-level fused step == autograd path
    auto = {k: p.grad.clone() for k, p in named.items() if p.grad is not None}
    for p in model.parameters():
        p.grad = None
    loss2 = model.train_step(x, tgt)
    assert abs(loss2.item() - loss.item()) <= 1e-4 * abs(loss.item())
    for k, g in auto.items():
        assert _cos(named[k].grad, g) >= (0.99999 if precision == "fp32" else 0.999), k
    # fused Gaussian targets from keypoints
    for p in model.parameters():
        p.grad = None
    loss3 = model.train_step(x, points=torch.from_numpy(pts).cuda())
    assert abs(loss3.item() - loss.item()) <= 1e-3 * abs(loss.item())


@pytest.mark.gpu
def test_fourcam_batch2_vs_oracle_and_data_parallel_step():
    """batch 2 at 64x64 against the oracle on the same weights (view/batch interleaving), then one
    DataParallelStep (flat buckets + fused Adam) against the oracle's Adam update."""
    from pose_estimation_amitai_b200 import CNNs, parallel
    joints, size = 8, 64
    torch.manual_seed(3)
    model = CNNs.FourCamerasBaseLine(dict(CFG, precision="fp32"), np.array((size, size, 16)), joints).cuda()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items() if ".bn" not in k}
    x = po.synthetic_crops(2, seed=4, cin=16, size=size)
    tgt = torch.from_numpy(po.gaussian_targets(po.synthetic_points(2, joints, seed=5, size=size), size=size))
    want_out, want_loss, want_grads = po.train_step_reference(sd, x, tgt, model="cnn4")
    with torch.no_grad():
        out = model(x.cuda())
    assert _rel(out.cpu(), want_out) <= 1e-4
    peaks = model.predict_peaks(x.cuda()).cpu().numpy()
    np.testing.assert_array_equal(peaks, po.find_peaks_argmax(out.cpu().permute(0, 2, 3, 1).contiguous()))
    dp = parallel.DataParallelStep(model, lr=1e-3)
    loss = dp.step(x.cuda(), tgt.cuda())
    assert abs(loss.item() - want_loss.item()) <= 1e-5 * want_loss.item()
    new_sd = model.state_dict()
    for k, g in want_grads.items():
        zeros = torch.zeros_like(sd[k])
        want_p, _, _ = po.adam_step(sd[k], g, zeros, zeros, 1)
        # first Adam step moves every weight by ~lr * sign(g): compare the update direction and size
        upd, want_upd = new_sd[k].cpu() - sd[k], want_p - sd[k]
        assert _cos(upd, want_upd) >= 0.99, k
