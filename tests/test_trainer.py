"""Trainer entry point (pose_estimation_amitai_b200/train_pytorch.py) against the reference's loop semantics
(pytorch/train_pytorch.py:99-194): scheduler, checkpoint / CSV schemas on the CPU; the loop itself on the GPU
against the oracle's CPU restatement of the same loop."""
import csv
import json
import os

import numpy as np
import pytest
import torch

from oracle import pose_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _config(tmp_path, **over):
    with open(os.path.join(ROOT, "pose_estimation_amitai_b200", "train_config.json")) as fh:
        cfg = json.load(fh)
    cfg.update({"base output path": str(tmp_path), "batch_size": 2, "epochs": 2, "batches per epoch": 4,
                "accumulation_steps": 3, "synthetic samples": 12, "val_fraction": 0.5, "clean": 1})
    cfg.update(over)
    return cfg


def test_config_has_every_key_the_reference_trainer_reads():
    with open(os.path.join(ROOT, "pose_estimation_amitai_b200", "train_config.json")) as fh:
        cfg = json.load(fh)
    # pytorch/train_pytorch.py:38-55, CNNs.py:164-170, VITs.py:206-218
    for k in ("batch_size", "epochs", "batches per epoch", "val_fraction", "debug mode", "accumulation_steps",
              "base output path", "do augmentations", "loss_function", "clean", "model type",
              "number of base filters", "convolution kernel size", "dilation rate", "dropout ratio", "optimizer",
              "patch size", "projection dim", "num heads", "dim head", "transformer layers"):
        assert k in cfg, k


def test_reduce_lr_on_plateau_matches_torch():
    from pose_estimation_amitai_b200.train_pytorch import ReduceLROnPlateau

    class Opt:
        lr = 1e-3

    rs = np.random.RandomState(0)
    series = np.concatenate([np.linspace(1.0, 0.5, 6), 0.5 + 1e-7 * rs.rand(12), np.linspace(0.5, 0.49999, 9),
                             0.6 + 0.0 * rs.rand(30)])
    p = torch.nn.Parameter(torch.zeros(1))
    topt = torch.optim.Adam([p], lr=1e-3)
    tsch = torch.optim.lr_scheduler.ReduceLROnPlateau(topt, mode='min', factor=0.1, patience=3, threshold=1e-5,
                                                      threshold_mode='rel', cooldown=0, min_lr=1e-10)
    mine_opt = Opt()
    mine = ReduceLROnPlateau(mine_opt, mode='min', factor=0.1, patience=3, threshold=1e-5, threshold_mode='rel',
                             cooldown=0, min_lr=1e-10)
    for v in series:
        tsch.step(float(v))
        mine.step(float(v))
        assert mine_opt.lr == pytest.approx(topt.param_groups[0]["lr"], rel=1e-12)
    assert mine_opt.lr < 1e-3


def test_fused_adam_state_dict_is_torch_adam_schema():
    """the checkpoint's optimizer_state_dict loads into torch.optim.Adam(model.parameters()) and back."""
    from pose_estimation_amitai_b200 import parallel
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(2, 3, 3), torch.nn.BatchNorm2d(3), torch.nn.Conv2d(3, 1, 3))
    model[1].weight.requires_grad_(False)
    model[1].bias.requires_grad_(False)
    ordered = [(n, p) for n, p in reversed(list(model.named_parameters())) if p.requires_grad]
    fb = parallel.FlatBuckets(ordered)
    opt = parallel.FusedAdam(fb, lr=3e-4)
    assert opt.torch_state_dict(model)["state"] == {}
    opt.step_count = 7
    opt.exp_avg.copy_(torch.arange(fb.total, dtype=torch.float32))
    opt.exp_avg_sq.copy_(torch.arange(fb.total, dtype=torch.float32) * 2)
    sd = opt.torch_state_dict(model)
    ref = torch.optim.Adam(model.parameters(), lr=1e-3)
    ref.load_state_dict(sd)
    assert ref.param_groups[0]["lr"] == 3e-4
    params = list(model.parameters())
    assert set(sd["state"]) == {i for i, p in enumerate(params) if p.requires_grad}
    for i, st in sd["state"].items():
        assert st["exp_avg"].shape == params[i].shape and float(st["step"]) == 7.0
        assert torch.equal(ref.state[params[i]]["exp_avg_sq"], st["exp_avg_sq"])
    opt2 = parallel.FusedAdam(fb, lr=1.0)
    opt2.load_torch_state_dict(model, ref.state_dict())
    assert opt2.step_count == 7 and opt2.lr == 3e-4
    for p, o in zip(fb.params, fb.offsets):
        assert torch.equal(opt2.exp_avg[o:o + p.numel()], opt.exp_avg[o:o + p.numel()])


def test_trainer_without_gpu_fails_loudly(tmp_path):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pose_estimation_amitai_b200.train_pytorch import Trainer
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Trainer(_config(tmp_path))


@pytest.mark.gpu
def test_trainer_loop_matches_reference_loop(tmp_path):
    """fp32 mode, 2 epochs x 4 micro-batches, accumulation 3: the optimiser steps after micro-batch 3 of each
    epoch, and micro-batch 4's gradient leaks into the next epoch's step exactly as in the reference loop
    (pytorch/train_pytorch.py:125-144), restated below on the CPU with the oracle's forward."""
    from pose_estimation_amitai_b200.train_pytorch import Trainer
    cfg = _config(tmp_path, precision="fp32", **{"number of output channels": 5})
    tr = Trainer(cfg)
    gen = tr.data_generator
    init = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    # record the batches the trainer will draw (same RandomState stream), then rewind
    state = gen._rs.get_state()
    batches = []
    for _ in range(cfg["epochs"]):
        gen.shuffle_train_indices()
        for _ in range(cfg["batches per epoch"]):
            x, pts = gen.get_next_train_batch()
            batches.append((x.cpu(), pts.cpu().numpy()))
    gen._rs.set_state(state)
    hist = tr.train()

    params = {k: v.clone().requires_grad_(True) for k, v in init.items() if v.dtype.is_floating_point}
    live = [v for k, v in params.items() if ".bn" not in k]
    opt = torch.optim.Adam(live, lr=cfg["learning rate"])
    acc = cfg["accumulation_steps"]
    want_train = []
    it = iter(batches)
    for _ in range(cfg["epochs"]):
        running = 0.0
        for b in range(cfg["batches per epoch"]):
            x, pts = next(it)
            tgt = torch.from_numpy(po.gaussian_targets(pts))
            loss = po.mse_loss(po.basicnet_forward(params, x), tgt) / acc
            loss.backward()
            if (b + 1) % acc == 0:
                opt.step()
                opt.zero_grad()
            running += loss.item() * x.shape[0]
        want_train.append(running / (cfg["batches per epoch"] * cfg["batch_size"]))
    np.testing.assert_allclose(hist["train_losses"], want_train, rtol=2e-4)
    got = tr.model.state_dict()
    for k, v in params.items():
        if ".bn" in k:
            continue
        # Adam's first steps are ~ lr * sign(g): elements whose gradient is ~0 amplify fp32 summation-order
        # noise, so the update is compared as a whole (direction and size), not element by element
        d_got, d_want = (got[k].cpu() - init[k]).double().flatten(), (v.detach() - init[k]).double().flatten()
        cos = (torch.dot(d_got, d_want) / (d_got.norm() * d_want.norm() + 1e-300)).item()
        assert cos >= 0.999, (k, cos)
        assert abs(d_got.norm().item() / d_want.norm().item() - 1.0) <= 1e-2, k

    # artefacts: reference schemas
    ck = torch.load(os.path.join(tr.run_path, "checkpoint.pth"), weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "loss"} and ck["epoch"] == 1
    torch.optim.Adam(tr.model.parameters()).load_state_dict(ck["optimizer_state_dict"])
    rows = list(csv.reader(open(os.path.join(tr.run_path, "losses.csv"))))
    assert rows[0] == ['Epoch', 'Train Loss', 'Val Loss', 'L2 Loss', 'L2 Std', 'L2 Max Outlier'] and len(rows) == 3
    assert json.load(open(os.path.join(tr.run_path, "configuration.json")))["model type"] == cfg["model type"]
    # validation numbers == the oracle's on the final weights
    final = {k: v.cpu() for k, v in got.items()}
    tot, n_val, want_pts, want_tgt_pts = 0.0, 0, [], []
    for vx, vt in gen.val_batches():
        vo = po.basicnet_forward(final, vx.cpu())
        tot += po.mse_loss(vo, vt.cpu()).item() * vx.shape[0]
        n_val += vx.shape[0]
        want_pts.append(po.find_peaks_argmax(vo.detach().permute(0, 2, 3, 1)))
        want_tgt_pts.append(po.find_peaks_argmax(vt.cpu().permute(0, 2, 3, 1)))
    val_loss, l2_all, per_point = tr.validate()
    assert per_point.shape == (5, gen.num_val()) and n_val == gen.num_val()
    assert abs(val_loss - tot / n_val) <= 1e-4 * tot / n_val
    # target peaks are exact; predicted peaks of a barely-trained net sit on near-flat maps, so only the
    # target side and the distance arithmetic are compared bit-exactly
    got_tgt = np.concatenate([tr.get_points_from_confmaps(vt) for _, vt in gen.val_batches()])
    np.testing.assert_array_equal(got_tgt, np.concatenate(want_tgt_pts))


@pytest.mark.gpu
def test_trainer_bf16_loss_decreases(tmp_path):
    from pose_estimation_amitai_b200.train_pytorch import Trainer
    cfg = _config(tmp_path, epochs=3, accumulation_steps=1, **{"batches per epoch": 6, "batch_size": 4,
                                                                "synthetic samples": 16})
    tr = Trainer(cfg)
    hist = tr.train()
    assert hist["train_losses"][-1] < hist["train_losses"][0]
    assert np.isfinite(hist["l2_losses"]).all()


@pytest.mark.gpu
def test_trainer_with_cuda_graph_step_equals_eager_trainer(tmp_path):
    """config key "b200 cuda graph": the Trainer replays the optimisation step from one CUDA graph per batch shape
    (accumulation_steps == 1).  Same losses per epoch, same final parameters as the kernel-by-kernel Trainer on the
    same synthetic data; ReduceLROnPlateau's learning rate reaches the replays through the device-side scalar."""
    from pose_estimation_amitai_b200.train_pytorch import Trainer
    runs = []
    for graph in (0, 1):
        cfg = _config(tmp_path / f"g{graph}", epochs=3, accumulation_steps=1,
                      **{"batches per epoch": 6, "batch_size": 4, "synthetic samples": 16, "b200 cuda graph": graph})
        tr = Trainer(cfg)
        hist = tr.train()
        torch.cuda.synchronize()
        runs.append((hist["train_losses"], tr.dp.buckets.flat_param.clone(), len(tr.dp._graphs), tr.dp.opt.step_count))
    (l0, p0, g0, s0), (l1, p1, g1, s1) = runs
    assert g0 == 0 and g1 >= 1 and s0 == s1 == 18
    np.testing.assert_allclose(l1, l0, rtol=1e-5)
    np.testing.assert_allclose(p1.cpu().numpy(), p0.cpu().numpy(), rtol=1e-5, atol=1e-7)


class _ArrayPreprocessor:
    """stands in for the reference's HDF5 preprocessor (pytorch/preprocessor.py): the three methods DataGenerator calls."""

    def __init__(self, n, joints, seed=0):
        rs = np.random.RandomState(seed)
        self.box = rs.randint(0, 256, size=(n, 192, 192, 4)).astype(np.uint8)
        pts = rs.randint(24, 168, size=(n, joints, 2)).astype(np.float32)
        self.confmaps = np.ascontiguousarray(np.moveaxis(po.gaussian_targets(pts), 1, -1))   # (n, H, W, joints)

    def get_box(self):
        return self.box

    def get_confmaps(self):
        return self.confmaps

    def get_num_frames(self):
        return len(self.box)


@pytest.mark.gpu
def test_trainer_with_device_data_generator(tmp_path):
    """the reference's own data path on the device: DataGenerator(config, preprocessor) (pytorch/Datagenerators.py)
    feeding Trainer.train -- augmented uint8 crops + confidence-map targets; the first batch is the oracle's
    per-sample __getitem__ loop under the same np.random seed, bit for bit."""
    from pose_estimation_amitai_b200.Datagenerators import DataGenerator
    from pose_estimation_amitai_b200.train_pytorch import Trainer
    joints = 5
    cfg = _config(tmp_path, epochs=2, accumulation_steps=1, val_fraction=0.25,
                  **{"batches per epoch": 3, "batch_size": 2, "number of output channels": joints, "do augmentations": 1})
    pre = _ArrayPreprocessor(8, joints)
    np.random.seed(21)
    gen = DataGenerator(cfg, pre)
    assert len(gen.train_dataset) == 6 and gen.num_val() == 2
    # first batch vs the oracle restatement of DefaultDataset.__getitem__ (same global np.random stream)
    state = np.random.get_state()
    gen.shuffle_train_indices()
    order = gen.train_indices.copy()
    x, t = gen.get_next_train_batch()
    np.random.set_state(state)
    np.random.shuffle(np.arange(6))               # consume what shuffle_train_indices consumed
    for i in range(2):
        src = gen.train_inds[order[i]]
        wb, wc = po.dataset_getitem(pre.box[src], pre.confmaps[src], cfg, True, np.random)
        np.testing.assert_array_equal(x[i].cpu().numpy(), wb)
        np.testing.assert_array_equal(t[i].cpu().numpy(), wc)
    assert x.shape == (2, 4, 192, 192) and t.shape == (2, joints, 192, 192) and x.dtype == torch.float32
    tr = Trainer(cfg, data_generator=gen)
    hist = tr.train()
    assert len(hist["train_losses"]) == 2 and np.isfinite(hist["train_losses"]).all()
    assert np.isfinite(hist["val_losses"]).all() and np.isfinite(hist["l2_losses"]).all()


def test_scripted_archive_round_trip_on_cpu(tmp_path):
    """best_model.pth is a TorchScript archive like the reference's (pytorch/train_pytorch.py:177-181): scripting,
    saving and torch.jit.load work without a GPU; calling it on CPU tensors fails loudly (no CPU fallback)."""
    import pose_estimation_amitai_b200  # noqa: F401  (registers poseb200::heatmaps / ::peaks)
    from pose_estimation_amitai_b200 import CNNs, scripted
    cfg = {"model type": "MODEL_18_POINTS_PER_WING", "number of base filters": 64, "convolution kernel size": 3,
           "dilation rate": 2, "dropout ratio": 0.5}
    torch.manual_seed(0)
    model = CNNs.BasicNet(cfg, np.array((192, 192, 4)), 18)
    path = os.path.join(tmp_path, "best_model.pth")
    scripted.save(model, path)
    loaded = torch.jit.load(path)
    spec = json.loads(loaded.spec)
    assert spec["num_output_channels"] == 18 and spec["image_size"] == [192, 192, 4]
    assert len(spec["tensors"]) == len(model.state_dict()) == 91
    n_weights = sum(int(np.prod(shape)) if shape else 1 for _, _, shape, _ in spec["tensors"])
    assert loaded.flat_weights.numel() == n_weights
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        loaded(torch.zeros(1, 4, 192, 192))


@pytest.mark.gpu
def test_scripted_model_equals_eager_model(tmp_path):
    from pose_estimation_amitai_b200 import CNNs, scripted
    cfg = {"model type": "MODEL_18_POINTS_PER_WING", "number of base filters": 64, "convolution kernel size": 3,
           "dilation rate": 2, "dropout ratio": 0.5, "precision": "fp16"}
    torch.manual_seed(0)
    model = CNNs.BasicNet(cfg, np.array((192, 192, 4)), 18).cuda().eval()
    path = os.path.join(tmp_path, "best_model.pth")
    scripted.save(model, path)
    loaded = torch.jit.load(path, map_location="cuda")
    x = po.synthetic_crops(3, seed=4).cuda()
    with torch.no_grad():
        want = model(x)
    assert torch.equal(loaded(x), want)
    assert torch.equal(loaded.peaks(x), model.predict_peaks(x))


@pytest.mark.gpu
def test_raw_fused_adam_step_is_seen_by_the_next_forward():
    """parameters written through the C ABI (pb_adam_step) do not move torch's version counter; the packed tensor-core
    operands are tagged with the package's weight generation instead, so a FusedAdam WITHOUT any on_update hook still
    changes the next forward (stale packed weights)."""
    from pose_estimation_amitai_b200 import CNNs, parallel
    cfg = {"model type": "MODEL_18_POINTS_PER_WING", "number of base filters": 64, "convolution kernel size": 3,
           "dilation rate": 2, "dropout ratio": 0.5}
    torch.manual_seed(0)
    model = CNNs.BasicNet(cfg, np.array((192, 192, 4)), 18).cuda()
    x = po.synthetic_crops(2, seed=1).cuda()
    pts = torch.from_numpy(po.synthetic_points(2, 18, seed=2)).cuda()
    buckets = parallel.FlatBuckets(parallel.reverse_execution_order(model))
    opt = parallel.FusedAdam(buckets, lr=1e-2)          # no on_update
    with torch.no_grad():
        before = model(x).clone()
    model.train_step(x, points=pts)
    opt.step()
    with torch.no_grad():
        after = model(x).clone()
        model.invalidate_packed_weights()
        fresh = model(x)
    assert not torch.equal(before, after)
    assert torch.equal(after, fresh)
    # in-place writes through .data are invisible to both counters: the documented remedy is the explicit call
    model.encoder.conv1.weight.data.mul_(0.5)
    model.invalidate_packed_weights()
    with torch.no_grad():
        assert not torch.equal(model(x), fresh)


@pytest.mark.gpu
def test_trainer_resumes_from_checkpoint_and_reads_npz(tmp_path):
    """checkpoint.pth written by one Trainer restores weights, Adam moments / step, scheduler and best loss into a
    second one, which continues at the next epoch; the data comes from an .npz export (box / confmaps arrays) through
    the reference-surface DataGenerator; best_model.pth loads with torch.jit.load."""
    from pose_estimation_amitai_b200.train_pytorch import Trainer
    joints = 5
    pre = _ArrayPreprocessor(8, joints)
    npz = os.path.join(tmp_path, "data.npz")
    np.savez(npz, box=pre.box, confmaps=pre.confmaps)
    cfg = _config(tmp_path, epochs=2, accumulation_steps=1, val_fraction=0.25, data_path=npz,
                  **{"batches per epoch": 3, "batch_size": 2, "number of output channels": joints})
    first = Trainer(cfg)
    first.train()
    ck_path = os.path.join(first.run_path, "checkpoint.pth")
    ck = torch.load(ck_path, weights_only=True)
    assert ck["epoch"] == 1 and {"model_state_dict", "optimizer_state_dict", "loss"} <= set(ck)
    best = torch.jit.load(os.path.join(first.run_path, "best_model.pth"), map_location="cuda")
    assert best(torch.rand(1, 4, 192, 192, device="cuda")).shape == (1, joints, 192, 192)
    second = Trainer(dict(cfg, epochs=3, clean=0))      # clean=1 would wipe the first run's folder (same name / date)
    second.load_checkpoint(ck_path)
    assert second.start_epoch == 2 and second.dp.opt.step_count == first.dp.opt.step_count == 6
    assert torch.equal(second.dp.opt.exp_avg_sq, first.dp.opt.exp_avg_sq)
    for (k, a), (_, b) in zip(first.model.state_dict().items(), second.model.state_dict().items()):
        assert torch.equal(a, b), k
    hist = second.train()
    assert len(hist["train_losses"]) == 1 and second.dp.opt.step_count == 9
