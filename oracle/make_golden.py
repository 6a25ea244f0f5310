"""Generate tests/golden/*.npz from the REAL reference modules (build container only).

TEST INFRASTRUCTURE.  Run as ``python oracle/make_golden.py`` where /root/reference is
mounted.  The vectors travel to the GPU box with the repo; the reference does not.

What is recorded (everything produced by code imported from /root/reference):
  basicnet_c36.npz / vit_c36.npz
      weight checksums of ``torch.manual_seed(0); Model(cfg, (192,192,4), 36)``,
      the model output on seeded crops (batch 2; a channel subsample is stored in full,
      plus whole-tensor statistics), the MSE loss against sigma=3 Gaussian targets,
      per-parameter gradient norms and the small gradient tensors in full, the argmax
      peaks of the output, and the set of parameters whose grad is None.
  augment.npz
      DefaultDataset.__getitem__ (pytorch/Datagenerators.py:130-186: ToTensor + augment_view once / twice)
      on seeded uint8 crops + Gaussian confidence maps, and F.affine known-answer cases.
  fourcam_c72.npz
      FourCamerasBaseLine (multi-camera baseline), same recipe at 96x96x16, 72 joints, batch 1.
  multicam_next.npz
      FourCamerasDisentanglement / VIT4CamerasBaseLine forward outputs (the 8f2 models still to be built): pins the
      oracle restatements that round 2 will check the kernels against.
  kat.npz
      known-answer cases for argmax peaks (ties, NaNs, negatives), soft-argmax and the
      Gaussian target renderer.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import pose_oracle as po  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CH_STEP = 6  # channels stored in full: 0,6,...,30


def _model_fixture(kind: str, joints: int = 36, batch: int = 2) -> dict:
    CNNs, VITs, Augmentor = ref_shim.load_modules()
    cfg = ref_shim.load_config("MODEL_18_POINTS_PER_WING" if kind == "cnn" else "MODEL_18_POINTS_PER_WING_VIT")
    torch.manual_seed(0)
    cls = CNNs.BasicNet if kind == "cnn" else VITs.VIT_encoder_CNN_decoder
    model = cls(cfg, np.array((192, 192, 4)), joints)
    model.train()
    x = po.synthetic_crops(batch, seed=1)
    pts = po.synthetic_points(batch, joints, seed=2)
    tgt = torch.from_numpy(po.gaussian_targets(pts))
    out = model(x)
    loss = torch.nn.MSELoss()(out, tgt)
    loss.backward()
    fx: dict = {"joints": joints, "batch": batch}
    sd = model.state_dict()
    keys = [k for k, v in sd.items() if v.is_floating_point()]
    fx["param_keys"] = np.array(keys)
    fx["param_sum"] = np.array([sd[k].double().sum().item() for k in keys])
    fx["param_abs_sum"] = np.array([sd[k].double().abs().sum().item() for k in keys])
    fx["state_dict_len"] = len(sd)
    o = out.detach()
    fx["out_sub"] = o[:, ::CH_STEP].numpy()
    fx["out_stats"] = np.array([o.mean().item(), o.std().item(), o.min().item(), o.max().item()])
    fx["out_abs_sum_per_channel"] = o.double().abs().sum(dim=(0, 2, 3)).numpy()
    fx["loss"] = np.array(loss.item())
    fx["x_sum"] = np.array(x.double().sum().item())
    fx["points"] = pts
    named = dict(model.named_parameters())
    gkeys = [k for k, p in named.items() if p.grad is not None]
    fx["grad_keys"] = np.array(gkeys)
    fx["grad_none_keys"] = np.array([k for k, p in named.items() if p.grad is None])
    fx["grad_norm"] = np.array([named[k].grad.double().norm().item() for k in gkeys])
    for k in gkeys:
        if named[k].grad.numel() <= 4096:
            fx["grad::" + k] = named[k].grad.numpy()
    peaks = Augmentor.Augmentor.tf_find_peaks(o.permute(0, 2, 3, 1).contiguous().numpy())
    fx["peaks"] = peaks.cpu().numpy()
    return fx


def _four_cam_fixture(joints: int = 72, size: int = 96) -> dict:
    """FourCamerasBaseLine (pytorch/CNNs.py:189-237) from the real reference: seeded init, forward on seeded
    16-channel crops (batch 1, 96x96 to keep the CPU cost of the 1280-channel decoder low), MSE + backward."""
    CNNs, _, _ = ref_shim.load_modules()
    cfg = ref_shim.load_config("ALL_CAMS_18_POINTS")
    torch.manual_seed(0)
    model = CNNs.FourCamerasBaseLine(cfg, np.array((size, size, 16)), joints)
    model.train()
    x = po.synthetic_crops(1, seed=1, cin=16, size=size)
    pts = po.synthetic_points(1, joints, seed=2, size=size)
    tgt = torch.from_numpy(po.gaussian_targets(pts, size=size))
    out = model(x)
    loss = torch.nn.MSELoss()(out, tgt)
    loss.backward()
    sd = model.state_dict()
    fx: dict = {"joints": joints, "size": size, "state_dict_len": len(sd)}
    keys = [k for k, v in sd.items() if v.is_floating_point()]
    fx["param_keys"] = np.array(keys)
    fx["param_shapes"] = np.array([",".join(str(d) for d in sd[k].shape) for k in keys])
    fx["param_sum"] = np.array([sd[k].double().sum().item() for k in keys])
    o = out.detach()
    fx["out_sub"] = o[:, ::9].numpy()
    fx["out_stats"] = np.array([o.mean().item(), o.std().item(), o.min().item(), o.max().item()])
    fx["loss"] = np.array(loss.item())
    named = dict(model.named_parameters())
    gkeys = [k for k, p in named.items() if p.grad is not None]
    fx["grad_keys"] = np.array(gkeys)
    fx["grad_norm"] = np.array([named[k].grad.double().norm().item() for k in gkeys])
    for k in gkeys:
        if named[k].grad.numel() <= 4096:
            fx["grad::" + k] = named[k].grad.numpy()
    return fx


def _remaining_multicam_fixture() -> dict:
    """FourCamerasDisentanglement (train-mode BatchNorm, FTL / InvFTL; pytorch/CNNs.py:240-345) and
    VIT4CamerasBaseLine (pytorch/VITs.py:253-306) from the real reference: seeded init checksums and a strided
    subsample of the forward output -- the checker for the SURVEY 8f2 models that are not built yet."""
    CNNs, VITs, _ = ref_shim.load_modules()
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 16, 192, 192, generator=g)
    cams = torch.randn(2, 4, 3, 4, generator=g)
    cams_inv = torch.randn(2, 4, 4, 3, generator=g)
    fx: dict = {"cams": cams.numpy(), "cams_inv": cams_inv.numpy(), "x_sum": np.array(x.double().sum().item())}
    torch.manual_seed(5)
    m = CNNs.FourCamerasDisentanglement(ref_shim.load_config("ALL_CAMS_DISENTANGLED_PER_WING_CNN"),
                                        np.array((192, 192, 16)), 72).train()
    sd = m.state_dict()
    keys = [k for k, v in sd.items() if v.is_floating_point()]
    fx["dis_param_keys"] = np.array(keys)
    fx["dis_param_sum"] = np.array([sd[k].double().sum().item() for k in keys])
    with torch.no_grad():
        out = m(x, cams, cams_inv)
    fx["dis_out_sub"] = out[:, ::24, ::3, ::3].numpy()
    fx["dis_out_stats"] = np.array([out.mean().item(), out.std().item(), out.min().item(), out.max().item()])
    torch.manual_seed(6)
    v = VITs.VIT4CamerasBaseLine(ref_shim.load_config("ALL_CAMS_18_POINTS_VIT"), np.array((192, 192, 4)), 72).eval()
    sd = v.state_dict()
    keys = [k for k, t in sd.items() if t.is_floating_point()]
    fx["vit4_param_keys"] = np.array(keys)
    fx["vit4_param_sum"] = np.array([sd[k].double().sum().item() for k in keys])
    with torch.no_grad():
        out = v(x[:1])
    fx["vit4_out_sub"] = out[:, ::24, ::3, ::3].numpy()
    fx["vit4_out_stats"] = np.array([out.mean().item(), out.std().item(), out.min().item(), out.max().item()])
    return fx


def _kat_fixture() -> dict:
    _, _, Augmentor = ref_shim.load_modules()
    soft = ref_shim.soft_argmax_fn()
    gauss = ref_shim.gaussian_fn()
    g = torch.Generator().manual_seed(7)
    fx: dict = {}
    # argmax peaks: random maps, crafted ties, NaNs, all-negative, constant
    hm = torch.rand(3, 24, 20, 5, generator=g) - 0.3
    hm[0, 3, 4, 0] = 2.0
    hm[0, 17, 9, 0] = 2.0          # tie -> lowest flat index wins
    hm[1, 5, 6, 1] = float("nan")
    hm[1, 2, 19, 1] = float("nan")  # two NaNs -> first NaN
    hm[2, :, :, 2] = -1.5           # constant map -> index 0
    hm[2, :, :, 3] = -torch.rand(24, 20, generator=g) - 1.0  # all negative
    fx["argmax_in"] = hm.numpy()
    fx["argmax_out"] = Augmentor.Augmentor.tf_find_peaks(hm.numpy()).cpu().numpy()
    big = torch.rand(2, 192, 192, 4, generator=g)
    big[0, 191, 191, 0] = 3.0
    big[1, 0, 0, 1] = 3.0
    big[1, 100, 7, 2] = 3.0
    big[1, 100, 8, 2] = 3.0
    fx["argmax_big_seed"] = np.array(7)
    fx["argmax_big_in_sum"] = np.array(big.double().sum().item())
    fx["argmax_big_in"] = big.numpy().astype(np.float16)  # values exactly representable? no -> store rounded
    big16 = torch.from_numpy(fx["argmax_big_in"].astype(np.float32))
    fx["argmax_big_out"] = Augmentor.Augmentor.tf_find_peaks(big16.numpy()).cpu().numpy()
    # soft argmax
    sm = torch.zeros(2, 192, 192, 3)
    sm[0, 50, 120, 0] = 1.0
    sm[0, 10, 20, 1] = 1.0
    sm[0, 12, 20, 1] = 1.0
    sm[0, :, :, 2] = torch.rand(192, 192, generator=g)
    sm[1] = torch.rand(192, 192, 3, generator=g) - 0.2
    fx["soft_in"] = sm.numpy().astype(np.float16)
    fx["soft_out"] = soft(fx["soft_in"].astype(np.float32))
    # gaussian renderer
    means = np.array([[120.0, 50.0], [0.0, 0.0], [191.0, 191.0], [95.5, 17.25], [8.0, 183.0]])
    fx["gauss_means"] = means
    fx["gauss_out"] = np.stack([gauss(m) for m in means]).astype(np.float64)
    fx["gauss_sigma6"] = gauss(means[0], sigma=6).astype(np.float64)
    return fx


def _augment_fixture() -> dict:
    """DefaultDataset.__getitem__ / augment_view (pytorch/Datagenerators.py:130-186) run for real on seeded
    crops + confidence maps with the reference train_config.json augmentation keys; inputs are stored as
    uint8 (ToTensor's /255 path) and float32 so both dataset dtypes are pinned."""
    DefaultDataset = ref_shim.default_dataset_cls()
    cfg = ref_shim.load_config("MODEL_18_POINTS_PER_WING")
    g = np.random.RandomState(11)
    n, hw, cin, cj = 4, 192, 4, 5
    box = g.randint(0, 256, size=(n, hw, hw, cin)).astype(np.uint8)
    pts = g.randint(20, hw - 20, size=(n, cj, 2)).astype(np.float32)
    conf = np.moveaxis(po.gaussian_targets(pts), 1, -1).copy()          # (n, H, W, cj) float32
    fx: dict = {"box_u8": box, "points": pts, "np_seed": np.array(1234)}
    fx["config_keys"] = np.array(["rotation range", "augmentation shift x y", "horizontal flip", "vertical flip"])
    fx["config_vals"] = np.array([cfg["rotation range"], cfg["augmentation shift x y"], cfg["horizontal flip"],
                                  cfg["vertical flip"]], dtype=np.float64)
    fx["zoom_range"] = np.array(cfg["zoom range"], dtype=np.float64)
    for tag, aug in (("train", True), ("val", False)):
        ds = DefaultDataset(cfg, box=box, confmaps=conf, do_augmentations=aug)
        np.random.seed(1234)
        items = [ds[i] for i in range(n)]
        out = np.stack([b.numpy() for b, _ in items])
        q = np.rint(out * 255).astype(np.uint8)      # every output value is an input value (u8/255) or zero
        assert (q.astype(np.float32) / np.float32(255) == out).all()
        fx[f"{tag}_box_u8"] = q
        fx[f"{tag}_conf_sum"] = np.stack([c.double().sum(dim=(1, 2)).numpy() for _, c in items])
        fx[f"{tag}_conf_sub"] = np.stack([c.numpy()[:2] for _, c in items])
    # single augment_view calls with hand-picked parameters (ties at .5, 90-degree turns, big shifts)
    import torchvision.transforms.functional as TF
    img = torch.rand(3, 33, 47, generator=torch.Generator().manual_seed(5))
    cases = [(0.0, (0.5, -0.5), 1.0), (90.0, (0.0, 0.0), 1.0), (45.0, (3.0, -2.0), 1.0), (-180.0, (1.5, 2.5), 0.5),
             (12.25, (40.0, -60.0), 1.3), (0.0, (0.0, 0.0), 1.0)]
    fx["kat_img"] = img.numpy()
    fx["kat_params"] = np.array([[a, t[0], t[1], s] for a, t, s in cases])
    fx["kat_out"] = np.stack([TF.affine(img, angle=a, translate=t, scale=s, shear=0).numpy() for a, t, s in cases])
    fx["kat_theta"] = np.array([TF._get_inverse_affine_matrix([0.0, 0.0], a, list(t), s, [0.0, 0.0])
                                for a, t, s in cases])
    return fx


def main() -> None:
    if not ref_shim.available():
        raise SystemExit("reference not mounted at " + ref_shim.REF_ROOT)
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if "--augment-only" not in sys.argv and "--new-only" not in sys.argv:
        np.savez_compressed(os.path.join(OUT, "kat.npz"), **_kat_fixture())
    np.savez_compressed(os.path.join(OUT, "augment.npz"), **_augment_fixture())
    if "--augment-only" in sys.argv:
        return
    np.savez_compressed(os.path.join(OUT, "fourcam_c72.npz"), **_four_cam_fixture())
    np.savez_compressed(os.path.join(OUT, "multicam_next.npz"), **_remaining_multicam_fixture())
    if "--new-only" in sys.argv:
        return
    np.savez_compressed(os.path.join(OUT, "basicnet_c36.npz"), **_model_fixture("cnn"))
    np.savez_compressed(os.path.join(OUT, "vit_c36.npz"), **_model_fixture("vit"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
