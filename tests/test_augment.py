"""Input-pipeline augmentation: pb_affine_nearest / Datagenerators.DefaultDataset against the oracle
restatement and the vectors produced by the reference's own DefaultDataset + torchvision (tests/golden/augment.npz)."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import pose_oracle as po
from oracle import ref_shim


def _fx(golden_dir):
    return np.load(os.path.join(golden_dir, "augment.npz"), allow_pickle=False)


def _config(fx, model_type="MODEL_18_POINTS_PER_WING"):
    cfg = {str(k): (int(v) if float(v).is_integer() else float(v)) for k, v in zip(fx["config_keys"], fx["config_vals"])}
    cfg["zoom range"] = [float(v) for v in fx["zoom_range"]]
    cfg["model type"] = model_type
    return cfg


# ------------------------------------------------------------------------------------------ host logic (CPU)
def test_inverse_matrix_matches_oracle_and_golden(golden_dir):
    from pose_estimation_amitai_b200 import Datagenerators as dg
    fx = _fx(golden_dir)
    for (angle, tx, ty, sc), theta in zip(fx["kat_params"], fx["kat_theta"]):
        assert dg.inverse_affine_matrix(float(angle), (float(tx), float(ty)), float(sc)) == [float(v) for v in theta]
    rng = np.random.RandomState(0)
    for _ in range(500):
        a, tx, ty, sc = rng.uniform(-180, 180), rng.uniform(-30, 30), rng.uniform(-30, 30), rng.uniform(0.5, 2.0)
        assert dg.inverse_affine_matrix(a, (tx, ty), sc) == po.inverse_affine_matrix(a, (tx, ty), sc)


class _FakePre:
    def __init__(self, n):
        self.box = np.zeros((n, 4, 4, 1), np.uint8)
        self.conf = np.zeros((n, 4, 4, 1), np.float32)

    def get_box(self):
        return self.box

    def get_confmaps(self):
        return self.conf

    def get_num_frames(self):
        return len(self.box)


def _bare_generator(n_train, batch):
    """A DataGenerator without its device datasets: only the index logic under test."""
    from pose_estimation_amitai_b200 import Datagenerators as dg
    g = object.__new__(dg.DataGenerator)
    g.batch_size, g.train_indices, g.current_train_index = batch, np.arange(n_train), 0
    return g


def test_next_train_indices_wraps_like_the_reference():
    g = _bare_generator(10, 4)
    assert g.next_train_indices() == [0, 1, 2, 3]
    assert g.next_train_indices() == [4, 5, 6, 7]
    assert g.next_train_indices() == [8, 9, 0, 1]        # wraps to the start of the order
    assert g.current_train_index == 2
    g = _bare_generator(3, 8)                              # batch larger than the split: several wraps
    assert g.next_train_indices() == [0, 1, 2, 0, 1, 2, 0, 1]
    assert g.current_train_index == 2


@pytest.mark.skipif(not ref_shim.available(), reason="reference not mounted (GPU box)")
def test_index_logic_against_live_reference():
    """shuffle_train_indices / get_next_train_batch / get_train_val_split of the reference's DataGenerator
    (pytorch/Datagenerators.py:39-65,105-112), compiled from its source, under the same numpy seed."""
    path = os.path.join(ref_shim.REF_PT, "Datagenerators.py")
    with open(path) as fh:
        tree = ast.parse(fh.read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DataGenerator")
    keep = {"shuffle_train_indices", "get_next_train_batch", "get_train_val_split"}
    cls.body = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in keep]
    mod = ast.Module(body=[cls], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"np": np, "torch": torch}
    exec(compile(mod, path, "exec"), ns)
    ref = object.__new__(ns["DataGenerator"])
    n_train, batch = 23, 5
    ref.batch_size, ref.train_indices, ref.current_train_index, ref.val_fraction = batch, np.arange(n_train), 0, 0.1
    ref.train_dataset = [(torch.tensor(i), torch.tensor(i)) for i in range(n_train)]
    ours = _bare_generator(n_train, batch)
    ours.val_fraction = 0.1
    np.random.seed(5)
    want = []
    for epoch in range(3):
        ref.shuffle_train_indices()
        want += [ref.get_next_train_batch()[0].tolist() for _ in range(7)]
    want_split = ref.get_train_val_split(57)
    np.random.seed(5)
    got = []
    for epoch in range(3):
        ours.shuffle_train_indices()
        got += [ours.next_train_indices() for _ in range(7)]
    got_split = ours.get_train_val_split(57)
    assert got == want
    np.testing.assert_array_equal(got_split[0], want_split[0])
    np.testing.assert_array_equal(got_split[1], want_split[1])


# ------------------------------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
def test_affine_kernel_kat(golden_dir):
    """C-ABI kernel vs torchvision F.affine outputs (ties at .5, quarter turns, off-image shifts): bit-exact."""
    from pose_estimation_amitai_b200 import ops
    fx = _fx(golden_dir)
    img = torch.from_numpy(fx["kat_img"]).cuda()
    n = len(fx["kat_theta"])
    theta = torch.from_numpy(fx["kat_theta"].astype(np.float32)).cuda()
    got = ops.affine_nearest(img[None].expand(n, -1, -1, -1).contiguous(), theta).cpu().numpy()
    np.testing.assert_array_equal(got, fx["kat_out"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["train", "val"])
def test_default_dataset_matches_reference_batches(golden_dir, tag):
    """DefaultDataset.get_batch on the device == the reference's per-sample __getitem__ loop under the same
    np.random seed (ToTensor /255, draws in order, augment twice for train / once for val): bit-exact."""
    from pose_estimation_amitai_b200 import Datagenerators as dg
    fx = _fx(golden_dir)
    conf = np.moveaxis(po.gaussian_targets(fx["points"]), 1, -1)
    ds = dg.DefaultDataset(_config(fx), fx["box_u8"], conf, do_augmentations=(tag == "train"))
    assert len(ds) == fx["box_u8"].shape[0] and ds.box.dtype == torch.uint8
    np.random.seed(int(fx["np_seed"]))
    box, cm = ds.get_batch(range(len(ds)))
    np.testing.assert_array_equal(box.cpu().numpy(), fx[f"{tag}_box_u8"].astype(np.float32) / np.float32(255))
    np.testing.assert_array_equal(cm.cpu().numpy()[:, :2], fx[f"{tag}_conf_sub"])
    np.testing.assert_allclose(cm.double().sum(dim=(2, 3)).cpu().numpy(), fx[f"{tag}_conf_sum"], rtol=1e-12)
    # per-sample __getitem__ consumes the same random stream
    np.random.seed(int(fx["np_seed"]))
    b0, c0 = ds[0]
    np.testing.assert_array_equal(b0.cpu().numpy(), box[0].cpu().numpy())
    np.testing.assert_array_equal(c0.cpu().numpy(), cm[0].cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(5, 3, 33, 47), (3, 2, 192, 192), (2, 7, 64, 40), (0, 1, 8, 8)])
def test_affine_kernel_vs_oracle_random(shape):
    """random rotations / shifts / zooms / flips / gathers at odd and even sizes, fp32 and uint8 sources."""
    from pose_estimation_amitai_b200 import ops
    b, c, h, w = shape
    rng = np.random.RandomState(b * 100 + h)
    nsrc = b + 2
    for u8 in (False, True):
        data = rng.randint(0, 256, size=(nsrc, c, h, w)).astype(np.uint8) if u8 \
            else rng.rand(nsrc, c, h, w).astype(np.float32)
        theta = np.zeros((b, 6), np.float32)
        flips = rng.randint(0, 4, size=b).astype(np.int32)
        src = rng.randint(0, nsrc, size=b).astype(np.int32)
        mats = []
        for i in range(b):
            m = po.inverse_affine_matrix(rng.uniform(-180, 180), (rng.uniform(-20, 20), rng.uniform(-20, 20)),
                                         rng.uniform(0.6, 1.5))
            mats.append(m)
            theta[i] = m
        got = ops.affine_nearest(torch.from_numpy(data).cuda(), torch.from_numpy(theta).cuda(),
                                 torch.from_numpy(flips).cuda(), src_index=torch.from_numpy(src).cuda())
        assert got.shape == (b, c, h, w) and got.dtype == torch.float32
        for i in range(b):
            img = data[src[i]].astype(np.float32) / np.float32(255) if u8 else data[src[i]]
            want = po.affine_nearest(img, mats[i], bool(flips[i] & 1), bool(flips[i] & 2))
            np.testing.assert_array_equal(got[i].cpu().numpy(), want)


@pytest.mark.gpu
def test_affine_full_size_properties():
    """bench-size batch (64 x 36 x 192^2): identity matrix is a copy; a double h+v flip of the identity is a
    180-degree turn; every output value is an input value of the same sample/channel or zero."""
    from pose_estimation_amitai_b200 import ops
    b, c, h, w = 64, 36, 192, 192
    x = torch.rand(b, c, h, w, device="cuda")
    ident = torch.tensor([[1.0, 0, 0, 0, 1.0, 0]], device="cuda").repeat(b, 1)
    assert torch.equal(ops.affine_nearest(x, ident), x)
    both = torch.full((b,), 3, dtype=torch.int32, device="cuda")
    assert torch.equal(ops.affine_nearest(x, ident, both), torch.flip(x, dims=(2, 3)))
    rng = np.random.RandomState(1)
    theta = torch.tensor(np.array([po.inverse_affine_matrix(rng.uniform(-30, 30), (rng.uniform(-10, 10),
                         rng.uniform(-10, 10)), 1.0) for _ in range(b)], dtype=np.float32), device="cuda")
    y = ops.affine_nearest(x, theta)
    # a pure gather keeps per-(sample, channel) extrema inside the source's range and never invents values
    assert (y.amax(dim=(2, 3)) <= x.amax(dim=(2, 3))).all()
    nz = y[0, 0][y[0, 0] != 0]
    assert torch.isin(nz, x[0, 0].flatten()).all()


@pytest.mark.gpu
def test_affine_rejects_cpu_tensors():
    from pose_estimation_amitai_b200 import ops
    with pytest.raises(RuntimeError):
        ops.affine_nearest(torch.zeros(1, 1, 4, 4), torch.zeros(1, 6))


def test_draw_batch_follows_the_reference_draw_order(golden_dir):
    """DefaultDataset.draw_batch (host side of get_batch) against the oracle's restatement of augment_view's draws:
    sample-major, both passes of a training sample before the next sample, flips packed as bit 0 = h, bit 1 = v."""
    from pose_estimation_amitai_b200 import Datagenerators as dg
    fx = _fx(golden_dir)
    cfg = _config(fx)
    for aug in (True, False):
        ds = object.__new__(dg.DefaultDataset)      # the host logic only: no device tensors
        ds.xy_shifts, ds.rotation_range = cfg["augmentation shift x y"], cfg["rotation range"]
        ds.do_horizontal_flip, ds.do_vertical_flip = bool(cfg["horizontal flip"]), bool(cfg["vertical flip"])
        ds.scale_range, ds.do_augmentations = cfg["zoom range"], aug
        np.random.seed(77)
        theta, flips = ds.draw_batch(5)
        assert theta.shape == (2 if aug else 1, 5, 6) and theta.dtype == np.float32 and flips.dtype == np.int32
        rng = np.random.RandomState(77)
        for i in range(5):
            for p in range(2 if aug else 1):
                d = po.draw_augmentation(cfg, rng)
                m = po.inverse_affine_matrix(d["angle"], d["translate"], d["scale"])
                np.testing.assert_array_equal(theta[p, i], np.asarray(m, dtype=np.float32))
                assert flips[p, i] == int(d["hflip"]) | (int(d["vflip"]) << 1)
    # zero ranges draw nothing for angle / shifts (pytorch/Datagenerators.py:154-164) but still toss both coins
    ds.rotation_range = ds.xy_shifts = 0
    np.random.seed(3)
    theta, flips = ds.draw_batch(1)
    rng = np.random.RandomState(3)
    h, v = rng.rand() < 0.5, rng.rand() < 0.5
    assert flips[0, 0] == int(h) | (int(v) << 1)
    np.testing.assert_array_equal(theta[0, 0], np.array([1, 0, 0, 0, 1, 0], np.float32))
