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
    g.rank, g.world, g._rng, g._split_rng = 0, 1, np.random, np.random
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


def test_module_seeds_numpy_like_the_reference():
    """pytorch/Datagenerators.py:14 calls np.random.seed(0) at import; a fresh interpreter importing ours must leave
    the global stream in the same state (so an unseeded run draws the reference's split and augmentations)."""
    import subprocess
    import sys
    code = ("import numpy as np, sys; sys.path.insert(0, %r); import pose_estimation_amitai_b200.Datagenerators; "
            "print(np.random.randint(0, 2**31 - 1))" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    got = int(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout.split()[-1])
    assert got == np.random.RandomState(0).randint(0, 2 ** 31 - 1)


def test_data_parallel_split_is_rank_consistent():
    """every rank draws the SAME train / val permutation (private RandomState(seed)), whatever the process drew from
    the global stream before, and the rank shards are disjoint and cover both halves: no training row of one rank is
    a validation row of another."""
    from pose_estimation_amitai_b200 import Datagenerators as dg
    n, world = 103, 4
    shards = []
    for rank in range(world):
        np.random.seed(1000 + rank)          # ranks arrive with different global streams
        np.random.rand(rank * 7 + 1)
        g = object.__new__(dg.DataGenerator)
        g.rank, g.world, g.val_fraction = rank, world, 0.25
        g._split_rng, g._rng = np.random.RandomState(11), np.random.RandomState(12 + rank)
        shards.append(g.split_and_shard(n))
    train = np.concatenate([t for t, _ in shards])
    val = np.concatenate([v for _, v in shards])
    assert len(val) == round(n * 0.25) and len(train) + len(val) == n
    assert len(np.intersect1d(train, val)) == 0
    assert sorted(np.concatenate([train, val]).tolist()) == list(range(n))
    ref = np.arange(n)
    np.random.RandomState(11).shuffle(ref)
    np.testing.assert_array_equal(val, ref[:len(val)])
    np.testing.assert_array_equal(train, ref[len(val):])


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
def test_streaming_dataset_equals_resident(golden_dir):
    """a dataset kept in pinned host memory and staged batch by batch (one async copy per tensor and batch, next batch
    in flight while the current one is consumed) yields the same bits as the HBM-resident dataset, in the reference's
    random order; the prefetched batches are the ones consumed."""
    from pose_estimation_amitai_b200 import Datagenerators as dg
    fx = _fx(golden_dir)
    cfg = dict(_config(fx), batch_size=3, val_fraction=0.25, **{"do augmentations": 1})
    conf = np.moveaxis(po.gaussian_targets(fx["points"]), 1, -1)
    reps = 5
    box = np.concatenate([np.roll(fx["box_u8"], k, axis=1) for k in range(reps)])
    conf = np.concatenate([np.roll(conf, k, axis=2) for k in range(reps)]).astype(np.float32)

    class Pre:
        def get_box(self): return box
        def get_confmaps(self): return conf
        def get_num_frames(self): return len(box)

    out = {}
    for resident in (True, False):
        np.random.seed(21)
        gen = dg.DataGenerator(cfg, Pre(), resident=resident)
        assert gen.train_dataset.resident == resident
        got = []
        for epoch in range(2):
            gen.shuffle_train_indices()
            got += [tuple(t.clone() for t in gen.get_next_train_batch()) for _ in range(4)]
        got += [tuple(t.clone() for t in b) for b in gen.val_batches()]
        out[resident] = got
        if not resident:
            st = gen.train_dataset._stage
            assert st.copies >= 2 * 8 and st.hits >= 6       # every batch but the first of an epoch was staged ahead
            assert gen.train_dataset.box.is_pinned() and not gen.train_dataset.box.is_cuda
    assert len(out[True]) == len(out[False])
    for (b0, c0), (b1, c1) in zip(out[True], out[False]):
        assert torch.equal(b0, b1) and torch.equal(c0, c1)


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
        ds.scale_range, ds.do_augmentations, ds.rng = cfg["zoom range"], aug, np.random
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
