"""Device-resident input pipeline with the reference's surface (pytorch/Datagenerators.py:16-186).

``DataGenerator(config, preprocessor)`` and ``DefaultDataset(config, box, confmaps, do_augmentations)`` keep the
reference's constructor arguments, index logic (``get_train_val_split``, ``shuffle_train_indices``,
``get_next_train_batch``) and, under the same ``np.random`` seed, produce bit-identical batches -- but the whole
dataset lives in HBM (uint8 crops stay uint8: 147 KB per 192x192x4 sample) and a batch is assembled by ONE launch per
tensor and augmentation pass of ``pb_affine_nearest``: batch gather + ``ToTensor`` (/255 for uint8) +
``F.affine(nearest)`` + flips.  Only the six matrix entries and the flip bits of every sample are
computed on the host (python doubles, exactly as torchvision's ``_get_inverse_affine_matrix`` does) and copied in.

Reference behaviour kept on purpose (pytorch/Datagenerators.py:130-151): ``augment_view`` runs through
``cast_as_float`` TWICE per training sample when ``do augmentations`` is set and ONCE otherwise -- the validation
split is therefore augmented too.  Each pass is its own nearest-neighbour resampling (they do not compose exactly),
so each is its own launch.

Out of scope here: the HDF5 ``preprocessor`` (h5py; any object with ``get_box() / get_confmaps() /
get_num_frames()`` is accepted), the multi-camera branches (``ALL_CAMS_*``: SURVEY.md 8f2).
"""
from __future__ import annotations

import math
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .constants import ALL_CAMS_18_POINTS


def inverse_affine_matrix(angle: float, translate: Sequence[float], scale: float) -> List[float]:
    """What F.affine hands to the resampler for a tensor image: torchvision's inverse matrix for
    centre (0, 0) and zero shear, python doubles (pytorch/Datagenerators.py:170-173)."""
    rot = math.radians(angle)
    cs, sn = math.cos(rot), math.sin(rot)
    # RSS^-1 / scale, rows [d, -b, 0], [-c, a, 0] with a = d = cos, b = -sin (shear terms are exactly 0), c = sin
    m = [(-sn * 0.0 + cs) / scale, -(-cs * 0.0 - sn) / scale, 0.0, -sn / scale, cs / scale, 0.0]
    tx, ty = float(translate[0]), float(translate[1])
    m[2] += m[0] * (-0.0 - tx) + m[1] * (-0.0 - ty)
    m[5] += m[3] * (-0.0 - tx) + m[4] * (-0.0 - ty)
    return m


class DefaultDataset:
    """pytorch/Datagenerators.py:115-186 for the single-view models, on the device.

    ``box`` (N,H,W,Cin) and ``confmaps`` (N,H,W,C) are numpy arrays or tensors in the reference's channel-last
    layout (uint8 or float); they are moved to ``device`` once, as NCHW."""

    def __init__(self, config: dict, box, confmaps, do_augmentations: bool = False, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("DefaultDataset: no CUDA device -- the B200 input pipeline has no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.model_type = config["model type"]
        if self.model_type == ALL_CAMS_18_POINTS:
            raise NotImplementedError("multi-camera augmentation (ALL_CAMS_18_POINTS) is outside this build (8f2)")
        self.xy_shifts = config["augmentation shift x y"]
        self.rotation_range = config["rotation range"]
        self.do_horizontal_flip = bool(config["horizontal flip"])
        self.do_vertical_flip = bool(config["vertical flip"])
        self.scale_range = config["zoom range"]
        self.do_augmentations = do_augmentations
        self.box = self._to_device_nchw(box)
        self.confmaps = self._to_device_nchw(confmaps)
        self.image_size = self.box.shape[-1]

    def _to_device_nchw(self, a) -> torch.Tensor:
        t = torch.as_tensor(np.ascontiguousarray(a)) if not torch.is_tensor(a) else a
        if t.dtype != torch.uint8:
            t = t.to(torch.float32)
        return t.to(self.device).permute(0, 3, 1, 2).contiguous()

    def __len__(self) -> int:
        return self.box.shape[0]

    # -- the reference's random draws, in its order (pytorch/Datagenerators.py:154-169) ---------------------
    def draw_view(self) -> Tuple[List[float], int]:
        angle = np.random.uniform(-self.rotation_range, self.rotation_range) if self.rotation_range != 0 else 0
        if self.xy_shifts != 0:
            shift_y = np.random.uniform(-self.xy_shifts, self.xy_shifts)
            shift_x = np.random.uniform(-self.xy_shifts, self.xy_shifts)
        else:
            shift_y = shift_x = 0
        hflip = np.random.rand() < 0.5 and self.do_horizontal_flip
        vflip = np.random.rand() < 0.5 and self.do_vertical_flip
        scaling = np.random.uniform(self.scale_range[0], self.scale_range[1])
        return inverse_affine_matrix(angle, (shift_x, shift_y), scaling), int(bool(hflip)) | (int(bool(vflip)) << 1)

    @property
    def passes(self) -> int:
        return 2 if self.do_augmentations else 1

    def draw_batch(self, n: int) -> Tuple[np.ndarray, np.ndarray]:
        """theta [passes, n, 6] fp32, flips [passes, n] int32; drawn sample-major like a DataLoader walking
        ``__getitem__`` (all passes of sample 0, then sample 1, ...)."""
        theta = np.empty((self.passes, n, 6), dtype=np.float32)
        flips = np.empty((self.passes, n), dtype=np.int32)
        for i in range(n):
            for p in range(self.passes):
                m, f = self.draw_view()
                theta[p, i] = m     # torch.tensor(matrix, dtype=float32): double -> float rounding
                flips[p, i] = f
        return theta, flips

    def get_batch(self, indices: Iterable[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        """([B,Cin,H,W], [B,C,H,W]) fp32 for dataset rows ``indices`` == torch.stack of ``__getitem__`` results."""
        idx = np.asarray(list(indices), dtype=np.int32)
        n = int(idx.shape[0])
        theta, flips = self.draw_batch(n)
        params = torch.from_numpy(np.concatenate([theta.reshape(-1), flips.view(np.float32).reshape(-1),
                                                  idx.view(np.float32)])).to(self.device, non_blocking=True)
        p = self.passes
        th = params[:p * n * 6].view(p, n, 6)
        fl = params[p * n * 6:p * n * 7].view(torch.int32).view(p, n)
        src = params[p * n * 7:].view(torch.int32)
        out = []
        for data in (self.box, self.confmaps):
            cur = ops.affine_nearest(data, th[0], fl[0], src_index=src)
            for k in range(1, p):
                cur = ops.affine_nearest(cur, th[k], fl[k])
            out.append(cur)
        return out[0], out[1]

    def __getitem__(self, idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        b, c = self.get_batch([int(idx)])
        return b[0], c[0]


class _BatchLoader:
    """DataLoader(dataset, batch_size, shuffle=False, drop_last=False) over a DefaultDataset."""

    def __init__(self, dataset: DefaultDataset, batch_size: int):
        self.dataset, self.batch_size = dataset, int(batch_size)

    def __len__(self) -> int:
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        n = len(self.dataset)
        for b0 in range(0, n, self.batch_size):
            yield self.dataset.get_batch(range(b0, min(n, b0 + self.batch_size)))


class DataGenerator:
    """pytorch/Datagenerators.py:16-112 (single-view model types)."""

    def __init__(self, config: dict, preprocessor, device=None, rank: int = 0, world: int = 1):
        """rank / world: data-parallel ranks (one process per GPU) keep disjoint contiguous shards of both splits --
        the split itself is drawn from the same ``np.random`` stream on every rank, as in the reference."""
        self.rank, self.world = int(rank), int(world)
        self.config = config
        self.model_type = self.config["model type"]
        self.val_fraction = config["val_fraction"]
        self.do_augmentations = self.config["do augmentations"]
        self.batch_size = config["batch_size"]
        self.preprocessor = preprocessor
        self.num_frames = self.preprocessor.get_num_frames()
        self.device = device
        self.train_dataset, self.val_dataset = self.config_data_generator()
        self.train_dataloader = _BatchLoader(self.train_dataset, self.batch_size)
        self.val_dataloader = _BatchLoader(self.val_dataset, self.batch_size)
        self.vis_sample = (self.box[self.val_inds[0]], self.confmaps[self.val_inds[0]])
        self.train_indices = np.arange(len(self.train_dataset))
        self.current_train_index = 0

    def shuffle_train_indices(self) -> None:
        np.random.shuffle(self.train_indices)
        self.current_train_index = 0

    def next_train_indices(self) -> List[int]:
        """index selection of get_next_train_batch (pytorch/Datagenerators.py:43-58): wraps to the start of the
        shuffled order when the epoch's indices run out."""
        if len(self.train_indices) == 0:
            raise RuntimeError("DataGenerator: empty training split")
        batch_indices: List[int] = []
        while len(batch_indices) < self.batch_size:
            remaining = self.batch_size - len(batch_indices)
            start_index = self.current_train_index
            end_index = start_index + remaining
            if end_index > len(self.train_indices):
                batch_indices.extend(self.train_indices[start_index:].tolist())
                self.current_train_index = 0
            else:
                batch_indices.extend(self.train_indices[start_index:end_index].tolist())
                self.current_train_index = end_index
        return batch_indices[:self.batch_size]

    def get_next_train_batch(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.train_dataset.get_batch(self.next_train_indices())

    def config_data_generator(self):
        if self.model_type == "ALL_CAMS_DISENTANGLED_PER_WING_CNN":
            raise NotImplementedError("CameraMatrixGenerator (multi-camera) is outside this build (SURVEY.md 8f2)")
        self.box = self.preprocessor.get_box()
        self.confmaps = self.preprocessor.get_confmaps()
        self.num_samples = len(self.confmaps)
        self.train_inds, self.val_inds = self.get_train_val_split(self.num_samples)
        if self.world > 1:
            from .parallel import shard_range
            t0, t1 = shard_range(len(self.train_inds), self.rank, self.world)
            v0, v1 = shard_range(len(self.val_inds), self.rank, self.world)
            self.train_inds, self.val_inds = self.train_inds[t0:t1], self.val_inds[v0:v1]
        train = DefaultDataset(self.config, box=self.box[self.train_inds], confmaps=self.confmaps[self.train_inds],
                               do_augmentations=self.do_augmentations, device=self.device)
        val = DefaultDataset(self.config, box=self.box[self.val_inds], confmaps=self.confmaps[self.val_inds],
                             do_augmentations=False, device=self.device)
        return train, val

    def get_train_dataloader(self):
        return self.train_dataloader

    def get_val_dataloader(self):
        return self.val_dataloader

    def get_vis_sample(self):
        return self.vis_sample

    def get_train_val_split(self, num_samples: int):
        all_inds = np.arange(num_samples)
        np.random.shuffle(all_inds)
        val_size = round(num_samples * self.val_fraction)
        return all_inds[val_size:], all_inds[:val_size]

    # -- what this package's Trainer additionally calls (train_pytorch.py) ----------------------------------
    def val_batches(self):
        return iter(self.val_dataloader)

    def num_val(self) -> int:
        return len(self.val_dataset)
