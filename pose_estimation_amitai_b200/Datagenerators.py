"""Device-resident input pipeline with the reference's surface (pytorch/Datagenerators.py:16-186).

``DataGenerator(config, preprocessor)`` and ``DefaultDataset(config, box, confmaps, do_augmentations)`` keep the
reference's constructor arguments, index logic (``get_train_val_split``, ``shuffle_train_indices``,
``get_next_train_batch``) and, under the same ``np.random`` seed (the module seeds the global stream with 0 at import
exactly as the reference does, Datagenerators.py:14), produce bit-identical batches.  A batch is assembled by ONE
launch per tensor and augmentation pass of ``pb_affine_nearest``: batch gather + ``ToTensor`` (/255 for uint8) +
``F.affine(nearest)`` + flips.  Only the six matrix entries and the flip bits of every sample are
computed on the host (python doubles, exactly as torchvision's ``_get_inverse_affine_matrix`` does) and copied in.

Where the dataset lives.  *Resident* (default while it fits): the whole split sits in HBM (uint8 crops stay uint8:
147 KB per 192x192x4 sample) and the gather is folded into the kernel.  *Streaming* (``resident=False``, or
automatically when the split is larger than the ``"hbm dataset budget GB"`` config key / a quarter of free HBM): the
split stays in PINNED host memory; a batch's rows are gathered into one of two pinned staging buffers and moved with
one ``cudaMemcpyAsync`` per tensor on a copy stream, and ``DataGenerator`` stages batch k+1 (whose indices the index
logic already determines) while batch k trains -- the reference's per-sample ``__getitem__`` + ``torch.stack`` +
implicit H2D of every batch (Datagenerators.py:43-65) without its per-sample work.  Both modes produce the same bits.

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

np.random.seed(0)   # as the reference does at import (pytorch/Datagenerators.py:14, train_pytorch.py:34)


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

    def __init__(self, config: dict, box, confmaps, do_augmentations: bool = False, device=None,
                 resident: Optional[bool] = None, rng=None):
        """resident: True = dataset in HBM, False = pinned host memory + staged copies, None = decide by size.
        rng: the stream augmentation parameters are drawn from (default: the global ``np.random``, like the
        reference; data-parallel ranks pass their own RandomState)."""
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
        self.rng = rng if rng is not None else np.random
        host = [self._to_host_nchw(box), self._to_host_nchw(confmaps)]
        if host[0].shape[0] != host[1].shape[0]:
            raise ValueError("DefaultDataset: box and confmaps hold a different number of samples")
        nbytes = sum(t.numel() * t.element_size() for t in host)
        if resident is None:
            budget = config.get("hbm dataset budget GB")
            budget = float(budget) * 2 ** 30 if budget is not None else torch.cuda.mem_get_info(self.device)[0] / 4
            resident = nbytes <= budget
        self.resident = bool(resident)
        if self.resident:
            self.box, self.confmaps = (t.to(self.device) for t in host)
            self._stage = None
        else:
            self.box, self.confmaps = (t if t.is_pinned() else t.pin_memory() for t in host)
            self._stage = _HostStage(self.device)
        self.image_size = self.box.shape[-1]

    @staticmethod
    def _to_host_nchw(a) -> torch.Tensor:
        t = torch.as_tensor(np.ascontiguousarray(a)) if not torch.is_tensor(a) else a.detach().cpu()
        if t.dtype != torch.uint8:
            t = t.to(torch.float32)
        return t.permute(0, 3, 1, 2).contiguous()

    def __len__(self) -> int:
        return self.box.shape[0]

    # -- the reference's random draws, in its order (pytorch/Datagenerators.py:154-169) ---------------------
    def draw_view(self) -> Tuple[List[float], int]:
        rng = self.rng
        angle = rng.uniform(-self.rotation_range, self.rotation_range) if self.rotation_range != 0 else 0
        if self.xy_shifts != 0:
            shift_y = rng.uniform(-self.xy_shifts, self.xy_shifts)
            shift_x = rng.uniform(-self.xy_shifts, self.xy_shifts)
        else:
            shift_y = shift_x = 0
        hflip = rng.rand() < 0.5 and self.do_horizontal_flip
        vflip = rng.rand() < 0.5 and self.do_vertical_flip
        scaling = rng.uniform(self.scale_range[0], self.scale_range[1])
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

    def prefetch(self, indices: Iterable[int]) -> None:
        """streaming mode: start gathering + copying the raw rows of a FUTURE batch (no random draws happen here, so
        the ``np.random`` order of the batches is untouched).  No-op when the dataset is resident."""
        if self._stage is not None:
            self._stage.start(tuple(int(i) for i in indices), (self.box, self.confmaps))

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
        if self._stage is None:
            sources, src = (self.box, self.confmaps), params[p * n * 7:].view(torch.int32)
        else:
            # staged rows are already in batch order: the kernel's gather is the identity
            sources, src = self._stage.take(tuple(int(i) for i in idx), (self.box, self.confmaps)), None
        out = []
        for data in sources:
            cur = ops.affine_nearest(data, th[0], fl[0], src_index=src)
            for k in range(1, p):
                cur = ops.affine_nearest(cur, th[k], fl[k])
            out.append(cur)
        if self._stage is not None:
            self._stage.release()
        return out[0], out[1]

    def __getitem__(self, idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        b, c = self.get_batch([int(idx)])
        return b[0], c[0]


class _HostStage:
    """Two pinned staging buffers + two device buffers per tensor and a copy stream: rows of a host-resident dataset
    are gathered into pinned memory (one multi-threaded ``index_select``) and cross PCIe as ONE async copy per tensor
    and batch, overlapping the training step that is still consuming the previous buffer."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [dict(key=None, pinned=None, dev=None, ready=None, free=None) for _ in range(2)]
        self.turn = 0
        self.copies = 0          # async H2D copies issued (tests / bench count them)
        self.hits = 0            # batches that were already in flight when they were asked for

    def _buffers(self, slot, n, sources):
        if slot["pinned"] is None or slot["pinned"][0].shape[0] < n:
            slot["pinned"] = [torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory() for t in sources]
            slot["dev"] = [torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device) for t in sources]
        return slot["pinned"], slot["dev"]

    def start(self, key, sources):
        if any(s["key"] == key for s in self.slots):
            return
        slot = self.slots[self.turn]
        self.turn ^= 1
        if slot["free"] is not None:
            slot["free"].synchronize()            # kernels that read this slot's device buffers have finished
        if slot["ready"] is not None:
            slot["ready"].synchronize()           # ... and so has the previous copy out of its pinned buffers
        n = len(key)
        pinned, dev = self._buffers(slot, n, sources)
        index = torch.as_tensor(key, dtype=torch.int64)
        with torch.cuda.stream(self.stream):
            for t, pbuf, dbuf in zip(sources, pinned, dev):
                torch.index_select(t, 0, index, out=pbuf[:n])
                dbuf[:n].copy_(pbuf[:n], non_blocking=True)
                self.copies += 1
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(self.stream)
        slot["key"], slot["n"] = key, n

    def take(self, key, sources):
        slot = next((s for s in self.slots if s["key"] == key), None)
        if slot is None:
            self.start(key, sources)
            slot = next(s for s in self.slots if s["key"] == key)
        else:
            self.hits += 1
        torch.cuda.current_stream().wait_event(slot["ready"])
        out = [d[:slot["n"]] for d in slot["dev"]]
        slot["key"] = None                         # consumed: refillable once release() has marked its readers
        self._taken = slot
        return out

    def release(self):
        """call once the kernels reading the last take()'s buffers have been enqueued: records the event a later
        start() on that slot waits for before overwriting the device buffers."""
        slot, self._taken = getattr(self, "_taken", None), None
        if slot is not None:
            slot["free"] = torch.cuda.Event()
            slot["free"].record(torch.cuda.current_stream())


class _BatchLoader:
    """DataLoader(dataset, batch_size, shuffle=False, drop_last=False) over a DefaultDataset."""

    def __init__(self, dataset: DefaultDataset, batch_size: int):
        self.dataset, self.batch_size = dataset, int(batch_size)

    def __len__(self) -> int:
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        n = len(self.dataset)
        for b0 in range(0, n, self.batch_size):
            batch = self.dataset.get_batch(range(b0, min(n, b0 + self.batch_size)))
            if b0 + self.batch_size < n:          # streaming datasets: stage the next batch behind this one
                self.dataset.prefetch(range(b0 + self.batch_size, min(n, b0 + 2 * self.batch_size)))
            yield batch


class DataGenerator:
    """pytorch/Datagenerators.py:16-112 (single-view model types)."""

    def __init__(self, config: dict, preprocessor, device=None, rank: int = 0, world: int = 1,
                 resident: Optional[bool] = None):
        """rank / world: data-parallel ranks (one process per GPU) keep disjoint contiguous shards of both splits.
        A single process draws the split, the shuffles and the augmentations from the global ``np.random`` stream
        (seeded 0 at import) exactly like the reference.  With world > 1 the train / val split is drawn from a
        private ``RandomState(config["seed"])`` -- the SAME permutation on every rank whatever else the processes
        drew before, so no rank's training rows can sit in another rank's validation shard -- and each rank's
        shuffles / augmentations come from its own ``RandomState(seed + 1 + rank)``.
        resident: see DefaultDataset."""
        self.rank, self.world = int(rank), int(world)
        self.resident = resident
        seed = int(config.get("seed", 0))
        self._split_rng = np.random.RandomState(seed) if self.world > 1 else np.random
        self._rng = np.random.RandomState(seed + 1 + self.rank) if self.world > 1 else np.random
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
        self._rng.shuffle(self.train_indices)
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
        out = self.train_dataset.get_batch(self.next_train_indices())
        if not self.train_dataset.resident:
            # the index logic already fixes the next batch (unless a shuffle intervenes, in which case the staged rows
            # are simply not used): stage it while this one trains
            cursor = self.current_train_index
            self.train_dataset.prefetch(self.next_train_indices())
            self.current_train_index = cursor
        return out

    def config_data_generator(self):
        if self.model_type == "ALL_CAMS_DISENTANGLED_PER_WING_CNN":
            raise NotImplementedError("CameraMatrixGenerator (multi-camera) is outside this build (SURVEY.md 8f2)")
        self.box = self.preprocessor.get_box()
        self.confmaps = self.preprocessor.get_confmaps()
        self.num_samples = len(self.confmaps)
        self.train_inds, self.val_inds = self.split_and_shard(self.num_samples)
        train = DefaultDataset(self.config, box=self.box[self.train_inds], confmaps=self.confmaps[self.train_inds],
                               do_augmentations=self.do_augmentations, device=self.device, resident=self.resident,
                               rng=self._rng)
        val = DefaultDataset(self.config, box=self.box[self.val_inds], confmaps=self.confmaps[self.val_inds],
                             do_augmentations=False, device=self.device, resident=self.resident, rng=self._rng)
        return train, val

    def split_and_shard(self, num_samples: int):
        """this rank's (train rows, val rows): the reference's split, then contiguous shards of both halves."""
        train_inds, val_inds = self.get_train_val_split(num_samples)
        if self.world > 1:
            from .parallel import shard_range
            t0, t1 = shard_range(len(train_inds), self.rank, self.world)
            v0, v1 = shard_range(len(val_inds), self.rank, self.world)
            train_inds, val_inds = train_inds[t0:t1], val_inds[v0:v1]
        return train_inds, val_inds

    def get_train_dataloader(self):
        return self.train_dataloader

    def get_val_dataloader(self):
        return self.val_dataloader

    def get_vis_sample(self):
        return self.vis_sample

    def get_train_val_split(self, num_samples: int):
        all_inds = np.arange(num_samples)
        self._split_rng.shuffle(all_inds)
        val_size = round(num_samples * self.val_fraction)
        return all_inds[val_size:], all_inds[:val_size]

    # -- what this package's Trainer additionally calls (train_pytorch.py) ----------------------------------
    def val_batches(self):
        return iter(self.val_dataloader)

    def num_val(self) -> int:
        return len(self.val_dataset)
