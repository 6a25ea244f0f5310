"""Training entry point with the reference's surface (pytorch/train_pytorch.py):

    python -m pose_estimation_amitai_b200.train_pytorch train_config.json
    torchrun --nproc-per-node 8 -m pose_estimation_amitai_b200.train_pytorch train_config.json   (data parallel)

``Trainer(configuration_path).train()`` reads the same ``train_config.json`` keys, builds the model through
``Network.Network(config, image_size, num_output_channels).get_model()``, and keeps the reference's loop
semantics -- ``loss / accumulation_steps``, optimiser step every ``accumulation_steps`` micro-batches
(pytorch/train_pytorch.py:125-144), ``ReduceLROnPlateau(factor 0.1, patience 3, threshold 1e-5 rel, min_lr 1e-10)``
on the validation loss (:112-114,175), ``checkpoint.pth`` / ``losses.csv`` / ``configuration.json`` with the same
schemas (:253-283,347-349) -- while every tensor operation runs on the B200 path:

* forward + MSE + backward is ``model.train_step`` (one fused pass, gradients written into the flat buckets),
  the optimiser is the fused Adam kernel, data-parallel ranks exchange gradients with the bucketed all-reduce;
* bf16 has fp32's exponent range, so the reference's fp16 ``GradScaler`` (:115,137,140-141) has nothing to do:
  ``loss_scale`` stays 1 and no step is ever skipped;
* validation (:155-170,199-213) never copies a heatmap to the host: loss, arg-max peaks of prediction and target
  and their pixel L2 distances are computed on the device; only the (N, C) distances come back for the CSV.

Data.  The reference reads crops and confidence maps from an HDF5 file through ``preprocessor`` (h5py).  Data loading
is outside the hot path; this module takes ``"data_path": "synthetic"`` (or a missing file
together with ``"allow synthetic": 1``) and then draws crops ~ U[0,1) and integer keypoints whose sigma-3 Gaussian
confidence maps are rendered on the device (tensorflow/simple_data_generator.py:119-136).  Any object with the
``DataGenerator`` methods used below can be passed as ``data_generator=`` instead.

Known reference quirk kept on purpose: with ``batches per epoch`` not a multiple of ``accumulation_steps`` the
trailing micro-batches' gradients are neither stepped nor zeroed and are added to the next epoch's first step
(pytorch/train_pytorch.py:139-142).
"""
from __future__ import annotations

import csv
import json
import os
import shutil
import sys
from datetime import date
from time import time
from typing import Optional

import numpy as np
import torch

from . import Network, ops, parallel


class ReduceLROnPlateau:
    """torch.optim.lr_scheduler.ReduceLROnPlateau for an optimiser exposing ``.lr`` (the fused Adam):
    same bookkeeping as torch's (mode, rel/abs threshold, patience, cooldown, min_lr, eps)."""

    def __init__(self, optimizer, mode="min", factor=0.1, patience=3, threshold=1e-5, threshold_mode="rel",
                 cooldown=0, min_lr=1e-10, eps=1e-8, verbose=False):
        if factor >= 1.0:
            raise ValueError("Factor should be < 1.0.")
        self.optimizer, self.mode, self.factor, self.patience = optimizer, mode, factor, patience
        self.threshold, self.threshold_mode, self.cooldown, self.min_lr, self.eps = \
            threshold, threshold_mode, cooldown, min_lr, eps
        self.verbose = verbose
        self.best = float("inf") if mode == "min" else -float("inf")
        self.num_bad_epochs, self.cooldown_counter, self.last_epoch = 0, 0, 0

    def _is_better(self, a: float) -> bool:
        if self.mode == "min" and self.threshold_mode == "rel":
            return a < self.best * (1.0 - self.threshold)
        if self.mode == "min":
            return a < self.best - self.threshold
        if self.threshold_mode == "rel":
            return a > self.best * (self.threshold + 1.0)
        return a > self.best + self.threshold

    def step(self, metric: float) -> None:
        current = float(metric)
        self.last_epoch += 1
        if self._is_better(current):
            self.best, self.num_bad_epochs = current, 0
        else:
            self.num_bad_epochs += 1
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            old = float(self.optimizer.lr)
            new = max(old * self.factor, self.min_lr)
            if old - new > self.eps:
                self.optimizer.lr = new
                if self.verbose:
                    print(f"Epoch {self.last_epoch}: reducing learning rate to {new:.4e}.", flush=True)
            self.cooldown_counter = self.cooldown
            self.num_bad_epochs = 0


class SyntheticDataGenerator:
    """Stands in for Datagenerators.DataGenerator (pytorch/Datagenerators.py) with the methods the trainer calls.
    Crops and keypoints live on the device; confidence maps are rendered there on demand."""

    def __init__(self, config: dict, device, rank: int = 0, world: int = 1, image_size=(192, 192, 4),
                 num_output_channels: int = 18, sigma: float = 3.0):
        self.batch_size = int(config["batch_size"])
        self.val_fraction = 0.5 if bool(config.get("debug mode", 0)) else float(config["val_fraction"])
        n = int(config.get("synthetic samples", 64))
        self.h, self.w, self.cin = (int(v) for v in image_size)
        self.c, self.sigma, self.device = int(num_output_channels), float(sigma), device
        seed = int(config.get("seed", 1))
        g = torch.Generator().manual_seed(seed)
        box = torch.rand(n, self.cin, self.h, self.w, generator=g)
        pts = torch.randint(8, min(self.h, self.w) - 8, (n, self.c, 2), generator=g).float()
        n_val = int(np.round(n * self.val_fraction))
        idx = np.random.RandomState(seed).permutation(n)
        val_idx, train_idx = idx[:n_val], idx[n_val:]
        # data-parallel ranks own disjoint contiguous shards of both splits (batch-sharded)
        t0, t1 = parallel.shard_range(len(train_idx), rank, world)
        v0, v1 = parallel.shard_range(len(val_idx), rank, world)
        self.train_idx, self.val_idx = train_idx[t0:t1], val_idx[v0:v1]
        self.box, self.points = box.to(device), pts.to(device)
        self._order = np.array(self.train_idx)
        self._cursor = 0
        self._rs = np.random.RandomState(seed + 17 + rank)

    # -- the reference's DataGenerator surface used by Trainer.train ------------------------------------------
    def shuffle_train_indices(self) -> None:
        self._order = self._rs.permutation(self.train_idx)
        self._cursor = 0

    def get_next_train_batch(self):
        """(inputs [B,4,H,W], targets) -- targets are keypoints [B,C,2]; the trainer renders / fuses the maps."""
        if len(self._order) == 0:
            raise RuntimeError("SyntheticDataGenerator: this rank owns no training samples")
        take = [self._order[(self._cursor + i) % len(self._order)] for i in range(self.batch_size)]
        self._cursor = (self._cursor + self.batch_size) % len(self._order)
        sel = torch.as_tensor(np.array(take), device=self.device)
        return self.box[sel], self.points[sel]

    def val_batches(self):
        for b0 in range(0, len(self.val_idx), self.batch_size):
            sel = torch.as_tensor(np.array(self.val_idx[b0:b0 + self.batch_size]), device=self.device)
            yield self.box[sel], ops.gaussian_heatmaps(self.points[sel], sigma=self.sigma, size=(self.h, self.w))

    def num_val(self) -> int:
        return len(self.val_idx)

    def get_vis_sample(self):
        return self.box[:1], self.points[:1]


class ArrayPreprocessor:
    """The three methods DataGenerator calls on the reference's preprocessor, over arrays exported to an .npz
    (keys ``box`` [N,H,W,Cin], ``confmaps`` [N,H,W,C]; uint8 or float) or a pair of .npy files
    (``<path>`` = box, ``<path minus .npy>_confmaps.npy``), memory-mapped where numpy allows."""

    def __init__(self, path: str):
        if path.endswith(".npz"):
            z = np.load(path)
            self.box, self.confmaps = z["box"], z["confmaps"]
        else:
            self.box = np.load(path, mmap_mode="r")
            self.confmaps = np.load(path[:-4] + "_confmaps.npy", mmap_mode="r")
        if self.box.ndim != 4 or self.confmaps.ndim != 4 or len(self.box) != len(self.confmaps):
            raise ValueError("ArrayPreprocessor: expected box [N,H,W,Cin] and confmaps [N,H,W,C] of equal length")

    def get_box(self):
        return self.box

    def get_confmaps(self):
        return self.confmaps

    def get_num_frames(self):
        return len(self.box)


class Trainer:
    def __init__(self, configuration_path, data_generator=None):
        if isinstance(configuration_path, dict):
            config = dict(configuration_path)
        else:
            with open(configuration_path) as C:
                config = json.load(C)
        self.config = config
        self.batch_size = config['batch_size']
        self.num_epochs = config['epochs']
        self.batches_per_epoch = config['batches per epoch']
        self.val_fraction = config['val_fraction']
        self.debug_mode = bool(config["debug mode"])
        self.accumulation_steps = config['accumulation_steps']
        if self.debug_mode:
            self.val_fraction = 0.5
        self.base_output_path = config["base output path"]
        self.do_augmentations = bool(config.get("do augmentations", 0))
        self.viz_idx = 1
        self.loss_function = config["loss_function"]
        if self.loss_function != "mean_squared_error":
            raise ValueError("only the reference's 'mean_squared_error' heatmap loss is implemented")
        self.clean = bool(config["clean"])
        self.model_type = config["model type"]
        # the reference builds Adam(lr=0.001) and never reads config["learning rate"] (train_pytorch.py:111): the
        # key is ignored here too, so editing it changes nothing in either code base.  "b200 learning rate" (absent
        # from the reference's config) is this package's explicit override.
        self.learning_rate = float(config.get("b200 learning rate", 0.001))

        if not torch.cuda.is_available():
            raise RuntimeError("Trainer: no CUDA device -- the B200 hot path has no CPU fallback")
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = torch.device("cuda", local_rank)
        torch.cuda.set_device(self.device)
        if self.world > 1 and not torch.distributed.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.distributed.init_process_group("nccl", device_id=self.device)
        print("**************** CUDA is available. Using GPU. ****************", flush=True)

        self.num_output_channels = int(config.get("number of output channels", 18))
        self.img_size = np.array(config.get("image size", (192, 192, 4)))
        self.data_generator = data_generator
        if data_generator is None:
            path = str(config.get("data_path", "synthetic"))
            if path != "synthetic" and os.path.exists(path) and path.endswith((".npz", ".npy")):
                # arrays exported from the reference's preprocessor (box [N,H,W,Cin], confmaps [N,H,W,C]): the
                # reference-surface DataGenerator with its device-side augmentation; datasets larger than the HBM
                # budget stream from pinned host memory
                from . import Datagenerators
                self.data_generator = Datagenerators.DataGenerator(
                    config, ArrayPreprocessor(path), device=None, rank=int(os.environ.get("RANK", "0")),
                    world=int(os.environ.get("WORLD_SIZE", "1")))
            elif path != "synthetic" and os.path.exists(path):
                raise NotImplementedError(
                    "HDF5 datasets go through the reference's preprocessor (h5py is not in this image, SURVEY.md 8f3): "
                    "export its get_box() / get_confmaps() arrays to an .npz (keys 'box', 'confmaps') and point "
                    "data_path at it, pass data_generator=..., or set \"data_path\": \"synthetic\"")
            elif path != "synthetic" and not config.get("allow synthetic", 0):
                raise FileNotFoundError(f"data_path {path!r} not found (set \"data_path\": \"synthetic\" for the "
                                        "synthetic-data mode)")
            if getattr(self, "data_generator", None) is None:
                self.data_generator = SyntheticDataGenerator(config, self.device, self.rank, self.world,
                                                             self.img_size, self.num_output_channels)

        self.run_name = f"{self.model_type}_{date.today().strftime('%b %d')}"
        self.run_path = self.create_run_folders() if self.rank == 0 else None
        if self.rank == 0:
            self.save_configuration()

        torch.manual_seed(0)   # same random init on every rank
        self.network = Network.Network(config, image_size=self.img_size,
                                       num_output_channels=self.num_output_channels)
        self.model = self.network.get_model()
        print("img_size:", self.img_size, flush=True)
        print("num_output_channels:", self.num_output_channels, flush=True)
        self.dp: Optional[parallel.DataParallelStep] = None
        self.scheduler: Optional[ReduceLROnPlateau] = None
        self.start_epoch, self.best_loss = 0, float('inf')

    def _ensure_optimizer(self):
        """flat parameter / gradient buckets, fused Adam and the LR scheduler: built once, by whichever of train() /
        load_checkpoint() runs first, so a loaded optimiser state is the one training continues from."""
        if self.dp is None:
            self.model = self.model.to(self.device)
            self.dp = parallel.DataParallelStep(self.model, lr=self.learning_rate)
            if self.config.get("b200 cuda graph", 0):
                # replay the optimisation step from one CUDA graph per batch shape (accumulation_steps == 1 only;
                # with gradient accumulation the step keeps launching kernel by kernel)
                self.dp.enable_graph()
            self.scheduler = ReduceLROnPlateau(self.dp.opt, mode='min', factor=0.1, patience=3, verbose=True,
                                               threshold=1e-5, threshold_mode='rel', cooldown=0, min_lr=1e-10)
        return self.dp.opt

    # ------------------------------------------------------------------------------------------ training
    def train(self):
        t0_train = time()
        print("Using device", self.device, flush=True)
        optimizer = self._ensure_optimizer()
        best_loss = self.best_loss
        train_losses, val_losses, l2_losses, l2_losses_per_point, l2_stds, l2_max_outlier = [], [], [], [], [], []
        pending_micro = 0   # micro-batches whose gradients sit in the buckets, not yet stepped (never reset
        #                     at an epoch boundary: pytorch/train_pytorch.py:139-142)
        for epoch in range(self.start_epoch, self.num_epochs):
            print(f"Epoch {epoch + 1}/{self.num_epochs}", flush=True)
            self.model.train()
            self.data_generator.shuffle_train_indices()
            loss_acc = torch.zeros(1, device=self.device)
            for batch_num in range(self.batches_per_epoch):
                inputs, targets = self.data_generator.get_next_train_batch()
                batch_size = targets.size(0)
                if batch_num % 10 == 0:
                    print(f"Batch number is {batch_num + 1}, batch size is {batch_size}", flush=True)
                do_step = (batch_num + 1) % self.accumulation_steps == 0
                kw = {"points": targets} if targets.dim() == 3 else {}
                loss = self.dp.step(inputs, None if targets.dim() == 3 else targets,
                                    accumulation_steps=self.accumulation_steps, accumulate=pending_micro > 0,
                                    do_step=do_step, **kw)
                pending_micro = 0 if do_step else pending_micro + 1
                loss_acc += loss * batch_size        # stays on the device: no host sync inside the epoch
            if self.world > 1:
                torch.distributed.all_reduce(loss_acc)      # the logged loss is the mean over ALL ranks' samples
                loss_acc /= self.world
            epoch_loss = loss_acc.item() / (self.batches_per_epoch * self.batch_size)
            print(f'Train Loss: {epoch_loss:.7f}', flush=True)
            train_losses.append(epoch_loss)

            val_loss, l2_all, l2_per_point = self.validate()
            print(f'Val Loss: {val_loss:.4f}', flush=True)
            self.scheduler.step(val_loss)
            val_losses.append(val_loss)
            l2_stds.append(float(np.std(l2_all)))
            l2_losses.append(float(np.mean(l2_all)))
            l2_losses_per_point.append(l2_per_point)
            l2_max_outlier.append(float(np.max(l2_all)))
            if self.rank == 0:
                if val_loss < best_loss:
                    best_loss = val_loss
                    self.save_best_model()
                self.save_checkpoint(epoch, val_loss, self.model, optimizer, best_loss)
                self.save_losses_to_csv(len(train_losses) - 1, train_losses, val_losses, l2_losses, l2_stds,
                                        l2_max_outlier)
        elapsed_train = time() - t0_train
        print("Total runtime first loss: %.1f mins" % (elapsed_train / 60), flush=True)
        return {"train_losses": train_losses, "val_losses": val_losses, "l2_losses": l2_losses,
                "l2_stds": l2_stds, "l2_max_outlier": l2_max_outlier, "lr": optimizer.lr}

    @torch.no_grad()
    def validate(self):
        """pytorch/train_pytorch.py:150-170 on the device; with several ranks the sums are all-reduced."""
        self.model.eval()
        sq_sum = torch.zeros(1, device=self.device)
        n_seen = 0
        dists = []
        for inputs, confmaps in self.data_generator.val_batches():
            outputs = self.model(inputs)
            loss_sum, _, _ = ops.mse_loss_fwd_bwd(outputs.float().contiguous(), confmaps.float().contiguous())
            sq_sum += loss_sum / float(outputs[0].numel())     # sum over the batch of per-sample mean losses
            n_seen += outputs.size(0)
            dists.append(self.find_l2_val_loss(outputs, confmaps, as_numpy=False)[1])
        d = torch.cat(dists, dim=1) if dists else torch.zeros(self.num_output_channels, 0, device=self.device)
        count = torch.tensor([float(n_seen)], device=self.device)
        if self.world > 1:
            torch.distributed.all_reduce(sq_sum)
            torch.distributed.all_reduce(count)
            gathered = [None] * self.world
            torch.distributed.all_gather_object(gathered, d.cpu().numpy())
            per_point = np.concatenate(gathered, axis=1)
        else:
            per_point = d.cpu().numpy()
        val_loss = (sq_sum / count.clamp_min(1.0)).item()
        return val_loss, per_point.flatten(), per_point

    @staticmethod
    def find_l2_val_loss(output_confmaps, input_confmaps, as_numpy: bool = True):
        """pixel distance between predicted and target peaks (:199-205): (dists_flatten, dists_per_point[C, N])."""
        output_points = Trainer.get_points_from_confmaps(output_confmaps, as_numpy=False)
        input_points = Trainer.get_points_from_confmaps(input_confmaps, as_numpy=False)
        dists_per_point = torch.linalg.norm(output_points - input_points, dim=-1).T
        if as_numpy:
            dists_per_point = dists_per_point.cpu().numpy()
        return dists_per_point.flatten(), dists_per_point

    @staticmethod
    def get_points_from_confmaps(confmaps, as_numpy: bool = True):
        """(N, C, H, W) heatmaps -> (N, C, 2) peaks; the reference (:207-213) copies the maps to the host and
        transposes them first -- here only the peaks ever leave the device."""
        pts = ops.peaks_argmax(confmaps.detach().contiguous())
        return pts.cpu().numpy() if as_numpy else pts

    @staticmethod
    def find_points(confmaps):
        """(N, H, W, C) numpy or tensor -> (N, C, 2), pytorch/train_pytorch.py:327-331."""
        from . import utils
        return utils.torch_find_peaks_argmax(confmaps)

    # ------------------------------------------------------------------------------------------ artefacts
    def save_best_model(self):
        """the reference writes ``torch.jit.script(model)`` to best_model.pth (:177-181).  Here best_model.pth is a
        TorchScript archive too -- ``torch.jit.load`` works once this package is imported (it registers the
        ``poseb200::heatmaps`` custom op the scripted module calls) -- and the plain weights go next to it."""
        from . import scripted
        torch.save(self.model.state_dict(), os.path.join(self.run_path, 'best_model_state_dict.pth'))
        scripted.save(self.model, os.path.join(self.run_path, 'best_model.pth'))

    def save_checkpoint(self, epoch, epoch_loss, model, optimizer, best_loss=None):
        """checkpoint.pth holds exactly the reference's four keys (:253-260); what resuming needs beyond them
        (scheduler bookkeeping, best loss) goes to checkpoint_resume.pth next to it."""
        save_path = os.path.join(self.run_path, 'checkpoint.pth')
        torch.save({
            'epoch': epoch,
            'model_state_dict': model.state_dict(),
            'optimizer_state_dict': optimizer.torch_state_dict(model),
            'loss': epoch_loss,
        }, save_path)
        sched = self.scheduler
        torch.save({
            'best_loss': float(best_loss if best_loss is not None else epoch_loss),
            'scheduler_state': {"best": sched.best, "num_bad_epochs": sched.num_bad_epochs,
                                "cooldown_counter": sched.cooldown_counter, "last_epoch": sched.last_epoch}
            if sched is not None else None,
        }, os.path.join(self.run_path, 'checkpoint_resume.pth'))

    def load_checkpoint(self, path):
        """resume: weights, Adam moments / step / lr from checkpoint.pth (the reference's schema -- a checkpoint written
        by the reference loads too), scheduler bookkeeping and best loss from checkpoint_resume.pth when it exists;
        train() then continues at the epoch after the saved one.  Both files hold tensors and plain Python values
        only (weights_only=True)."""
        ck = torch.load(path, map_location=self.device, weights_only=True)
        opt = self._ensure_optimizer()
        self.model.load_state_dict(ck['model_state_dict'])
        if hasattr(self.model, "invalidate_packed_weights"):
            self.model.invalidate_packed_weights()
        opt.load_torch_state_dict(self.model, ck['optimizer_state_dict'])
        self.start_epoch = int(ck['epoch']) + 1
        self.best_loss = float(ck['loss'])
        extra_path = os.path.join(os.path.dirname(path), 'checkpoint_resume.pth')
        if os.path.exists(extra_path):
            extra = torch.load(extra_path, map_location="cpu", weights_only=True)
            self.best_loss = float(extra.get('best_loss', self.best_loss))
            for k, v in (extra.get('scheduler_state') or {}).items():
                setattr(self.scheduler, k, v)
        return ck['epoch'], ck['loss']

    def save_losses_to_csv(self, epoch, train_losses, val_losses, l2_losses, l2_stds, l2_max_outlier):
        csv_save_path = os.path.join(self.run_path, 'losses.csv')

        def format_significant(value, precision):
            return f"{value:.{precision}g}"

        with open(csv_save_path, 'w', newline='') as file:
            writer = csv.writer(file)
            writer.writerow(['Epoch', 'Train Loss', 'Val Loss', 'L2 Loss', 'L2 Std', 'L2 Max Outlier'])
            for i in range(epoch + 1):
                writer.writerow([i + 1, format_significant(train_losses[i], 4), format_significant(val_losses[i], 4),
                                 format_significant(l2_losses[i], 4), format_significant(l2_stds[i], 4),
                                 format_significant(l2_max_outlier[i], 4)])

    def save_configuration(self):
        with open(f"{self.run_path}/configuration.json", 'w') as file:
            json.dump(self.config, file, indent=4)

    def create_run_folders(self):
        run_path = os.path.join(self.base_output_path, self.run_name)
        if not self.clean:
            initial_run_path = run_path
            i = 1
            while os.path.exists(run_path):
                run_path = "%s_%02d" % (initial_run_path, i)
                i += 1
        if os.path.exists(run_path):
            shutil.rmtree(run_path)
        os.makedirs(run_path)
        # the reference also creates histograms/, viz_pred/, l2_histograms/, l2_histograms_per_point/ for its
        # matplotlib figures (:293-301); plotting is outside the hot path (no matplotlib here), so they are not made
        os.makedirs(os.path.join(run_path, "weights"))
        print("Created folder:", run_path)
        return run_path


if __name__ == '__main__':
    config_path = sys.argv[1]
    print(config_path)
    trainer = Trainer(config_path)
    trainer.train()
