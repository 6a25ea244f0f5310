"""CPU oracle for the heatmap-regression hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The shipped package
(``pose_estimation_amitai_b200``) never imports anything from ``oracle/``.

It is a *functional* restatement (plain torch fp32 on the CPU, driven by a
state_dict) of what the reference's nn.Modules compute.  The arithmetic itself
lives in a third-party dependency of the reference -- PyTorch (no version pinned
by the reference; this container has torch 2.11.0+cu128, CPU kernels from
oneDNN/MKL) -- so the restatement calls the same ATen CPU ops
(conv2d / conv_transpose2d / linear / layer_norm / softmax / gelu) in the order
the reference modules do, and is pinned against the real reference modules by
``oracle/make_golden.py`` (run in the build container where /root/reference is
mounted; vectors committed under ``tests/golden/``).

Parity status: PINNED against the reference's own modules imported from
/root/reference (the reference ships no tests / golden vectors of its own, see
SURVEY.md section 4 and 8c), via tests/golden/*.npz.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.1  # nn.LeakyReLU(0.1): pytorch/CNNs.py:21,104 ; pytorch/VITs.py:35

StateDict = Dict[str, torch.Tensor]


def _act(t: torch.Tensor) -> torch.Tensor:
    return F.leaky_relu(t, LEAKY_SLOPE)


# --------------------------------------------------------------------------- #
# BasicNet  (pytorch/CNNs.py)
# --------------------------------------------------------------------------- #
def encoder2d_atrous_forward(sd: StateDict, x: torch.Tensor, prefix: str = "encoder.",
                             dilation: int = 2, padding: int = 2) -> torch.Tensor:
    """Encoder2DAtrous.forward, pytorch/CNNs.py:73-88.

    Three stages of dilated 3x3 convs (pad 2: CNNs.py:18,45-49), LeakyReLU after each
    conv, residual add inside a stage, 2x2 max-pool followed by a *second* LeakyReLU
    between stages (CNNs.py:77,82).  Dropout has p=float(int(0.5))=0 (CNNs.py:14,22).
    BatchNorm layers exist as parameters only (CNNs.py:25-43) and are never applied.
    """
    def conv(i: int, t: torch.Tensor) -> torch.Tensor:
        return F.conv2d(t, sd[f"{prefix}conv{i}.weight"], sd[f"{prefix}conv{i}.bias"],
                        stride=1, padding=padding, dilation=dilation)

    t = x
    for stage in range(3):
        a = _act(conv(3 * stage + 1, t))
        b = _act(conv(3 * stage + 2, a)) + a
        c = _act(conv(3 * stage + 3, b)) + b
        t = _act(F.max_pool2d(c, kernel_size=2, stride=2)) if stage < 2 else c
    return t


def decoder2d_forward(sd: StateDict, x: torch.Tensor, prefix: str = "decoder.") -> torch.Tensor:
    """Decoder2d.forward, pytorch/CNNs.py:151-157 (layer definitions :108-128).

    convT(s2,p1,op1) -> 2x [convT(s1,p1) + residual] -> convT(s2,p1,op1); LeakyReLU after
    every layer including the last; no min/max normalisation (commented out at :156).
    """
    def up(i: int, t: torch.Tensor, stride: int) -> torch.Tensor:
        return F.conv_transpose2d(t, sd[f"{prefix}conv2dTranspose{i}.weight"],
                                  sd[f"{prefix}conv2dTranspose{i}.bias"], stride=stride,
                                  padding=1, output_padding=1 if stride == 2 else 0)

    d1 = _act(up(1, x, 2))
    d2 = _act(up(2, d1, 1)) + d1
    d3 = _act(up(3, d2, 1)) + d2
    return _act(up(4, d3, 2))


def basicnet_forward(sd: StateDict, x: torch.Tensor, dilation: int = 2) -> torch.Tensor:
    """BasicNet.forward, pytorch/CNNs.py:183-186: decoder(encoder(x))."""
    return decoder2d_forward(sd, encoder2d_atrous_forward(sd, x, dilation=dilation))


def round_to(t: torch.Tensor, fmt: Optional[str]) -> torch.Tensor:
    """t rounded (to nearest even) to a 16-bit storage format and widened back to fp32; None / 'fp32' = identity."""
    if fmt in (None, "fp32"):
        return t
    return t.to({"bf16": torch.bfloat16, "fp16": torch.float16}[fmt]).float()


def basicnet_forward_operand_rounded(sd: StateDict, x: torch.Tensor, fmt: str, dilation: int = 2,
                                     weights_only: bool = False) -> torch.Tensor:
    """BasicNet.forward (pytorch/CNNs.py:73-88,151-157,183-186) as ANY implementation with 16-bit tensor-core
    operands must compute it: the same ATen ops in fp32, with the crop, every weight tensor and every tensor that is
    read back as a contraction operand (each layer's output after activation + residual add, each pooled map) rounded
    once to `fmt` ('bf16' | 'fp16').  Accumulation, bias, activation and residual adds stay fp32, the heatmaps are
    never rounded.  This is the parity FLOOR of a 16-bit-operand forward against the fp32 reference: tests require
    the CUDA path to sit at this floor (its error against the fp32 reference is not larger than this function's),
    i.e. that the kernels add nothing beyond the operand rounding the format itself imposes.
    weights_only: round the weight tensors alone (activations stay fp32) -- the smallest error any implementation
    holding its weights in `fmt` can have."""
    rw = lambda t: round_to(t, fmt)
    r = (lambda t: t) if weights_only else rw

    def conv(i: int, t: torch.Tensor) -> torch.Tensor:
        return F.conv2d(t, rw(sd[f"encoder.conv{i}.weight"]), sd[f"encoder.conv{i}.bias"], stride=1, padding=2,
                        dilation=dilation)

    def up(i: int, t: torch.Tensor, stride: int) -> torch.Tensor:
        return F.conv_transpose2d(t, rw(sd[f"decoder.conv2dTranspose{i}.weight"]), sd[f"decoder.conv2dTranspose{i}.bias"],
                                  stride=stride, padding=1, output_padding=1 if stride == 2 else 0)

    t = r(x)
    for stage in range(3):
        a = r(_act(conv(3 * stage + 1, t)))
        b = r(_act(conv(3 * stage + 2, a)) + a)
        c = r(_act(conv(3 * stage + 3, b)) + b)
        t = r(_act(F.max_pool2d(c, kernel_size=2, stride=2))) if stage < 2 else c
    d1 = r(_act(up(1, t, 2)))
    d2 = r(_act(up(2, d1, 1)) + d1)
    d3 = r(_act(up(3, d2, 1)) + d2)
    return _act(up(4, d3, 2))


def heatmap_parity(got: torch.Tensor, ref: torch.Tensor) -> Dict[str, float]:
    """The heatmap parity figures of DESIGN.md section 5, all against the fp32 reference `ref`:
      worst    max|err| / max|ref|                                 (worst element on the heatmap's scale)
      rms      ||err||_2 / ||ref||_2
      floor10  max |err| / (|ref| + 0.1 max|ref|)                  (element-wise relative, 10 % floor: THE GATE, 2e-2)
      s8d      max |err| / max(|ref|, 1e-3 max|ref|)               (SURVEY.md 8d's denominator; reported, see DESIGN)"""
    got, ref = got.double(), ref.double()
    err, scale = (got - ref).abs(), ref.abs().max()
    return {"worst": (err.max() / scale).item(), "rms": ((got - ref).norm() / ref.norm()).item(),
            "floor10": (err / (ref.abs() + 0.1 * scale)).max().item(),
            "s8d": (err / torch.clamp(ref.abs(), min=1e-3 * scale.item())).max().item()}


def four_cameras_baseline_forward(sd: StateDict, x: torch.Tensor, dilation: int = 2) -> torch.Tensor:
    """FourCamerasBaseLine.forward, pytorch/CNNs.py:220-237.  x [B,16,H,W] = four 4-channel
    views; ONE shared encoder per view, the four encodings concatenated and mixed by a 1x1 conv with a
    residual add (no activation), ONE shared decoder applied to cat(view encoding, mixed encodings)
    (256 + 1024 = 1280 channels), the four [B,C/4,H,W] outputs concatenated along channels."""
    views = torch.split(x, 4, dim=1)
    enc = [encoder2d_atrous_forward(sd, v, prefix="shared_encoder.", dilation=dilation) for v in views]
    all_enc = torch.cat(enc, dim=1)
    all_enc = F.conv2d(all_enc, sd["shared_conv2d.weight"], sd["shared_conv2d.bias"]) + all_enc
    dec = [decoder2d_forward(sd, torch.cat((e, all_enc), dim=1), prefix="shared_decoder.") for e in enc]
    return torch.cat(dec, dim=1)


def ftl(x: torch.Tensor, P: torch.Tensor) -> torch.Tensor:
    """FTL.forward, pytorch/CNNs.py:322-331: the raw reinterpretation of an NCHW (B,400,48,48) tensor as
    (B,48,48,100,4,1) -- no permute, so spatial and channel indices are scrambled exactly as in the reference --
    times the (3,4) camera matrix of each sample, reinterpreted back as (B,300,48,48)."""
    z = torch.reshape(x, (-1, 48, 48, 100, 4, 1))
    return torch.reshape(torch.reshape(P, (-1, 1, 1, 1, 3, 4)) @ z, (-1, 300, 48, 48))


def inv_ftl(x: torch.Tensor, inv_P: torch.Tensor) -> torch.Tensor:
    """InvFTL.forward, pytorch/CNNs.py:335-345: (B,300,48,48) -> (B,48,48,100,3,1), (4,3) @ -> (B,400,48,48)."""
    z = torch.reshape(x, (-1, 48, 48, 100, 3, 1))
    return torch.reshape(torch.reshape(inv_P, (-1, 1, 1, 1, 4, 3)) @ z, (-1, 400, 48, 48))


def four_cameras_disentanglement_forward(sd: StateDict, x: torch.Tensor, camera_matrices: torch.Tensor,
                                         camera_matrices_inv: torch.Tensor, dilation: int = 2,
                                         training: bool = True) -> torch.Tensor:
    """FourCamerasDisentanglement.forward, pytorch/CNNs.py:281-319 (not built on the B200 yet --
    restated so that the checker exists first).  Shared encoder per view, 1x1 conv to 300 channels, InvFTL into a
    canonical frame with each view's inverse camera matrix, two 1x1 fusion convs with BatchNorm (batch statistics
    in training mode; the SAME batch_norm3 module normalises all four re-projected views, each with its own batch
    statistics) + ReLU, FTL back into each view, 1x1 conv to 256 channels, encoder skip add, shared decoder."""
    def bn(key: str, t: torch.Tensor) -> torch.Tensor:
        if training:
            return F.batch_norm(t, None, None, sd[key + ".weight"], sd[key + ".bias"], True, 0.1, 1e-5)
        return F.batch_norm(t, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"],
                            sd[key + ".bias"], False, 0.1, 1e-5)

    views = torch.split(x, 4, dim=1)
    first = [encoder2d_atrous_forward(sd, v, prefix="shared_encoder.", dilation=dilation) for v in views]
    enc = [F.conv2d(e, sd["rearrange_layer_1.weight"], sd["rearrange_layer_1.bias"]) for e in first]
    canonic = [inv_ftl(e, camera_matrices_inv[:, i]) for i, e in enumerate(enc)]
    fus = torch.cat(canonic, dim=1)
    fus = F.relu(bn("batch_norm1", F.conv2d(fus, sd["fusion_layer_1.weight"], sd["fusion_layer_1.bias"])))
    fus = F.relu(bn("batch_norm2", F.conv2d(fus, sd["fusion_layer_2.weight"], sd["fusion_layer_2.bias"])))
    outs = []
    for i in range(4):
        ent = F.relu(bn("batch_norm3", ftl(fus, camera_matrices[:, i])))
        ent = F.conv2d(ent, sd["rearrange_layer_2.weight"], sd["rearrange_layer_2.bias"])
        outs.append(decoder2d_forward(sd, ent + first[i], prefix="shared_decoder."))
    return torch.cat(outs, dim=1)


# --------------------------------------------------------------------------- #
# ViT encoder + conv-transpose decoder  (pytorch/pytorch_vit_encoder.py, pytorch/VITs.py)
# --------------------------------------------------------------------------- #
def patchify(img: torch.Tensor, patch: int) -> torch.Tensor:
    """CustomViT.forward patch extraction, pytorch/pytorch_vit_encoder.py:135-138.

    (B,C,H,W) -> (B, (H/p)*(W/p), C*p*p) with the feature axis ordered (c, ph, pw).
    """
    b, c, h, w = img.shape
    t = img.reshape(b, c, h // patch, patch, w // patch, patch)
    return t.permute(0, 2, 4, 1, 3, 5).reshape(b, (h // patch) * (w // patch), c * patch * patch)


def _ln(sd: StateDict, key: str, t: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(t, (t.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def attention_forward(sd: StateDict, p: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """Attention.forward, pytorch/pytorch_vit_encoder.py:59-78 (pre-norm; qkv has no bias :52).

    scale = dim_head**-0.5 (:45); softmax over the key axis (:49); heads merged back as
    (b, n, heads*dim_head) before the output projection (:76-77).
    """
    b, n, _ = x.shape
    h = _ln(sd, p + "norm", x)
    qkv = F.linear(h, sd[p + "to_qkv.weight"])
    inner = qkv.shape[-1] // 3
    dh = inner // heads
    q, k, v = (t.reshape(b, n, heads, dh).transpose(1, 2) for t in qkv.split(inner, dim=-1))
    probs = torch.softmax((q @ k.transpose(-1, -2)) * (dh ** -0.5), dim=-1)
    o = (probs @ v).transpose(1, 2).reshape(b, n, inner)
    return F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def feedforward_forward(sd: StateDict, p: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward.forward, pytorch/pytorch_vit_encoder.py:18-28: LN -> Linear -> GELU(erf) -> Linear."""
    h = _ln(sd, p + "net.0", x)
    h = F.gelu(F.linear(h, sd[p + "net.1.weight"], sd[p + "net.1.bias"]))
    return F.linear(h, sd[p + "net.4.weight"], sd[p + "net.4.bias"])


def custom_vit_forward(sd: StateDict, img: torch.Tensor, patch: int, heads: int, depth: int,
                       prefix: str = "vit_encoder.") -> torch.Tensor:
    """CustomViT.forward, pytorch/pytorch_vit_encoder.py:131-149 + Transformer.forward :98-105.

    patchify -> Linear -> LayerNorm -> + pos_embedding (cls_token is never used) ->
    depth x [attn(x)+x ; ff(x)+x] -> final LayerNorm.
    """
    t = patchify(img, patch)
    t = F.linear(t, sd[prefix + "patch_to_embedding.weight"], sd[prefix + "patch_to_embedding.bias"])
    t = _ln(sd, prefix + "norm", t)
    t = t + sd[prefix + "pos_embedding"][:, : t.shape[1]]
    for layer in range(depth):
        lp = f"{prefix}transformer.layers.{layer}."
        t = attention_forward(sd, lp + "0.", t, heads) + t
        t = feedforward_forward(sd, lp + "1.", t) + t
    return _ln(sd, prefix + "transformer.norm", t)


def cnn_decoder_forward(sd: StateDict, tokens: torch.Tensor, dim: int,
                        prefix: str = "cnn_decoder.") -> torch.Tensor:
    """CNN_Decoder.forward, pytorch/VITs.py:38-46.

    The token matrix (B,144,dim) is *reinterpreted* (no transpose) as (B,dim,12,12) (:39),
    then 4 x [convT(k3,s2,p1,op1) + LeakyReLU] and a min/max normalisation taken over the
    whole batch tensor (:45,55-58).
    """
    t = tokens.reshape(-1, dim, 12, 12)
    for i in range(1, 5):
        t = _act(F.conv_transpose2d(t, sd[f"{prefix}deconv{i}.weight"], sd[f"{prefix}deconv{i}.bias"],
                                    stride=2, padding=1, output_padding=1))
    lo, hi = t.min(), t.max()
    return (t - lo) / (hi - lo)


def vit_forward(sd: StateDict, x: torch.Tensor, patch: int = 16, heads: int = 12, depth: int = 8,
                dim: int = 256) -> torch.Tensor:
    """VIT_encoder_CNN_decoder.forward, pytorch/VITs.py:226-229."""
    return cnn_decoder_forward(sd, custom_vit_forward(sd, x, patch, heads, depth), dim)


def cross_attention_forward(sd: StateDict, p: str, x: torch.Tensor, heads: int = 4) -> torch.Tensor:
    """CrossAttention.forward, pytorch/VITs.py:235-250: Transformer(dim, depth 1, 4 heads) -> LayerNorm -> Linear
    (dim -> projection dim) -> GELU."""
    t = attention_forward(sd, p + "layers.0.layers.0.0.", x, heads) + x
    t = feedforward_forward(sd, p + "layers.0.layers.0.1.", t) + t
    t = _ln(sd, p + "layers.0.norm", t)
    t = _ln(sd, p + "layers.1", t)
    return F.gelu(F.linear(t, sd[p + "layers.2.weight"], sd[p + "layers.2.bias"]))


def vit_four_cameras_forward(sd: StateDict, x: torch.Tensor, patch: int = 16, heads: int = 12, depth: int = 8,
                             dim: int = 256, cross_layers: int = 4) -> torch.Tensor:
    """VIT4CamerasBaseLine.forward, pytorch/VITs.py:287-306 (checker only so far): the shared ViT
    encoder on each 4-channel view; four rounds in which every view's tokens, concatenated with ALL four ORIGINAL
    encodings (``encodings`` is built once, :295), pass the round's CrossAttention and are added back; the encoder
    output is added once more as a skip; the shared CNN decoder (with its per-call min/max normalisation)."""
    views = torch.split(x, 4, dim=1)
    skip = [custom_vit_forward(sd, v, patch, heads, depth, prefix="shared_vit_encoder.") for v in views]
    enc = list(skip)
    encodings = torch.cat(skip, dim=-1)
    for i in range(cross_layers):
        for v in range(4):     # enc[v] is updated in place of the list: later views see the same `encodings`
            enc[v] = cross_attention_forward(sd, f"cross_attentions.{i}.", torch.cat([enc[v], encodings], dim=-1)) + enc[v]
    outs = [cnn_decoder_forward(sd, enc[v] + skip[v], dim, prefix="shared_cnn_decoder.") for v in range(4)]
    return torch.cat(outs, dim=1)


# --------------------------------------------------------------------------- #
# loss / peaks / targets
# --------------------------------------------------------------------------- #
def mse_loss(outputs: torch.Tensor, targets: torch.Tensor, accumulation_steps: int = 1) -> torch.Tensor:
    """torch.nn.MSELoss() then / accumulation_steps, pytorch/train_pytorch.py:110,134-135."""
    d = outputs.float() - targets.float()
    return (d * d).mean() / accumulation_steps


def mse_loss_grad(outputs: torch.Tensor, targets: torch.Tensor, accumulation_steps: int = 1,
                  scale: float = 1.0) -> torch.Tensor:
    """d(scale * mse/acc)/d(outputs) -- what scaler.scale(loss).backward() seeds
    (pytorch/train_pytorch.py:137)."""
    n = outputs.numel()
    return (outputs.float() - targets.float()) * (2.0 * scale / (n * accumulation_steps))


def find_peaks_argmax(confmaps_nhwc) -> np.ndarray:
    """Augmentor.tf_find_peaks, pytorch/Augmentor.py:105-148 (== preprocessor.py:630-668,
    utils.py:6-44).  (N,H,W,C) -> (N,C,2) float32 [x=col, y=row]; flat-index tie break =
    lowest index; NaN compares as the maximum (torch.max semantics).
    """
    t = torch.as_tensor(np.asarray(confmaps_nhwc)) if not torch.is_tensor(confmaps_nhwc) else confmaps_nhwc
    n, h, w, c = t.shape
    _, flat_idx = torch.max(t.reshape(n, h * w, c), dim=1)
    xs = (flat_idx % w).to(torch.float32)
    ys = (flat_idx // w).to(torch.float32)
    return torch.stack([xs, ys], dim=-1).cpu().numpy()


def find_peaks_soft_argmax(confmaps_nhwc) -> np.ndarray:
    """find_peaks_soft_argmax, pytorch/utils.py:47-83: intensity centroid on a [0,1] grid,
    rescaled by (W-1)/(H-1) and clamped.  No softmax, no clipping of negative weights.
    """
    hm = torch.as_tensor(np.asarray(confmaps_nhwc)).float().permute(0, 3, 1, 2)
    _, _, h, w = hm.shape
    gy = torch.linspace(0, 1, steps=h).view(h, 1).expand(h, w)
    gx = torch.linspace(0, 1, steps=w).view(1, w).expand(h, w)
    total = hm.sum(dim=[2, 3])
    cx = torch.clamp((gx * hm).sum(dim=[2, 3]) / total * (w - 1), 0, w - 1)
    cy = torch.clamp((gy * hm).sum(dim=[2, 3]) / total * (h - 1), 0, h - 1)
    return torch.stack([cx, cy], dim=-1).numpy()


def gaussian_heatmap(mean_xy: Sequence[float], sigma: float = 3.0,
                     grid_size: Tuple[int, int] = (192, 192)) -> np.ndarray:
    """SimpleDataGenerator.get_gaussian, tensorflow/simple_data_generator.py:119-125.
    float64; indexed [y, x]; peak value 1, not normalised."""
    xs = np.arange(grid_size[0])[None, :]
    ys = np.arange(grid_size[1])[:, None]
    r2 = (xs - mean_xy[0]) ** 2 + (ys - mean_xy[1]) ** 2
    return np.exp(-(np.sqrt(r2) ** 2 / (2.0 * sigma ** 2)))


def gaussian_targets(points_xy: np.ndarray, sigma: float = 3.0, size: int = 192) -> np.ndarray:
    """ensure_sigma-style rendering (tensorflow/simple_data_generator.py:127-136) for a whole
    batch: (B,C,2) [x,y] -> (B,C,size,size) float32 (NCHW, the layout the trainer feeds the loss)."""
    pts = np.asarray(points_xy, dtype=np.float64)
    b, c, _ = pts.shape
    out = np.empty((b, c, size, size), dtype=np.float32)
    for i in range(b):
        for j in range(c):
            out[i, j] = gaussian_heatmap(pts[i, j], sigma, (size, size))
    return out


# --------------------------------------------------------------------------- #
# input-pipeline augmentation: DefaultDataset.augment_view,
# pytorch/Datagenerators.py:153-186  ->  torchvision F.affine (nearest, zero fill) + flips.
# The resampling arithmetic lives in torchvision 0.26 / torch 2.11 (not vendored by the
# reference): transforms/functional.py::_get_inverse_affine_matrix (python doubles),
# _functional_tensor.py::_gen_affine_grid (fp32 linspace base grid, bmm with theta^T / (W/2, H/2))
# and ATen grid_sampler_2d (nearest, zeros, align_corners=False:
# ix = ((g + 1) * W - 1) / 2, nearbyint, bounds test).  Restated here step by step in fp32.
# --------------------------------------------------------------------------- #
def inverse_affine_matrix(angle: float, translate: Sequence[float], scale: float) -> list:
    """torchvision _get_inverse_affine_matrix(center=[0,0], angle, translate, scale, shear=[0,0]) as
    F.affine calls it for tensors (pytorch/Datagenerators.py:170-173); python-double arithmetic in
    the same order, output pixel -> input pixel in centre-relative coordinates."""
    rot = math.radians(angle)
    a = math.cos(rot) / math.cos(0.0)
    b = -math.cos(rot) * math.tan(0.0) / math.cos(0.0) - math.sin(rot)
    c = math.sin(rot) / math.cos(0.0)
    d = -math.sin(rot) * math.tan(0.0) / math.cos(0.0) + math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m = [v / scale for v in m]
    tx, ty = float(translate[0]), float(translate[1])
    m[2] += m[0] * (-0.0 - tx) + m[1] * (-0.0 - ty)
    m[5] += m[3] * (-0.0 - tx) + m[4] * (-0.0 - ty)
    m[2] += 0.0
    m[5] += 0.0
    return m


def _fma32(a: np.ndarray, b, c: np.ndarray) -> np.ndarray:
    """fp32 fused multiply-add: the product of two fp32 values is exact in float64."""
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def affine_source_index(matrix: Sequence[float], h: int, w: int) -> Tuple[np.ndarray, np.ndarray]:
    """For every output pixel the flat source index (row*W+col) F.affine(nearest) samples and whether it
    is inside the image.  fp32 throughout, in the order torch's CPU kernels evaluate it."""
    f = np.float32
    th = np.asarray(matrix, dtype=np.float32)          # torch.tensor(matrix, dtype=float32)
    hw, hh = f(0.5 * w), f(0.5 * h)
    r = [th[0] / hw, th[1] / hw, th[2] / hw, th[3] / hh, th[4] / hh, th[5] / hh]   # theta^T / (W/2, H/2)
    bx = (np.arange(w, dtype=np.float32) + f(0.5 - 0.5 * w))[None, :].repeat(h, 0)  # linspace(-W/2+.5, W/2-.5, W)
    by = (np.arange(h, dtype=np.float32) + f(0.5 - 0.5 * h))[:, None].repeat(w, 1)
    gx = _fma32(by, r[1], bx * r[0]) + r[2]            # [x, y, 1] @ rescaled_theta, k = 0, 1, 2 in order
    gy = _fma32(by, r[4], bx * r[3]) + r[5]
    ix = ((gx + f(1)) * f(w) - f(1)) / f(2)            # grid_sampler_unnormalize, align_corners=False
    iy = ((gy + f(1)) * f(h) - f(1)) / f(2)
    xr, yr = np.rint(ix), np.rint(iy)                  # nearbyint: half to even
    ok = (xr >= 0) & (xr < w) & (yr >= 0) & (yr < h)
    src = np.where(ok, yr.astype(np.int64) * w + xr.astype(np.int64), 0)
    return src, ok


def affine_nearest(img: np.ndarray, matrix: Sequence[float], hflip: bool = False, vflip: bool = False) -> np.ndarray:
    """F.affine(img[C,H,W], ...) with the inverse matrix already computed, then F.hflip / F.vflip
    (pytorch/Datagenerators.py:170-182).  A pure gather: values are copied or zero."""
    img = np.asarray(img)
    c, h, w = img.shape
    src, ok = affine_source_index(matrix, h, w)
    out = np.where(ok[None], img.reshape(c, h * w)[:, src.reshape(-1)].reshape(c, h, w), 0).astype(img.dtype)
    if hflip:
        out = out[:, :, ::-1]
    if vflip:
        out = out[:, ::-1, :]
    return np.ascontiguousarray(out)


def draw_augmentation(config: dict, rng=np.random) -> dict:
    """The random draws of DefaultDataset.augment_view, pytorch/Datagenerators.py:154-169, in the
    reference's order (rotation, shift_y, shift_x, hflip coin, vflip coin, scaling) from numpy's legacy
    global generator (or any RandomState)."""
    rot_range, shifts = config["rotation range"], config["augmentation shift x y"]
    angle = rng.uniform(-rot_range, rot_range) if rot_range != 0 else 0
    if shifts != 0:
        shift_y = rng.uniform(-shifts, shifts)
        shift_x = rng.uniform(-shifts, shifts)
    else:
        shift_y = shift_x = 0
    hflip = bool(rng.rand() < 0.5 and bool(config["horizontal flip"]))
    vflip = bool(rng.rand() < 0.5 and bool(config["vertical flip"]))
    zoom = config["zoom range"]
    scale = rng.uniform(zoom[0], zoom[1])
    return {"angle": angle, "translate": (shift_x, shift_y), "scale": scale, "hflip": hflip, "vflip": vflip}


def augment_view(box: np.ndarray, confmaps: np.ndarray, config: dict, rng=np.random):
    """DefaultDataset.augment_view (pytorch/Datagenerators.py:153-186) on CHW arrays: one set of draws,
    applied to the crop and to its confidence maps."""
    p = draw_augmentation(config, rng)
    m = inverse_affine_matrix(p["angle"], p["translate"], p["scale"])
    return (affine_nearest(box, m, p["hflip"], p["vflip"]), affine_nearest(confmaps, m, p["hflip"], p["vflip"]), p)


def dataset_getitem(box_hwc: np.ndarray, confmaps_hwc: np.ndarray, config: dict, do_augmentations: bool,
                    rng=np.random):
    """DefaultDataset.__getitem__ for the single-view models (pytorch/Datagenerators.py:130-151):
    ToTensor (HWC -> CHW; uint8 -> /255), then augment_view through cast_as_float -- TWICE when
    do_augmentations is set (:144 and :149) and ONCE when it is not (:149 runs unconditionally, so the
    validation set is augmented too)."""
    def to_tensor(a):
        a = np.asarray(a)
        t = np.ascontiguousarray(np.moveaxis(a, -1, 0))
        return t.astype(np.float32) / np.float32(255) if a.dtype == np.uint8 else t
    b, c = to_tensor(box_hwc), to_tensor(confmaps_hwc)
    for _ in range(2 if do_augmentations else 1):
        b, c, _p = augment_view(b, c, config, rng)
    return b.astype(np.float32), c.astype(np.float32)


# --------------------------------------------------------------------------- #
# optimiser step (torch.optim.Adam defaults, pytorch/train_pytorch.py:111)
# --------------------------------------------------------------------------- #
def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int,
              lr: float = 1e-3, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """One torch.optim.Adam update (no weight decay, no amsgrad), returns (p, m, v)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    denom = v.sqrt() / math.sqrt(1 - b2 ** step) + eps
    return p - (lr / (1 - b1 ** step)) * m / denom, m, v


# --------------------------------------------------------------------------- #
# seeded parameter construction (same RNG stream as the reference constructors)
# --------------------------------------------------------------------------- #
def basicnet_state_dict(num_out: int = 36, filters: int = 64, cin: int = 4, seed: int = 0) -> StateDict:
    """Random-init parameters exactly as ``torch.manual_seed(seed); CNNs.BasicNet(cfg,(192,192,4),C)``
    would draw them: nn.Conv2d / nn.ConvTranspose2d default init, modules created in the
    order of pytorch/CNNs.py:25-43 (conv_i then bn_i) and :108-129.  BatchNorm init draws no
    random numbers, so only the conv order matters.  Checked against the reference by
    tests/golden (weight checksums)."""
    from torch import nn
    torch.manual_seed(seed)
    sd: StateDict = {}
    chans = [(cin, filters), (filters, filters), (filters, filters),
             (filters, 2 * filters), (2 * filters, 2 * filters), (2 * filters, 2 * filters),
             (2 * filters, 4 * filters), (4 * filters, 4 * filters), (4 * filters, 4 * filters)]
    for i, (ci, co) in enumerate(chans, 1):
        m = nn.Conv2d(ci, co, 3, padding=2, dilation=2)
        sd[f"encoder.conv{i}.weight"], sd[f"encoder.conv{i}.bias"] = m.weight.detach(), m.bias.detach()
    f4 = 4 * filters
    for i, (ci, co, s) in enumerate([(f4, f4 // 2, 2), (f4 // 2, f4 // 2, 1), (f4 // 2, f4 // 2, 1),
                                     (f4 // 2, num_out, 2)], 1):
        m = nn.ConvTranspose2d(ci, co, 3, stride=s, padding=1, output_padding=1 if s == 2 else 0)
        sd[f"decoder.conv2dTranspose{i}.weight"] = m.weight.detach()
        sd[f"decoder.conv2dTranspose{i}.bias"] = m.bias.detach()
    return sd


def four_cameras_state_dict(num_out: int = 72, filters: int = 64, seed: int = 0) -> StateDict:
    """Random-init parameters as ``torch.manual_seed(seed); CNNs.FourCamerasBaseLine(cfg,(H,W,16),C)`` draws them:
    shared_encoder (conv_i, bn_i), shared_conv2d (1x1, 4*256 -> 4*256), shared_decoder on 5*256 channels with C/4
    outputs -- pytorch/CNNs.py:201-218.  Checked against the reference by tests/golden/fourcam_c72.npz."""
    from torch import nn
    torch.manual_seed(seed)
    sd: StateDict = {}
    chans = [(4, filters), (filters, filters), (filters, filters),
             (filters, 2 * filters), (2 * filters, 2 * filters), (2 * filters, 2 * filters),
             (2 * filters, 4 * filters), (4 * filters, 4 * filters), (4 * filters, 4 * filters)]
    for i, (ci, co) in enumerate(chans, 1):
        m = nn.Conv2d(ci, co, 3, padding=2, dilation=2)
        sd[f"shared_encoder.conv{i}.weight"], sd[f"shared_encoder.conv{i}.bias"] = m.weight.detach(), m.bias.detach()
    w = 16 * filters
    m = nn.Conv2d(w, w, 1)
    sd["shared_conv2d.weight"], sd["shared_conv2d.bias"] = m.weight.detach(), m.bias.detach()
    cin = 20 * filters
    for i, (ci, co, s) in enumerate([(cin, cin // 2, 2), (cin // 2, cin // 2, 1), (cin // 2, cin // 2, 1),
                                     (cin // 2, num_out // 4, 2)], 1):
        m = nn.ConvTranspose2d(ci, co, 3, stride=s, padding=1, output_padding=1 if s == 2 else 0)
        sd[f"shared_decoder.conv2dTranspose{i}.weight"] = m.weight.detach()
        sd[f"shared_decoder.conv2dTranspose{i}.bias"] = m.bias.detach()
    return sd


def vit_state_dict(num_out: int = 36, dim: int = 256, heads: int = 12, depth: int = 8, dim_head: int = 256,
                   patch: int = 16, cin: int = 4, image: int = 192, seed: int = 0) -> StateDict:
    """Random-init parameters in the RNG order of
    ``torch.manual_seed(seed); VITs.VIT_encoder_CNN_decoder(cfg,(192,192,4),C)``:
    CustomViT.__init__ (pytorch/pytorch_vit_encoder.py:122-129: patch Linear, LayerNorm,
    pos_embedding randn, cls_token randn, then per layer Attention(:47-57: LN, to_qkv, to_out)
    and FeedForward(:18-25)), then the four deconvs (pytorch/VITs.py:23-34)."""
    from torch import nn
    torch.manual_seed(seed)
    sd: StateDict = {}
    p = "vit_encoder."
    n_patches = (image // patch) ** 2
    lin = nn.Linear(cin * patch * patch, dim)
    sd[p + "patch_to_embedding.weight"], sd[p + "patch_to_embedding.bias"] = lin.weight.detach(), lin.bias.detach()
    sd[p + "norm.weight"], sd[p + "norm.bias"] = torch.ones(dim), torch.zeros(dim)
    sd[p + "pos_embedding"] = torch.randn(1, n_patches, dim)
    sd[p + "cls_token"] = torch.randn(1, 1, dim)
    # Transformer.__init__: self.norm first (no RNG), then the layers (:89-96)
    sd[p + "transformer.norm.weight"], sd[p + "transformer.norm.bias"] = torch.ones(dim), torch.zeros(dim)
    inner = heads * dim_head
    for layer in range(depth):
        a = f"{p}transformer.layers.{layer}.0."
        f = f"{p}transformer.layers.{layer}.1."
        sd[a + "norm.weight"], sd[a + "norm.bias"] = torch.ones(dim), torch.zeros(dim)
        qkv = nn.Linear(dim, inner * 3, bias=False)
        sd[a + "to_qkv.weight"] = qkv.weight.detach()
        out = nn.Linear(inner, dim)
        sd[a + "to_out.0.weight"], sd[a + "to_out.0.bias"] = out.weight.detach(), out.bias.detach()
        sd[f + "net.0.weight"], sd[f + "net.0.bias"] = torch.ones(dim), torch.zeros(dim)
        l1 = nn.Linear(dim, 4 * dim)
        sd[f + "net.1.weight"], sd[f + "net.1.bias"] = l1.weight.detach(), l1.bias.detach()
        l2 = nn.Linear(4 * dim, dim)
        sd[f + "net.4.weight"], sd[f + "net.4.bias"] = l2.weight.detach(), l2.bias.detach()
    for i in range(1, 5):
        co = dim if i < 4 else num_out
        m = nn.ConvTranspose2d(dim, co, 3, stride=2, padding=1, output_padding=1)
        sd[f"cnn_decoder.deconv{i}.weight"], sd[f"cnn_decoder.deconv{i}.bias"] = m.weight.detach(), m.bias.detach()
    return sd


def _transformer_params(sd: StateDict, p: str, dim: int, depth: int, heads: int, dim_head: int, mlp_dim: int) -> None:
    """Transformer.__init__ (pytorch/pytorch_vit_encoder.py:82-96) in its RNG order: norm (no draws), then per layer
    Attention (LN, to_qkv, to_out) and FeedForward (LN, Linear, Linear)."""
    from torch import nn
    sd[p + "norm.weight"], sd[p + "norm.bias"] = torch.ones(dim), torch.zeros(dim)
    inner = heads * dim_head
    for layer in range(depth):
        a, f = f"{p}layers.{layer}.0.", f"{p}layers.{layer}.1."
        sd[a + "norm.weight"], sd[a + "norm.bias"] = torch.ones(dim), torch.zeros(dim)
        sd[a + "to_qkv.weight"] = nn.Linear(dim, inner * 3, bias=False).weight.detach()
        out = nn.Linear(inner, dim)
        sd[a + "to_out.0.weight"], sd[a + "to_out.0.bias"] = out.weight.detach(), out.bias.detach()
        sd[f + "net.0.weight"], sd[f + "net.0.bias"] = torch.ones(dim), torch.zeros(dim)
        l1 = nn.Linear(dim, mlp_dim)
        sd[f + "net.1.weight"], sd[f + "net.1.bias"] = l1.weight.detach(), l1.bias.detach()
        l2 = nn.Linear(mlp_dim, dim)
        sd[f + "net.4.weight"], sd[f + "net.4.bias"] = l2.weight.detach(), l2.bias.detach()


def vit_four_cameras_state_dict(num_out: int = 72, dim: int = 256, heads: int = 12, depth: int = 8, dim_head: int = 256,
                                patch: int = 16, image: int = 192, seed: int = 0) -> StateDict:
    """``torch.manual_seed(seed); VITs.VIT4CamerasBaseLine(cfg, (192,192,4), C)`` in its RNG order
    (pytorch/VITs.py:253-285): shared CustomViT, four CrossAttention blocks (Transformer(5*dim, depth 1, 4 heads,
    dim_head = mlp_dim = dim), LayerNorm, Linear(5*dim -> dim)), the shared CNN decoder with C/4 outputs."""
    from torch import nn
    torch.manual_seed(seed)
    sd: StateDict = {}
    p = "shared_vit_encoder."
    lin = nn.Linear(4 * patch * patch, dim)
    sd[p + "patch_to_embedding.weight"], sd[p + "patch_to_embedding.bias"] = lin.weight.detach(), lin.bias.detach()
    sd[p + "norm.weight"], sd[p + "norm.bias"] = torch.ones(dim), torch.zeros(dim)
    sd[p + "pos_embedding"] = torch.randn(1, (image // patch) ** 2, dim)
    sd[p + "cls_token"] = torch.randn(1, 1, dim)
    _transformer_params(sd, p + "transformer.", dim, depth, heads, dim_head, 4 * dim)
    for i in range(4):
        c = f"cross_attentions.{i}.layers."
        _transformer_params(sd, c + "0.", 5 * dim, 1, 4, dim, dim)
        sd[c + "1.weight"], sd[c + "1.bias"] = torch.ones(5 * dim), torch.zeros(5 * dim)
        lin = nn.Linear(5 * dim, dim)
        sd[c + "2.weight"], sd[c + "2.bias"] = lin.weight.detach(), lin.bias.detach()
    for i in range(1, 5):
        m = nn.ConvTranspose2d(dim, dim if i < 4 else num_out // 4, 3, stride=2, padding=1, output_padding=1)
        sd[f"shared_cnn_decoder.deconv{i}.weight"] = m.weight.detach()
        sd[f"shared_cnn_decoder.deconv{i}.bias"] = m.bias.detach()
    return sd


def four_cameras_disentanglement_state_dict(num_out: int = 72, filters: int = 64, seed: int = 0) -> StateDict:
    """``torch.manual_seed(seed); CNNs.FourCamerasDisentanglement(cfg, (H,W,16), C)`` in its RNG order
    (pytorch/CNNs.py:251-279): shared encoder, rearrange_layer_1 (256 -> 300), fusion layers (1600 -> 400 -> 400),
    three BatchNorms (no draws; running statistics at their initial 0 / 1), rearrange_layer_2 (300 -> 256), shared
    decoder on 256 channels with C/4 outputs."""
    from torch import nn
    torch.manual_seed(seed)
    sd: StateDict = {}
    chans = [(4, filters), (filters, filters), (filters, filters),
             (filters, 2 * filters), (2 * filters, 2 * filters), (2 * filters, 2 * filters),
             (2 * filters, 4 * filters), (4 * filters, 4 * filters), (4 * filters, 4 * filters)]
    for i, (ci, co) in enumerate(chans, 1):
        m = nn.Conv2d(ci, co, 3, padding=2, dilation=2)
        sd[f"shared_encoder.conv{i}.weight"], sd[f"shared_encoder.conv{i}.bias"] = m.weight.detach(), m.bias.detach()
    f4 = 4 * filters
    for name, ci, co in (("rearrange_layer_1", f4, 300), ("fusion_layer_1", 1600, 400), ("fusion_layer_2", 400, 400)):
        m = nn.Conv2d(ci, co, 1)
        sd[name + ".weight"], sd[name + ".bias"] = m.weight.detach(), m.bias.detach()
    for name, c in (("batch_norm1", 400), ("batch_norm2", 400), ("batch_norm3", 300)):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(c), torch.zeros(c)
        sd[name + ".running_mean"], sd[name + ".running_var"] = torch.zeros(c), torch.ones(c)
    m = nn.Conv2d(300, f4, 1)
    sd["rearrange_layer_2.weight"], sd["rearrange_layer_2.bias"] = m.weight.detach(), m.bias.detach()
    for i, (ci, co, st) in enumerate([(f4, f4 // 2, 2), (f4 // 2, f4 // 2, 1), (f4 // 2, f4 // 2, 1),
                                      (f4 // 2, num_out // 4, 2)], 1):
        m = nn.ConvTranspose2d(ci, co, 3, stride=st, padding=1, output_padding=1 if st == 2 else 0)
        sd[f"shared_decoder.conv2dTranspose{i}.weight"] = m.weight.detach()
        sd[f"shared_decoder.conv2dTranspose{i}.bias"] = m.bias.detach()
    return sd


# --------------------------------------------------------------------------- #
# synthetic workload
# --------------------------------------------------------------------------- #
def synthetic_crops(batch: int, seed: int = 1, cin: int = 4, size: int = 192) -> torch.Tensor:
    return torch.rand(batch, cin, size, size, generator=torch.Generator().manual_seed(seed))


def synthetic_points(batch: int, joints: int, seed: int = 2, size: int = 192) -> np.ndarray:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(8, size - 8, (batch, joints, 2), generator=g).numpy().astype(np.float32)


def train_step_reference(sd: StateDict, x: torch.Tensor, target: torch.Tensor, model: str = "cnn",
                         accumulation_steps: int = 1) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, torch.Tensor]]:
    """forward + MSE + backward with autograd on the functional restatement; returns
    (outputs, loss, grads-by-key) -- mirrors pytorch/train_pytorch.py:132-137 without AMP."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    fwd = {"cnn": basicnet_forward, "vit": vit_forward, "cnn4": four_cameras_baseline_forward}[model]
    out = fwd(leaves, x)
    loss = mse_loss(out, target, accumulation_steps)
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    return out.detach(), loss.detach(), grads
