"""Import shim for the *real* reference modules -- build container only.

TEST INFRASTRUCTURE.  /root/reference is mounted read-only in the build container
and does not exist on the GPU box, so nothing under ``tests -m gpu``, ``smoke()``
or ``bench.py`` may import this file; it is used by ``oracle/make_golden.py``
(to generate tests/golden/*.npz) and by the CPU tests that are skipped when the
mount is absent.

The reference imports ``torchsummary`` at module scope (pytorch/CNNs.py:2,
VITs.py:4, pytorch_vit_encoder.py:3, Network.py:4) but only uses it in __main__
blocks / Network.get_model, so a stub is installed.  pytorch/utils.py imports
tensorflow (:1) which is absent; ``find_peaks_soft_argmax`` (pure torch, :47-83)
and ``SimpleDataGenerator.get_gaussian`` (pure numpy,
tensorflow/simple_data_generator.py:119-125) are therefore compiled from their
source text in isolation.
"""
from __future__ import annotations

import ast
import json
import os
import sys
import types

REF_ROOT = os.environ.get("POSE_REFERENCE_ROOT", "/root/reference")
REF_PT = os.path.join(REF_ROOT, "pytorch")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_PT, "CNNs.py"))


def _install_stubs() -> None:
    if "torchsummary" not in sys.modules:
        stub = types.ModuleType("torchsummary")
        stub.summary = lambda *a, **k: None
        sys.modules["torchsummary"] = stub
    if REF_PT not in sys.path:
        sys.path.insert(0, REF_PT)


def load_modules():
    """returns (CNNs, VITs, Augmentor) reference modules."""
    _install_stubs()
    import CNNs  # type: ignore
    import VITs  # type: ignore
    import Augmentor  # type: ignore
    return CNNs, VITs, Augmentor


def load_config(model_type: str | None = None) -> dict:
    with open(os.path.join(REF_PT, "train_config.json")) as fh:
        cfg = json.load(fh)
    if model_type is not None:
        cfg["model type"] = model_type
    return cfg


def _function_from_source(path: str, name: str, namespace: dict):
    """Compile one (possibly nested-in-class) function out of a source file without
    executing the file's imports."""
    with open(path) as fh:
        tree = ast.parse(fh.read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, path, "exec"), namespace)
            return namespace[name]
    raise KeyError(name)


def soft_argmax_fn():
    import numpy as np
    import torch
    return _function_from_source(os.path.join(REF_PT, "utils.py"), "find_peaks_soft_argmax",
                                 {"np": np, "torch": torch})


def gaussian_fn():
    import numpy as np
    return _function_from_source(os.path.join(REF_ROOT, "tensorflow", "simple_data_generator.py"),
                                 "get_gaussian", {"np": np})


def default_dataset_cls():
    """DefaultDataset of pytorch/Datagenerators.py:115-186, rebuilt from its own method sources
    (__init__, __len__, __getitem__, cast_as_float, augment_view): the module itself does not
    import here (h5py, matplotlib absent).  Namespace = what those methods touch."""
    import numpy as np
    import torch
    import torchvision.transforms.functional as F
    from torchvision import transforms
    _install_stubs()
    import constants  # type: ignore
    ns = {"np": np, "torch": torch, "F": F, "transforms": transforms}
    ns.update({k: v for k, v in vars(constants).items() if k.isupper()})
    path = os.path.join(REF_PT, "Datagenerators.py")
    with open(path) as fh:
        tree = ast.parse(fh.read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DefaultDataset")
    keep = {"__init__", "__len__", "__getitem__", "cast_as_float", "augment_view"}
    cls.body = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in keep]
    cls.bases = []
    mod = ast.Module(body=[cls], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, path, "exec"), ns)
    return ns["DefaultDataset"]
