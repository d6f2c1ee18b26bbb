"""Import the reference's INFERNO modules (third_party/inferno) in the build container - TEST INFRASTRUCTURE, used only by
oracle/make_golden.py. The modules need packages that are absent here (pytorch_lightning, omegaconf, munch, pytorch3d,
matplotlib, skimage, ...); none of them takes part in the arithmetic of the EMOTE inference path, so they are replaced by
stubs: a dict-with-attributes for Munch / DictConfig, nn.Module for pl.LightningModule, MagicMock for the rest."""
from __future__ import annotations

import contextlib
import importlib
import sys
import types
from unittest.mock import MagicMock

import torch.nn as nn

REF = "/root/reference"


class Munch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def munchify(x):
    if isinstance(x, dict):
        return Munch({k: munchify(v) for k, v in x.items()})
    if isinstance(x, (list, tuple)):
        return type(x)(munchify(v) for v in x)
    return x


def install_stubs():
    # transformers probes optional packages with importlib.util.find_spec, which chokes on spec-less stubs: import what the
    # reference needs from it BEFORE the stubs go in
    from transformers import Wav2Vec2Model, Wav2Vec2Processor  # noqa: F401
    for p in (REF, REF + "/third_party/inferno"):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "munch" not in sys.modules:
        m = types.ModuleType("munch")
        m.Munch, m.munchify = Munch, munchify
        sys.modules["munch"] = m
    if "omegaconf" not in sys.modules or isinstance(sys.modules["omegaconf"], MagicMock):
        oc = types.ModuleType("omegaconf")

        class OmegaConf:
            @staticmethod
            def to_container(c, **k):
                return {k2: (OmegaConf.to_container(v) if isinstance(v, dict) else v) for k2, v in dict(c).items()}

        class DictConfig(Munch):
            pass

        @contextlib.contextmanager
        def open_dict(c):
            yield c

        oc.OmegaConf, oc.DictConfig, oc.open_dict = OmegaConf, DictConfig, open_dict
        sys.modules["omegaconf"] = oc
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = nn.Module
        pl.LightningDataModule = object
        sys.modules["pytorch_lightning"] = pl
    for s in ["pytorch3d", "pytorch3d.transforms", "pytorch3d.io", "pytorch3d.structures", "pytorch3d.renderer",
              "pytorch3d.renderer.mesh", "skimage", "skimage.io", "skimage.transform", "kornia", "hydra", "hydra.experimental",
              "pytorch_lightning.loggers", "pytorch_lightning.callbacks", "wandb", "chumpy", "matplotlib", "matplotlib.pyplot",
              "matplotlib.cm", "matplotlib.colors", "seaborn", "face_alignment", "adabound",
              "compress_pickle", "imgaug", "skvideo", "skvideo.io"]:
        if s not in sys.modules:
            sys.modules[s] = MagicMock(name=s)


def load(*names):
    """Import reference modules; any further missing third-party package met on the way is stubbed too (it can only be one
    the EMOTE inference arithmetic does not touch: a MagicMock result would break the golden comparison loudly)."""
    install_stubs()
    out = []
    for n in names:
        for _ in range(64):
            try:
                out.append(importlib.import_module(n))
                break
            except ModuleNotFoundError as e:
                if e.name is None or e.name.startswith("inferno"):
                    raise
                sys.modules[e.name] = MagicMock(name=e.name)
        else:
            raise ImportError(n)
    return out
