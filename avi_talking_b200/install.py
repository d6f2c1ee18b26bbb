"""Make the reference's own scripts run on the B200 path without editing them.

    import avi_talking_b200.install as avi
    avi.install()                      # BEFORE `import train_diffusion_prior` / `from models.faceformer_disentangle import Faceformer`

`install()` registers drop-in modules under the import names the reference uses, so that e.g.
`from inferno_apps.TalkingHead.evaluation.TalkingHeadWrapper import TalkingHeadWrapper` (train_diffusion_prior.py:11) or
`from models.diffusion_prior import InstructDiffusionPrior, VersatileDiffusionPriorNetwork, BrainNetwork, FrozenCLIPEmbedder`
(:10) resolve to the classes of this package. Two mechanisms, chosen per module:

  * REPLACE  - the reference module cannot even be imported without the heavy / private dependencies this path does not need
               (pytorch_lightning, omegaconf, pytorch3d, clip, dalle2_pytorch ...), or every public name of it is rebuilt here:
               a synthetic module object holding the drop-in names is put into `sys.modules` (parent packages are created as empty
               namespace packages when the reference tree is not on sys.path).
  * PATCH    - the reference module is importable and only SOME of its names are rebuilt (e.g. `lbs` inside `inferno.utils.lbs`, whose
               small helpers `batch_rodrigues`, `vertices2landmarks` ... stay upstream's): the named attributes are rebound in the real
               module, and in every already-imported module that had copied them with `from ... import name`.

`uninstall()` restores what was there. Nothing here computes; it only binds names.
"""
from __future__ import annotations

import importlib
import sys
import types

# reference import name -> (drop-in module of this package, {reference attribute: drop-in attribute})
TABLE = {
    "models.lib.wav2vec": ("avi_talking_b200.wav2vec", {"Wav2Vec2Model": "Wav2Vec2Model", "linear_interpolation": "linear_interpolation"}),
    "models.faceformer_disentangle": ("avi_talking_b200.faceformer", {
        "Faceformer": "Faceformer", "init_biased_mask": "init_biased_mask", "enc_dec_mask": "enc_dec_mask", "mask_lip": "mask_lip",
        "PeriodicPositionalEncoding": "PeriodicPositionalEncoding"}),
    "models.faceformer_vert": ("avi_talking_b200.faceformer", {
        "Faceformer": "FaceformerVert", "init_biased_mask": "init_biased_mask", "enc_dec_mask": "enc_dec_mask", "mask_lip": "mask_lip",
        "PeriodicPositionalEncoding": "PeriodicPositionalEncoding"}),
    "models.diffusion_prior": ("avi_talking_b200.diffusion_prior", {
        "BrainNetwork": "BrainNetwork", "VersatileDiffusionPriorNetwork": "VersatileDiffusionPriorNetwork",
        "InstructDiffusionPrior": "InstructDiffusionPrior", "FrozenCLIPEmbedder": "FrozenCLIPEmbedder"}),
    "inferno.models.DecaFLAME": ("avi_talking_b200.flame", {"FLAME": "FLAME", "FLAME_mediapipe": "FLAME_mediapipe"}),
    "gdl.models.DecaFLAME": ("avi_talking_b200.flame", {"FLAME": "FLAME", "FLAME_mediapipe": "FLAME_mediapipe"}),
    "inferno.utils.lbs": ("avi_talking_b200.flame", {"lbs": "lbs"}),
    "gdl.utils.lbs": ("avi_talking_b200.flame", {"lbs": "lbs"}),
    "inferno_apps.TalkingHead.evaluation.TalkingHeadWrapper": ("avi_talking_b200.talking_head", {"TalkingHeadWrapper": "TalkingHeadWrapper"}),
    "inferno.models.IO": ("avi_talking_b200.talking_head", {"locate_checkpoint": "locate_checkpoint"}),
    "inferno_apps.TalkingHead.utils.load": ("avi_talking_b200.talking_head", {"load_model": "load_model"}),
}
# modules whose remaining names must stay upstream's when the module is importable (PATCH); everything else is REPLACEd
PATCH_IF_IMPORTABLE = ("inferno.utils.lbs", "gdl.utils.lbs", "inferno.models.IO", "inferno.models.DecaFLAME", "gdl.models.DecaFLAME")

_saved: list = []          # undo log: (kind, module name / module object, attribute, old value)
_installed = False


def _ensure_parents(name: str) -> None:
    parts = name.split(".")
    for i in range(1, len(parts)):
        pkg = ".".join(parts[:i])
        if pkg in sys.modules:
            continue
        try:
            importlib.import_module(pkg)
        except Exception:  # noqa: BLE001 - reference tree absent (or its package __init__ needs missing deps): empty namespace package
            mod = types.ModuleType(pkg)
            mod.__path__ = []          # marks it as a package
            mod.__avi_b200_stub__ = True
            sys.modules[pkg] = mod
            _saved.append(("module", pkg, None, None))
    for i in range(1, len(parts)):     # bind children on their parents so `import a.b.c; a.b.c.X` works as well as `from a.b.c import X`
        parent, child = ".".join(parts[:i]), ".".join(parts[:i + 1])
        if child in sys.modules and not hasattr(sys.modules[parent], parts[i]):
            setattr(sys.modules[parent], parts[i], sys.modules[child])


def _rebind_copies(old, new) -> int:
    """Modules that did `from x import name` before install() hold their own reference to the old object: rebind those too."""
    n = 0
    for mod in list(sys.modules.values()):
        d = getattr(mod, "__dict__", None)
        if not isinstance(d, dict) or getattr(mod, "__name__", "").startswith("avi_talking_b200"):
            continue
        for k, v in list(d.items()):
            if v is old and v is not new:
                _saved.append(("attr", mod, k, old))
                d[k] = new
                n += 1
    return n


def install(verbose: bool = False) -> dict:
    """Bind every drop-in of TABLE under the reference's import names. Idempotent. Returns {reference module: 'replaced' | 'patched'}."""
    global _installed
    report = {}
    if _installed:
        return {k: "already installed" for k in TABLE}
    for ref_name, (our_name, names) in TABLE.items():
        ours = importlib.import_module(our_name)
        real = None
        if ref_name in PATCH_IF_IMPORTABLE:
            try:
                real = sys.modules.get(ref_name) or importlib.import_module(ref_name)
                if getattr(real, "__avi_b200_stub__", False):
                    real = None
            except Exception:  # noqa: BLE001
                real = None
        if real is not None:
            for ref_attr, our_attr in names.items():
                new = getattr(ours, our_attr)
                old = getattr(real, ref_attr, None)
                _saved.append(("attr", real, ref_attr, old))
                setattr(real, ref_attr, new)
                if old is not None:
                    _rebind_copies(old, new)
            report[ref_name] = "patched"
        else:
            mod = types.ModuleType(ref_name)
            mod.__avi_b200_stub__ = True
            mod.__doc__ = f"avi_talking_b200 drop-in for the reference module {ref_name} (names from {our_name})"
            for ref_attr, our_attr in names.items():
                setattr(mod, ref_attr, getattr(ours, our_attr))
            _saved.append(("module", ref_name, None, sys.modules.get(ref_name)))
            sys.modules[ref_name] = mod
            _ensure_parents(ref_name)
            setattr(sys.modules[ref_name.rsplit(".", 1)[0]], ref_name.rsplit(".", 1)[1], mod)
            report[ref_name] = "replaced"
        if verbose:
            print(f"avi_talking_b200.install: {ref_name} {report[ref_name]} ({', '.join(names)})")
    _installed = True
    return report


def uninstall() -> None:
    global _installed
    while _saved:
        kind, where, attr, old = _saved.pop()
        if kind == "module":
            if old is None:
                sys.modules.pop(where, None)
            else:
                sys.modules[where] = old
        elif old is None:
            try:
                delattr(where, attr)
            except AttributeError:
                pass
        else:
            setattr(where, attr, old)
    _installed = False
