"""TEST INFRASTRUCTURE (never imported by the product): mints tests/golden/reference_signatures.json - the parameter lists of every
class / method / function of the reference that the drop-in boundary mirrors, read from the reference SOURCES with
`ast` (the modules themselves do not import here: easydict, omegaconf, pytorch_lightning, clip, dalle2_pytorch ... are absent).

    python -m oracle.make_signatures            # needs /root/reference (build container only)

tests/test_boundary.py compares `inspect.signature` of the drop-ins with this fixture, and re-derives the fixture from
/root/reference when that tree is present, so a stale fixture fails here.
"""
from __future__ import annotations

import ast
import json
import os

REF = os.environ.get("AVI_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_signatures.json")

INFERNO = "third_party/inferno"
# (reference file, {class or "" for module level: [function names]}, drop-in module, {reference class name: drop-in class name})
TARGETS = [
    ("models/lib/wav2vec.py", {"": ["linear_interpolation"], "Wav2Vec2Model": ["forward"]}),
    ("models/faceformer_disentangle.py", {"": ["init_biased_mask", "enc_dec_mask", "mask_lip"],
                                          "PeriodicPositionalEncoding": ["__init__", "forward"],
                                          "Faceformer": ["__init__", "forward", "predict", "forward_ff", "convert_coeff2verts"]}),
    ("models/faceformer_vert.py", {"Faceformer": ["__init__", "forward", "predict", "forward_ff", "convert_coeff2verts"]}),
    (INFERNO + "/inferno/models/DecaFLAME.py", {"FLAME": ["__init__", "forward"], "FLAME_mediapipe": ["__init__", "forward"]}),
    (INFERNO + "/inferno/utils/lbs.py", {"": ["lbs", "batch_rodrigues", "vertices2landmarks", "blend_shapes", "vertices2joints",
                                              "batch_rigid_transform", "transform_mat", "rot_mat_to_euler", "find_dynamic_lmk_idx_and_bcoords"]}),
    (INFERNO + "/inferno_apps/TalkingHead/evaluation/TalkingHeadWrapper.py",
     {"TalkingHeadWrapper": ["__init__", "forward", "get_num_intensities", "get_num_emotions", "get_num_identities", "get_subject_labels",
                             "set_neutral_mesh"]}),
    (INFERNO + "/inferno/models/IO.py", {"": ["locate_checkpoint"]}),
    (INFERNO + "/inferno_apps/TalkingHead/utils/load.py", {"": ["load_model"]}),
    ("models/diffusion_prior.py", {"FrozenCLIPEmbedder": ["__init__", "forward", "encode"],
                                   "BrainNetwork": ["__init__", "forward"],
                                   "VersatileDiffusionPriorNetwork": ["__init__", "forward", "forward_with_cond_scale"],
                                   "InstructDiffusionPrior": ["__init__", "p_sample", "p_sample_loop_ddpm", "p_losses", "forward"]}),
    ("train_diffusion_prior.py", {"": ["voxel2style_emb", "soft_clip_loss"]}),
]


def _params(fn: ast.FunctionDef) -> list:
    a = fn.args
    pos = [x.arg for x in a.posonlyargs + a.args]
    defaults = [None] * (len(pos) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    out = [{"name": n, "default": d} for n, d in zip(pos, defaults)]
    if a.vararg:
        out.append({"name": "*" + a.vararg.arg, "default": None})
    for x, d in zip(a.kwonlyargs, a.kw_defaults):
        out.append({"name": x.arg, "default": None if d is None else ast.unparse(d), "kwonly": True})
    if a.kwarg:
        out.append({"name": "**" + a.kwarg.arg, "default": None})
    return out


def extract(root: str = REF) -> dict:
    sigs = {}
    for path, wanted in TARGETS:
        with open(os.path.join(root, path)) as fh:
            tree = ast.parse(fh.read())
        found = {}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in wanted.get("", []):
                found[node.name] = {"line": node.lineno, "params": _params(node)}
            elif isinstance(node, ast.ClassDef) and node.name in wanted:
                for sub in node.body:
                    if isinstance(sub, ast.FunctionDef) and sub.name in wanted[node.name]:
                        found[f"{node.name}.{sub.name}"] = {"line": sub.lineno, "params": _params(sub)}
        missing = [f"{c}.{m}" if c else m for c, ms in wanted.items() for m in ms if (f"{c}.{m}" if c else m) not in found]
        # methods a class inherits (e.g. FLAME_mediapipe.__init__ written out, or not) are simply absent from the source: record that
        sigs[path] = {"functions": found, "absent_in_source": missing}
    return sigs


def main():
    sigs = extract()
    with open(OUT, "w") as fh:
        json.dump(sigs, fh, indent=1, sort_keys=True)
    n = sum(len(v["functions"]) for v in sigs.values())
    print(f"{OUT}: {n} signatures from {len(sigs)} reference files")
    for k, v in sigs.items():
        if v["absent_in_source"]:
            print("  not in source:", k, v["absent_in_source"])


if __name__ == "__main__":
    main()
