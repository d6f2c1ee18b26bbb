"""Drop-in FLAME / FLAME_mediapipe / lbs for {gdl,inferno}.models.DecaFLAME and {gdl,inferno}.utils.lbs.

Same constructor config, registered buffer names, forward signature and return tuple as
third_party/inferno/inferno/models/DecaFLAME.py:50-106,222-297 and third_party/inferno/inferno/utils/lbs.py:142-234;
the arithmetic runs in libavi_b200.so (avi_flame_pack / avi_flame_lbs_fwd / avi_flame_landmarks).
Inference only (no autograd through the CUDA calls). There is no CPU path.
"""
from __future__ import annotations

import pickle

import numpy as np
import torch
import torch.nn as nn

from . import ops

_N_POSE_FEAT = 36


def default_precision() -> str:
    import os
    p = os.environ.get("AVI_B200_PRECISION", "bf16").lower()
    if p not in ("bf16", "fp32"):
        raise ValueError("AVI_B200_PRECISION must be bf16 or fp32")
    return p


def _k_pad(nb: int) -> int:
    return ((nb + _N_POSE_FEAT + 1 + 15) // 16) * 16


class _PackCache:
    """Packed static tensors, rebuilt when any source buffer is replaced or mutated in place
    (callers do mutate v_template in place: TalkingHeadWrapper.py:140-158, FaceFormerDecoder.py:1163-1181)."""

    def __init__(self):
        self.key = None
        self.dirs = self.jreg = self.dirs16 = None
        self.hoist = None   # (n_shape, dirs16 of the expression + pose rows, shape directions [V*3, n_shape], template [V*3])

    def get(self, shapedirs, posedirs, v_template, J_regressor):
        key = tuple((t.data_ptr(), t._version, t.device) for t in (shapedirs, posedirs, v_template, J_regressor))
        if key != self.key:
            nb = shapedirs.shape[2]
            self.dirs, self.jreg = ops.flame_pack(shapedirs.contiguous(), posedirs.contiguous(), v_template.contiguous(),
                                                  J_regressor.contiguous(), _k_pad(nb))
            self.dirs16 = ops.flame_pack_tc(self.dirs, shapedirs.shape[0], nb) if ops.flame_tc_supported(nb) else None
            self.hoist = None
            self.key = key
        return self.dirs, self.jreg


def _run(cache, precision, betas, full_pose, shapedirs, posedirs, v_template, J_regressor, lbs_weights, **kw):
    """precision 'bf16' -> tcgen05 blend (fp16 operands, ~1e-5 m); 'fp32' -> CUDA-core blend (exact to ~1e-7 m)."""
    dirs, jreg = cache.get(shapedirs, posedirs, v_template, J_regressor)
    V, nb = shapedirs.shape[0], shapedirs.shape[2]
    padded = kw.pop("padded", False)   # 16-byte aligned frame stride (vertices-only callers; the landmark gathers want dense frames)
    if precision == "bf16" and cache.dirs16 is not None:
        return ops.flame_lbs_tc(betas, full_pose, cache.dirs16, jreg, lbs_weights.contiguous(), v_template.contiguous(), V, nb,
                                _k_pad(nb), padded=padded, **kw)
    return ops.flame_lbs(betas, full_pose, dirs, jreg, lbs_weights.contiguous(), V, nb, _k_pad(nb), **kw)


_lbs_cache = _PackCache()


def _refuse_grad(*tensors):
    """The kernels are forward-only: a caller that differentiates through FLAME (MotionPrior.postprocess(with_grad=True),
    MotionPrior.py:345 - EMOTE training, not one of the BASELINE configs) must not silently get a detached mesh."""
    if torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors):
        raise NotImplementedError("avi_talking_b200 FLAME / lbs are forward-only (no backward kernel): detach the inputs or wrap the "
                                  "call in torch.no_grad()")


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, pose2rot=True,
        dtype=torch.float32, detach_pose_correctives=False, _cache=None):
    """lbs.py:142-234. v_template may be [V,3] or the reference's expanded [B,V,3] (row 0 is used: the reference expands one
    template over the batch, DecaFLAME.py:243). pose: [B, 15] axis-angle, or with pose2rot=False the rotation matrices themselves
    ([B, 5, 3, 3] or any shape with 45 values per frame, lbs.py:205-209). Returns (verts [B,V,3], J_transformed [B,5,3])."""
    _refuse_grad(betas, pose, v_template)
    if not pose2rot:
        pose = pose.reshape(pose.shape[0], -1)
        if pose.shape[1] != 45:
            raise ValueError(f"pose2rot=False expects 5 rotation matrices per frame (45 values), got {pose.shape[1]}")
    if J_regressor.shape[0] != 5 or [int(p) for p in parents] != [-1, 0, 1, 1, 1]:
        raise NotImplementedError("only the FLAME kinematic tree (5 joints, parents [-1,0,1,1,1]) is supported")
    vt = v_template[0] if v_template.dim() == 3 else v_template
    B = max(betas.shape[0], pose.shape[0])
    betas = betas.expand(B, -1).contiguous().float()
    pose = pose.expand(B, -1).contiguous().float()
    cache = _cache if _cache is not None else _lbs_cache
    verts, joints, _ = _run(cache, default_precision(), betas, pose, shapedirs, posedirs, vt, J_regressor, lbs_weights,
                            want_joints=True, rotmat=not pose2rot)
    return verts, joints


class _Cfg:
    pass


class FLAME(nn.Module):
    """Given FLAME parameters, outputs the mesh and the 2D/3D facial landmarks (DecaFLAME.py:44-269)."""

    def __init__(self, config):
        super().__init__()
        with open(config.flame_model_path, "rb") as fh:
            model = pickle.load(fh, encoding="latin1")
        get = (lambda k: model[k]) if isinstance(model, dict) else (lambda k: getattr(model, k))

        def arr(x, dt=np.float32):
            if "scipy.sparse" in str(type(x)):
                x = x.todense()
            return np.array(x, dtype=dt)

        self.cfg = config
        self.dtype = torch.float32
        self.register_buffer("faces_tensor", torch.from_numpy(arr(get("f"), np.int64)))
        self.register_buffer("v_template", torch.from_numpy(arr(get("v_template"))))
        sd = torch.from_numpy(arr(get("shapedirs")))
        self.register_buffer("shapedirs", torch.cat([sd[:, :, :config.n_shape], sd[:, :, 300:300 + config.n_exp]], 2))
        pd = arr(get("posedirs"))
        self.register_buffer("posedirs", torch.from_numpy(np.reshape(pd, [-1, pd.shape[-1]]).T.copy()))
        self.register_buffer("J_regressor", torch.from_numpy(arr(get("J_regressor"))))
        parents = torch.from_numpy(arr(get("kintree_table"), np.int64)[0].copy())
        parents[0] = -1
        self.register_buffer("parents", parents)
        self.register_buffer("lbs_weights", torch.from_numpy(arr(get("weights"))))
        self.register_parameter("eye_pose", nn.Parameter(torch.zeros(1, 6), requires_grad=False))
        self.register_parameter("neck_pose", nn.Parameter(torch.zeros(1, 3), requires_grad=False))
        emb = np.load(config.flame_lmk_embedding_path, allow_pickle=True, encoding="latin1")[()]
        self.register_buffer("lmk_faces_idx", torch.as_tensor(emb["static_lmk_faces_idx"], dtype=torch.long))
        self.register_buffer("lmk_bary_coords", torch.as_tensor(emb["static_lmk_bary_coords"], dtype=torch.float32))
        self.register_buffer("dynamic_lmk_faces_idx", torch.as_tensor(emb["dynamic_lmk_faces_idx"], dtype=torch.long))
        self.register_buffer("dynamic_lmk_bary_coords", torch.as_tensor(emb["dynamic_lmk_bary_coords"], dtype=torch.float32))
        self.register_buffer("full_lmk_faces_idx", torch.as_tensor(emb["full_lmk_faces_idx"], dtype=torch.long))
        self.register_buffer("full_lmk_bary_coords", torch.as_tensor(emb["full_lmk_bary_coords"], dtype=torch.float32))
        chain, cur = [], 1
        while cur != -1:
            chain.append(cur)
            cur = int(self.parents[cur])
        self.register_buffer("neck_kin_chain", torch.tensor(chain, dtype=torch.long))
        if chain != [1, 0]:
            raise NotImplementedError("unexpected neck kinematic chain")
        self._pack = _PackCache()
        self.precision = default_precision()

    # -- pieces -------------------------------------------------------------------------------------------------
    def _full_pose(self, batch_size, pose_params, eye_pose_params):
        if pose_params is None:
            pose_params = self.eye_pose.expand(batch_size, -1)
        if eye_pose_params is None:
            eye_pose_params = self.eye_pose.expand(batch_size, -1)
        return torch.cat([pose_params[:, :3], self.neck_pose.expand(batch_size, -1), pose_params[:, 3:], eye_pose_params], dim=1)

    def _run_lbs(self, shape_params, expression_params, pose_params, eye_pose_params, want_rows=True, padded=False):
        _refuse_grad(shape_params, expression_params, pose_params, eye_pose_params)
        B = shape_params.shape[0]
        if expression_params is None:
            expression_params = torch.zeros(B, self.cfg.n_exp, device=shape_params.device)
        betas = torch.cat([shape_params, expression_params], dim=1).contiguous().float()
        full_pose = self._full_pose(B, pose_params, eye_pose_params).contiguous().float()
        nb = self.shapedirs.shape[2]
        if betas.shape[1] != nb:
            raise ValueError(f"expected {nb} shape+expression coefficients, got {betas.shape[1]}")
        verts, _, rows = _run(self._pack, self.precision, betas, full_pose, self.shapedirs, self.posedirs, self.v_template,
                              self.J_regressor, self.lbs_weights, want_dyn_rows=want_rows, padded=padded)
        return verts, rows

    def _landmarks(self, verts, rows):
        B = verts.shape[0]
        rows = rows.long()
        idx = torch.cat([self.dynamic_lmk_faces_idx.index_select(0, rows), self.lmk_faces_idx[None].expand(B, -1)], 1).contiguous()
        bc = torch.cat([self.dynamic_lmk_bary_coords.index_select(0, rows), self.lmk_bary_coords[None].expand(B, -1, -1)],
                       1).contiguous()
        lmk2d = ops.flame_landmarks(verts, self.faces_tensor, idx, bc, per_frame=True)
        lmk3d = ops.flame_landmarks(verts, self.faces_tensor, self.full_lmk_faces_idx[0].contiguous(),
                                    self.full_lmk_bary_coords[0].contiguous(), per_frame=False)
        return lmk2d, lmk3d

    def seletec_3d68(self, vertices):  # (sic) name kept from DecaFLAME.py:216
        return ops.flame_landmarks(vertices.contiguous(), self.faces_tensor, self.full_lmk_faces_idx[0].contiguous(),
                                   self.full_lmk_bary_coords[0].contiguous(), per_frame=False)

    @torch.no_grad()
    def vertices_only(self, shape_params, expression_params=None, pose_params=None, eye_pose_params=None):
        """The mesh without the landmark gathers (what the audio->vertex hot path consumes)."""
        return self._run_lbs(shape_params, expression_params, pose_params, eye_pose_params, want_rows=False, padded=True)[0]

    @torch.no_grad()
    def vertices_sequence(self, shape, exp, pose):
        """Meshes of G clips x T frames that share one shape per clip (how FlamePreprocessor / BertPriorDecoder / convert_coeff2verts
        call FLAME: Preprocessors.py:136-150): shape [G, n_shape], exp [G, T, n_exp], pose [G, T, 6] -> [G, T, V*3].
        bf16 mode hoists the shape blendshapes out of the per-frame contraction: one small GEMM gives the G shaped templates and the
        tensor-core blend covers expression + pose correctives only (60 500 B written per frame, 4*(n_exp+6) B read)."""
        G, T = exp.shape[:2]
        V, nb = self.shapedirs.shape[0], self.shapedirs.shape[2]
        ns = shape.shape[1]
        pose = pose.reshape(G * T, -1).float()
        betas = torch.cat([shape[:, None].expand(G, T, ns), exp], dim=2).reshape(G * T, nb).contiguous().float()
        full_pose = self._full_pose(G * T, pose, None).contiguous().float()
        n_dirs = nb - ns + _N_POSE_FEAT
        if self.precision != "bf16" or n_dirs > 192:
            verts, _, _ = _run(self._pack, self.precision, betas, full_pose, self.shapedirs, self.posedirs, self.v_template,
                               self.J_regressor, self.lbs_weights)
            return verts.view(G, T, V * 3)
        dirs, jreg = self._pack.get(self.shapedirs, self.posedirs, self.v_template, self.J_regressor)
        if self._pack.hoist is None or self._pack.hoist[0] != ns:
            w_shape = self.shapedirs[:, :, :ns].reshape(V * 3, ns).contiguous().float()
            self._pack.hoist = (ns, ops.flame_pack_tc_rows(dirs, V, ns, n_dirs), w_shape, self.v_template.reshape(-1).contiguous().float())
        _, dirs16_exp, w_shape, tmpl = self._pack.hoist
        templates = ops.linear(shape.contiguous().float(), w_shape, tmpl)                     # [G, V*3] = v_template + S shape
        verts = ops.flame_lbs_tc_grouped(betas, full_pose, dirs16_exp, jreg, self.lbs_weights.contiguous(), templates, V, nb, ns,
                                         _k_pad(nb), T)
        return verts.view(G, T, V * 3)

    @torch.no_grad()
    def forward(self, shape_params=None, expression_params=None, pose_params=None, eye_pose_params=None):
        verts, rows = self._run_lbs(shape_params, expression_params, pose_params, eye_pose_params)
        lmk2d, lmk3d = self._landmarks(verts, rows)
        return verts, lmk2d, lmk3d


class FLAME_mediapipe(FLAME):
    """DecaFLAME.py:272-297: adds the 105 static mediapipe landmarks."""

    def __init__(self, config):
        super().__init__(config)
        emb = np.load(config.flame_mediapipe_lmk_embedding_path, allow_pickle=True, encoding="latin1")
        self.register_buffer("lmk_faces_idx_mediapipe", torch.as_tensor(emb["lmk_face_idx"].astype(np.int64), dtype=torch.long))
        self.register_buffer("lmk_bary_coords_mediapipe", torch.as_tensor(emb["lmk_b_coords"], dtype=torch.float32))

    @torch.no_grad()
    def forward(self, shape_params=None, expression_params=None, pose_params=None, eye_pose_params=None):
        verts, rows = self._run_lbs(shape_params, expression_params, pose_params, eye_pose_params)
        lmk2d, lmk3d = self._landmarks(verts, rows)
        lmk_mp = ops.flame_landmarks(verts, self.faces_tensor, self.lmk_faces_idx_mediapipe, self.lmk_bary_coords_mediapipe,
                                     per_frame=False)
        return verts, lmk2d, lmk3d, lmk_mp
