"""Oracle: FLAME blendshapes + linear blend skinning (TEST INFRASTRUCTURE; see oracle/__init__.py).

fp32 torch-on-CPU restatement of
  third_party/inferno/inferno/utils/lbs.py      lbs :142-234, blend_shapes :280-301,
      vertices2joints :260-277, batch_rodrigues :304-335, transform_mat :338-348,
      batch_rigid_transform :351-408, vertices2landmarks :103-139, rot_mat_to_euler
  third_party/inferno/inferno/models/DecaFLAME.py  FLAME.forward :222-269,
      _find_dynamic_lmk_idx_and_bcoords :110-149, FLAME_mediapipe.forward :285-297
(the BlendshapeVisualizer/EMOCA/gdl copies are byte-identical in the arithmetic).
Pinned by tests/golden/flame_*.npz (minted from the reference classes by oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def batch_rodrigues(rot_vecs: torch.Tensor) -> torch.Tensor:
    # lbs.py:304-335 ; note the +1e-8 inside the norm
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.cos(angle)[:, None]
    sin = torch.sin(angle)[:, None]
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=rot_vecs.dtype)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(n, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype)[None]
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def batch_rigid_transform(rot_mats, joints, parents):
    # lbs.py:351-408
    joints = joints[..., None]
    rel = joints.clone()
    rel[:, 1:] -= joints[:, parents[1:]]
    nj = joints.shape[1]
    tm = torch.cat([F.pad(rot_mats.reshape(-1, 3, 3), [0, 0, 0, 1]),
                    F.pad(rel.reshape(-1, 3, 1), [0, 0, 0, 1], value=1)], dim=2).reshape(-1, nj, 4, 4)
    chain = [tm[:, 0]]
    for i in range(1, nj):
        chain.append(torch.matmul(chain[int(parents[i])], tm[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed = transforms[:, :, :3, 3]
    jh = F.pad(joints, [0, 0, 0, 1])
    rel_tf = transforms - F.pad(torch.matmul(transforms, jh), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed, rel_tf


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, pose2rot=True):
    """lbs.py:142-234. v_template [B,V,3] or [V,3]; pose2rot=False: pose holds the rotation matrices (:205-209)."""
    B = max(betas.shape[0], pose.shape[0])
    if v_template.dim() == 2:
        v_template = v_template[None].expand(B, -1, -1)
    v_shaped = v_template + torch.einsum("bl,mkl->bmk", betas, shapedirs)          # :188
    J = torch.einsum("bik,ji->bjk", v_shaped, J_regressor)                           # :192
    rot = batch_rodrigues(pose.reshape(-1, 3)).view(B, -1, 3, 3) if pose2rot else pose.reshape(B, -1, 3, 3)   # :198 / :207
    pose_feature = (rot[:, 1:] - torch.eye(3)).reshape(B, -1)                        # :201
    v_posed = v_shaped + torch.matmul(pose_feature, posedirs).view(B, -1, 3)         # :203,215
    J_tf, A = batch_rigid_transform(rot, J, parents)                                 # :217
    nj = J_regressor.shape[0]
    T = torch.matmul(lbs_weights[None].expand(B, -1, -1), A.view(B, nj, 16)).view(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1)], dim=2)
    verts = torch.matmul(T, vh[..., None])[:, :, :3, 0]                              # :230-232
    return verts, J_tf


def vertices2landmarks(vertices, faces, lmk_faces_idx, lmk_bary_coords):
    # lbs.py:103-139
    B, V = vertices.shape[:2]
    lmk_faces = torch.index_select(faces, 0, lmk_faces_idx.reshape(-1)).view(B, -1, 3)
    lmk_faces = lmk_faces + torch.arange(B, dtype=torch.long).view(-1, 1, 1) * V
    lv = vertices.reshape(-1, 3)[lmk_faces].view(B, -1, 3, 3)
    return torch.einsum("blfi,blf->bli", lv, lmk_bary_coords)


def rot_mat_to_euler(rot_mats):
    # lbs.py:28-36
    sy = torch.sqrt(rot_mats[:, 0, 0] * rot_mats[:, 0, 0] + rot_mats[:, 1, 0] * rot_mats[:, 1, 0])
    return torch.atan2(-rot_mats[:, 2, 0], sy)


def dynamic_lmk_rows(full_pose, neck_kin_chain):
    """DecaFLAME.py:110-149: which of the 79 contour rows each frame uses."""
    B = full_pose.shape[0]
    aa = torch.index_select(full_pose.view(B, -1, 3), 1, neck_kin_chain)
    rm = batch_rodrigues(aa.reshape(-1, 3)).view(B, -1, 3, 3)
    rel = torch.eye(3)[None].expand(B, -1, -1)
    for idx in range(len(neck_kin_chain)):
        rel = torch.bmm(rm[:, idx], rel)
    y = torch.round(torch.clamp(rot_mat_to_euler(rel) * 180.0 / np.pi, max=39)).to(torch.long)
    neg = y.lt(0).to(torch.long)
    mask = y.lt(-39).to(torch.long)
    neg_vals = mask * 78 + (1 - mask) * (39 - y)
    return neg * neg_vals + (1 - neg) * y


def flame_forward(buf: dict, shape_params, expression_params=None, pose_params=None, eye_pose_params=None,
                  mediapipe: bool = False):
    """FLAME.forward (DecaFLAME.py:222-269) / FLAME_mediapipe.forward (:285-297) on a buffer dict."""
    B = shape_params.shape[0]
    if pose_params is None:
        pose_params = buf["eye_pose"].expand(B, -1)
    if eye_pose_params is None:
        eye_pose_params = buf["eye_pose"].expand(B, -1)
    if expression_params is None:
        expression_params = torch.zeros(B, buf["shapedirs"].shape[2] - shape_params.shape[1])
    betas = torch.cat([shape_params, expression_params], dim=1)
    full_pose = torch.cat([pose_params[:, :3], buf["neck_pose"].expand(B, -1), pose_params[:, 3:],
                           eye_pose_params], dim=1)
    verts, _ = lbs(betas, full_pose, buf["v_template"], buf["shapedirs"], buf["posedirs"],
                   buf["J_regressor"], buf["parents"], buf["lbs_weights"])
    rows = dynamic_lmk_rows(full_pose, buf["neck_kin_chain"])
    dyn_idx = torch.index_select(buf["dynamic_lmk_faces_idx"], 0, rows)
    dyn_bc = torch.index_select(buf["dynamic_lmk_bary_coords"], 0, rows)
    idx = torch.cat([dyn_idx, buf["lmk_faces_idx"][None].expand(B, -1)], 1)
    bc = torch.cat([dyn_bc, buf["lmk_bary_coords"][None].expand(B, -1, -1)], 1)
    lmk2d = vertices2landmarks(verts, buf["faces_tensor"], idx, bc)
    lmk3d = vertices2landmarks(verts, buf["faces_tensor"], buf["full_lmk_faces_idx"].repeat(B, 1),
                               buf["full_lmk_bary_coords"].repeat(B, 1, 1))
    if not mediapipe:
        return verts, lmk2d, lmk3d
    lmk_mp = vertices2landmarks(verts, buf["faces_tensor"],
                                buf["lmk_faces_idx_mediapipe"][None].expand(B, -1).contiguous(),
                                buf["lmk_bary_coords_mediapipe"][None].expand(B, -1, -1).contiguous())
    return verts, lmk2d, lmk3d, lmk_mp


def convert_coeff2verts(buf, coeff_mean, coeff_std, gt_coeff, gt_pose, gt_shape):
    """models/faceformer_disentangle.py:425-433 (mutates gt_pose[..., :3] in place, as upstream does)."""
    c = gt_coeff * coeff_std[: gt_coeff.shape[-1]] + coeff_mean[: gt_coeff.shape[-1]]
    gt_pose[..., :3] = 0.0
    return flame_forward(buf, gt_shape, c[:, :50], gt_pose, mediapipe=True)[0]
