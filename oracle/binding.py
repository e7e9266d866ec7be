"""Torch restatement of GaussianAvatars' mesh binding (the eager PyTorch path the fused kernel replaces).

TEST INFRASTRUCTURE (oracle).  Runs on CPU tensors in float32 or float64 and is differentiable through
autograd exactly like the reference, which has no hand-written backward for this part.

Follows, line by line in behaviour:
  - per-face frame            scene/flame_gaussian_model.py:137-147, utils/graphics_utils.py:90-135
  - splat getters             scene/gaussian_model.py:113-160 (activations :29-44: exp / sigmoid / normalize)
  - roma helpers (roma is not installed here): quaternions are xyzw,
    quat_product = Hamilton product, rotmat_to_unitquat = SciPy's largest-of-(diag, trace) branch + normalise.
These are pinned against the real reference modules by tests/golden/make_golden.py (which imports
/root/reference with a roma shim built from THIS file for the three roma calls -- the reference's own
arithmetic around them is real).
"""
from __future__ import annotations

import torch


# ----- utils/graphics_utils.py:90-100 ------------------------------------------------------------
def dot(x, y):
    return torch.sum(x * y, -1, keepdim=True)


def length(x, eps: float = 1e-20):
    return torch.sqrt(torch.clamp(dot(x, x), min=eps))


def safe_normalize(x, eps: float = 1e-20):
    return x / length(x, eps)


# ----- utils/graphics_utils.py:116-135 -----------------------------------------------------------
def compute_face_orientation(verts, faces):
    """verts (V,3), faces (F,3) long -> orientation (F,3,3) with columns [a0 a1 a2], scale (F,1)."""
    i0, i1, i2 = faces[..., 0].long(), faces[..., 1].long(), faces[..., 2].long()
    v0, v1, v2 = verts[..., i0, :], verts[..., i1, :], verts[..., i2, :]
    a0 = safe_normalize(v1 - v0)
    a1 = safe_normalize(torch.cross(a0, v2 - v0, dim=-1))
    a2 = -safe_normalize(torch.cross(a1, a0, dim=-1))
    orientation = torch.cat([a0[..., None], a1[..., None], a2[..., None]], dim=-1)
    s0 = length(v1 - v0)
    s1 = dot(a2, (v2 - v0)).abs()
    scale = (s0 + s1) / 2
    return orientation, scale


# ----- roma semantics (xyzw) ----------------------------------------------------------------------
def quat_xyzw_to_wxyz(q):
    return torch.cat((q[..., 3:4], q[..., 0:3]), dim=-1)


def quat_wxyz_to_xyzw(q):
    return torch.cat((q[..., 1:4], q[..., 0:1]), dim=-1)


def quat_product(p, q):
    """Hamilton product, xyzw."""
    vector = p[..., None, 3] * q[..., :3] + q[..., None, 3] * p[..., :3] + torch.cross(p[..., :3], q[..., :3], dim=-1)
    last = p[..., 3] * q[..., 3] - torch.sum(p[..., :3] * q[..., :3], dim=-1)
    return torch.cat((vector, last[..., None]), dim=-1)


def rotmat_to_unitquat(R):
    """(N,3,3) -> (N,4) xyzw, sign not canonicalised."""
    m = R.reshape(-1, 3, 3)
    diag = torch.stack((m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]), dim=1)
    trace = diag.sum(dim=1)
    decision = torch.cat((diag, trace[:, None]), dim=1)
    choice = decision.argmax(dim=1)
    quats = []
    for i in range(3):
        j, k = (i + 1) % 3, (i + 2) % 3
        comp = [None] * 4
        comp[i] = 1 - trace + 2 * m[:, i, i]
        comp[j] = m[:, j, i] + m[:, i, j]
        comp[k] = m[:, k, i] + m[:, i, k]
        comp[3] = m[:, k, j] - m[:, j, k]
        quats.append(torch.stack(comp, dim=1))
    quats.append(torch.stack((m[:, 2, 1] - m[:, 1, 2], m[:, 0, 2] - m[:, 2, 0], m[:, 1, 0] - m[:, 0, 1], 1 + trace),
                             dim=1))
    q = torch.zeros_like(quats[0])
    for c in range(4):
        q = torch.where((choice == c)[:, None], quats[c], q)
    q = q / torch.norm(q, dim=1)[:, None]
    return q.reshape(R.shape[:-2] + (4,))


# ----- scene/flame_gaussian_model.py:137-147 ------------------------------------------------------
def update_mesh_properties(verts, faces):
    """verts (1,V,3) or (V,3); faces (F,3).  Returns dict of the per-face frame tensors."""
    v = verts.reshape(-1, 3)
    triangles = v[faces]  # (F,3,3)
    face_center = triangles.mean(dim=-2)
    face_orien_mat, face_scaling = compute_face_orientation(v, faces)
    face_orien_quat = quat_xyzw_to_wxyz(rotmat_to_unitquat(face_orien_mat))
    return dict(face_center=face_center, face_orien_mat=face_orien_mat, face_scaling=face_scaling,
                face_orien_quat=face_orien_quat)


# ----- scene/gaussian_model.py:113-160 ------------------------------------------------------------
def get_scaling(_scaling, binding=None, face_scaling=None):
    s = torch.exp(_scaling)
    return s if binding is None else s * face_scaling[binding]


def get_rotation(_rotation, binding=None, face_orien_quat=None):
    rot = torch.nn.functional.normalize(_rotation)
    if binding is None:
        return rot
    fq = torch.nn.functional.normalize(face_orien_quat[binding])
    return quat_xyzw_to_wxyz(quat_product(quat_wxyz_to_xyzw(fq), quat_wxyz_to_xyzw(rot)))


def get_xyz(_xyz, binding=None, face_center=None, face_orien_mat=None, face_scaling=None):
    if binding is None:
        return _xyz
    xyz = torch.bmm(face_orien_mat[binding], _xyz[..., None]).squeeze(-1)
    return xyz * face_scaling[binding] + face_center[binding]


def get_opacity(_opacity):
    return torch.sigmoid(_opacity)


def get_features(features_dc, features_rest):
    return torch.cat((features_dc, features_rest), dim=1)
