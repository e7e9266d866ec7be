"""Generates tests/golden/reference_vectors.npz by IMPORTING the real reference (/root/reference, read-only) in this
container.  The fixtures travel to the GPU box; /root/reference does not.

What is real reference code here and what is not:
  REAL  utils/sh_utils.eval_sh, utils/graphics_utils.{compute_face_orientation,getProjectionMatrix,getWorld2View2},
        utils/general_utils.{build_rotation,build_scaling_rotation,strip_symmetric} (device="cuda" literals are
        neutralised by dropping the `device` kwarg of torch.zeros for the duration of the call),
        scene/gaussian_model.GaussianModel getters (get_xyz/get_rotation/get_scaling/get_opacity/get_features/
        get_covariance), scene/flame_gaussian_model.FlameGaussianModel.update_mesh_properties,
        scene/cameras.Camera, utils/viewer_utils.OrbitCamera (+ fps_benchmark_demo.prepare_camera's transposes).
  SHIM  roma (not installed): quat_product / rotmat_to_unitquat / quat_xyzw_to_wxyz / quat_wxyz_to_xyzw come from
        oracle/binding.py (SURVEY.md Appendix C semantics); plyfile, simple_knn, dearpygui, matplotlib are empty stubs
        (never called on these paths).
The rasterizer proper (submodules/diff-gaussian-rasterization) is absent from /root/reference: no vector for it can
be generated -- parity for it stays "unpinned" (see oracle/splat_oracle.c header).

    python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import binding as ob  # noqa: E402

# ---- stubs / shims ---------------------------------------------------------------------------------------------
roma = types.ModuleType("roma")
roma.quat_product = ob.quat_product
roma.quat_xyzw_to_wxyz = ob.quat_xyzw_to_wxyz
roma.quat_wxyz_to_xyzw = ob.quat_wxyz_to_xyzw
roma.rotmat_to_unitquat = ob.rotmat_to_unitquat
sys.modules["roma"] = roma
for name in ("plyfile", "simple_knn", "simple_knn._C", "dearpygui", "dearpygui.dearpygui", "matplotlib",
             "matplotlib.pyplot", "diff_gaussian_rasterization", "nvdiffrast", "nvdiffrast.torch", "pytorch3d",
             "pytorch3d.io", "iopath", "iopath.common", "iopath.common.file_io", "chumpy", "lpips", "tqdm_unused"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["plyfile"].PlyData = sys.modules["plyfile"].PlyElement = object
sys.modules["simple_knn._C"].distCUDA2 = None
sys.modules["pytorch3d.io"].load_obj = None
sys.modules["iopath.common.file_io"].PathManager = object
sys.modules["diff_gaussian_rasterization"].GaussianRasterizationSettings = object
sys.modules["diff_gaussian_rasterization"].GaussianRasterizer = object

_zeros = torch.zeros


def _zeros_cpu(*a, **k):
    k.pop("device", None)
    return _zeros(*a, **k)


def on_cpu(fn, *a, **k):
    torch.zeros = _zeros_cpu
    try:
        return fn(*a, **k)
    finally:
        torch.zeros = _zeros


def main():
    g = torch.Generator().manual_seed(1234)
    out = {}

    # ---- SH (utils/sh_utils.py:57-112) ----
    from utils.sh_utils import eval_sh
    dirs = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=1)
    sh = torch.randn(64, 3, 16, generator=g)
    out["sh_dirs"], out["sh_coeffs"] = dirs.numpy(), sh.numpy()
    for deg in range(4):
        out[f"sh_eval_deg{deg}"] = eval_sh(deg, sh, dirs).numpy()

    # ---- cov3D / rotation (utils/general_utils.py:64-110, scene/gaussian_model.py:30-34) ----
    from utils import general_utils as gu
    scal = torch.exp(0.5 * torch.randn(48, 3, generator=g))
    rot = torch.randn(48, 4, generator=g)  # build_rotation normalises internally
    L = on_cpu(gu.build_scaling_rotation, 1.7 * scal, rot)
    cov = on_cpu(gu.strip_symmetric, L @ L.transpose(1, 2))
    out["cov_scaling"], out["cov_rotation"], out["cov_modifier"] = scal.numpy(), rot.numpy(), np.float32(1.7)
    out["cov_R"] = on_cpu(gu.build_rotation, rot).numpy()
    out["cov_sym6"] = cov.numpy()

    # ---- face frame (utils/graphics_utils.py:116-135 + scene/flame_gaussian_model.py:137-147) ----
    from gaussianavatars_b200 import synthetic as syn
    verts, faces = syn.head_mesh(n_lat=9, n_lon=14, seed=2)
    verts = syn.pose_mesh(verts, 3)
    from utils.graphics_utils import compute_face_orientation
    R, s = compute_face_orientation(verts, faces, return_scale=True)
    out["ff_verts"], out["ff_faces"], out["ff_R"], out["ff_scale"] = verts.numpy(), faces.numpy(), R.numpy(), s.numpy()
    import scene.flame_gaussian_model as fgm

    class Dummy:
        pass

    d = Dummy()
    d.flame_model = Dummy()
    d.flame_model.faces = faces
    fgm.FlameGaussianModel.update_mesh_properties(d, verts[None], verts[None])
    out["ff_center"], out["ff_quat_wxyz"] = d.face_center.numpy(), d.face_orien_quat.numpy()
    assert np.allclose(d.face_orien_mat.numpy(), out["ff_R"]) and np.allclose(d.face_scaling.numpy(), out["ff_scale"])

    # ---- binding getters (scene/gaussian_model.py:113-163) ----
    from scene.gaussian_model import GaussianModel
    P, F = 200, faces.shape[0]
    gm = GaussianModel(3)
    gm._xyz = torch.randn(P, 3, generator=g)
    gm._rotation = torch.randn(P, 4, generator=g) * 2.0
    gm._scaling = 0.5 * torch.randn(P, 3, generator=g) - 1.0
    gm._opacity = 2.0 * torch.randn(P, 1, generator=g)
    gm._features_dc = torch.randn(P, 1, 3, generator=g)
    gm._features_rest = torch.randn(P, 15, 3, generator=g)
    gm.binding = torch.randint(0, F, (P,), generator=g)
    gm.face_center, gm.face_orien_mat = d.face_center, d.face_orien_mat
    gm.face_scaling, gm.face_orien_quat = d.face_scaling, d.face_orien_quat
    for k in ("_xyz", "_rotation", "_scaling", "_opacity", "_features_dc", "_features_rest", "binding"):
        out["gm" + k] = getattr(gm, k).numpy()
    out["gm_get_xyz"] = gm.get_xyz.numpy()
    out["gm_get_rotation"] = gm.get_rotation.numpy()
    out["gm_get_scaling"] = gm.get_scaling.numpy()
    out["gm_get_opacity"] = gm.get_opacity.numpy()
    out["gm_get_features"] = gm.get_features.numpy()
    out["gm_get_covariance_mod1p3"] = on_cpu(gm.get_covariance, 1.3).numpy()  # quirk: ignores the face rotation
    gm.binding = None
    out["gm_unbound_get_rotation"] = gm.get_rotation.numpy()
    out["gm_unbound_get_scaling"] = gm.get_scaling.numpy()

    # ---- cameras (scene/cameras.py:44-47, utils/graphics_utils.py:38-71) ----
    from scene.cameras import Camera
    Rm = torch.linalg.qr(torch.randn(3, 3, generator=g))[0].numpy().astype(np.float64)
    if np.linalg.det(Rm) < 0:
        Rm[:, 0] *= -1
    T = np.array([0.1, -0.2, 2.5])
    cam = Camera(colmap_id=0, R=Rm, T=T, FoVx=0.8, FoVy=0.6, bg=None, image_width=64, image=None, image_height=48,
                 image_path=None, image_name="x", uid=0)
    out["cam_R"], out["cam_T"] = Rm, T
    out["cam_world_view"], out["cam_full_proj"] = cam.world_view_transform.numpy(), cam.full_proj_transform.numpy()
    out["cam_center"] = cam.camera_center.numpy()

    # ---- OrbitCamera as used by fps_benchmark_demo.prepare_camera (utils/viewer_utils.py:73-170) ----
    cwd = os.getcwd()
    os.chdir("/tmp")  # OrbitCamera.load() looks for ./camera.json
    from utils.viewer_utils import OrbitCamera
    for (W, H) in ((550, 802), (1920, 1080)):
        oc = OrbitCamera(W, H, r=1, fovy=20, convention="opencv")
        out[f"orbit_{W}x{H}_world_view_T"] = np.asarray(oc.world_view_transform, np.float32).T
        out[f"orbit_{W}x{H}_full_proj_T"] = np.asarray(oc.full_proj_transform, np.float32).T
        out[f"orbit_{W}x{H}_campos"] = np.asarray(oc.pose[:3, 3], np.float32)
        out[f"orbit_{W}x{H}_fov"] = np.array([np.radians(oc.fovx), np.radians(oc.fovy)], np.float64)
    os.chdir(cwd)

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
