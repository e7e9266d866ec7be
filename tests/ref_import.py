"""Imports the REAL reference (/root/reference, read-only) into a CPU process: the modules it needs but this image
lacks are stubbed (never executed on the paths the tests touch); roma's three functions come from oracle/binding.py.
Used by the golden generators' twins in tests/ -- only where /root/reference is mounted (this container, not the GPU
box): callers must skip when `available()` is False."""
import os
import sys
import types

REF = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "scene"))


def prepare():
    """Puts /root/reference on sys.path and installs the stubs.  Idempotent."""
    if not available():
        raise RuntimeError("/root/reference is not mounted")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from oracle import binding as ob

    if "roma" not in sys.modules:
        roma = types.ModuleType("roma")
        roma.quat_product = ob.quat_product
        roma.quat_xyzw_to_wxyz = ob.quat_xyzw_to_wxyz
        roma.quat_wxyz_to_xyzw = ob.quat_wxyz_to_xyzw
        roma.rotmat_to_unitquat = ob.rotmat_to_unitquat
        sys.modules["roma"] = roma
    for name in ("plyfile", "simple_knn", "simple_knn._C", "dearpygui", "dearpygui.dearpygui", "matplotlib",
                 "matplotlib.pyplot", "diff_gaussian_rasterization", "nvdiffrast", "nvdiffrast.torch", "pytorch3d",
                 "pytorch3d.io", "iopath", "iopath.common", "iopath.common.file_io", "chumpy", "lpips"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["plyfile"].PlyData = sys.modules["plyfile"].PlyElement = object
    sys.modules["simple_knn._C"].distCUDA2 = None
    sys.modules["pytorch3d.io"].load_obj = None
    sys.modules["iopath.common.file_io"].PathManager = object
    sys.modules["diff_gaussian_rasterization"].GaussianRasterizationSettings = object
    sys.modules["diff_gaussian_rasterization"].GaussianRasterizer = object


def on_cpu(fn, *a, **k):
    """Runs a reference function whose only CUDA dependence is a literal `device="cuda"` in torch.zeros(...)."""
    import torch

    real = torch.zeros

    def zeros_cpu(*args, **kw):
        kw.pop("device", None)
        return real(*args, **kw)

    torch.zeros = zeros_cpu
    try:
        return fn(*a, **k)
    finally:
        torch.zeros = real
