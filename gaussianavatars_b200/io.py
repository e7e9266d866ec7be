"""On-disk format either side of the path: the splat PLY the reference writes and
reads (`GaussianModel.save_ply` / `load_ply`, scene/gaussian_model.py:253-275, :282-332), without `plyfile`.

The file is `binary_little_endian 1.0`, one `vertex` element, every property `float`:
    x y z | nx ny nz (zeros) | f_dc_0..2 | f_rest_0..(3(D+1)^2-4) | opacity | scale_0..2 | rot_0..3 | [binding_0]
SH coefficients are stored CHANNEL-major (`_features_dc.transpose(1, 2).flatten(1)`): f_rest_{c*15+k} is coefficient
k+1 of colour channel c.  `binding_0` (the face index) is stored as float32 like everything else
(scene/gaussian_model.py:268-270) -- exact below 2^24 faces.

`load_ply` returns the dict `MeshBoundGaussians` / `synthetic.avatar_splats` use (raw, un-activated parameters on the
CPU; the record array is memory-mapped, each tensor is one strided gather out of it)."""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
              "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
              "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


def attribute_names(sh_rest_coeffs: int = 45, with_binding: bool = True):
    """`GaussianModel.construct_list_of_attributes` (scene/gaussian_model.py:234-251)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(3)]
    names += [f"f_rest_{i}" for i in range(sh_rest_coeffs)]
    names += ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]
    if with_binding:
        names.append("binding_0")
    return names


def ply_header(num_vertices: int, names) -> bytes:
    """The header `plyfile.PlyData([PlyElement.describe(elements, 'vertex')]).write()` produces for an all-float32
    record (byte-identical to the reference's files: tests/test_io.py pins it against media/306/point_cloud.ply)."""
    lines = ["ply", "format binary_little_endian 1.0", f"element vertex {int(num_vertices)}"]
    lines += [f"property float {n}" for n in names]
    lines.append("end_header")
    return ("\n".join(lines) + "\n").encode("ascii")


def _read_header(f):
    if f.readline().strip() != b"ply":
        raise ValueError("not a PLY file")
    fmt, count, props, in_vertex = None, None, [], False
    while True:
        line = f.readline()
        if not line:
            raise ValueError("PLY header is not terminated by end_header")
        tok = line.decode("ascii", "replace").split()
        if not tok or tok[0] == "comment" or tok[0] == "obj_info":
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            in_vertex = count is None  # only the first element is read, like the reference's `plydata.elements[0]`
            if in_vertex:
                count = int(tok[2])
        elif tok[0] == "property" and in_vertex:
            if tok[1] == "list":
                raise ValueError("list properties are not part of the splat format")
            if tok[1] not in _PLY_TYPES:
                raise ValueError(f"unknown PLY property type {tok[1]}")
            props.append((tok[2], _PLY_TYPES[tok[1]]))
        elif tok[0] == "end_header":
            break
    if fmt != "binary_little_endian":
        raise ValueError(f"only binary_little_endian PLY files are supported (the reference writes no other), got {fmt}")
    if count is None:
        raise ValueError("PLY file has no element")
    return count, props, f.tell()


def _indexed(names, prefix):
    sel = [n for n in names if n.startswith(prefix)]
    return sorted(sel, key=lambda n: int(n.split("_")[-1]))


def load_ply(path: str, max_sh_degree: int = 3) -> Dict[str, Optional[torch.Tensor]]:
    """`GaussianModel.load_ply` (scene/gaussian_model.py:282-332): raw parameters, float32, CPU.
    Keys: _xyz (P,3), _features_dc (P,1,3), _features_rest (P,(D+1)^2-1,3), _opacity (P,1), _scaling (P,3),
    _rotation (P,4), binding (P,) int32 or None."""
    with open(path, "rb") as f:
        count, props, offset = _read_header(f)
    rec = np.memmap(path, dtype=np.dtype(props), mode="r", offset=offset, shape=(count,)) if count else \
        np.zeros((0,), dtype=np.dtype(props))
    names = [n for n, _ in props]

    def cols(sel):
        out = np.empty((count, len(sel)), dtype=np.float32)
        for j, n in enumerate(sel):
            out[:, j] = rec[n]
        return torch.from_numpy(out)

    rest_names = _indexed(names, "f_rest_")
    n_rest = (max_sh_degree + 1) ** 2 - 1
    if len(rest_names) != 3 * n_rest:   # the reference's assert (scene/gaussian_model.py:297)
        raise ValueError(f"{path}: {len(rest_names)} f_rest_* properties, expected {3 * n_rest} for SH degree {max_sh_degree}")
    scale_names, rot_names = _indexed(names, "scale_"), _indexed(names, "rot")
    out = {
        "_xyz": cols(["x", "y", "z"]),
        "_features_dc": cols(["f_dc_0", "f_dc_1", "f_dc_2"]).reshape(count, 3, 1).transpose(1, 2).contiguous(),
        "_features_rest": cols(rest_names).reshape(count, 3, n_rest).transpose(1, 2).contiguous(),
        "_opacity": cols(["opacity"]),
        "_scaling": cols(scale_names),
        "_rotation": cols(rot_names),
        "binding": None,
    }
    binding_names = _indexed(names, "binding")
    if binding_names:
        b = np.empty((count, len(binding_names)), dtype=np.int32)
        for j, n in enumerate(binding_names):
            b[:, j] = rec[n]          # float32 -> int32, as the reference's assignment into an int32 array does
        out["binding"] = torch.from_numpy(b).squeeze(-1)
    return out


def save_ply(path: str, params: Dict[str, torch.Tensor], binding: Optional[torch.Tensor] = None) -> None:
    """`GaussianModel.save_ply` (scene/gaussian_model.py:253-275).  `params` holds the raw tensors under the names
    `load_ply` returns (or `xyz/features_dc/...` without the underscore); `binding` defaults to params["binding"]."""
    def get(k):
        v = params.get("_" + k, params.get(k))
        if v is None:
            raise KeyError(f"save_ply: missing parameter '{k}'")
        return v.detach().to("cpu", torch.float32)

    xyz = get("xyz")
    P = xyz.shape[0]
    f_dc = get("features_dc").transpose(1, 2).flatten(start_dim=1)
    f_rest = get("features_rest").transpose(1, 2).flatten(start_dim=1)
    if binding is None:
        binding = params.get("binding")
    blocks = [xyz, torch.zeros_like(xyz), f_dc, f_rest, get("opacity").reshape(P, 1), get("scaling"), get("rotation")]
    if binding is not None:
        blocks.append(binding.detach().to("cpu", torch.float32).reshape(P, 1))
    table = torch.cat(blocks, dim=1).contiguous().numpy().astype("<f4", copy=False)
    names = attribute_names(f_rest.shape[1], with_binding=binding is not None)
    assert table.shape[1] == len(names)
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        f.write(ply_header(P, names))
        f.write(table.tobytes())


# ================================================================================================================
# flame_param.npz (scene/flame_gaussian_model.py:219-258; layout: SURVEY.md Appendix D)
# ================================================================================================================
FLAME_STATIC_KEYS = ("shape", "static_offset")
FLAME_DYNAMIC_KEYS = ("translation", "rotation", "neck_pose", "jaw_pose", "eyes_pose", "expr", "dynamic_offset")


def _npz_members(path: str):
    """(name, dtype, shape, fortran, data offset) of every member of an UNCOMPRESSED .npz (what np.savez writes):
    the arrays can then be memory-mapped in place instead of being copied out of the zip."""
    import struct
    import zipfile

    from numpy.lib import format as npf

    out = []
    with zipfile.ZipFile(path) as zf, open(path, "rb") as f:
        for info in zf.infolist():
            if info.compress_type != zipfile.ZIP_STORED or not info.filename.endswith(".npy"):
                return None
            f.seek(info.header_offset)
            local = f.read(30)
            if local[:4] != b"PK\x03\x04":
                return None
            n_name, n_extra = struct.unpack("<HH", local[26:30])
            start = info.header_offset + 30 + n_name + n_extra
            f.seek(start)
            major, minor = npf.read_magic(f)
            shape, fortran, dtype = npf.read_array_header_1_0(f) if (major, minor) == (1, 0) else npf.read_array_header_2_0(f)
            if dtype.hasobject:
                return None
            out.append((info.filename[:-4], dtype, shape, fortran, f.tell()))
    return out


def load_flame_param(path: str, motion_path: Optional[str] = None, device=None, mmap: bool = True) -> Dict[str, torch.Tensor]:
    """`flame_param.npz` next to a point_cloud.ply -> {shape (300,), expr (T,100), rotation / neck_pose / jaw_pose /
    translation (T,3), eyes_pose (T,6), static_offset (1,V,3), dynamic_offset (T,V,3)} as tensors
    (FlameGaussianModel.load_ply, scene/flame_gaussian_model.py:229-236).  With `motion_path`, the static entries
    (shape, static_offset) are kept and the dynamic ones are replaced by the float32 arrays of that file (:238-256).
    mmap: the arrays of an uncompressed .npz are mapped in place (the 69 MB demo file opens without being read)."""

    def read(p):
        members = _npz_members(p) if mmap else None
        if members is None:
            z = np.load(p)
            return {k: z[k] for k in z.files}
        d = {}
        for name, dtype, shape, fortran, off in members:
            n = int(np.prod(shape)) if len(shape) else 1
            arr = np.memmap(p, dtype=dtype, mode="r", offset=off, shape=(n,)) if n else np.zeros((0,), dtype)
            d[name] = arr.reshape(shape, order="F" if fortran else "C")
        return d

    def to_t(a):
        t = torch.from_numpy(np.array(a, copy=True) if device is None else np.asarray(a))
        return t if device is None else t.to(device)

    fp = {k: to_t(v) for k, v in read(path).items()}
    if motion_path is not None:
        mo = {k: to_t(v) for k, v in read(motion_path).items() if v.dtype == np.float32}
        fp = {**{k: fp[k] for k in FLAME_STATIC_KEYS}, **{k: mo[k] for k in FLAME_DYNAMIC_KEYS}}
    return fp


def save_flame_param(path: str, flame_param: Dict[str, torch.Tensor]) -> str:
    """FlameGaussianModel.save_ply's second half (scene/flame_gaussian_model.py:219-224): np.savez of the tensors
    (moved to the CPU), key order preserved.  `path` is either the .npz itself or the point_cloud.ply it sits beside."""
    if path.endswith(".ply"):
        path = os.path.join(os.path.dirname(path), "flame_param.npz")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    arrays = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in flame_param.items()}
    np.savez(path, **arrays)
    return path
