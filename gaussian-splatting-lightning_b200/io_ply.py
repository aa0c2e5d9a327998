"""3DGS ``.ply`` point-cloud files <-> the model's parameter tensors, without the ``plyfile`` dependency.

The on-disk format is the one the reference reads and writes in ``internal/utils/gaussian_utils.py:51-255``
(``GaussianPlyUtils.load_from_ply`` / ``save_to_ply``): one ``vertex`` element, ``binary_little_endian``, float32
properties ``x y z [nx ny nz] f_dc_0..2 f_rest_0..3(K-1)-1 opacity scale_0..2 rot_0..3`` (+ optional uint8 ``red green
blue``).  ``f_rest_*`` is CHANNEL-major — ``features_rest[N, 3, K-1]`` flattened (``:66,:198-200``) — while the model keeps
``shs_rest[N, K-1, 3]`` (``vanilla_gaussian.py``), hence the transposes below.  All values are raw (pre-activation), exactly
what a checkpoint's ``gaussian_model.gaussians.*`` holds.  Properties are looked up by NAME, so files written by the original
3DGS code (which carry normals) load as well.
"""
from typing import Dict, Optional

import numpy as np
import torch

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4", "double": "f8",
              "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"}


def _read_vertex_table(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, current = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                current = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(current)
            elif tok[0] == "property":
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not part of the 3DGS layout")
                current["props"].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: only binary PLY files are supported (format {fmt})")
        order = "<" if fmt == "binary_little_endian" else ">"
        for el in elements:
            dtype = np.dtype([(n, order + t) for n, t in el["props"]])
            data = np.fromfile(f, dtype=dtype, count=el["count"])
            if data.shape[0] != el["count"]:
                raise ValueError(f"{path}: element {el['name']} is truncated")
            if el["name"] == "vertex":
                return data
    raise ValueError(f"{path}: no vertex element")


def _stack(table: np.ndarray, prefix: str, required: bool = True) -> np.ndarray:
    names = sorted((n for n in table.dtype.names if n.startswith(prefix)), key=lambda n: int(n.split("_")[-1]))
    if not names:
        if required:
            raise RuntimeError(f"'{prefix}' not found in ply")
        return np.empty((table.shape[0], 0), dtype=np.float32)
    return np.stack([np.asarray(table[n], dtype=np.float32) for n in names], axis=1)


def load_ply(path: str, sh_degree: int = -1) -> Dict[str, torch.Tensor]:
    """-> {"means" [N,3], "shs_dc" [N,1,3], "shs_rest" [N,K-1,3], "opacities" [N,1], "scales" [N,3], "rotations" [N,4]} float32 (raw
    parameters) + "sh_degree" (int).  sh_degree < 0: derived from the number of f_rest_* properties, like the reference."""
    t = _read_vertex_table(path)
    n = t.shape[0]
    means = np.stack([t["x"], t["y"], t["z"]], axis=1).astype(np.float32)
    dc = np.stack([t["f_dc_0"], t["f_dc_1"], t["f_dc_2"]], axis=1).astype(np.float32)            # [N,3]
    rest = _stack(t, "f_rest_", required=False).reshape(n, 3, -1)                                    # channel-major on disk
    k_rest = rest.shape[-1]
    if sh_degree >= 0:
        if k_rest != (sh_degree + 1) ** 2 - 1:
            raise ValueError(f"{path}: {3 * k_rest} f_rest properties do not match sh_degree {sh_degree}")
    else:
        sh_degree = next((d for d in range(5) if (d + 1) ** 2 - 1 == k_rest), -1)
        if sh_degree < 0:
            raise ValueError(f"{path}: cannot derive the SH degree from {3 * k_rest} f_rest properties")
    out = {
        "means": torch.from_numpy(means),
        "shs_dc": torch.from_numpy(np.ascontiguousarray(dc[:, None, :])),
        "shs_rest": torch.from_numpy(np.ascontiguousarray(rest.transpose(0, 2, 1))),
        "opacities": torch.from_numpy(np.asarray(t["opacity"], dtype=np.float32)[:, None].copy()),
        "scales": torch.from_numpy(_stack(t, "scale_")),
        "rotations": torch.from_numpy(_stack(t, "rot_")),
    }
    out["sh_degree"] = sh_degree
    return out


def save_ply(path: str, params: Dict[str, torch.Tensor], with_normals: bool = False, colors: Optional[torch.Tensor] = None) -> None:
    """Writes the layout of GaussianPlyUtils.save_to_ply (gaussian_utils.py:187-246).  with_normals adds zero nx ny nz after xyz (the
    original 3DGS writer's layout); colors: optional uint8 [N,3] -> red green blue."""
    g = {k: v.detach().cpu().numpy().astype(np.float32) for k, v in params.items() if isinstance(v, torch.Tensor)}
    n = g["means"].shape[0]
    cols = [("x", g["means"][:, 0]), ("y", g["means"][:, 1]), ("z", g["means"][:, 2])]
    if with_normals:
        cols += [(k, np.zeros(n, np.float32)) for k in ("nx", "ny", "nz")]
    dc = g["shs_dc"].reshape(n, 3)
    cols += [(f"f_dc_{i}", dc[:, i]) for i in range(3)]
    rest = g["shs_rest"].transpose(0, 2, 1).reshape(n, -1)          # [N,K-1,3] -> channel-major [N,3*(K-1)]
    cols += [(f"f_rest_{i}", rest[:, i]) for i in range(rest.shape[1])]
    cols.append(("opacity", g["opacities"].reshape(n)))
    cols += [(f"scale_{i}", g["scales"][:, i]) for i in range(g["scales"].shape[1])]
    cols += [(f"rot_{i}", g["rotations"][:, i]) for i in range(g["rotations"].shape[1])]
    dtype = [(name, "<f4") for name, _ in cols]
    if colors is not None:
        c = colors.detach().cpu().numpy().astype(np.uint8)
        cols += [("red", c[:, 0]), ("green", c[:, 1]), ("blue", c[:, 2])]
        dtype += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    table = np.empty(n, dtype=np.dtype(dtype))
    for name, v in cols:
        table[name] = v
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"]
    header += [f"property {'uchar' if t == 'u1' else 'float'} {name}" for name, t in dtype]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        table.tofile(f)
