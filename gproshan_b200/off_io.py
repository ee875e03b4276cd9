"""OFF mesh I/O with the reference's conventions (src/che_off.cpp:28-100).

read_off : OFF / COFF / NOFF headers (colour or normal columns are skipped); when the FIRST face is a quad every
           face is read as a quad (a b c d) and split into the half-edges (a b c) (d a c), exactly like
           che_off::read_file (:52-77).
write_off: "OFF", "V F 0", positions, "3 a b c" rows (che_off::write_file :82-100). The reference prints positions
           with the stream default of 6 significant digits, which does not round-trip; here `digits=17` (the
           default) writes round-trip-exact doubles and `digits=6` mimics the reference.
Host-side input handling only (BASELINE config C1 names an OFF mesh); nothing here runs on the GPU path.
"""
from __future__ import annotations

import numpy as np


def read_off(path: str, dtype=np.float64):
    """-> (xyz[V,3], faces[3*F] uint32)"""
    with open(path, "rb") as f:
        tok = f.read().split()
    kind = tok[0].decode()
    if not kind.endswith("OFF"):
        raise ValueError(f"{path}: not an OFF file (header {kind!r})")
    n_v, n_f = int(tok[1]), int(tok[2])
    per_vertex = 3 + (4 if kind[0] == "C" else 3 if kind[0] == "N" else 0)
    pos = 4
    vert = np.array(tok[pos:pos + per_vertex * n_v], dtype=np.float64).reshape(n_v, per_vertex)[:, :3]
    pos += per_vertex * n_v
    if n_f == 0:
        return vert.astype(dtype), np.zeros(0, dtype=np.uint32)
    first = int(tok[pos])
    if first not in (3, 4):
        raise ValueError("only triangle and quad faces are supported (like the reference)")
    rows = np.array(tok[pos:pos + (first + 1) * n_f], dtype=np.int64).reshape(n_f, first + 1)
    if (rows[:, 0] != first).any():
        raise ValueError("mixed face sizes are not supported (the reference sizes its tables from the first face)")
    idx = rows[:, 1:]
    if first == 4:  # (a b c d) -> (a b c) (d a c)
        idx = np.concatenate([idx[:, [0, 1, 2]], idx[:, [3, 0, 2]]], axis=1).reshape(-1, 3)
    return vert.astype(dtype), np.ascontiguousarray(idx.reshape(-1), dtype=np.uint32)


def write_off(path: str, xyz, faces, digits: int = 17):
    xyz = np.asarray(xyz, dtype=np.float64)
    tri = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    with open(path, "w") as f:
        f.write("OFF\n")
        f.write(f"{xyz.shape[0]} {tri.shape[0]} 0\n")
        np.savetxt(f, xyz, fmt=f"%.{digits}g")
        np.savetxt(f, np.concatenate([np.full((tri.shape[0], 1), 3), tri], axis=1), fmt="%d")
