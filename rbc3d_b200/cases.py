"""Input formats and example configurations of the reference, for the harness.

* ``read_wall_mesh``  -- ReadWallMesh (ModIO.F90:547-605): Tri3 Exodus II wall meshes (NetCDF classic: dimensions
  num_nodes / num_elem / num_nod_per_el1, variables coordx / coordy / coordz / connect1).
* ``read_tube_in``    -- the head of Input/tube.in as ReadConfig consumes it (ModConf.F90:241-273): alpha_Ewd, eps_Ewd,
  PBspln_Ewd, nCellTypes, viscRat(:), refRad, Deflate, the pressure-gradient lines, Nt, Ts.
* ``minicase``        -- examples/minicase/minit.F90 restated: the cylinder mesh mapped to radius 5 and length 8, the box
  Lb = (10.5, 10.5, 8), two unrotated biconcave cells at the init program's positions, everything recentred.

Nothing here is on the product path; the files live under /root/reference (not present on the GPU box), so only CPU
tests use the readers and they skip when the files are absent.
"""
from __future__ import annotations

import numpy as np

from . import synth


def read_wall_mesh(path: str):
    """-> x (3, nvert) float64, e2v (3, nele) int32 with 1-based vertex numbers (wall%x, wall%e2v)."""
    from scipy.io import netcdf_file
    f = netcdf_file(path, "r", mmap=False)
    try:
        if f.dimensions["num_nod_per_el1"] != 3:
            raise ValueError("the input mesh is not of Tri3 type")            # ModIO.F90:567-571
        x = np.stack([np.array(f.variables[k].data, dtype=np.float64) for k in ("coordx", "coordy", "coordz")])
        conn = np.array(f.variables["connect1"].data, dtype=np.int32)         # (nele, 3), 1-based
    finally:
        f.close()
    return np.ascontiguousarray(x), np.ascontiguousarray(conn.T)


def read_tube_in(path: str) -> dict:
    """Leading entries of tube.in (one value per line, '!' starts a comment)."""
    vals = []
    with open(path) as fh:
        for line in fh:
            t = line.split("!")[0].strip()
            if t:
                vals.append(t)
    it = iter(vals)
    out = {"alpha_Ewd": float(next(it).replace("D", "E")), "eps_Ewd": float(next(it).replace("D", "E")),
           "PBspln_Ewd": int(next(it))}
    ntypes = int(next(it))
    out["nCellTypes"] = ntypes
    out["viscRat"] = [float(next(it).replace("D", "E")) for _ in range(ntypes)]
    out["refRad"] = float(next(it).replace("D", "E"))
    out["Deflate"] = next(it).lower().startswith(".t")
    return out


def minicase(mesh_path: str, nlat0: int = 12, dealias: int = 3, seed: int = 161269):
    """-> (suspension, walls, vBkg) of examples/minicase (minit.F90:34-116)."""
    x, e2v = read_wall_mesh(mesh_path)
    actlen, lengtube, tube_rad = 13.33, 8.0, 5.0
    th = np.arctan2(x[0], x[1])                                   # ATAN2(wall%x(i,1), wall%x(i,2)), minit.F90:53
    xw = np.stack([tube_rad * np.cos(th), tube_rad * np.sin(th), lengtube / actlen * x[2]])
    Lb = np.array([xw[0].max() - xw[0].min() + 0.5, 0.0, xw[2].max() - xw[2].min()])
    Lb[1] = Lb[0]
    centers = np.array([[-0.5, -0.5, 4.0], [0.5, 0.5, 1.0]])      # minit.F90:80-96
    centers[:, 0] += 0.5 * Lb[0]                                  # Recenter_Cells_and_Walls
    centers[:, 1] += 0.5 * Lb[1]
    xc = 0.5 * (xw.min(axis=1) + xw.max(axis=1))
    xw = xw + (0.5 * Lb - xc)[:, None]
    sus = synth.make_suspension(1, nlat0=nlat0, dealias=dealias, seed=seed, L=1.0, centers=centers, rotate=False,
                                visc_ratio=1.0)
    sus.Lb = Lb
    W = synth.Walls(np.array([xw.shape[1]], np.int32), np.array([e2v.shape[1]], np.int32), np.ascontiguousarray(xw),
                    np.ascontiguousarray(e2v), None, None, np.zeros_like(xw))
    W.area, W.epsDist = synth.wall_geometry(W.x, W.e2v_global())
    return sus, W, np.array([0.0, 0.0, 8.0])


def carotid_web_walls(input_dir: str):
    """-> (walls, Lb) of examples/carotid_web (carotid_initcond.F90:47-70, 118-138): carotid.e + web.e, shifted so that
    the first wall's coordinates start at 0 (recenterWalls), Lb = (max x + 0.5, max y + 0.5, 30)."""
    import os
    xs, es = [], []
    for name in ("carotid.e", "web.e"):
        x, e = read_wall_mesh(os.path.join(input_dir, name))
        xs.append(x)
        es.append(e)
    off = -xs[0].min(axis=1)
    xs = [x + off[:, None] for x in xs]
    Lb = np.array([xs[0][0].max() + 0.5, xs[0][1].max() + 0.5, 30.0])
    x = np.ascontiguousarray(np.concatenate(xs, axis=1))
    e2v = np.ascontiguousarray(np.concatenate(es, axis=1))
    W = synth.Walls(np.array([a.shape[1] for a in xs], np.int32), np.array([a.shape[1] for a in es], np.int32), x, e2v,
                    None, None, np.zeros_like(x))
    W.area, W.epsDist = synth.wall_geometry(W.x, W.e2v_global())
    return W, Lb
