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
    """Input/tube.in as ReadConfig consumes it (ModConf.F90:239-273): list-directed reads, one READ statement per
    record -- a scalar takes the first value of the next non-blank record ('!' ends the data of a line), an array of
    nCellTypes values (viscRat, refRad) takes as many values as it needs, over several records if necessary."""
    recs = []
    with open(path) as fh:
        for line in fh:
            t = line.split("!")[0].replace(",", " ").split()
            if t:
                recs.append(t)
    it = iter(recs)

    def num(tok):
        return float(tok.upper().replace("D", "E"))

    def logical(tok):
        return tok.strip(".").lower().startswith("t")

    def scalar(conv):
        return conv(next(it)[0])

    def array(n, conv):
        vals = []
        while len(vals) < n:
            vals += next(it)
        return [conv(v) for v in vals[:n]]

    out = {"alpha_Ewd": scalar(num), "eps_Ewd": scalar(num), "PBspln_Ewd": scalar(lambda t: int(num(t)))}
    n = out["nCellTypes"] = scalar(lambda t: int(num(t)))
    out["viscRat"] = array(n, num)
    out["refRad"] = array(n, num)
    out["Deflate"] = scalar(logical)
    out["pGradTar"] = [scalar(num) for _ in range(3)]
    out["Nt"] = scalar(lambda t: int(num(t)))
    out["Ts"] = scalar(num)
    for k in ("cell_out", "wall_out", "pgrad_out", "flow_out", "ftot_out", "restart_out"):
        out[k] = scalar(lambda t: int(num(t)))
    out["restart_file"] = scalar(lambda t: t.strip("'\""))
    out["epsDist"], out["ForceCoef"], out["viscRatThresh"] = scalar(num), scalar(num), scalar(num)
    out["rigidsep"] = scalar(logical)
    out["fmags"] = scalar(num)
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


# ---------------------------------------------------------------------------------------------------------------------
# Binary restart files (ModIO.F90:742-784 WriteRestart, :792-921 ReadRestart): Fortran unformatted sequential records
# (gfortran: 4-byte length before and after each record), real(WP) = float64, default integer = int32, arrays in Fortran
# element order -- rbc%x(nlat, nlon, 3) is (3, nlon, nlat) in C order, wall%x(nvert, 3) is (3, nvert), e2v(nele, 3) is
# (3, nele): exactly the SoA layouts of the boundary.
def write_restart(path: str, Lb, lt: int, time: float, vbkg, cells, walls) -> None:
    """cells: list of dicts(nlat0, nlon0, nlat, nlon, celltype, starting_area, x (3, nlon, nlat));
    walls: list of dicts(x (3, nvert), f (3, nvert), e2v (3, nele) int32 1-based)."""
    from scipy.io import FortranFile
    f = FortranFile(path, "w")
    try:
        f.write_record(np.asarray(Lb, dtype=np.float64))
        f.write_record(np.array([lt], dtype=np.int32))
        f.write_record(np.array([time], dtype=np.float64))
        f.write_record(np.asarray(vbkg, dtype=np.float64))
        f.write_record(np.array([len(cells)], dtype=np.int32))
        for c in cells:
            f.write_record(np.array([c["nlat0"], c["nlon0"]], dtype=np.int32))
            f.write_record(np.array([c["nlat"], c["nlon"]], dtype=np.int32))
            f.write_record(np.array([c["celltype"]], dtype=np.int32))
            f.write_record(np.array([c["starting_area"]], dtype=np.float64))
            x = np.ascontiguousarray(c["x"], dtype=np.float64)
            assert x.shape == (3, c["nlon"], c["nlat"])
            f.write_record(x)
        f.write_record(np.array([len(walls)], dtype=np.int32))
        for w in walls:
            x = np.ascontiguousarray(w["x"], dtype=np.float64)
            e2v = np.ascontiguousarray(w["e2v"], dtype=np.int32)
            f.write_record(np.array([x.shape[1], e2v.shape[1]], dtype=np.int32))
            f.write_record(x)
            f.write_record(np.ascontiguousarray(w["f"], dtype=np.float64))
            f.write_record(e2v)
    finally:
        f.close()


def read_restart(path: str) -> dict:
    """-> dict(Lb, Nt0, time0, vBkg, cells=[...], walls=[...]) with the array layouts of write_restart."""
    from scipy.io import FortranFile
    f = FortranFile(path, "r")
    try:
        out = {"Lb": f.read_reals(np.float64), "Nt0": int(f.read_ints(np.int32)[0]),
               "time0": float(f.read_reals(np.float64)[0]), "vBkg": f.read_reals(np.float64), "cells": [], "walls": []}
        for _ in range(int(f.read_ints(np.int32)[0])):
            nlat0, nlon0 = (int(v) for v in f.read_ints(np.int32))
            nlat, nlon = (int(v) for v in f.read_ints(np.int32))
            celltype = int(f.read_ints(np.int32)[0])
            area0 = float(f.read_reals(np.float64)[0])
            x = f.read_reals(np.float64)
            if x.size != 3 * nlat * nlon:
                raise ValueError("invalid array dimension")                      # ModIO.F90:867-875
            out["cells"].append(dict(nlat0=nlat0, nlon0=nlon0, nlat=nlat, nlon=nlon, celltype=celltype,
                                     starting_area=area0, x=x.reshape(3, nlon, nlat)))
        for _ in range(int(f.read_ints(np.int32)[0])):
            nvert, nele = (int(v) for v in f.read_ints(np.int32))
            x = f.read_reals(np.float64).reshape(3, nvert)
            fw = f.read_reals(np.float64).reshape(3, nvert)
            e2v = f.read_ints(np.int32).reshape(3, nele)
            out["walls"].append(dict(x=x, f=fw, e2v=e2v))
    finally:
        f.close()
    return out


def state_from_restart(rst: dict, visc_ratio=1.0, seed: int = 161269):
    """restart contents -> (Suspension, Walls | None, vBkg): geometry by RBC_ComputeGeometry's spectral tangents
    (synth.suspension_from_shapes); all cells must share one mesh size, as they do in every shipped example."""
    cells = rst["cells"]
    sus = None
    if cells:
        c0 = cells[0]
        if any((c["nlat"], c["nlon"], c["nlat0"]) != (c0["nlat"], c0["nlon"], c0["nlat0"]) for c in cells):
            raise ValueError("cells with different mesh sizes are not supported by the harness")
        x = np.stack([c["x"] for c in cells])
        sus = synth.suspension_from_shapes(x, rst["Lb"], nlat0=c0["nlat0"], dealias=c0["nlat"] // c0["nlat0"],
                                           visc_ratio=visc_ratio, seed=seed)
    W = None
    if rst["walls"]:
        ws = rst["walls"]
        x = np.ascontiguousarray(np.concatenate([w["x"] for w in ws], axis=1))
        e2v = np.ascontiguousarray(np.concatenate([w["e2v"] for w in ws], axis=1))
        fw = np.ascontiguousarray(np.concatenate([w["f"] for w in ws], axis=1))
        W = synth.Walls(np.array([w["x"].shape[1] for w in ws], np.int32), np.array([w["e2v"].shape[1] for w in ws], np.int32),
                        x, e2v, None, None, fw)
        W.area, W.epsDist = synth.wall_geometry(W.x, W.e2v_global())
    return sus, W, np.asarray(rst["vBkg"], dtype=float)


# ---------------------------------------------------------------------------------------------------------------------
# Tecplot output of the reference (ModIO.F90:180-227 WriteManyRBCs, :389-423 WriteManyWalls), so that states stepped by
# the harness can be looked at with the reference's own visualisation recipe (docs/visualization.md).
def write_many_rbcs(path: str, sus) -> None:
    """x_<step>.dat: every cell filtered to degree < nlat0 and synthesised on nlat + 1 equally spaced colatitudes
    0..pi (ShAnalGau + ShFilter + ShSynthEqu), one ZONE per cell, the first meridian repeated to close the surface."""
    from . import sphere
    nlat, nlon, nc = sus.nlat, sus.nlon, sus.ncell
    if nc <= 0:
        return
    proj = sphere.SphereProjector(nlat, nlon, sus.nlat0, np.linspace(0.0, np.pi, nlat + 1))
    x = proj(np.ascontiguousarray(sus.x.reshape(3, nc, nlon, nlat).transpose(1, 0, 2, 3)))    # (nc, 3, nlon, nlat + 1)
    with open(path, "w") as fh:
        fh.write("VARIABLES = X, Y, Z\n")
        for c in range(nc):
            fh.write("ZONE I=%9d  J=%9d  F=POINT\n" % (nlat + 1, nlon + 1))
            for j in list(range(nlon)) + [0]:
                for i in range(nlat + 1):
                    fh.write("%20.10f%20.10f%20.10f\n" % tuple(x[c, :, j, i]))


def write_many_walls(path: str, W) -> None:
    """wall_<step>.dat: FEPOINT / TRIANGLE zones, vertex numbers local to each wall (1-based)."""
    if W is None or W.nwall <= 0:
        return
    vo, eo = W.voff(), W.eoff()
    with open(path, "w") as fh:
        fh.write("VARIABLES = X, Y, Z\n")
        for w in range(W.nwall):
            fh.write("ZONE N = %9d E = %9d F=FEPOINT ET=TRIANGLE\n" % (W.nvert[w], W.nele[w]))
            for v in range(vo[w], vo[w + 1]):
                fh.write("%20.10f%20.10f%20.10f\n" % tuple(W.x[:, v]))
            for e in range(eo[w], eo[w + 1]):
                fh.write("%9d%9d%9d\n" % tuple(W.e2v[:, e]))
