"""Input formats and example configurations of the reference, for the harness.

* ``read_wall_mesh``  -- ReadWallMesh (ModIO.F90:547-605): Tri3 Exodus II wall meshes (NetCDF classic: dimensions
  num_nodes / num_elem / num_nod_per_el1, variables coordx / coordy / coordz / connect1).
* ``read_tube_in``    -- the head of Input/tube.in as ReadConfig consumes it (ModConf.F90:241-273): alpha_Ewd, eps_Ewd,
  PBspln_Ewd, nCellTypes, viscRat(:), refRad, Deflate, the pressure-gradient lines, Nt, Ts.
* ``minicase``        -- examples/minicase/minit.F90 restated: the cylinder mesh mapped to radius 5 and length 8, the box
  Lb = (10.5, 10.5, 8), two unrotated biconcave cells at the init program's positions, everything recentred.

* ``case`` / ``carotid_web`` -- examples/case, case_sickles (initcond.F90, sickle_initcond.F90) and
  examples/carotid_web (carotid_initcond.F90: 72 cells in the carotid vessel with its web, two walls) restated.

Nothing here is on the product path.  The wall meshes the BASELINE configs use are committed as input data under
rbc3d_b200/data/meshes/ (scripts/make_golden_meshes.py, SHA-256 in MANIFEST.json), so the GPU box -- where /root/reference does
not exist -- runs the operator on the reference's own geometries; ``mesh_file`` resolves a name to the fixture or, failing
that, to the reference tree.
"""
from __future__ import annotations

import os

import numpy as np

from . import synth


_HERE = os.path.dirname(os.path.abspath(__file__))
DATA_DIR = os.path.join(_HERE, "data")          # input data of the example configurations (no expected results here)
CAROTID_CELLS = os.path.join(DATA_DIR, "carotid_web_cells.npz")
MESH_DIRS = (os.path.join(DATA_DIR, "meshes"),
             "/root/reference/examples/minicase/Input", "/root/reference/examples/carotid_web/Input")


def mesh_file(name: str) -> str:
    """path of a wall mesh of the BASELINE configs: the committed fixture, else the reference tree"""
    for d in MESH_DIRS:
        p = os.path.join(d, name)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(name)


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
    def expand(tokens):
        # list-directed input: 'r*c' repeats the constant c r times; a '/' ends the record's data
        out = []
        for tok in tokens:
            if tok.startswith("/"):
                break
            head, star, tail = tok.partition("*")
            if star and head.isdigit() and tail:
                out += [tail] * int(head)
            else:
                out.append(tok.split("/")[0] if "/" in tok and not tok.startswith(("'", '"')) else tok)
                if "/" in tok and not tok.startswith(("'", '"')):
                    break
        return out

    recs = []
    with open(path) as fh:
        for line in fh:
            t = expand(line.split("!")[0].replace(",", " ").split())
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


def case(mesh_path: str | None = None, nrbc: int = 8, sickles: bool = False, visc_ratio: float = 1.0,
         sickle_x: np.ndarray | None = None, seed: int = 161269):
    """-> (suspension, walls, vBkg) of examples/case (initcond.F90:36-106) and examples/case_sickles
    (sickle_initcond.F90): the cylinder mesh mapped to radius 5 and length nrbc / 0.7, Lb = (10.5, 10.5, length), nrbc
    cells on the axis at z = (iz - 1/2) length / nrbc -- all biconcave, or every second one the imported SickleCell.dat --
    and everything moved to the middle of the box (Recenter_Cells_and_Walls)."""
    from . import mtube, sphere
    x, e2v = read_wall_mesh(mesh_path or mesh_file("new_cyl_D6_L13_33.e"))
    actlen, lengtube = 13.33, nrbc / 0.7
    th = np.arctan2(x[0], x[1])
    xw = np.stack([5.0 * np.cos(th), 5.0 * np.sin(th), lengtube / actlen * x[2]])
    Lb = np.array([xw[0].max() - xw[0].min() + 0.5, 0.0, xw[2].max() - xw[2].min()])
    Lb[1] = Lb[0]
    spacing = Lb[2] / nrbc
    xc_w = 0.5 * (xw.min(axis=1) + xw.max(axis=1))
    xw = xw + (0.5 * Lb - xc_w)[:, None]
    thg, phig, _ = sphere.gauss_grid(36, 72)
    xb, _, _ = sphere.biconcave_unit(thg, phig, 1.0)
    if sickles and sickle_x is None:
        sickle_x = np.load(mtube.GOLDEN_SICKLE)["x"]
    xs = []
    for iz in range(1, nrbc + 1):
        xc = np.array([0.5 * Lb[0], 0.5 * Lb[1], spacing * (iz - 0.5)])
        xs.append(mtube.import_read_rbc(sickle_x, xc) if (sickles and iz % 2 == 0) else xb + xc[:, None, None])
    sus = synth.suspension_from_shapes(np.stack(xs), Lb, nlat0=12, dealias=3, visc_ratio=visc_ratio, seed=seed)
    W = synth.Walls(np.array([xw.shape[1]], np.int32), np.array([e2v.shape[1]], np.int32), np.ascontiguousarray(xw),
                    np.ascontiguousarray(e2v), None, None, np.zeros_like(xw))
    W.area, W.epsDist = synth.wall_geometry(W.x, W.e2v_global())
    return sus, W, np.array([0.0, 0.0, 8.0])


def carotid_web_walls(input_dir: str | None = None):
    """-> (walls, Lb) of examples/carotid_web (carotid_initcond.F90:47-70, 118-138): carotid.e + web.e, shifted so that
    the first wall's coordinates start at 0 (recenterWalls), Lb = (max x + 0.5, max y + 0.5, 30)."""
    xs, es = [], []
    for name in ("carotid.e", "web.e"):
        x, e = read_wall_mesh(os.path.join(input_dir, name) if input_dir else mesh_file(name))
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


def carotid_place_cells(W, nrbc: int, seed: int = 112, tubelen: float = 30.0, max_attempts: int = 2_000_000,
                        progress: bool = False, stall: int = 4000, n_random: int = 8):
    """The rejection sampling of carotid_initcond.F90:186-268 -> (centres (nrbc, 3), rotations (nrbc, 3, 3)): a random
    rotation (rotate_cell, :272-299), a random point of the vessel cross-section at a random height that is not inside
    the web (choose_point, :131-182), rejected when the cell leaves [0, tubelen) in z, comes closer than 0.3 to a wall
    vertex (check_wall_collision, :302-336) or touches an earlier cell (check_cell_collision, :341-407).  Differences,
    stated: the reference draws from its own generator (ModBasicMath RandomNumber, seed 112), this uses NumPy's PCG64
    with the same seed, and the cell-cell test (overlap of the bounding boxes of diagonal mesh neighbours, centres closer
    than 4) is restated as "no two mesh points closer than 0.15" (about the diagonal-neighbour spacing of the mesh,
    median 0.13) with a k-d tree; and because the random proposals need millions of attempts beyond ~50 cells (87 000 for
    the 53rd; the vessel narrows to a radius of 3), only the first ``n_random`` cells are drawn that way; the others come
    from a systematic sweep of the vessel with discs roughly across its axis, accepted by the same three tests -- the
    same cell count and wall meshes, not the same positions."""
    from scipy.spatial import cKDTree
    from . import sphere
    rng = np.random.Generator(np.random.PCG64(seed))
    thg, phig, _ = sphere.gauss_grid(36, 72)
    xb, _, _ = sphere.biconcave_unit(thg, phig, 1.0)                      # (3, nlon, nlat)
    xb_pts = xb.reshape(3, -1)
    vo = W.voff()
    w1, w2 = W.x[:, vo[0]:vo[1]], W.x[:, vo[1]:vo[2]]
    o1, o2 = np.argsort(w1[2]), np.argsort(w2[2])
    w1, w2 = w1[:, o1], w2[:, o2]
    wall_tree = cKDTree(W.x.T)
    diag = 0.15

    def choose_point():
        while True:
            ln = rng.random() * tubelen
            lo, hi = np.searchsorted(w1[2], [ln - 0.2, ln + 0.2], side="left")[0], np.searchsorted(w1[2], ln + 0.2, side="right")
            seg = w1[:2, lo:hi]
            mx, mn = seg.max(), seg.min()
            rad = 0.5 * (mx - mn)
            ang, r = rng.random() * 2 * np.pi, np.sqrt(rng.random()) * rad
            pt = np.array([r * np.cos(ang) + rad + mn, r * np.sin(ang) + rad + mn, ln])
            lo2, hi2 = np.searchsorted(w2[2], ln - 0.5, side="left"), np.searchsorted(w2[2], ln + 0.5, side="right")
            if hi2 > lo2 and w2[1, lo2:hi2].min() <= pt[1] <= w2[1, lo2:hi2].max():
                continue                                                  # the point is inside the web
            return pt

    centres, rots, all_pts, tree = [], [], np.zeros((0, 3)), None

    def try_place(R, xc):
        nonlocal all_pts, tree
        pts = (R @ xb_pts + xc[:, None]).T
        if (pts[:, 2] >= tubelen).any() or (pts[:, 2] < 0).any():
            return False
        if np.isfinite(wall_tree.query(pts, k=1, distance_upper_bound=0.3)[0]).any():
            return False
        if tree is not None and np.isfinite(tree.query(pts, k=1, distance_upper_bound=diag)[0]).any():
            return False
        centres.append(xc)
        rots.append(R)
        all_pts = np.concatenate([all_pts, pts])
        tree = cKDTree(all_pts)
        if progress:
            print("placed", len(centres), flush=True)
        return True

    fails = 0
    for _ in range(max_attempts):
        if len(centres) >= min(nrbc, n_random) or fails > stall:
            break
        while True:
            v1, v2 = rng.random(2) * 2 - 1
            vsq = v1 * v1 + v2 * v2
            if vsq < 1:
                break
        zv = np.array([v1 * 2 * np.sqrt(1 - vsq), v2 * 2 * np.sqrt(1 - vsq), 1 - 2 * vsq])
        fails = 0 if try_place(sphere.rotate_matrix(zv), choose_point()) else fails + 1
    if len(centres) < nrbc:
        # the random proposals have stalled: sweep the vessel systematically with discs across the axis (same tests)
        zs = np.arange(0.75, tubelen - 0.75, 0.125)            # ascending: a dense stack
        for ln in zs:
            lo, hi = np.searchsorted(w1[2], ln - 0.2, side="left"), np.searchsorted(w1[2], ln + 0.2, side="right")
            seg = w1[:2, lo:hi]
            mx, mn = seg.max(), seg.min()
            for dx in np.arange(mn + 1.0, mx - 1.0 + 1e-9, 0.25):
                for dy in np.arange(mn + 1.0, mx - 1.0 + 1e-9, 0.25):
                    if len(centres) == nrbc:
                        break
                    tilt = rng.normal(scale=0.15, size=2)
                    zv = np.array([tilt[0], tilt[1], 1.0])
                    try_place(sphere.rotate_matrix(zv / np.linalg.norm(zv)), np.array([dx, dy, ln]))
    if len(centres) < nrbc:
        raise RuntimeError("carotid_place_cells: placed %d of %d cells" % (len(centres), nrbc))
    return np.array(centres), np.array(rots)


def carotid_web(input_dir: str | None = None, nrbc: int | None = None, visc_ratio: float = 1.0, seed: int = 112,
                hematocrit: float = 0.2, placement=None):
    """-> (suspension, walls, Lb, vBkg) of examples/carotid_web (carotid_initcond.F90): the two walls and
    nrbc = 3 * tubelen * tuber^2 * hematocrit / 4 = 72 biconcave cells (:38).  ``placement`` = (centres, rotations) as
    returned by carotid_place_cells (the committed rbc3d_b200/data/carotid_web_cells.npz holds one such draw: the sampling
    takes minutes at 72 cells, as the init program's does); None = sample now."""
    from . import sphere
    W, Lb = carotid_web_walls(input_dir)
    tuber, tubelen = 4.0, 30.0
    if nrbc is None:
        nrbc = int((3 * (tubelen * tuber ** 2 * hematocrit)) / 4)        # carotid_initcond.F90:38 -> 72
    if placement is None:
        placement = carotid_place_cells(W, nrbc, seed, tubelen)
    centres, rots = placement
    centres, rots = np.asarray(centres)[:nrbc], np.asarray(rots)[:nrbc]
    thg, phig, _ = sphere.gauss_grid(36, 72)
    xb, _, _ = sphere.biconcave_unit(thg, phig, 1.0)
    x = np.einsum("cij,jlk->cilk", rots, xb) + centres[:, :, None, None]
    sus = synth.suspension_from_shapes(x, Lb, nlat0=12, dealias=3, visc_ratio=visc_ratio, seed=161269)
    return sus, W, Lb, np.array([0.0, 0.0, 8.0])


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
