"""Restart files of the reference (WriteRestart / ReadRestart, ModIO.F90:742-921) in the harness (rbc3d_b200/cases.py):
record structure byte for byte, round trip, and a restart of the minicase-like state through the oracle."""
import struct

import numpy as np

from rbc3d_b200 import cases, mtube


def minicase_restart(tmp_path):
    sus, W = mtube.minicase_like(nlat0=6)
    npc = sus.nlat * sus.nlon
    cells = [dict(nlat0=sus.nlat0, nlon0=2 * sus.nlat0, nlat=sus.nlat, nlon=sus.nlon, celltype=1,
                  starting_area=float(sus.area[c]), x=sus.x[:, c * npc:(c + 1) * npc].reshape(3, sus.nlon, sus.nlat))
             for c in range(sus.ncell)]
    walls = [dict(x=W.x, f=W.f + 0.25, e2v=W.e2v)]
    path = str(tmp_path / "restart.LATEST.dat")
    cases.write_restart(path, sus.Lb, 7, 0.0056, [0.0, 0.0, 8.0], cells, walls)
    return path, sus, W


def test_record_structure_is_fortran_unformatted_sequential(tmp_path):
    path, sus, W = minicase_restart(tmp_path)
    raw = open(path, "rb").read()
    # walk the 4-byte record markers: lengths of the records WriteRestart emits, in order
    pos, lens = 0, []
    while pos < len(raw):
        n = struct.unpack("<i", raw[pos:pos + 4])[0]
        assert struct.unpack("<i", raw[pos + 4 + n:pos + 8 + n])[0] == n
        lens.append(n)
        pos += 8 + n
    npc8 = 3 * sus.nlat * sus.nlon * 8
    cell = [8, 8, 4, 8, npc8]
    wall = [8, W.NV * 24, W.NV * 24, W.NE * 12]
    assert lens == [24, 4, 8, 24, 4] + cell * 2 + [4] + wall
    assert struct.unpack("<3d", raw[4:28]) == tuple(sus.Lb)
    # x(nlat, nlon, 3) in Fortran element order: the first nlat doubles are the first meridian of the x component
    off = sum(8 + n for n in lens[:9]) + 4
    first = np.frombuffer(raw[off:off + 8 * sus.nlat], dtype="<f8")
    assert np.array_equal(first, sus.x[0, :sus.nlat])


def test_round_trip_and_state(tmp_path, oracle_lib):
    path, sus, W = minicase_restart(tmp_path)
    rst = cases.read_restart(path)
    assert rst["Nt0"] == 7 and rst["time0"] == 0.0056 and np.array_equal(rst["vBkg"], [0.0, 0.0, 8.0])
    assert np.array_equal(rst["Lb"], sus.Lb) and len(rst["cells"]) == 2 and len(rst["walls"]) == 1
    sus2, W2, vbkg = cases.state_from_restart(rst)
    assert np.array_equal(sus2.x, sus.x) and np.array_equal(W2.x, W.x) and np.array_equal(W2.e2v, W.e2v)
    assert np.array_equal(W2.f, W.f + 0.25)
    # geometry rebuilt from x alone (spectral tangents) agrees with the analytic biconcave geometry
    assert np.abs(sus2.a3 - sus.a3).max() < 1e-9 and np.abs(sus2.detj - sus.detj).max() < 1e-9 * sus.detj.max()
    assert np.abs(sus2.spx - sus.spx).max() < 1e-9
    # and runs: one time step's boundary-integral work from the restart
    W2.f[:] = 0.0
    from oracle import harness
    r = mtube.bi_timestep(harness.OracleStep(oracle_lib.Oracle(sus2.Lb), sus2, W2, vbkg))
    assert 0 < r["wall_iterations"] <= 60 and r["history"][-1] < 1e-3 * r["history"][0]


def test_tecplot_writers(tmp_path):
    """WriteManyRBCs / WriteManyWalls (ModIO.F90:180-227, 389-423): zone headers, point counts, closed surfaces."""
    sus, W = mtube.minicase_like(nlat0=6)
    pc, pw = str(tmp_path / "x000000000.dat"), str(tmp_path / "wall000000000.dat")
    cases.write_many_rbcs(pc, sus)
    cases.write_many_walls(pw, W)
    lines = open(pc).read().splitlines()
    nlat, nlon = sus.nlat, sus.nlon
    per = 1 + (nlat + 1) * (nlon + 1)
    assert lines[0] == "VARIABLES = X, Y, Z" and len(lines) == 1 + sus.ncell * per
    assert lines[1] == "ZONE I=%9d  J=%9d  F=POINT" % (nlat + 1, nlon + 1)
    zone = np.array([[float(t) for t in ln.split()] for ln in lines[2:1 + per]]).reshape(nlon + 1, nlat + 1, 3)
    assert np.array_equal(zone[0], zone[-1])                              # the first meridian closes the surface
    # equally spaced colatitudes include the poles: one point per pole, the same on every meridian
    assert np.abs(zone[:, 0] - zone[0, 0]).max() < 1e-9 and np.abs(zone[:, -1] - zone[0, -1]).max() < 1e-9
    # the biconcave cell: poles on the axis through the centre, at half the dimple thickness 0.5 * 1.3858189 * 0.207
    assert np.allclose(zone[0, 0, :2], sus.centers[0][:2], atol=1e-9)
    assert abs(abs(zone[0, 0, 2] - sus.centers[0][2]) - 0.5 * 1.3858189 * 0.207) < 1e-9
    wl = open(pw).read().splitlines()
    assert wl[1] == "ZONE N = %9d E = %9d F=FEPOINT ET=TRIANGLE" % (W.NV, W.NE) and len(wl) == 2 + W.NV + W.NE
    assert [int(t) for t in wl[2 + W.NV].split()] == list(W.e2v[:, 0])
