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
    r = mtube.bi_timestep(mtube.OracleStep(oracle_lib.Oracle(sus2.Lb), sus2, W2, vbkg))
    assert 0 < r["wall_iterations"] <= 60 and r["history"][-1] < 1e-3 * r["history"][0]
