"""Generate rbc3d_b200/data/ref_sickle_cell.npz from the reference's SickleCell.dat.

SickleCell.dat (examples/case_sickles/Input, sample_files/sample_cells) is the only OUTPUT OF THE REFERENCE CODE that
the reference tree ships: a cell surface written by ExportWriteRBC (ModIO.F90:607-629) at the end of a simulation, i.e.
after FilterRbcs -- SPHEREPACK shags analysis, truncation to degree < nlat0, shsgs synthesis on the reference's own
36 x 72 Gauss grid.  It is read back by ImportReadRBC (ModIO.F90:631-683): list-directed `nlat0 nlon0 / nlat nlon /
celltype / x(nlat, nlon, 3)` in Fortran array-element order (ilat fastest, then ilon, then component).

The fixture holds the header and x as (3, nlon, nlat) float64 exactly as parsed (no recentring), plus the SHA-256 of
the source file.  tests/test_reference_golden.py uses it as a golden vector of the reference for the Gauss grid, the
point ordering and the spherical-harmonic truncation this project restates (rbc3d_b200/sphere.py, gmres.py, solver.cu),
and as the heterogeneous cell shape of BASELINE.json configs[3].

    python scripts/make_golden_sickle.py      (needs /root/reference; the GPU box only sees the committed .npz)
"""
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/sample_files/sample_cells/SickleCell.dat"


def parse(path):
    tok = open(path).read().split()
    nlat0, nlon0, nlat, nlon, celltype = (int(t) for t in tok[:5])
    vals = np.array(tok[5:], dtype=np.float64)
    assert vals.size == 3 * nlat * nlon
    return (nlat0, nlon0, nlat, nlon, celltype), np.ascontiguousarray(vals.reshape(3, nlon, nlat))


def main():
    hdr, x = parse(SRC)
    digest = hashlib.sha256(open(SRC, "rb").read()).hexdigest()
    path = os.path.join(ROOT, "rbc3d_b200", "data", "ref_sickle_cell.npz")
    np.savez_compressed(path, header=np.array(hdr, dtype=np.int32), x=x, sha256=np.array(digest))
    print("wrote", path, os.path.getsize(path), "bytes", hdr, digest)


if __name__ == "__main__":
    main()
