"""Copy the wall meshes the BASELINE.json configs use from the reference tree into rbc3d_b200/data/meshes/ (input DATA of
the reference, not source code), with their SHA-256 in a manifest, so that the GPU box -- where /root/reference does not
exist -- runs the operator on the reference's own geometries (tests/test_gpu_reference_configs.py).

    python scripts/make_golden_meshes.py          # run in the build container, where /root/reference is mounted
"""
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "rbc3d_b200", "data", "meshes")
FILES = {
    "new_cyl_D6_L13_33.e": "/root/reference/examples/minicase/Input/new_cyl_D6_L13_33.e",   # minicase, case, case_sickles
    "carotid.e": "/root/reference/examples/carotid_web/Input/carotid.e",                   # carotid_web wall 1
    "web.e": "/root/reference/examples/carotid_web/Input/web.e",                           # carotid_web wall 2
}

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for name, src in FILES.items():
        dst = os.path.join(OUT, name)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        manifest[name] = {"source": src.replace("/root/reference/", ""), "bytes": os.path.getsize(dst),
                          "sha256": hashlib.sha256(open(dst, "rb").read()).hexdigest()}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1))
