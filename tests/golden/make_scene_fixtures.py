"""Regenerates tests/golden/{cornellbox_blacklight,colorballs}.npz from the reference's example scenes.

Run in the build container (needs /root/reference):  python tests/golden/make_scene_fixtures.py
The .npz hold the patch arrays MeshS::loadFromFile produces (checked against the reference's own loader compiled
in oracle/_ref) plus the raw Kd/Ke/Ks of every MTL entry; the GPU box has no /root/reference, so tests and the
benchmark read these instead of the OBJ files."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from daisyriot_b200 import scenes  # noqa: E402

REF = "/root/reference/example_scenes"
for name in ("cornellbox_blacklight", "colorballs"):
    sc = scenes.load_obj(os.path.join(REF, name + ".obj"), REF)
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    scenes.save_scene_npz(sc, out)
    print(name, sc.numtriangles, os.path.getsize(out))
