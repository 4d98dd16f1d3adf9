"""The C++ drop-in classes (shim/OptixPrimeFunctionality.h, shim/Lightning.h) driven the way the reference's main()
drives them, against the Python mirror on the same scene and sample pattern."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
DEMO = os.path.join(ROOT, "shim", "_build", "shim_demo")


@pytest.mark.skipif(not os.path.exists(DEMO), reason="shim demo is built where /root/reference exists (make -C shim)")
@pytest.mark.parametrize("method", [2, 1])
def test_cpp_shim_matches_python_mirror(tmp_path, coeff_model, uv50, method):
    import daisyriot_b200 as dz
    from daisyriot_b200 import scenes
    model, cwd = coeff_model
    sc = scenes.cornell_box(2048)
    if method == 1:  # give the RGB flavour something to emit (the UV lamp has no RGB emission)
        sc.materials[3]["Ke"] = np.array([0.2, 0.1, 0.05], np.float32)
    obj, _ = scenes.write_obj(sc, cwd, "shim_scene%d" % method)
    rands = os.path.join(cwd, "rands.bin")
    uv50.astype(np.float32).tofile(rands)
    ev = 7.0 if method == 2 else 500.0
    out = subprocess.run([DEMO, obj, cwd + "/", str(method), str(ev), "1", rands], cwd=cwd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"RESULT passes=(\d+) sumB=(\S+) color=(\S+),(\S+),(\S+)", out.stdout)
    assert m, out.stdout[-2000:]
    wl = np.arange(200, 601, 50).astype(np.float32)
    mesh = dz.MeshS.from_scene(sc, wl, model)
    p = dz.OptixPrimeFunctionality(mesh, rands=uv50)
    lt = dz.Lightning.get_lightning(method, mesh, p, ev, wl, True, None)
    B = lt.lightningvalues
    assert int(m.group(1)) == lt.numpasses and lt.numpasses > 0
    assert abs(float(m.group(2)) - B.astype(np.float64).sum()) <= 1e-6 * abs(B.astype(np.float64).sum())
    c = lt.get_color_of_patch(sc.numtriangles // 2)
    assert np.allclose([float(m.group(i)) for i in (3, 4, 5)], c, rtol=1e-4, atol=1e-5)
    lt.close(); p.close()


@pytest.mark.skipif(not os.path.exists(DEMO), reason="shim demo is built where /root/reference exists (make -C shim)")
def test_cpp_shim_matrix_cache_roundtrip(tmp_path, coeff_model, uv50):
    """initMatFromFile (reference Lightning.h:84-96; main.cpp:108 always passes a matfile): the first run builds the matrix
    and writes the cache in the reference's byte layout, the second run loads it instead of tracing -- same result; and the
    file the C++ side wrote is what the Python mirror reads and what the GPU matrix converts to."""
    import daisyriot_b200 as dz
    from daisyriot_b200 import api, scenes
    model, cwd = coeff_model
    sc = scenes.cornell_box(1024)
    obj, _ = scenes.write_obj(sc, cwd, "shim_cache_scene")
    rands = os.path.join(cwd, "rands_cache.bin")
    uv50.astype(np.float32).tofile(rands)
    matfile = str(tmp_path / "radmat.bin")
    outs = []
    for run in range(2):
        out = subprocess.run([DEMO, obj, cwd + "/", "2", "7.0", "1", rands, matfile], cwd=cwd, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        assert ("Loaded & Serialized matrix" if run == 0 else "Deserialized matrix") in out.stdout
        outs.append(re.search(r"RESULT (passes=\d+ sumB=\S+ color=\S+)", out.stdout).group(1))
    assert outs[0] == outs[1]
    # the file against the Python mirror
    F_file = api.deserialize_mat(matfile)
    wl = np.arange(200, 601, 50).astype(np.float32)
    p = dz.OptixPrimeFunctionality(dz.MeshS.from_scene(sc, wl, model), rands=uv50)
    F_gpu = p.cudaCalculateRadiosityMatrix().rows()
    assert np.array_equal(np.asarray(F_file, np.float32).view(np.uint32), F_gpu.view(np.uint32))
    # calculateAllVisibility through the shim: one triplet per non-zero entry, same total
    m = re.search(r"tripl=(\d+),(\S+)", out.stdout)
    assert int(m.group(1)) == int(np.count_nonzero(F_gpu))
    assert abs(float(m.group(2)) - F_gpu.astype(np.float64).sum()) <= 1e-9 * F_gpu.astype(np.float64).sum()
    # per-pair debugging entry points: C++ shim == Python mirror
    m = re.search(r"nusselt=(\S+) p2p=(\S+) shoot=(\d)", out.stdout)
    pa, pb = 3, sc.numtriangles // 2 + 5
    assert np.isclose(float(m.group(1)), p.p2pFormfactorNusselt(pa, pb), rtol=2e-6, atol=1e-12)
    assert np.isclose(float(m.group(2)), p.p2pFormfactor(pa, pb), rtol=2e-6, atol=1e-12)
    picks = np.zeros(2, api.HIT_DTYPE)
    picks["triangleId"] = [pa, pb]
    picks["u"] = [0.25, 0.3]
    picks["v"] = [0.5, 0.3]
    assert int(m.group(3)) == int(p.shootPatchRay(picks))
    p.close()


@pytest.mark.skipif(not os.path.exists(DEMO), reason="shim demo is built where /root/reference exists (make -C shim)")
def test_cpp_shim_trace_screen_png_and_picking(tmp_path, coeff_model, uv50):
    """traceScreen / intersectMouse through the C++ shim with the reference's own Camera.h (the step right after the solve,
    main.cpp:113): the frame the demo saves as PNG equals the Python mirror's traceScreen of the same camera, the pick in the
    middle of the screen is the triangle the closest-hit query reports."""
    import struct
    import zlib
    import daisyriot_b200 as dz
    from daisyriot_b200 import api, scenes
    model, cwd = coeff_model
    sc = scenes.cornell_box(2048)
    obj, _ = scenes.write_obj(sc, cwd, "shim_png_scene")
    rands = os.path.join(cwd, "rands_png.bin")
    uv50.astype(np.float32).tofile(rands)
    png = str(tmp_path / "view.png")
    out = subprocess.run([DEMO, obj, cwd + "/", "2", "7.0", "1", rands], cwd=cwd, capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, DAISY_DEMO_PNG=png))
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"IMAGE sum=(\S+) pick=(-?\d+) onscreen=(\d+)", out.stdout)
    assert m, out.stdout[-2000:]
    # decode the PNG (8-bit RGB, stored deflate blocks, bottom row first)
    raw = open(png, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    w, h = struct.unpack(">II", raw[16:24])
    assert (w, h) == (160, 120)
    pos, idat = 8, b""
    while pos < len(raw):
        n, tag = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        assert zlib.crc32(tag + body) & 0xFFFFFFFF == struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0]
        if tag == b"IDAT":
            idat += body
        pos += 12 + n
    pix = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)[:, 1:].reshape(h, w, 3)[::-1]
    # the same frame through the Python mirror
    wl = np.arange(200, 601, 50).astype(np.float32)
    mesh = dz.MeshS.from_scene(sc, wl, model)
    p = dz.OptixPrimeFunctionality(mesh, rands=uv50)
    lt = dz.Lightning.get_lightning(2, mesh, p, 7.0, wl, True, None)
    cam = api.Camera(160, 120, 4)
    lo, hi = sc.vertices.min(0), sc.vertices.max(0)
    mid = (np.float32(0.5) * (lo + hi)).astype(np.float32)
    cam.dir = mid.copy()
    cam.eye = np.array([mid[0], mid[1], hi[2] + np.float32(1.6) * (hi[2] - lo[2])], np.float32)
    colors = np.stack([lt.get_color_of_patch(i) for i in range(sc.numtriangles)]).astype(np.float32)
    img = api.traceScreen(p, cam, colors, True, True)
    assert img.sum() > 100 and abs(float(m.group(1)) - float(img.astype(np.float64).sum())) <= 1e-4 * float(img.sum())
    assert np.abs(pix.astype(np.int32) - (np.clip(img, 0, 1) * 255 + 0.5).astype(np.int32)).max() <= 1
    assert int(m.group(3)) == sc.numtriangles
    lt.close(); p.close()
