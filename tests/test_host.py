"""CPU tests of the host side: the C-ABI library loads and exports what include/rt_b200.h declares, the OBJ/MTL
importer, the error behaviour, the tile sharding map, and the C++ drop-in headers (compiled and run without a GPU)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(rtb):
    lib = rtb.lib()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/rt_b200.h but not exported"
    bound = {s[0] for s in rtb.SYMBOLS}
    assert set(names) <= bound, f"python binding misses {set(names) - bound}"
    assert lib.rt_version().decode().startswith("rt_b200")


def test_checked_library_has_the_same_abi(rtb):
    """librtb200_checked.so (csrc/rt_types.h: RT_CHECKED; built by __graft_entry__.build()) is the same library with the kernels'
    bounds checks live: every function of the header, and it says which of the two it is."""
    import ctypes
    path = os.path.join(ROOT, "raytracer-group27_b200", "librtb200_checked.so")
    assert os.path.exists(path), "missing: __graft_entry__.build() makes it"
    chk = ctypes.CDLL(path)
    for n in declared_functions():
        assert hasattr(chk, n), f"{n} is declared in include/rt_b200.h but not exported by the checked build"
    assert chk.rt_checked_build() == 1 and rtb.lib().rt_checked_build() == 0
    assert len(rtb.CHECK_SITES) == 10


def test_no_cpu_fallback_without_gpu(rtb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rtb.RtError) as e:
        rtb.Context(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


OBJ = """# two objects, three materials, quads, a concave quad, no normals in the second object
mtllib t.mtl
o first
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vn 0 0 1
usemtl red
f 1//1 2//1 3//1 4//1
usemtl glass
f 1//1 3//1 4//1
o second
v 0 0 1
v 2 0 1
v 0.5 0.5 1
v 0 2 1
usemtl red
f 5 6 7 8
v 3 3 3
v 4 3 3
v 3 4 3
f 9 10 11
"""
MTL = """newmtl red
Kd 0.8 0.1 0.1
Ks 0.5 0.5 0.5
Ns 32
d 1
newmtl glass
Kd 0.1 0.1 0.1
Tr 0.75
"""


def test_obj_importer(rtb, tmp_path):
    (tmp_path / "t.obj").write_text(OBJ)
    (tmp_path / "t.mtl").write_text(MTL)
    sc = rtb.load_obj(str(tmp_path / "t.obj"))
    # objects come out last-first (the reference pops assimp's children from a stack); meshes inside an object in order
    assert list(np.bincount(sc.mesh_id)) == [3, 2, 1]          # second/red: quad + tri; first/red: quad; first/glass: tri
    assert len(sc.mats) == 3
    np.testing.assert_allclose(sc.mats["kd"][0], [0.8, 0.1, 0.1])
    assert sc.mats["shininess"][0] == 32 and sc.mats["transparency"][0] == 1
    np.testing.assert_allclose(sc.mats["transparency"][2], 0.25)   # Tr 0.75 -> opacity 0.25
    assert sc.mats["shininess"][2] == 0                            # importer default
    # the concave quad (corner 2 at (0.5,0.5) is reflex) is fanned from its concave corner: both triangles share it
    q = sc.pos[:2].reshape(2, 3, 3)
    assert np.allclose(q[0, 0], [0.5, 0.5, 1]) and np.allclose(q[1, 0], [0.5, 0.5, 1])
    # generated face normals for the object without vn: unit length, +-z for the planar quad
    n = sc.nrm[:2].reshape(-1, 3)
    assert np.allclose(np.abs(n[:, 2]), 1) and np.allclose(n[:, :2], 0)
    # normals given in the file are kept
    assert np.allclose(sc.nrm[3].reshape(3, 3), [[0, 0, 1]] * 3)
    # centre + scale: all corner positions end up inside the unit sphere, the farthest one on it
    scn = rtb.load_obj(str(tmp_path / "t.obj"), normalize=True)
    r = np.linalg.norm(scn.pos.reshape(-1, 3), axis=1)
    assert r.max() == pytest.approx(1.0, abs=1e-6)


def test_obj_importer_errors(rtb, tmp_path):
    with pytest.raises(rtb.RtError) as e:
        rtb.load_obj(str(tmp_path / "missing.obj"))
    assert e.value.code == rtb.RT_ERR_IO
    (tmp_path / "empty.obj").write_text("v 0 0 0\n")
    with pytest.raises(rtb.RtError):
        rtb.load_obj(str(tmp_path / "empty.obj"))
    (tmp_path / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nf 1 2 7\n")
    with pytest.raises(rtb.RtError):
        rtb.load_obj(str(tmp_path / "bad.obj"))


def test_importer_matches_fixture_geometry(rtb):
    """Where the reference's data directory is reachable (this container), the importer reproduces the geometry stored in
    the golden fixtures bit for bit."""
    data = "/root/reference/data"
    if not os.path.isdir(data):
        pytest.skip("reference data not present on this box")
    from util import Golden
    for fixture, obj in (("cornell_c1_256", "CornellBox-Mirror-Rotated.obj"), ("monkey_192", "monkey-rotated.obj"), ("teapot_c2_256x144", "teapot.obj")):
        g = Golden(fixture)
        sc = rtb.load_obj(os.path.join(data, obj), normalize=True)
        assert np.array_equal(sc.pos.view(np.int32), g.scene.pos.view(np.int32))
        assert np.array_equal(sc.mesh_id, g.scene.mesh_id)
        assert np.array_equal(sc.mats["transparency"], g.scene.mats["transparency"])


def test_owner_map_partitions_the_image(rtb):
    for (w, h, world) in ((3840, 2160, 8), (70, 45, 4), (33, 17, 3), (16, 16, 2)):
        m = rtb.owner_map(w, h, world)
        assert m.shape == (h, w) and m.min() >= 0 and m.max() < world
        tx, ty = rtb.tile_grid(w, h)
        assert sum(rtb.local_tile_count(w, h, r, world) for r in range(world)) == tx * ty
        # bottom-left tile belongs to rank 0 (tile id 0), and Screen rows are flipped
        assert m[h - 1, 0] == 0


CPP_TEST = r"""
#include "scene.h"
#include "screen.h"
#include "trackball.h"
#include <cmath>
#include <cstdio>
int main(int argc, char** argv) {
    Window window{"t", glm::ivec2(64, 32), OpenGLVersion::GL2};
    Trackball camera{&window, glm::radians(50.0f), 3.0f};
    camera.setCamera(glm::vec3(0.0f), glm::radians(glm::vec3(20.0f, 20.0f, 0.0f)), 3.0f);
    const glm::vec3 p = camera.position();
    if (std::fabs(glm::length(p) - 3.0f) > 1e-5f) return 1;
    const Ray r = camera.generateRay(glm::vec2(0.0f, 0.0f));
    // the central ray looks at the look-at point
    const glm::vec3 to = glm::normalize(glm::vec3(0.0f) - p);
    if (glm::length(r.direction - to) > 1e-5f) return 2;
    Screen screen(glm::ivec2(4, 2));
    screen.clear(glm::vec3(0.0f));
    screen.setPixel(0, 0, glm::vec3(2.0f, 0.5f, -1.0f));   // bottom-left -> last row, first column
    if (!(screen.pixels()[4] == glm::vec3(2.0f, 0.5f, -1.0f))) return 3;
    screen.writeBitmapToFile(argv[1]);
    Scene scene = loadScene(Cube, argv[2]);
    if (scene.meshes.size() != 6 || scene.pointLights.size() != 1 || scene.spotLight.size() != 1) return 4;
    std::printf("%zu meshes\n", scene.meshes.size());
    return 0;
}
"""


def test_cpp_host_api_compiles_and_runs(tmp_path):
    host = os.path.join(ROOT, "raytracer-group27_b200", "host")
    lib = os.path.join(ROOT, "raytracer-group27_b200")
    (tmp_path / "t.cpp").write_text(CPP_TEST)
    (tmp_path / "cube.obj").write_text("mtllib cube.mtl\n" + "".join(f"v {x} {y} {z}\n" for x in (0, 1) for y in (0, 1) for z in (0, 1))
                                       + "".join(f"g face{i}\nusemtl m{i % 2}\nf {a} {b} {c}\n" for i, (a, b, c) in enumerate([(1, 2, 3), (2, 3, 4), (5, 6, 7), (6, 7, 8), (1, 5, 2), (3, 7, 4)])))
    (tmp_path / "cube.mtl").write_text("newmtl m0\nKd 1 0 0\nnewmtl m1\nKd 0 1 0\n")
    exe = tmp_path / "t"
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{host}", f"-I{os.path.join(ROOT, 'include')}", str(tmp_path / "t.cpp"), "-o", str(exe),
                        f"-L{lib}", "-lrtb200", f"-Wl,-rpath,{lib}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(tmp_path / "out.bmp"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    bmp = (tmp_path / "out.bmp").read_bytes()
    assert bmp[:2] == b"BM" and len(bmp) == 54 + 4 * 2 * 4
    # bottom-up BGRA rows: the first stored row is the image's last row, whose first pixel was set to (2, .5, -1) -> clamped (255, 127, 0)
    assert bmp[54:58] == bytes([0, 127, 255, 255])


def _mini_obj(path):
    """Independent, minimal OBJ reader for cross-checking: positions, and every face fanned to (n - 2) triangles; returns
    the set of position triples per triangle regardless of order (the importer may fan from another corner)."""
    v, faces = [], []
    with open(path, errors="replace") as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                v.append(tuple(float(x) for x in p[1:4]))
            elif p[0] == "f":
                idx = [int(tok.split("/")[0]) for tok in p[1:]]
                faces.append([i - 1 if i > 0 else len(v) + i for i in idx])
    return np.array(v, np.float64), faces


def test_importer_on_every_reference_obj(rtb):
    """All OBJ files the reference ships (its presets and the students' scenes): the importer loads each, produces exactly
    one triangle per fan step of every polygon, uses only corner positions the file declares, keeps mesh ids in range and
    unit normals; normalize=true centres and scales into the unit sphere (mesh.cpp:164-188)."""
    data = "/root/reference/data"
    if not os.path.isdir(data):
        pytest.skip("reference data not present on this box")
    names = sorted(n for n in os.listdir(data) if n.endswith(".obj"))
    assert len(names) >= 20
    for n in names:
        v, faces = _mini_obj(os.path.join(data, n))
        sc = rtb.load_obj(os.path.join(data, n), normalize=False)
        want_tris = sum(len(f) - 2 for f in faces if len(f) >= 3)
        assert sc.n_tris == want_tris, f"{n}: {sc.n_tris} triangles, the file's polygons fan to {want_tris}"
        assert sc.mesh_id.min() == 0 and sc.mesh_id.max() == len(sc.mats) - 1 and np.all(np.diff(sc.mesh_id) >= 0), n
        corners = sc.pos.reshape(-1, 3)
        declared = {tuple(np.float32(c)) for c in v}
        assert {tuple(c) for c in corners[:: max(1, len(corners) // 5000)]} <= declared, f"{n}: a corner is not a vertex of the file"
        # per polygon area: the triangles of the importer cover the same area as the fan of the file (planar polygons)
        def areas(tri):
            return 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
        fan = np.array([[v[f[0]], v[f[k]], v[f[k + 1]]] for f in faces if len(f) >= 3 for k in range(1, len(f) - 1)])
        quads_convex = abs(areas(fan).sum() - areas(sc.pos.reshape(-1, 3, 3).astype(np.float64)).sum()) <= 1e-3 * areas(fan).sum()
        assert quads_convex or any(len(f) > 3 for f in faces), n      # concave polygons are fanned from their reflex corner
        nn = np.linalg.norm(sc.nrm.reshape(-1, 3), axis=1)
        assert np.all((np.abs(nn - 1) < 1e-3) | (nn == 0) | ~np.isfinite(nn)), f"{n}: normals are not unit length"
        scn = rtb.load_obj(os.path.join(data, n), normalize=True)
        assert np.linalg.norm(scn.pos.reshape(-1, 3), axis=1).max() == pytest.approx(1.0, abs=2e-6), n


TEX_OBJ = """mtllib t.mtl
v 0 0 0
v 1 0 0
v 0 1 0
vt 0.25 0.5
vt 1.5 -0.25
vt 0 1
usemtl m
f 1/1 2/2 3/3
"""


def _load_textured(rtb, tmp_path, png_name):
    (tmp_path / "t.obj").write_text(TEX_OBJ)
    (tmp_path / "t.mtl").write_text(f"newmtl m\nKd 1 1 1\nmap_Kd {png_name}\n")
    return rtb.load_obj(str(tmp_path / "t.obj"))


def test_png_textures_decode_like_the_reference_image_reader(rtb, tmp_path):
    """Image::Image asks its image reader for 8-bit RGB (src/image.cpp:45): palette entries expanded, 16-bit samples reduced to
    their high byte, alpha dropped; files with fewer than 3 channels are refused (src/image.cpp:47-50).  Every scanline filter
    type and split IDAT chunks are covered by the writer."""
    from util import write_png
    rng = np.random.default_rng(4)
    rgb = rng.integers(0, 256, (13, 9, 3))
    write_png(tmp_path / "rgb8.png", rgb, 2)
    sc = _load_textured(rtb, tmp_path, "rgb8.png")
    assert np.array_equal(sc.textures[0], rgb) and list(sc.mesh_tex) == [0]
    np.testing.assert_array_equal(sc.uv, [[0.25, 0.5, 1.5, -0.25, 0, 1]])        # vt kept as written (no v flip)
    rgba = rng.integers(0, 256, (6, 11, 4))
    write_png(tmp_path / "rgba8.png", rgba, 6)
    assert np.array_equal(_load_textured(rtb, tmp_path, "rgba8.png").textures[0], rgba[..., :3])
    rgb16 = rng.integers(0, 65536, (5, 4, 3))
    write_png(tmp_path / "rgb16.png", rgb16, 2, depth=16)
    assert np.array_equal(_load_textured(rtb, tmp_path, "rgb16.png").textures[0], rgb16 >> 8)
    for depth in (1, 2, 4, 8):
        pal = rng.integers(0, 256, (1 << depth, 3))
        idx = rng.integers(0, 1 << depth, (7, 10, 1))
        write_png(tmp_path / f"pal{depth}.png", idx, 3, depth=depth, palette=pal)
        assert np.array_equal(_load_textured(rtb, tmp_path, f"pal{depth}.png").textures[0], pal[idx[..., 0]]), depth
    # grey: one channel -> refused like the reference does; loadMesh reports it and keeps the material's kd
    write_png(tmp_path / "grey.png", rng.integers(0, 256, (4, 4, 1)), 0)
    sc = _load_textured(rtb, tmp_path, "grey.png")
    assert sc.textures == [] and list(sc.mesh_tex) == [-1]
    # a file the MTL names but that is not there (the reference snapshot lacks two of its texture blobs)
    sc = _load_textured(rtb, tmp_path, "nowhere.png")
    assert sc.textures == [] and sc.n_tris == 1


def test_visible_rect_never_culls_a_ray_that_meets_the_box(rtb):
    """rt_visible_rect (the rectangle rt_render uses to answer primary rays as misses without tracing them) must be
    conservative: for random cameras and boxes, no camera ray of a pixel outside the rectangle — at the pixel corner or at any
    sub-pixel sample position (offsets below one pixel) — meets the box; and the rectangle is not trivially the whole image."""
    rng = np.random.default_rng(11)

    def rays(cam, w, h, px, py):  # Trackball::generateRay (framework/src/trackball.cpp:87-98) in float64
        ex, ey, ez = (0.5 * float(v) for v in cam.euler)
        cx, cy, cz, sx, sy, sz = np.cos(ex), np.cos(ey), np.cos(ez), np.sin(ex), np.sin(ey), np.sin(ez)
        qw, qv = cx * cy * cz + sx * sy * sz, np.array([sx * cy * cz - cx * sy * sz, cx * sy * cz + sx * cy * sz, cx * cy * sz - sx * sy * cz])

        def rot(v):
            uv = np.cross(qv, v)
            return v + (uv * qw + np.cross(qv, uv)) * 2.0
        o = np.array(cam.look_at, np.float64) + rot(np.array([0.0, 0.0, -float(cam.dist)]))
        half_h = np.tan(float(cam.fovy) / 2.0)
        half_w = w / h * half_h
        nx, ny = px / w * 2.0 - 1.0, py / h * 2.0 - 1.0
        c = np.stack([-nx * half_w, ny * half_h, np.ones_like(nx)], axis=-1)
        c /= np.linalg.norm(c, axis=-1, keepdims=True)
        return o, rot(c)

    def hits(o, d, lo, hi):  # slab test, origin outside or inside, t >= 0
        with np.errstate(divide="ignore", invalid="ignore"):
            t0, t1 = (lo - o) / d, (hi - o) / d
        tn, tf = np.minimum(t0, t1).max(axis=-1), np.maximum(t0, t1).min(axis=-1)
        return (tn <= tf) & (tf >= 0)

    culled_something = 0
    for trial in range(60):
        w, h = int(rng.integers(40, 400)), int(rng.integers(40, 300))
        centre = rng.uniform(-1.5, 1.5, 3)
        half = rng.uniform(0.01, 1.2, 3)
        lo, hi = centre - half, centre + half
        cam = rtb.make_camera(look_at=tuple(rng.uniform(-0.5, 0.5, 3)), euler_deg=tuple(rng.uniform(-180, 180, 3) * (1, 1, 0.2)),
                              dist=float(rng.uniform(0.3, 8.0)), fovy_deg=float(rng.uniform(20, 100)))
        x0, x1, y0, y1 = rtb.visible_rect(cam, w, h, lo, hi)
        assert 0 <= x0 <= x1 <= w and 0 <= y0 <= y1 <= h
        py, px = np.mgrid[0:h, 0:w]
        outside = (px < x0) | (px >= x1) | (py < y0) | (py >= y1)
        culled_something += int(outside.any())
        if not outside.any():
            continue
        for dx in (-0.9, 0.0, 0.9):
            for dy in (-0.9, 0.0, 0.9):
                o, d = rays(cam, w, h, px[outside] + dx, py[outside] + dy)
                assert not hits(o, d, lo, hi).any(), (trial, dx, dy)
        # and it is tight to a few pixels: the rays of the rectangle's four edges (3 pixels further in) do meet the box somewhere
        if x1 - x0 > 12 and y1 - y0 > 12 and x0 > 0 and x1 < w and y0 > 0 and y1 < h:
            ys, xs = np.mgrid[y0:y1, x0:x1]
            o, d = rays(cam, w, h, xs.astype(np.float64), ys.astype(np.float64))
            m = hits(o, d, lo, hi)
            assert m[:, :8].any() and m[:, -8:].any() and m[:8, :].any() and m[-8:, :].any(), trial
    assert culled_something >= 20
    # a corner behind the eye: the whole image
    assert rtb.visible_rect(rtb.make_camera(dist=0.5), 64, 48, (-1, -1, -1), (1, 1, 1)) == (0, 64, 0, 48)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the reference's own CPU code on the host cores; the one place besides the tests that runs the
    oracle) owes the driver exactly ONE line on stdout, a JSON object with the contract's keys; everything else goes to stderr."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    # the job description has the GPU arm's keys (the sample lives under cpu_baseline), and the product library stays out of this arm
    assert set(d["config"]) == {"workload", "bvh", "sharding", "gather", "l2"} and "sample" in d["cpu_baseline"]
    assert d["native_libraries_loaded"] and all(p.startswith("oracle" + os.sep) for p in d["native_libraries_loaded"]), d["native_libraries_loaded"]


def test_reference_arm_scene_equals_the_importers(rtb):
    """oracle/standin_np.py (numpy reader + centerAndScaleToUnitMesh, used by bench.py's reference arm so that it need not load
    the product library) gives bit for bit the arrays the product's OBJ importer gives for the stand-in."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import standin_np
    from rtb200 import standin
    a, b = standin_np.dragon_standin_scene(), standin.dragon_standin_scene()
    assert np.array_equal(a.pos, b.pos) and np.array_equal(a.nrm, b.nrm) and np.array_equal(a.mesh_id, b.mesh_id)
    assert a.mats.tobytes() == np.ascontiguousarray(b.mats).tobytes() and np.array_equal(a.point_lights, b.point_lights)
