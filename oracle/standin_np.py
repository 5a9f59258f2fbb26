"""TEST / BENCH INFRASTRUCTURE (oracle side): the C3 stand-in scene WITHOUT the product library.

`bench.py --impl reference` must not load librtb200.so, so it cannot use the product's OBJ importer.  This module
writes the same seeded stand-in OBJ (generator shared with raytracer-group27_b200/rtb200/standin.py, numpy only) and
reads it back with a minimal numpy reader for exactly that file's form (`v`, `vn`, `f a//a b//b c//c`, one object, no
material library), followed by the reference's centerAndScaleToUnitMesh (src/mesh.cpp:162-188): centre = sequential
float32 sum of the vertex positions / count, scale = 1 / max distance to the centre, over the imported vertex list (three
vertices per face in face order).  tests/test_host.py checks the arrays against the product importer's.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MATERIAL_DTYPE = np.dtype([("kd", np.float32, 3), ("ks", np.float32, 3), ("shininess", np.float32), ("transparency", np.float32)])


def read_simple_obj(path: str):
    """(pos (n,9) float32, nrm (n,9) float32) of an OBJ made of `v`, `vn` and `f a//a b//b c//c` lines, centred and unit-scaled."""
    v, vn, f = [], [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                v.append(line.split()[1:4])
            elif line.startswith("vn "):
                vn.append(line.split()[1:4])
            elif line.startswith("f "):
                f.append([int(tok.split("/")[0]) for tok in line.split()[1:4]])
    v = np.array(v, dtype=np.float64).astype(np.float32)   # strtod, then rounded to float as the importer's parser does
    vn = np.array(vn, dtype=np.float64).astype(np.float32)
    f = np.array(f, dtype=np.int64) - 1
    p = v[f.reshape(-1)]   # vertex list of the imported mesh: three vertices per face in face order (identical vertices are not joined)
    acc = np.zeros(3, np.float32)
    for axis in range(3):   # sequential float32 sum, like std::accumulate over glm::vec3
        acc[axis] = np.add.accumulate(p[:, axis], dtype=np.float32)[-1]
    centre = acc / np.float32(len(p))
    d = p - centre
    # glm::length: sqrt((x*x + y*y) + z*z) in float
    maxd = np.sqrt(((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32)).max()
    vs = ((v - centre) / np.float32(maxd)).astype(np.float32)
    return vs[f].reshape(-1, 9), vn[f].reshape(-1, 9)


def _standin_generator():
    """rtb200/standin.py's numpy generator, loaded without importing the rtb200 package (which would bind the product library)."""
    path = os.path.join(ROOT, "raytracer-group27_b200", "rtb200", "standin.py")
    src = open(path).read().replace("from . import MATERIAL_DTYPE, SceneData, load_obj", "")
    mod = types.ModuleType("_standin_generator")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


class Scene:
    """The arrays oracle.Oracle.render takes."""

    def __init__(self, pos, nrm, mesh_id, mats, point_lights, sphere_lights):
        self.pos, self.nrm, self.mesh_id, self.mats, self.point_lights, self.sphere_lights = pos, nrm, mesh_id, mats, point_lights, sphere_lights


def dragon_standin_scene(cache_dir: str | None = None) -> Scene:
    gen = _standin_generator()
    cache_dir = cache_dir or os.environ.get("RTB200_CACHE", "/tmp/rtb200_cache")
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, f"dragon_standin_{gen.NU}x{gen.NV}_s{gen.SEED}.obj")
    if not os.path.exists(path):
        tmp = path + f".{os.getpid()}.tmp"
        gen.write_obj(tmp, gen.NU, gen.NV, gen.SEED)
        os.replace(tmp, path)
    pos, nrm = read_simple_obj(path)
    mats = np.zeros(1, MATERIAL_DTYPE)   # harness material override of SURVEY section 8(d), as rtb200/standin.py applies it
    mats["kd"], mats["ks"], mats["shininess"], mats["transparency"] = 0.6, 0.5, 0.0, 1.0
    return Scene(pos, nrm, np.zeros(len(pos), np.int32), mats, np.array([[-1, 1, -1, 1, 1, 1]], np.float32), np.zeros((0, 7), np.float32))
