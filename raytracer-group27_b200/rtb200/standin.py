"""Procedural stand-in for the reference's missing data/dragon.obj (listed in .MISSING_LARGE_BLOBS).

The course's dragon has about 87 K triangles (assignment.html section 4.4).  Until the real file is supplied, the
dragon configs run on a seeded, closed, bumpy (2,3) torus-knot tube with 86 880 triangles and smooth vertex
normals: similar triangle count and depth complexity (the tube passes in front of itself several times from the
default camera).  EVERY number measured on it must be labelled "stand-in".  The mesh is written as a Wavefront OBJ
and read back through the library's own importer so it takes exactly the path any OBJ takes (centre + scale to
the unit sphere, src/mesh.cpp:164-188).
"""
from __future__ import annotations

import os

import numpy as np

from . import MATERIAL_DTYPE, SceneData, load_obj

SEED = 27
NU, NV = 724, 60  # 724 * 60 * 2 = 86 880 triangles


def knot_mesh(nu: int = NU, nv: int = NV, seed: int = SEED):
    """Vertices (nu*nv, 3), vertex normals, and triangle indices (2*nu*nv, 3) of the displaced torus-knot tube."""
    rng = np.random.default_rng(seed)
    u = np.linspace(0.0, 2.0 * np.pi, nu, endpoint=False)
    v = np.linspace(0.0, 2.0 * np.pi, nv, endpoint=False)
    p, q = 2, 3

    def curve(t):
        r = 2.0 + np.cos(q * t)
        return np.stack([r * np.cos(p * t), r * np.sin(p * t), -np.sin(q * t)], axis=-1)

    c = curve(u)
    eps = 1e-4
    tangent = curve(u + eps) - curve(u - eps)
    tangent /= np.linalg.norm(tangent, axis=1, keepdims=True)
    # frame by projecting the direction away from the knot's axis (never parallel to the tangent for a (2,3) knot)
    ref = np.stack([np.cos(p * u), np.sin(p * u), np.zeros_like(u)], axis=-1)
    n1 = ref - (ref * tangent).sum(1, keepdims=True) * tangent
    n1 /= np.linalg.norm(n1, axis=1, keepdims=True)
    n2 = np.cross(tangent, n1)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    # smooth low-frequency lobes + seeded band-limited bumps (periodic in both parameters)
    radius = 0.42 + 0.09 * np.sin(7 * uu + 3 * vv) + 0.05 * np.sin(13 * vv + 5 * uu)
    for _ in range(24):
        fu, fv = int(rng.integers(3, 40)), int(rng.integers(1, 9))
        radius += 0.012 * rng.standard_normal() * np.sin(fu * uu + fv * vv + rng.uniform(0, 2 * np.pi))
    pos = c[:, None, :] + radius[..., None] * (np.cos(vv)[..., None] * n1[:, None, :] + np.sin(vv)[..., None] * n2[:, None, :])
    # smooth normals from the parametric derivatives (central differences on the periodic grid)
    du = np.roll(pos, -1, 0) - np.roll(pos, 1, 0)
    dv = np.roll(pos, -1, 1) - np.roll(pos, 1, 1)
    nrm = np.cross(dv, du)
    nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
    idx = np.arange(nu * nv).reshape(nu, nv)
    a, b = idx, np.roll(idx, -1, 0)
    d, e = np.roll(idx, -1, 1), np.roll(np.roll(idx, -1, 0), -1, 1)
    tris = np.concatenate([np.stack([a, b, e], -1).reshape(-1, 3), np.stack([a, e, d], -1).reshape(-1, 3)], 0)
    return pos.reshape(-1, 3), nrm.reshape(-1, 3), tris


def write_obj(path: str, nu: int = NU, nv: int = NV, seed: int = SEED) -> int:
    """Write the stand-in as OBJ (one object, no mtllib -> importer default material).  Returns the triangle count."""
    pos, nrm, tris = knot_mesh(nu, nv, seed)
    with open(path, "w") as f:
        f.write("# procedural stand-in for dragon.obj (seed %d) - NOT the reference asset\no dragon_standin\n" % seed)
        np.savetxt(f, pos, fmt="v %.6f %.6f %.6f")
        np.savetxt(f, nrm, fmt="vn %.6f %.6f %.6f")
        t = tris + 1
        np.savetxt(f, np.stack([t[:, 0], t[:, 0], t[:, 1], t[:, 1], t[:, 2], t[:, 2]], 1), fmt="f %d//%d %d//%d %d//%d")
    return tris.shape[0]


def _override_material(sc: SceneData) -> SceneData:
    mats = np.zeros(len(sc.mats), MATERIAL_DTYPE)
    mats["kd"], mats["ks"], mats["shininess"], mats["transparency"] = 0.6, 0.5, 0.0, 1.0
    sc.mats = mats
    sc.point_lights = np.array([[-1, 1, -1, 1, 1, 1]], np.float32)  # src/scene.cpp:72
    return sc


def dragon_standin_scene(cache_dir: str | None = None, nu: int = NU, nv: int = NV) -> SceneData:
    """The C3 scene: stand-in mesh through the OBJ importer (centred, unit-scaled), harness material override
    kd 0.6 / ks 0.5 / shininess 0 / opaque and the Dragon preset's light (src/scene.cpp:72)."""
    cache_dir = cache_dir or os.environ.get("RTB200_CACHE", "/tmp/rtb200_cache")
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, f"dragon_standin_{nu}x{nv}_s{SEED}.obj")
    if not os.path.exists(path):
        tmp = path + f".{os.getpid()}.tmp"
        write_obj(tmp, nu, nv, SEED)
        os.replace(tmp, path)
    return _override_material(load_obj(path, normalize=True))


def dragon_lattice_scene(grid: int = 8, cache_dir: str | None = None, nu: int = NU, nv: int = NV) -> SceneData:
    """The C5 scene: grid^3 translated copies of the C3 mesh (spacing = 2 unit-sphere diameters) flattened into one
    mesh, then re-centred and re-scaled to the unit sphere.  Built directly as arrays (a 44 M-triangle OBJ would be
    gigabytes of text); the centre is the float64 mean of all corner positions instead of mesh.cpp's sequential
    float32 sum — a documented deviation that only moves the whole lattice by a few ulp."""
    base = dragon_standin_scene(cache_dir, nu, nv)
    n = base.n_tris
    spacing = 4.0  # unit sphere has diameter 2
    offs = np.array([(i, j, k) for i in range(grid) for j in range(grid) for k in range(grid)], np.float64) * spacing
    corners = base.pos.reshape(n, 3, 3).astype(np.float64)
    centre = corners.reshape(-1, 3).mean(0) + offs.mean(0)
    pos = np.empty((n * len(offs), 9), np.float32)
    maxd = 0.0
    for c, off in enumerate(offs):
        blk = corners + off - centre
        maxd = max(maxd, float(np.sqrt((blk ** 2).sum(-1)).max()))
        pos[c * n:(c + 1) * n] = blk.reshape(n, 9).astype(np.float32)
    pos /= np.float32(maxd)
    nrm = np.tile(base.nrm, (len(offs), 1))
    sc = SceneData(pos, nrm, np.zeros(n * len(offs), np.int32), base.mats.copy())
    return _override_material(sc)
