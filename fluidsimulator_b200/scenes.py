"""Scene descriptions in the reference's scene-JSON schema, and a float32 emitter.

The reference ships five scene files (reference scene/*.json) that are all the same
shape: six axis-aligned planes of a box, one fluid block, one or two cube emitters.
This module holds them as a compact table, writes them back out in the reference
schema (so the reference loader, the C++ loader in csrc/host/scene.cpp and this
module all read the same thing), and emits the initial particle lattice exactly as
fluid::init_scene_from_json does (reference core/src/init.cpp:158-418): float32
loop counters accumulating `x += spacing` (init.cpp:298-300), x outermost, z
innermost, spacing = cbrtf(mass/density) (init.cpp:224-226).

Synthetic blocks for the scaling runs use the same schema.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import json
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from .capi import PbfParams

F32 = np.float32

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.cbrtf.restype = ctypes.c_float
_libm.cbrtf.argtypes = [ctypes.c_float]


def cbrtf(x) -> np.float32:
    """glibc cbrtf — the function std::cbrt(float) resolves to (init.cpp:20,225)."""
    return F32(_libm.cbrtf(float(F32(x))))


@dataclass
class Scene:
    name: str
    box: tuple            # far corner of the box; near corner is the origin
    cubes: list           # [(origin xyz, size xyz[, velocity xyz])]
    spheres: list = field(default_factory=list)  # [(origin xyz, radius[, velocity])]
    fluid: dict = field(default_factory=lambda: dict(
        particle_mass=1, density=6000, h=0.1, epsilon=300, n=4, k=0.0005, c=0.0005))
    external_forces: tuple = (0, -9.8, 0)
    box_min: tuple = (0, 0, 0)

    # plane order of every reference scene: +y floor, +x, +z, then -y, -x, -z at the far corner
    def planes(self):
        lo, hi = list(self.box_min), list(self.box)
        return [(lo, [0, 1, 0]), (lo, [1, 0, 0]), (lo, [0, 0, 1]),
                (hi, [0, -1, 0]), (hi, [-1, 0, 0]), (hi, [0, 0, -1])]

    def to_json_obj(self) -> dict:
        shapes = []
        for c in self.cubes:
            e = {"type": "cube", "origin": list(c[0]), "size": list(c[1])}
            if len(c) > 2:
                e["velocity"] = list(c[2])
            shapes.append(e)
        for s in self.spheres:
            e = {"type": "sphere", "origin": list(s[0]), "radius": s[1]}
            if len(s) > 2:
                e["velocity"] = list(s[2])
            shapes.append(e)
        fluid = dict(self.fluid)
        fluid["shape"] = shapes
        return {
            "collisions": [{"type": "plane", "point": list(p), "normal": list(nrm), "friction": 0.5}
                           for p, nrm in self.planes()],
            "fluid": fluid,
            "external_forces": list(self.external_forces),
        }

    def write_json(self, path) -> Path:
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        path.write_text(json.dumps(self.to_json_obj(), indent=1))
        return path


# The five reference scenes (reference scene/*.json; particle counts SURVEY.md §6).
SCENES = {
    "fluid_large": Scene("fluid_large", (2, 3, 2), [((0.25, 1, 0.25), (1.5, 1.5, 1.5))]),
    "fluid_double_side": Scene("fluid_double_side", (2, 2, 5),
                               [((0, 0, 0), (2, 2, 2)), ((0, 0, 3), (2, 2, 2))]),
    "fluid_double_dem": Scene("fluid_double_dem", (4, 5, 4),
                              [((0, 1, 0), (2, 3, 2)), ((2, 1, 2), (2, 3, 2))]),
    "fluid_xlarge": Scene("fluid_xlarge", (5, 5, 5), [((1, 1, 1), (3, 3, 3))]),
    "fluid_million": Scene("fluid_million", (8, 8, 8), [((1, 2.3, 1), (5.5, 5.5, 5.5))]),
}


def synthetic_block(name: str, size_xyz, origin=(1, 1, 1), margin=1.0) -> Scene:
    """One cuboid of fluid with planes `margin` outside it."""
    box = tuple(float(o + s + margin) for o, s in zip(origin, size_xyz))
    return Scene(name, box, [(tuple(origin), tuple(size_xyz))])


def block_16m() -> Scene:
    """252^3 = 16 003 008 particles."""
    return Scene("block_16m", (16, 16, 16), [((1, 1, 1), (13.87, 13.87, 13.87))])


def weak_block(gpus: int) -> Scene:
    """~2 M particles per GPU, elongated along x."""
    return synthetic_block(f"weak_{gpus}", (6.94 * gpus, 6.94, 6.94))


def small_block(n_side: int = 12, name: str | None = None) -> Scene:
    """A small cube (n_side^3 particles, about) in a tight box — test-sized."""
    spacing = float(cbrtf(F32(1) / F32(6000)))
    size = spacing * n_side + 1e-3
    return Scene(name or f"small_{n_side}", (1.5, 2.0, 1.5), [((0.3, 0.4, 0.3), (size, size, size))])


def _axis(lo: np.float32, hi: np.float32, spacing: np.float32) -> np.ndarray:
    """for (float x = lo; x < hi; x += spacing) — float32 accumulation (init.cpp:298-300)."""
    vals = []
    x = F32(lo)
    while x < hi:
        vals.append(x)
        x = F32(x + spacing)
    return np.asarray(vals, dtype=F32)


def load_params(obj: dict, base: PbfParams | None = None):
    """fluid{...} / external_forces -> (Params, spacing) (init.cpp:183-242)."""
    p = base.copy() if base is not None else PbfParams.defaults()
    fl = obj["fluid"]
    if "particle_mass" in fl:
        p.particle_mass = F32(fl["particle_mass"])
    if "density" in fl:
        p.density = F32(fl["density"])
    if "h" in fl:
        p.h = F32(fl["h"])
    if "epsilon" in fl:
        p.epsilon = F32(fl["epsilon"])
    if "n" in fl:
        p.scorr_n = int(fl["n"])
    if "k" in fl:
        p.scorr_k = F32(fl["k"])
    if "c" in fl:
        p.visc_c = F32(fl["c"])
    if "external_forces" in obj:
        for k in range(3):
            p.external_force[k] = F32(obj["external_forces"][k])
    spacing = F32(0)
    from_mass = False
    if p.density > 0 and p.particle_mass > 0:
        spacing = cbrtf(F32(p.particle_mass) / F32(p.density))
        from_mass = True
    if spacing <= 0 and p.particle_radius > 0:
        spacing = F32(F32(p.particle_radius) * F32(2))
    if spacing <= 0:
        spacing = F32(0.02)
    if from_mass or p.particle_radius <= 0:
        p.particle_radius = F32(spacing * F32(0.5))
    if p.h <= 0:
        p.h = F32(F32(2.5) * spacing)
    return p, spacing


def emit(obj: dict, spacing: np.float32):
    """Particles of every shape, in file order (init.cpp:256-353).  Returns six
    float32 arrays (pos x,y,z, vel x,y,z)."""
    half = F32(spacing * F32(0.5))
    P = [[], [], []]
    V = [[], [], []]
    for shape in obj["fluid"]["shape"]:
        vel = [F32(v) for v in shape.get("velocity", (0, 0, 0))]
        if shape["type"] == "cube":
            o = [F32(v) for v in shape["origin"]]
            s = [F32(v) for v in shape["size"]]
            axes = [_axis(F32(o[k] + half), F32(o[k] + s[k]), spacing) for k in range(3)]
            gx, gy, gz = np.meshgrid(*axes, indexing="ij")  # x outermost, z innermost
            pts = [gx.ravel(), gy.ravel(), gz.ravel()]
        elif shape["type"] == "sphere":
            o = [F32(v) for v in shape["origin"]]
            r = F32(shape["radius"])
            axes = [_axis(F32(F32(o[k] - r) + half), F32(o[k] + r), spacing) for k in range(3)]
            gx, gy, gz = np.meshgrid(*axes, indexing="ij")
            dx, dy, dz = gx - o[0], gy - o[1], gz - o[2]
            keep = ((dx * dx + dy * dy) + dz * dz) < F32(r * r)
            pts = [gx[keep].ravel(), gy[keep].ravel(), gz[keep].ravel()]
        else:
            raise ValueError(f"Unsupported shape type: {shape['type']}")
        for k in range(3):
            P[k].append(pts[k].astype(F32))
            V[k].append(np.full(pts[k].shape[0], vel[k], dtype=F32))
    cat = lambda parts: (np.concatenate(parts) if parts else np.zeros(0, F32)).astype(F32)
    return [cat(P[0]), cat(P[1]), cat(P[2]), cat(V[0]), cat(V[1]), cat(V[2])]


def load_planes(obj: dict) -> np.ndarray:
    """collisions[] -> [P,4] float32 (nx, ny, nz, d), normalised, d = n.point (init.cpp:364-415)."""
    rows = []
    for e in obj.get("collisions", []):
        if e["type"] != "plane":
            raise ValueError(f"Unsupported collision type: {e['type']}")
        px, py, pz = (F32(v) for v in e["point"])
        nx, ny, nz = (F32(v) for v in e["normal"])
        len_sq = F32(F32(nx * nx + ny * ny) + nz * nz)
        inv = F32(F32(1) / np.sqrt(len_sq, dtype=F32))
        nxn, nyn, nzn = F32(nx * inv), F32(ny * inv), F32(nz * inv)
        d = F32(F32(nxn * px + nyn * py) + nzn * pz)
        rows.append((nxn, nyn, nzn, d))
    return np.asarray(rows, dtype=F32).reshape(-1, 4)


def load_scene(scene, base: PbfParams | None = None):
    """Scene | dict | path -> (PbfParams, planes[P,4], state6)."""
    if isinstance(scene, Scene):
        obj = scene.to_json_obj()
    elif isinstance(scene, dict):
        obj = scene
    else:
        obj = json.loads(Path(scene).read_text())
    params, spacing = load_params(obj, base)
    return params, load_planes(obj), emit(obj, spacing)
