"""ctypes binding of the C ABI in include/pbf_b200.h (libpbf_b200.so).

Python here is plumbing for tests and bench.py only: the product is the shared
library plus the C++ host layer in fluidsimulator_b200/csrc/host.  There is no CPU
fallback — if the library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libpbf_b200.so"

PBF_MODE_STRICT = 0
PBF_MODE_FAST = 1
PBF_BRICK_OFF, PBF_BRICK_PERSISTENT, PBF_BRICK_PER_CTA = 0, 1, 2

SCRATCH_IDS = {
    "pred_x": 0, "pred_y": 1, "pred_z": 2,
    "delta_x": 3, "delta_y": 4, "delta_z": 5,
    "lambda": 6, "rho": 7,
    "dv_x": 8, "dv_y": 9, "dv_z": 10,
    "omega_x": 11, "omega_y": 12, "omega_z": 13, "omega_mag": 14,
    "eta_x": 15, "eta_y": 16, "eta_z": 17,
}

STAGES = ["predict", "sort", "cells", "neighbors", "lambda", "delta", "xsph",
          "vort_omega", "vort_apply", "finalize", "exchange"]

PBF_COMM_ID_BYTES = 128


class PbfParams(C.Structure):
    """POD mirror of fluid::Params (reference core/include/fluid/core.h:11-43)."""

    _fields_ = [
        ("dt", C.c_float),
        ("density", C.c_float),
        ("particle_mass", C.c_float),
        ("h", C.c_float),
        ("particle_radius", C.c_float),
        ("epsilon", C.c_float),
        ("solver_iterations", C.c_int32),
        ("neighbor_reserve_factor", C.c_float),
        ("use_uniform_grid", C.c_int32),
        ("enable_scorr", C.c_int32),
        ("enable_xsph", C.c_int32),
        ("enable_vorticity", C.c_int32),
        ("scorr_k", C.c_float),
        ("scorr_n", C.c_int32),
        ("scorr_dq_coeff", C.c_float),
        ("visc_c", C.c_float),
        ("plane_restitution", C.c_float),
        ("plane_friction", C.c_float),
        ("vort_epsilon", C.c_float),
        ("vort_norm_eps", C.c_float),
        ("external_force", C.c_float * 3),
    ]

    @staticmethod
    def defaults() -> "PbfParams":
        """fluid::Params defaults (core.h:12-43)."""
        p = PbfParams()
        p.dt = np.float32(1.0) / np.float32(60.0)
        p.density = 6000.0
        p.particle_mass = 0.0
        p.h = 0.0
        p.particle_radius = 0.01
        p.epsilon = 600.0
        p.solver_iterations = 4
        p.neighbor_reserve_factor = 1.5
        p.use_uniform_grid = 1
        p.enable_scorr = 0
        p.enable_xsph = 0
        p.enable_vorticity = 0
        p.scorr_k = 0.00005
        p.scorr_n = 4
        p.scorr_dq_coeff = 0.3
        p.visc_c = 0.0002
        p.plane_restitution = 0.0
        p.plane_friction = 0.0
        p.vort_epsilon = 0.5
        p.vort_norm_eps = 1e-6
        p.external_force[0] = 0.0
        p.external_force[1] = -9.8
        p.external_force[2] = 0.0
        return p

    def copy(self) -> "PbfParams":
        q = PbfParams()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(PbfParams))
        return q

    def as_dict(self) -> dict:
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if name == "external_force" else v
        return out


_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


def fptr(a):
    """float32 C-contiguous numpy array -> float* (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_i32p)


# every symbol include/pbf_b200.h declares: name -> (restype, argtypes)
ABI = {
    "pbf_default_params": (None, [C.POINTER(PbfParams)]),
    "pbf_abi_version": (C.c_int, []),
    "pbf_device_count": (C.c_int, [C.POINTER(C.c_char_p)]),
    "pbf_create": (C.c_void_p, [C.c_int, C.c_size_t]),
    "pbf_destroy": (None, [C.c_void_p]),
    "pbf_last_error": (C.c_char_p, [C.c_void_p]),
    "pbf_set_params": (C.c_int, [C.c_void_p, C.POINTER(PbfParams)]),
    "pbf_set_planes": (C.c_int, [C.c_void_p, C.c_int, _f32p, _f32p, _f32p, _f32p]),
    "pbf_set_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "pbf_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pbf_set_graph": (C.c_int, [C.c_void_p, C.c_int]),
    "pbf_set_brick": (C.c_int, [C.c_void_p, C.c_int]),
    "pbf_strict_exact": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p)]),
    "pbf_brick_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "pbf_upload": (C.c_int, [C.c_void_p, C.c_size_t] + [_f32p] * 6),
    "pbf_download": (C.c_int, [C.c_void_p] + [_f32p] * 6),
    "pbf_batches_retried": (C.c_uint64, [C.c_void_p]),
    "pbf_snapshot_begin": (C.c_int, [C.c_void_p, C.c_int]),
    "pbf_snapshot_wait": (C.c_int, [C.c_void_p, C.c_int] + [C.POINTER(_f32p)] * 3 + [C.POINTER(C.c_size_t), C.POINTER(C.c_float)]),
    "pbf_step": (C.c_int, [C.c_void_p, C.c_int]),
    "pbf_host_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "pbf_host_unregister": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pbf_step_host": (C.c_int, [C.c_void_p, C.c_size_t] + [_f32p] * 6 + [C.c_int]),
    "pbf_count": (C.c_size_t, [C.c_void_p]),
    "pbf_time": (C.c_float, [C.c_void_p]),
    "pbf_set_time": (C.c_int, [C.c_void_p, C.c_float]),
    "pbf_debug_sizes": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "pbf_debug_grid": (C.c_int, [C.c_void_p] + [_i32p] * 7),
    "pbf_debug_neighbors": (C.c_int, [C.c_void_p, _i32p, _i32p]),
    "pbf_debug_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "pbf_debug_scratch": (C.c_int, [C.c_void_p, C.c_int, _f32p]),
    "pbf_stage_name": (C.c_char_p, [C.c_int]),
    "pbf_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "pbf_profile_reset": (C.c_int, [C.c_void_p]),
    "pbf_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "pbf_launch_count": (C.c_uint64, [C.c_void_p]),
    "pbf_comm_unique_id": (C.c_int, [C.c_void_p]),
    "pbf_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pbf_slab_upload": (C.c_int, [C.c_void_p, C.c_size_t] + [_f32p] * 6),
    "pbf_slab_owned": (C.c_size_t, [C.c_void_p]),
    "pbf_slab_download": (C.c_int, [C.c_void_p, _i64p] + [_f32p] * 6),
    "pbf_slab_upload_owned": (C.c_int, [C.c_void_p, C.c_size_t, _i64p] + [_f32p] * 6),
    "pbf_slab_set_p2p": (C.c_int, [C.c_void_p, C.c_int]),
    "pbf_slab_transport": (C.c_int, [C.c_void_p]),
    "pbf_slab_payload": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "pbf_slab_plan": (C.c_int, [C.c_size_t, _f32p, C.c_float, C.c_int, _i32p]),
    "pbf_slab_plan_hist": (C.c_int, [C.POINTER(C.c_uint64), C.c_int32, C.c_int32, C.c_int, C.c_float, _i32p]),
    "pbf_slab_cuts": (C.c_int, [C.c_void_p, _i32p, _i32p]),
    "pbf_slab_set_cuts": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "pbf_slab_set_rebalance": (C.c_int, [C.c_void_p, C.c_float]),
    "pbf_slab_rebalance_count": (C.c_uint64, [C.c_void_p]),
    "pbf_slab_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _i32p, _i32p]),
    "pbf_group_create": (C.c_void_p, [C.POINTER(C.c_void_p), C.c_int]),
    "pbf_group_destroy": (None, [C.c_void_p]),
    "pbf_group_upload": (C.c_int, [C.c_void_p, C.c_size_t] + [_f32p] * 6),
    "pbf_group_step": (C.c_int, [C.c_void_p, C.c_int]),
    "pbf_group_download": (C.c_int, [C.c_void_p] + [_f32p] * 6),
    "pbf_group_count": (C.c_size_t, [C.c_void_p]),
}

_lib = None


def _preload_bundled_nccl() -> None:
    """libpbf_b200.so needs libnccl.so.2; PyTorch ships a newer NCCL than the system one and resolves
    its own symbols against whichever libnccl.so.2 the process loaded FIRST.  Loading torch's copy
    before ours (same soname, newer minor version) keeps `import torch` working whether it happens
    before or after this library is opened.  Harmless when PyTorch / its NCCL wheel is absent."""
    try:
        import glob
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []) or []:
            for f in sorted(glob.glob(os.path.join(d, "lib", "libnccl.so*"))):
                C.CDLL(f, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """dlopen libpbf_b200.so and type every exported entry point.  Raises if the
    library or any declared symbol is missing (there is no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    if path is None and os.environ.get("PBF_B200_LIB"):
        path = os.environ["PBF_B200_LIB"]
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise RuntimeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C fluidsimulator_b200/csrc`). There is no CPU fallback.")
    _preload_bundled_nccl()
    lib = C.CDLL(str(p), mode=C.RTLD_GLOBAL)
    for name, (restype, argtypes) in ABI.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    if path is None:
        _lib = lib
    return lib


class PbfError(RuntimeError):
    pass


class Solver:
    """Thin object wrapper over one pbf_ctx (one GPU / one slab)."""

    def __init__(self, device: int = 0, capacity: int = 0, mode: int = PBF_MODE_STRICT):
        self.lib = load_library()
        self.ctx = self.lib.pbf_create(device, capacity)
        if not self.ctx:
            msg = self.lib.pbf_last_error(None)
            raise PbfError((msg or b"pbf_create failed").decode())
        self._check(self.lib.pbf_set_mode(self.ctx, mode))
        self.n = 0

    def _check(self, rc: int):
        if rc != 0:
            msg = self.lib.pbf_last_error(self.ctx)
            raise PbfError(f"pbf error {rc}: {(msg or b'?').decode()}")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.pbf_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration -----------------------------------------------------
    def set_params(self, params: PbfParams):
        self._check(self.lib.pbf_set_params(self.ctx, C.byref(params)))

    def set_planes(self, planes: np.ndarray):
        """planes: float32 [P,4] rows (nx, ny, nz, d)."""
        planes = np.ascontiguousarray(planes, dtype=np.float32).reshape(-1, 4)
        cols = [np.ascontiguousarray(planes[:, k]) for k in range(4)]
        self._check(self.lib.pbf_set_planes(self.ctx, planes.shape[0], *[fptr(c) for c in cols]))

    def set_mode(self, mode: int):
        self._check(self.lib.pbf_set_mode(self.ctx, mode))

    def set_stream(self, stream_ptr: int | None):
        self._check(self.lib.pbf_set_stream(self.ctx, C.c_void_p(stream_ptr or 0)))

    def set_graph(self, enabled: bool):
        self._check(self.lib.pbf_set_graph(self.ctx, int(enabled)))

    def strict_exact(self):
        """(True, None) when STRICT results are bit-identical to the reference for the current
        parameters, else (False, reason)."""
        why = C.c_char_p()
        rc = self.lib.pbf_strict_exact(self.ctx, C.byref(why))
        self._check(min(rc, 0))
        return rc == 1, (why.value.decode() if why.value else None)

    def set_brick(self, mode: int):
        """PBF_BRICK_OFF (global-gather kernels, default), PBF_BRICK_PERSISTENT or PBF_BRICK_PER_CTA
        (shared-memory staged bricks).  Same results bit for bit."""
        self._check(self.lib.pbf_set_brick(self.ctx, int(mode)))

    def brick_status(self) -> dict:
        """Whether the last batch ran on the brick path, batches replayed without it, largest tile."""
        fb, mt = C.c_uint64(), C.c_uint32()
        rc = self.lib.pbf_brick_status(self.ctx, C.byref(fb), C.byref(mt))
        self._check(min(rc, 0))
        return {"active": rc == 1, "fallbacks": fb.value, "max_tile": mt.value}

    # -- state ---------------------------------------------------------------
    def upload(self, state6):
        """state6: six float32 arrays (pos x,y,z, vel x,y,z) in original order."""
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in state6]
        n = arrs[0].shape[0]
        self._check(self.lib.pbf_upload(self.ctx, n, *[fptr(a) for a in arrs]))
        self.n = n

    def download(self):
        n = int(self.lib.pbf_count(self.ctx))
        out = [np.empty(n, dtype=np.float32) for _ in range(6)]
        self._check(self.lib.pbf_download(self.ctx, *[fptr(a) for a in out]))
        return out

    def batches_retried(self) -> int:
        """Batches replayed so far because a device table had to grow (transparent, but not free)."""
        return int(self.lib.pbf_batches_retried(self.ctx))

    def host_register(self, arr: np.ndarray):
        """Page-lock a host array in place (pbf_host_register): copies to and from it then run at
        full PCIe speed and pbf_step_host can use its single-graph path."""
        self._check(self.lib.pbf_host_register(self.ctx, C.c_void_p(arr.ctypes.data), arr.nbytes))

    def host_unregister(self, arr: np.ndarray):
        self._check(self.lib.pbf_host_unregister(self.ctx, C.c_void_p(arr.ctypes.data)))

    def snapshot_begin(self, slot: int):
        """Enqueue an asynchronous copy of the positions into the library's pinned buffer `slot`
        (0 or 1); it travels on its own stream under the next step() batch."""
        self._check(self.lib.pbf_snapshot_begin(self.ctx, slot))

    def snapshot_wait(self, slot: int):
        """(pos_x, pos_y, pos_z views of the pinned buffers, time) of the snapshot begun in `slot`."""
        ptrs = [_f32p() for _ in range(3)]
        n, t = C.c_size_t(0), C.c_float(0.0)
        self._check(self.lib.pbf_snapshot_wait(self.ctx, slot, *[C.byref(p) for p in ptrs], C.byref(n), C.byref(t)))
        arrs = [np.ctypeslib.as_array(p, shape=(n.value,)) if n.value else np.zeros(0, np.float32) for p in ptrs]
        return arrs, float(t.value)

    def step(self, nsteps: int = 1):
        self._check(self.lib.pbf_step(self.ctx, nsteps))

    def step_host(self, state6, nsteps: int = 1):
        """The cuda_step contract: host arrays in/out (modified in place)."""
        n = state6[0].shape[0]
        self._check(self.lib.pbf_step_host(self.ctx, n, *[fptr(a) for a in state6], nsteps))
        self.n = n

    @property
    def time(self) -> float:
        return float(self.lib.pbf_time(self.ctx))

    def set_time(self, t: float):
        self._check(self.lib.pbf_set_time(self.ctx, t))

    def count(self) -> int:
        return int(self.lib.pbf_count(self.ctx))

    # -- parity surface ------------------------------------------------------
    def debug_enable(self, on: bool = True):
        self._check(self.lib.pbf_debug_enable(self.ctx, int(on)))

    def debug_sizes(self):
        nc, nn = C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.pbf_debug_sizes(self.ctx, C.byref(nc), C.byref(nn)))
        return int(nc.value), int(nn.value)

    def debug_grid(self):
        n = self.count()
        nc, _ = self.debug_sizes()
        ecx, ecy, ecz, ep = (np.empty(n, dtype=np.int32) for _ in range(4))
        cxyz = np.empty(3 * nc, dtype=np.int32)
        cs, ce = (np.empty(nc, dtype=np.int32) for _ in range(2))
        self._check(self.lib.pbf_debug_grid(self.ctx, iptr(ecx), iptr(ecy), iptr(ecz), iptr(ep),
                                            iptr(cxyz), iptr(cs), iptr(ce)))
        return {"entry_cx": ecx, "entry_cy": ecy, "entry_cz": ecz, "entry_particle": ep,
                "cell_xyz": cxyz.reshape(-1, 3), "cell_start": cs, "cell_end": ce}

    def debug_neighbors(self):
        n = self.count()
        _, nn = self.debug_sizes()
        prefix = np.empty(n, dtype=np.int32)
        idx = np.empty(max(nn, 1), dtype=np.int32)
        self._check(self.lib.pbf_debug_neighbors(self.ctx, iptr(prefix), iptr(idx)))
        return prefix, idx[:nn]

    def debug_scratch(self, name: str) -> np.ndarray:
        out = np.empty(self.count(), dtype=np.float32)
        self._check(self.lib.pbf_debug_scratch(self.ctx, SCRATCH_IDS[name], fptr(out)))
        return out

    # -- measurement ---------------------------------------------------------
    def profile_enable(self, on: bool = True):
        self._check(self.lib.pbf_profile_enable(self.ctx, int(on)))

    def profile_reset(self):
        self._check(self.lib.pbf_profile_reset(self.ctx))

    def profile(self) -> dict:
        out = {}
        for k, name in enumerate(STAGES):
            ms, cnt = C.c_double(0), C.c_uint64(0)
            self._check(self.lib.pbf_profile_get(self.ctx, k, C.byref(ms), C.byref(cnt)))
            out[name] = {"ms": ms.value, "launches": int(cnt.value)}
        return out

    def launch_count(self) -> int:
        return int(self.lib.pbf_launch_count(self.ctx))


def device_count() -> int:
    lib = load_library()
    err = C.c_char_p()
    return int(lib.pbf_device_count(C.byref(err)))


def slab_plan(px: np.ndarray, h: float, nranks: int) -> np.ndarray:
    """pbf_slab_plan: x-cell cuts [nranks + 1] for `nranks` slabs (host only, no GPU needed)."""
    lib = load_library()
    px = np.ascontiguousarray(px, dtype=np.float32)
    cuts = np.empty(nranks + 1, dtype=np.int32)
    rc = lib.pbf_slab_plan(px.shape[0], fptr(px), C.c_float(h), nranks, iptr(cuts))
    if rc != 0:
        raise PbfError(f"pbf_slab_plan failed ({rc}): {(lib.pbf_last_error(None) or b'?').decode()}")
    return cuts


def slab_plan_hist(hist: np.ndarray, first_layer: int, nranks: int, ghost_weight: float) -> np.ndarray:
    """pbf_slab_plan_hist: the planner on an x-layer histogram with an explicit ghost weight."""
    lib = load_library()
    hist = np.ascontiguousarray(hist, dtype=np.uint64)
    cuts = np.empty(nranks + 1, dtype=np.int32)
    rc = lib.pbf_slab_plan_hist(hist.ctypes.data_as(C.POINTER(C.c_uint64)), hist.shape[0], int(first_layer), nranks,
                                C.c_float(ghost_weight), iptr(cuts))
    if rc != 0:
        raise PbfError(f"pbf_slab_plan_hist failed ({rc}): {(lib.pbf_last_error(None) or b'?').decode()}")
    return cuts


def comm_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(PBF_COMM_ID_BYTES)
    rc = lib.pbf_comm_unique_id(buf)
    if rc != 0:
        raise PbfError(f"pbf_comm_unique_id failed ({rc}): {(lib.pbf_last_error(None) or b'?').decode()}")
    return buf.raw


class SlabSolver(Solver):
    """One slab of a multi-process run: one rank per GPU, NCCL between x-neighbours."""

    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, PBF_COMM_ID_BYTES)
        self._check(self.lib.pbf_comm_init(self.ctx, rank, nranks, buf))

    def slab_upload(self, state6):
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in state6]
        self._check(self.lib.pbf_slab_upload(self.ctx, arrs[0].shape[0], *[fptr(a) for a in arrs]))

    def slab_upload_owned(self, gid, state6):
        """This rank's particles from host arrays.  gid None: the same particles in the same order as
        the last slab_download returned (positions and velocities only)."""
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in state6]
        if gid is None:
            self._check(self.lib.pbf_slab_upload_owned(self.ctx, arrs[0].shape[0], None, *[fptr(a) for a in arrs]))
            return
        gid = np.ascontiguousarray(gid, dtype=np.int64)
        self._check(self.lib.pbf_slab_upload_owned(self.ctx, gid.shape[0], gid.ctypes.data_as(_i64p),
                                                   *[fptr(a) for a in arrs]))

    def set_p2p(self, enabled: bool):
        self._check(self.lib.pbf_slab_set_p2p(self.ctx, int(enabled)))

    def payload_bytes(self) -> int:
        """Payload bytes sent to the neighbours during the last substep (capacity-independent)."""
        v = C.c_uint64(0)
        self._check(self.lib.pbf_slab_payload(self.ctx, C.byref(v)))
        return int(v.value)

    def transport(self) -> str:
        """Data plane of the halo exchanges of the last batch."""
        rc = int(self.lib.pbf_slab_transport(self.ctx))
        self._check(min(rc, 0))
        return {1: "local-copies", 2: "nccl-messages", 3: "peer-stores"}.get(rc, "unknown")

    def owned(self) -> int:
        return int(self.lib.pbf_slab_owned(self.ctx))

    def cuts(self):
        lo, hi = np.zeros(1, np.int32), np.zeros(1, np.int32)
        self._check(self.lib.pbf_slab_cuts(self.ctx, iptr(lo), iptr(hi)))
        return int(lo[0]), int(hi[0])

    def set_cuts(self, lo: int, hi: int):
        self._check(self.lib.pbf_slab_set_cuts(self.ctx, int(lo), int(hi)))

    def set_rebalance(self, threshold: float):
        self._check(self.lib.pbf_slab_set_rebalance(self.ctx, C.c_float(threshold)))

    def rebalance_count(self) -> int:
        return int(self.lib.pbf_slab_rebalance_count(self.ctx))

    def slab_download(self, out=None, ids=True):
        """(global ids, six SoA arrays) of the owned particles.  `out` = (int64 array, six float32
        arrays) of sufficient capacity (e.g. pinned) to receive them in place.  ids=False skips the
        ids (returns None for them)."""
        n = self.owned()
        if out is None:
            gid = np.empty(n, dtype=np.int64)
            arrs = [np.empty(n, dtype=np.float32) for _ in range(6)]
        else:
            gid, arrs = out[0][:n], [a[:n] for a in out[1]]
        self._check(self.lib.pbf_slab_download(self.ctx, gid.ctypes.data_as(_i64p) if ids else None,
                                               *[fptr(a) for a in arrs]))
        return (gid if ids else None), arrs

    def slab_stats(self) -> dict:
        ex, by = C.c_uint64(0), C.c_uint64(0)
        gh, hops = np.zeros(1, np.int32), np.zeros(1, np.int32)
        self._check(self.lib.pbf_slab_stats(self.ctx, C.byref(ex), C.byref(by), iptr(gh), iptr(hops)))
        return {"exchanges": int(ex.value), "bytes_sent": int(by.value), "ghosts": int(gh[0]), "hops": int(hops[0])}


class SlabGroup:
    """Several slabs driven by ONE process (pbf_group_*): contexts on the same device are
    "virtual ranks" (the 1-GPU parity tests), contexts on different devices a single-process
    multi-GPU run."""

    def __init__(self, devices, params: PbfParams, planes: np.ndarray, mode: int = PBF_MODE_STRICT,
                 p2p: bool = False):
        self.lib = load_library()
        self.slabs = [SlabSolver(d, 0, mode) for d in devices]
        for s in self.slabs:
            s.set_params(params)
            s.set_planes(planes)
        arr = (C.c_void_p * len(self.slabs))(*[s.ctx for s in self.slabs])
        self.group = self.lib.pbf_group_create(arr, len(self.slabs))
        if not self.group:
            raise PbfError("pbf_group_create failed")
        if p2p:
            for s in self.slabs:
                s.set_p2p(True)
        self.n = 0

    def _check(self, rc: int):
        if rc != 0:
            msgs = [(self.lib.pbf_last_error(s.ctx) or b"").decode() for s in self.slabs]
            raise PbfError(f"pbf group error {rc}: {[m for m in msgs if m]}")

    def upload(self, state6):
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in state6]
        self.n = arrs[0].shape[0]
        self._check(self.lib.pbf_group_upload(self.group, self.n, *[fptr(a) for a in arrs]))

    def step(self, nsteps: int = 1):
        self._check(self.lib.pbf_group_step(self.group, nsteps))

    def download(self):
        out = [np.empty(self.n, dtype=np.float32) for _ in range(6)]
        self._check(self.lib.pbf_group_download(self.group, *[fptr(a) for a in out]))
        return out

    def owned(self):
        return [s.owned() for s in self.slabs]

    def rebalance(self, h: float):
        """New cuts from the current positions (pbf_slab_plan) on every slab."""
        px = self.download()[0]
        cuts = slab_plan(px, h, len(self.slabs))
        for r, s in enumerate(self.slabs):
            s.set_cuts(cuts[r], cuts[r + 1])
        return cuts

    def close(self):
        if getattr(self, "group", None):
            self.lib.pbf_group_destroy(self.group)
            self.group = None
        for s in getattr(self, "slabs", []):
            s.close()
        self.slabs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
