"""ctypes loader for the two CPU oracles (see oracle/oracle_api.h).

TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never from fluidsimulator_b200/.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_ROOT = _HERE.parent
if str(_ROOT) not in sys.path:
    sys.path.insert(0, str(_ROOT))

from fluidsimulator_b200.capi import PbfParams, SCRATCH_IDS, fptr, iptr  # noqa: E402

REF_SO = _HERE / "_ref" / "libpbf_oracle_ref.so"
PORT_SO = _HERE / "_build" / "libpbf_oracle_port.so"
REFERENCE_ROOT = Path("/root/reference")
DROPIN_BIN = _HERE / "_ref" / "fluidsim_dropin"   # reference main + this repo's backend
REFAPP_BIN = _HERE / "_ref" / "fluidsim_ref"      # reference incl. its own CUDA backend (sm_100)

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


def build(which: str = "all", quiet: bool = True) -> None:
    """Compile the oracles (port always; ref only when /root/reference exists)."""
    targets = []
    if which in ("all", "port"):
        targets.append("port")
    if which in ("all", "ref") and (REFERENCE_ROOT / "core/src/core.cpp").exists():
        targets += ["ref", "dropin", "refapp"]
    for t in targets:
        subprocess.run(["make", "-C", str(_HERE), t], check=True,
                       stdout=subprocess.DEVNULL if quiet else None)


def _load(path: Path) -> C.CDLL:
    lib = C.CDLL(str(path))
    sig = {
        "oracle_kind": (C.c_char_p, []),
        "oracle_has_openmp": (C.c_int, []),
        "oracle_set_threads": (None, [C.c_int]),
        "oracle_max_threads": (C.c_int, []),
        "oracle_create": (C.c_void_p, []),
        "oracle_destroy": (None, [C.c_void_p]),
        "oracle_load_scene": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t]),
        "oracle_init_test_scene": (C.c_int, [C.c_void_p]),
        "oracle_set_params": (None, [C.c_void_p, C.POINTER(PbfParams)]),
        "oracle_get_params": (None, [C.c_void_p, C.POINTER(PbfParams)]),
        "oracle_set_planes": (None, [C.c_void_p, C.c_int] + [_f32p] * 4),
        "oracle_plane_count": (C.c_int, [C.c_void_p]),
        "oracle_get_planes": (None, [C.c_void_p] + [_f32p] * 4),
        "oracle_set_state": (None, [C.c_void_p, C.c_size_t] + [_f32p] * 6),
        "oracle_count": (C.c_size_t, [C.c_void_p]),
        "oracle_get_state": (None, [C.c_void_p] + [_f32p] * 6),
        "oracle_time": (C.c_float, [C.c_void_p]),
        "oracle_set_time": (None, [C.c_void_p, C.c_float]),
        "oracle_step": (None, [C.c_void_p, C.c_int]),
        "oracle_ncells": (C.c_size_t, [C.c_void_p]),
        "oracle_nneighbors": (C.c_size_t, [C.c_void_p]),
        "oracle_get_grid": (None, [C.c_void_p] + [_i32p] * 7),
        "oracle_get_neighbors": (None, [C.c_void_p, _i32p, _i32p]),
        "oracle_get_scratch": (None, [C.c_void_p, C.c_int, _f32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_libs: dict = {}


def available(kind: str) -> bool:
    return (REF_SO if kind == "reference" else PORT_SO).exists()


def best_kind() -> str:
    """'reference' when the compiled reference is present, else 'port'."""
    return "reference" if available("reference") else "port"


class Oracle:
    """One CPU simulation (Params + State) driven through oracle_api.h."""

    def __init__(self, kind: str = "port"):
        path = REF_SO if kind == "reference" else PORT_SO
        if not path.exists():
            build("ref" if kind == "reference" else "port")
        if not path.exists():
            raise RuntimeError(f"oracle library {path} is not built")
        if kind not in _libs:
            _libs[kind] = _load(path)
        self.lib = _libs[kind]
        assert self.lib.oracle_kind().decode() == kind
        self.kind = kind
        self.sim = self.lib.oracle_create()

    def close(self):
        if getattr(self, "sim", None):
            self.lib.oracle_destroy(self.sim)
            self.sim = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup ---------------------------------------------------------------
    def load_scene(self, json_path) -> None:
        err = C.create_string_buffer(512)
        rc = self.lib.oracle_load_scene(self.sim, str(json_path).encode(), err, 512)
        if rc != 0:
            raise RuntimeError(f"oracle_load_scene: {err.value.decode()}")

    def init_test_scene(self) -> None:
        if self.lib.oracle_init_test_scene(self.sim) != 0:
            raise RuntimeError("oracle_init_test_scene unavailable in this oracle")

    def set_threads(self, n: int) -> None:
        self.lib.oracle_set_threads(n)

    def max_threads(self) -> int:
        return int(self.lib.oracle_max_threads())

    def set_params(self, p: PbfParams) -> None:
        self.lib.oracle_set_params(self.sim, C.byref(p))

    def get_params(self) -> PbfParams:
        p = PbfParams()
        self.lib.oracle_get_params(self.sim, C.byref(p))
        return p

    def set_planes(self, planes: np.ndarray) -> None:
        planes = np.ascontiguousarray(planes, dtype=np.float32).reshape(-1, 4)
        cols = [np.ascontiguousarray(planes[:, k]) for k in range(4)]
        self.lib.oracle_set_planes(self.sim, planes.shape[0], *[fptr(c) for c in cols])

    def get_planes(self) -> np.ndarray:
        n = int(self.lib.oracle_plane_count(self.sim))
        cols = [np.zeros(max(n, 1), dtype=np.float32) for _ in range(4)]
        self.lib.oracle_get_planes(self.sim, *[fptr(c) for c in cols])
        return np.stack([c[:n] for c in cols], axis=1)

    def set_state(self, state6) -> None:
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in state6]
        self.lib.oracle_set_state(self.sim, arrs[0].shape[0], *[fptr(a) for a in arrs])

    def count(self) -> int:
        return int(self.lib.oracle_count(self.sim))

    def get_state(self):
        n = self.count()
        out = [np.empty(n, dtype=np.float32) for _ in range(6)]
        if n:
            self.lib.oracle_get_state(self.sim, *[fptr(a) for a in out])
        return out

    @property
    def time(self) -> float:
        return float(self.lib.oracle_time(self.sim))

    def set_time(self, t: float) -> None:
        self.lib.oracle_set_time(self.sim, t)

    # -- run -----------------------------------------------------------------
    def step(self, nsteps: int = 1) -> None:
        self.lib.oracle_step(self.sim, nsteps)

    # -- scratch -------------------------------------------------------------
    def grid(self) -> dict:
        n = self.count()
        nc = int(self.lib.oracle_ncells(self.sim))
        ecx, ecy, ecz, ep = (np.empty(n, dtype=np.int32) for _ in range(4))
        cxyz = np.empty(3 * nc, dtype=np.int32)
        cs, ce = (np.empty(nc, dtype=np.int32) for _ in range(2))
        self.lib.oracle_get_grid(self.sim, iptr(ecx), iptr(ecy), iptr(ecz), iptr(ep),
                                 iptr(cxyz), iptr(cs), iptr(ce))
        return {"entry_cx": ecx, "entry_cy": ecy, "entry_cz": ecz, "entry_particle": ep,
                "cell_xyz": cxyz.reshape(-1, 3), "cell_start": cs, "cell_end": ce}

    def neighbors(self):
        n = self.count()
        nn = int(self.lib.oracle_nneighbors(self.sim))
        prefix = np.empty(n, dtype=np.int32)
        idx = np.empty(max(nn, 1), dtype=np.int32)
        self.lib.oracle_get_neighbors(self.sim, iptr(prefix), iptr(idx))
        return prefix, idx[:nn]

    def scratch(self, name: str) -> np.ndarray:
        out = np.zeros(self.count(), dtype=np.float32)
        self.lib.oracle_get_scratch(self.sim, SCRATCH_IDS[name], fptr(out))
        return out
