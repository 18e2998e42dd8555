"""The C-ABI library loads on a CPU-only box and exports every symbol include/pbf_b200.h declares.
No compute call is made here (there is no GPU); the product must fail loudly without one."""
import ctypes as C
import re
from pathlib import Path

import pytest

from fluidsimulator_b200 import capi

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "pbf_b200.h"


def _declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(pbf_[a-z_0-9]+)\s*\(", text)))


def test_header_functions_all_exported(built):
    lib = capi.load_library()
    names = _declared_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in pbf_b200.h but not exported"


def test_python_binding_covers_header(built):
    assert sorted(capi.ABI) == _declared_functions()


def test_abi_version_and_defaults(built):
    lib = capi.load_library()
    assert lib.pbf_abi_version() == 1
    p = capi.PbfParams()
    lib.pbf_default_params(C.byref(p))
    q = capi.PbfParams.defaults()
    assert bytes(C.string_at(C.byref(p), C.sizeof(p))) == bytes(C.string_at(C.byref(q), C.sizeof(q)))
    assert [lib.pbf_stage_name(k).decode() for k in range(len(capi.STAGES))] == capi.STAGES


def test_params_struct_layout():
    """pbf_params is 23 4-byte fields, no padding (include/pbf_b200.h)."""
    assert C.sizeof(capi.PbfParams) == 23 * 4


def test_no_cpu_fallback(built):
    """Without a CUDA device pbf_create must fail with a reason, not fall back."""
    lib = capi.load_library()
    err = C.c_char_p()
    if lib.pbf_device_count(C.byref(err)) > 0:
        pytest.skip("a CUDA device is visible; covered by the gpu tests")
    assert not lib.pbf_create(0, 1024)
    msg = lib.pbf_last_error(None).decode()
    assert "no usable CUDA device" in msg and "no CPU fallback" in msg
    with pytest.raises(capi.PbfError):
        capi.Solver(0, 16)


def test_product_does_not_import_oracle():
    """Nothing under fluidsimulator_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    for path in (ROOT / "fluidsimulator_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".cpp", ".h", ".hpp"} or path.name == "Makefile":
            text = path.read_text(errors="ignore")
            assert "oracle_api" not in text and "oracle/" not in text and "pbf_oracle" not in text, path


def test_header_is_plain_c(tmp_path):
    """include/pbf_b200.h compiles as C11 (no C++-isms in the ABI) and a C program links the library."""
    import subprocess
    src = tmp_path / "abi_check.c"
    src.write_text(
        '#include "pbf_b200.h"\n'
        "#include <stdio.h>\n"
        "int main(void) {\n"
        "  pbf_params p;\n"
        "  pbf_default_params(&p);\n"
        "  const char* err = 0;\n"
        "  int n = pbf_device_count(&err);\n"
        '  printf("abi %d iters %d devices %d\\n", pbf_abi_version(), (int)p.solver_iterations, n);\n'
        "  return (pbf_abi_version() == PBF_ABI_VERSION && sizeof(pbf_params) == 23 * 4) ? 0 : 1;\n"
        "}\n")
    lib_dir = capi.LIB_PATH.parent
    exe = tmp_path / "abi_check"
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src),
                    "-o", str(exe), f"-L{lib_dir}", "-lpbf_b200", f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("abi 1 iters 4")
