"""Row f2 of SURVEY §8: the parity harness (tests/parity_tool.py — the "frame dumps + diff tool" the
reference's AGENTS.md:23-24 asks for) is itself exercised: STRICT must report bit-identity and exit
0, FAST must stay inside the tolerance BASELINE.json states, and a wrong answer must exit 1."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
TOOL = ROOT / "tests" / "parity_tool.py"


def tool(*args):
    return subprocess.run([sys.executable, str(TOOL), *map(str, args)], capture_output=True, text=True, cwd=ROOT,
                          timeout=600)


def test_parity_tool_strict_fluid_large(built):
    r = tool("--scene", "fluid_large", "--flags", "all", "--steps", 10, "--every", 5, "--mode", "strict")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = [l.split() for l in r.stdout.splitlines() if l.split() and l.split()[0].isdigit()]
    assert [row[0] for row in rows] == ["5", "10"]
    for row in rows:                      # step, tables, floats, max|dp|/h, rms|dp|/h, max|dv|
        assert row[1] == "identical" and row[2] == "bit-exact" and float(row[3]) == 0.0 and float(row[5]) == 0.0, row
    assert r.stdout.strip().endswith("PASS")


def test_parity_tool_fast_within_tolerance(built):
    r = tool("--scene", "fluid_large", "--flags", "stable", "--steps", 10, "--every", 5, "--mode", "fast")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.strip().endswith("PASS")
    rows = [l.split() for l in r.stdout.splitlines() if l.split() and l.split()[0].isdigit()]
    assert rows and all(float(row[3]) <= 1e-4 for row in rows)


def test_parity_tool_slabs(built):
    r = tool("--scene", "fluid_large", "--flags", "stable", "--steps", 6, "--every", 3, "--mode", "strict", "--slabs", 3)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_parity_tool_reports_failure(built):
    """FAST mode with vorticity diverges from the reference within a few substeps (chaotic system,
    SURVEY §0): the tool must say so and exit 1 under an (absurdly) tight tolerance."""
    r = tool("--scene", "fluid_large", "--flags", "all", "--steps", 10, "--every", 5, "--mode", "fast", "--tol", 1e-12)
    assert r.returncode == 1, r.stdout[-2000:]
    assert r.stdout.strip().endswith("FAIL")
