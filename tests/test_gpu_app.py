"""GPU tests through the applications: the reference's own main() with this backend linked in
(oracle/_ref/fluidsim_dropin), and this repo's fluidsim_b200."""
import filecmp
import subprocess
from pathlib import Path

import pytest

from fluidsimulator_b200 import scenes
from oracle import oracle_api

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
APP = ROOT / "fluidsimulator_b200" / "bin" / "fluidsim_b200"
FLAGS = ["--steps-per-sec", "120", "--enable-scorr", "--enable-xsph", "--enable-vorticity",
         "--plane-restitution", "0.05", "--plane-friction", "0.1"]


def run(binary, *args, env=None):
    import os
    r = subprocess.run([str(binary), *map(str, args)], capture_output=True, text=True,
                       env=None if env is None else dict(os.environ, **env))
    assert r.returncode == 0, r.stderr
    return r.stdout


def same_tree(a: Path, b: Path):
    names = sorted(p.name for p in a.iterdir())
    assert names == sorted(p.name for p in b.iterdir()) and names
    for n in names:
        assert filecmp.cmp(a / n, b / n, shallow=False), n
    return names


@pytest.mark.skipif(not oracle_api.DROPIN_BIN.exists(), reason="oracle/_ref/fluidsim_dropin not built")
def test_reference_app_cpu_vs_dropin_cuda(built, tmp_path):
    """The reference application, unmodified, with --backend=cpu and with --backend=cuda (= this repo's
    backend behind fluid::cuda_step): every output file is byte-identical, all flags on."""
    scene = scenes.small_block(14).write_json(tmp_path / "scene.json")
    outs = {}
    for backend in ("cpu", "cuda"):
        outs[backend] = tmp_path / backend
        stdout = run(oracle_api.DROPIN_BIN, f"--backend={backend}", "--scene", scene, "--steps", 12,
                     "--fps", 40, "--output-dir", outs[backend], *FLAGS)
        assert f"backend={backend}" in stdout
    names = same_tree(outs["cpu"], outs["cuda"])
    assert "series.pvd" in names and len(names) == 5   # frames after steps 1, 4, 7, 10


@pytest.mark.skipif(not oracle_api.DROPIN_BIN.exists(), reason="oracle/_ref/fluidsim_dropin not built")
def test_own_app_matches_reference_app(built, tmp_path):
    """fluidsim_b200 (device-resident stepping, own loader and writer) == reference app on CPU."""
    scene = scenes.small_block(12).write_json(tmp_path / "scene.json")
    ref, ours, host = tmp_path / "ref", tmp_path / "ours", tmp_path / "host"
    ref_stdout = run(oracle_api.DROPIN_BIN, "--backend=cpu", "--scene", scene, "--duration", "0.1", "--fps", 60,
                     "--output-dir", ref, *FLAGS)
    our_stdout = run(APP, "--scene", scene, "--duration", "0.1", "--fps", 60, "--output-dir", ours, *FLAGS)
    run(APP, "--scene", scene, "--duration", "0.1", "--fps", 60, "--output-dir", host, "--per-step-host", *FLAGS)
    same_tree(ref, ours)
    same_tree(ref, host)
    pick = lambda s: [l for l in s.splitlines() if l.split("=")[0] in ("particle_count", "end_time", "output_enabled", "core_version")]
    assert pick(ref_stdout) == pick(our_stdout)


def test_own_app_default_scene_and_iterations(built, tmp_path):
    out = run(APP, "--steps", 3, "--no-output", "--solver-iterations", 2, "--debug-print")
    assert "particle_count=46875" in out and "step_done=3" in out and "output_enabled=false" in out


def test_own_app_slabs_and_batched_output(built, tmp_path):
    """`--devices 0,0,0`: three x-slabs (here on one GPU) driven by the application; batched
    stepping between output frames and the background frame writer.  Files are byte-identical to
    the single-context run, which the test above ties to the reference application."""
    scene = scenes.SCENES["fluid_large"].write_json(tmp_path / "scene.json")
    one, three = tmp_path / "one", tmp_path / "three"
    stable = ["--steps-per-sec", "120", "--enable-scorr", "--enable-xsph", "--plane-restitution", "0.05",
              "--plane-friction", "0.1"]
    a = run(APP, "--scene", scene, "--steps", 14, "--fps", 30, "--output-dir", one, *stable)
    b = run(APP, "--scene", scene, "--steps", 14, "--fps", 30, "--output-dir", three, "--devices", "0,0,0", *stable)
    names = same_tree(one, three)
    assert len(names) == 5 and "slabs=3" in b     # frames after steps 1, 5, 9, 13 + series.pvd
    pick = lambda s: [l for l in s.splitlines() if l.split("=")[0] in ("particle_count", "end_time")]
    assert pick(a) == pick(b)


def test_own_app_overlapped_output_equals_blocking(built, tmp_path):
    """frames leave through two pinned buffers on a copy stream while the next batch of
    substeps runs (default) — byte-identical to the blocking download (PBF_BLOCKING_FRAMES=1), with
    enough frames that both buffers are reused several times and with a one-substep batch size."""
    scene = scenes.SCENES["fluid_large"].write_json(tmp_path / "scene.json")
    stable = ["--steps-per-sec", "120", "--enable-scorr", "--enable-xsph", "--plane-restitution", "0.05",
              "--plane-friction", "0.1"]
    for fps, steps, frames in ((40, 31, 11), (120, 9, 9)):
        over, block = tmp_path / f"over{fps}", tmp_path / f"block{fps}"
        a = run(APP, "--scene", scene, "--steps", steps, "--fps", fps, "--output-dir", over, *stable)
        b = run(APP, "--scene", scene, "--steps", steps, "--fps", fps, "--output-dir", block, *stable,
                env={"PBF_BLOCKING_FRAMES": "1"})
        names = same_tree(over, block)
        assert len(names) == frames + 1
        pick = lambda s: [l for l in s.splitlines() if l.split("=")[0] in ("particle_count", "end_time")]
        assert pick(a) == pick(b)
