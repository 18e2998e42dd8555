"""(Test infrastructure: compares with the oracle, like the tests it calls.)  Seconds-scale parity
check on the GPU box without pytest/torch start-up: smoke() and a
selection of the -m gpu tests called directly (oracle comparisons of grid, neighbour lists and
state; table overflow and replay; random clouds; slabs).  The full suite is `pytest tests -m gpu`."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))  # the test modules import helpers / golden_util by name

t_all = time.perf_counter()


def run(name, fn, *args):
    t0 = time.perf_counter()
    fn(*args)
    print(f"ok {name} ({time.perf_counter() - t0:.1f} s)", flush=True)


import __graft_entry__ as entry
run("smoke", entry.smoke)
import helpers as H
import test_gpu_parity as P
import test_gpu_random as R
import test_gpu_slabs as S

for flags, tag in ((H.NO_FLAGS, "none"), (H.STABLE_FLAGS, "stable"), (H.ALL_FLAGS, "all")):
    run(f"strict_bit_exact_small[{tag}]", P.test_strict_bit_exact_small, True, flags)
run("neighbor_capacity_growth", P.test_neighbor_capacity_growth_is_transparent, True)
for seed in (0, 1, 2, 3):
    run(f"random_clouds[{seed}]", R.test_strict_equals_oracle_on_random_clouds, True, seed)
run("slabs_against_oracle", S.test_virtual_slabs_against_oracle, True)
run("slabs_capacity_growth_and_hops", S.test_virtual_slabs_capacity_growth_and_hops, True)
run("step_host_contract", P.test_step_host_contract, True)
print(f"QUICK CHECK OK ({time.perf_counter() - t_all:.1f} s)", flush=True)
