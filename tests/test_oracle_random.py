"""CPU: the C restatement against the compiled reference on adversarial random clouds (cell-boundary
positions, coincident particles, ragged counts, random flags / iterations / planes) — every array the
reference leaves behind, bit for bit.  Widens the pin of the oracle beyond the lattice scenes."""
import numpy as np
import pytest

from oracle import oracle_api
from oracle.oracle_api import Oracle

import golden_util as G
import helpers as H

pytestmark = pytest.mark.skipif(not oracle_api.available("reference"), reason="oracle/_ref not built")


@pytest.mark.parametrize("seed", range(12))
def test_port_equals_reference_on_random_clouds(built, seed):
    params, planes, state, flags = H.random_cloud(seed)
    sims = []
    for kind in ("reference", "port"):
        orc = Oracle(kind)
        orc.set_params(params)
        orc.set_planes(planes)
        orc.set_state(state)
        sims.append(orc)
    for step in range(1, 4):
        for orc in sims:
            orc.step(1)
        a, b = (G.snapshot_of(orc, flags, False) for orc in sims)
        assert G.mismatches(b, a) == [], f"seed {seed} (n = {len(state[0])}), step {step}"
        assert np.float32(sims[0].time) == np.float32(sims[1].time)
