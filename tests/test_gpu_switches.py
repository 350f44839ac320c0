"""The library's alternative kernel paths stay correct: each environment switch (read once per process) is exercised in a
process of its own by running smoke() - cfg1 Glow forward / backward against the float32 oracle plus the tensor-core
ResidualBlock at cfg2's scale-1 channel plan against the float64 oracle (the thresholds are smoke()'s own)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SWITCHES = [
    {"INB_PLANE_LO8": "0"},        # 2-byte lo planes for the stored hidden tensors
    {"INB_COUPLING_PX": "0"},      # two-phase col2im + coupling kernel instead of the thread-per-pixel one
    {"INB_FUSE_COUPLING": "0"},    # separate col2im and coupling kernels
    {"INB_GRAPHS": "0"},           # direct launches instead of CUDA-graph replay
    {"INB_L2_HINTS": "3"},         # L2 eviction hints on the bulk stores
    {"INB_CHAIN_QSUM": "0"},       # full tap-expanded P instead of the tap-row planes
    {"INB_CHAIN_QWIDE": "0"},      # tap-row planes only where the whole tile can be staged at once (GEMM3 <= 128 columns)
]


@pytest.mark.parametrize("env", SWITCHES, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_smoke_under_switch(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "smoke: tensor-core ResidualBlock" in r.stdout
