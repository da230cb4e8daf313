"""The opt-in experiment switches of the forward NTT give the words of the default path: the NTT
parity tests re-run in a child process under each switch (the library reads its environment once).
PFHE_NTT_CLUSTER=16 | 8: one launch, a thread-block cluster per limb, the intermediate in distributed shared memory
(csrc/ntt_cluster.cu); PFHE_NTT_FUSED=1: one persistent launch with a ticket counter (k_fwd_fused)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["tests/test_gpu_parity.py::test_ntt_forward_inverse", "tests/test_gpu_parity.py::test_ntt_start_index_and_linearity",
         "tests/test_gpu_parity.py::test_ntt_config1_known_answer",
         "tests/test_gpu_parity.py::test_against_unmodified_reference"]   # 20 limbs of N = 2^16: enough tiles for the persistent form


@pytest.mark.parametrize("switch", ["PFHE_NTT_CLUSTER=16", "PFHE_NTT_CLUSTER=8", "PFHE_NTT_FUSED=1"])
def test_ntt_parity_under_switch(switch):
    name, value = switch.split("=")
    env = dict(os.environ, **{name: value})
    p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "--tb=short"] + CASES, cwd=ROOT,
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert " passed" in p.stdout and "failed" not in p.stdout
