"""The verbs stand-in under the reference stack (oracle/verbs_shim) must behave like an HCA where the reference's
transport code depends on it: RDMA WRITE/READ placement, completion rules (unsignalled successes are silent, errors
always complete, flush after error), access fencing through the responder's queue-pair state and PSN (DARE revokes
log access that way), remote-key range checks, UD unicast/multicast delivery behind a 40-byte GRH.
oracle/verbs_shim/shim_selftest.c exercises exactly that between two processes."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(os.path.dirname(HERE), "oracle", "verbs_shim")


def test_shim_selftest():
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "shim_selftest")
        subprocess.run(["gcc", "-O2", "-g", "-std=gnu99", "-Wall", "-I", SHIM, "-o", exe, os.path.join(SHIM, "shim_selftest.c"),
                        os.path.join(SHIM, "verbs_shim.c")], check=True)
        out = subprocess.run([exe], env=dict(os.environ, APUS_SHIM_DIR=os.path.join(d, "shim")), capture_output=True,
                             text=True, timeout=60)
        assert out.returncode == 0 and "selftest ok" in out.stdout, out.stdout + out.stderr
