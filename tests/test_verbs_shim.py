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


import pytest


@pytest.mark.parametrize("transport", ["vm", "shm"])
def test_shim_selftest(transport):
    """vm: process_vm_writev/readv per operation.  shm: the public record and (here, with the size floor lowered to two
    pages) the registered region are memfd mappings the peer shares -- the same rules must hold."""
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "shim_selftest")
        subprocess.run(["gcc", "-O2", "-g", "-std=gnu99", "-Wall", "-I", SHIM, "-o", exe, os.path.join(SHIM, "shim_selftest.c"),
                        os.path.join(SHIM, "verbs_shim.c")], check=True)
        env = dict(os.environ, APUS_SHIM_DIR=os.path.join(d, "shim"), APUS_SHIM_TRANSPORT=transport, APUS_SHIM_SHM_MIN="8192",
                   APUS_SHIM_TRACE="1")
        out = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=60)
        assert out.returncode == 0 and "selftest ok" in out.stdout, out.stdout + out.stderr
        assert ("re-backed by memfd" in out.stderr) == (transport == "shm"), out.stderr
