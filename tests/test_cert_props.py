"""The checksum behind the self-certifying publish -- apus_b200/csrc/apus_cert.h, the functions the
kernels themselves use, compiled as C: tests/hostlogic/cert_props.c checks that the chunk-wise sum a warp computes equals
the byte-wise definition, that every single-byte change inside the entry shows (odd weights), that bytes outside do not
matter, that stale bytes / the same bytes elsewhere / another publish's key do not verify.  No GPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_certificate_checksum_properties(tmp_path):
    exe = str(tmp_path / "cert_props")
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-Wall", "-o", exe, os.path.join(HERE, "hostlogic", "cert_props.c")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("cert ok"), out.stdout + out.stderr
