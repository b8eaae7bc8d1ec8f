"""The C++ host-side mirror of the reference's interface (include/fspann_host.hpp: ForwardSecureANNSystem, QueryTokenFactory,
PartitionedIndexService, QueryServiceImpl, KeyManager over the C ABI).  CPU: it compiles and links against the library.  GPU: the
test program (tests/cpp/host_mirror_test.cpp) runs the reference's facade flow, Rotate / Migrate / Retire and error behaviour, and
compares every search result bit for bit with the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fspann_query_system_b200", "csrc")
BIN = os.path.join(CSRC, "host_mirror_test")


def test_host_mirror_program_builds():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", CSRC, "host_mirror_test"], stdout=subprocess.DEVNULL)
    assert os.path.exists(BIN)
    out = subprocess.run(["ldd", BIN], capture_output=True, text=True).stdout
    assert "libfspann_gpu.so" in out and "not found" not in out.split("libfspann_gpu.so")[1].split("\n")[0]


@pytest.mark.gpu
def test_host_mirror_flow_matches_oracle_on_gpu():
    assert os.path.exists(BIN), "run __graft_entry__.build()"
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "HOST MIRROR OK" in r.stdout
