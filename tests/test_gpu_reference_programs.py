"""GPU acceptance test: the reference's own test programs, compiled UNMODIFIED against
include/pfft.h (tests/c/build_ref_tests.sh -> oracle/_ref/bin/, built in the container that
has /root/reference), run with `pfftrun -np 4` (2-D meshes) / `-np 8` (3-D meshes) on one
GPU.  Acceptance rule = the reference's: every printed `maxerror` below 1e-12
(tests/run_checks.sh:75); 1e-3 absolute for the single-precision program (not in run_checks.sh)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
pytestmark = pytest.mark.gpu


def _list(name):
    p = os.path.join(BIN, name)
    return open(p).read().split() if os.path.exists(p) else []


# everything in LIST_2D/LIST_3D is implemented; the pattern stays for rows that are not
NOT_YET = re.compile(r"^$")
PROGRAMS = [(n, 4) for n in _list("LIST_2D") if not n.startswith("time_")] + [(n, 8) for n in _list("LIST_3D")]


@pytest.mark.parametrize("name,np_", PROGRAMS, ids=[p[0] for p in PROGRAMS])
def test_reference_program(built_lib, name, np_):
    if NOT_YET.search(name):
        pytest.xfail("not implemented yet: " + name)
    exe = os.path.join(BIN, name)
    cmd = [os.path.join(ROOT, "pfft_b200", "bin", "pfftrun"), "-np", str(np_), "-timeout", "120", exe]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
    assert p.returncode == 0, (p.returncode, p.stdout[-1500:], p.stderr[-1500:])
    errs = [float(x) for x in re.findall(r"maxerror = ([^;]*);", p.stdout)]
    assert errs, p.stdout[-1500:]
    # the float program compares values up to 1500 in absolute terms: 1e-3 abs = 7e-7 relative
    tol = 1e-3 if "float" in name else 1e-12
    assert max(errs) < tol, (errs, p.stdout[-800:])


def test_programs_were_built():
    if not os.path.isdir(BIN):
        pytest.skip("oracle/_ref/bin missing (built only where /root/reference exists)")
    assert len(PROGRAMS) >= 40
