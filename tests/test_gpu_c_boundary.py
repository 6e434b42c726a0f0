"""The C ABI driven by a compiled C program (tests/c_boundary/boundary_test.c) - no Python on the call path."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _build():
    out = os.path.join(HERE, "c_boundary", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "boundary_test")
    libdir = os.path.join(ROOT, "dfki_quad_b200")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c11", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           os.path.join(HERE, "c_boundary", "boundary_test.c"), "-o", exe, "-L", libdir, "-lb200qp", "-lm",
                           "-Wl,-rpath," + libdir])
    return exe


def test_c_program_compiles_against_the_header():
    """CPU: include/b200qp.h is plain C and the program links against the shared library (no GPU needed to build)."""
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_compiled_c_caller_known_answers_on_the_gpu():
    r = subprocess.run([_build()], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert "all boundary checks passed" in r.stdout
