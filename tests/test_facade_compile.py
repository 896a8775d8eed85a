"""CPU: the facade headers must compile cleanly as C++11, C++14 and C++17 with warnings on, and every example must
link against the library (no GPU is needed to compile or link; running them is the job of tests/test_gpu_dropin.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIB = os.path.join(ROOT, "nanort_b200")
EXAMPLES = ["drop_in_check.cc", "ao_wavefront.cc", "dump_load.cc", "threads_check.cc", "nanosg_check.cc", "f64_check.cc"]


@pytest.mark.parametrize("std", ["c++11", "c++14", "c++17"])
def test_headers_compile_without_warnings(tmp_path, std):
    src = tmp_path / "tu.cc"
    src.write_text('#include "nanort.h"\n#include "nanosg.h"\n#include "nanort_b200.h"\nint main() { return 0; }\n')
    r = subprocess.run(["g++", f"-std={std}", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", f"-I{INC}", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_header_is_plain_c(tmp_path):
    src = tmp_path / "tu.c"
    src.write_text('#include "nanort_b200.h"\nint main(void) { return nrt_device_count() < 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", f"-I{INC}",
                        str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("example", EXAMPLES)
def test_examples_compile_and_link(tmp_path, example):
    if not os.path.exists(os.path.join(LIB, "libnanort_b200.so")):
        pytest.skip("library not built")
    out = tmp_path / "a.out"
    r = subprocess.run(["g++", "-std=c++11", "-O1", "-DNANORT_USE_CPP11_FEATURE", "-pthread", f"-I{INC}",
                        os.path.join(ROOT, "examples", example), "-o", str(out), f"-L{LIB}", "-lnanort_b200"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
