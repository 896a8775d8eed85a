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


# ---- the reference's own example programs against include/nanort.h (authoring container only: needs /root/reference)
REF = "/root/reference"
REF_EXAMPLES = [
    # (directory, sources, flags of the example's own Makefile)
    ("path_tracer", ["main.cc"], ["-std=c++11", "-DNANORT_USE_CPP11_FEATURE", "-pthread"]),  # examples/path_tracer/Makefile:2
    ("objrender", ["main.cc"], ["-std=c++11"]),
    ("double_precision", ["main.cc"], ["-std=c++11"]),
    ("bidir_path_tracer", ["main.cc"], ["-std=c++11", "-DNANORT_USE_CPP11_FEATURE", "-pthread"]),
    ("par_msquare", ["main.cc"], ["-std=c++11"]),
    ("vrcamera", ["main.cc"], ["-std=c++11"]),
]


@pytest.mark.parametrize("name,sources,flags", REF_EXAMPLES, ids=[e[0] for e in REF_EXAMPLES])
def test_reference_examples_compile_against_the_facade(name, sources, flags):
    d = os.path.join(REF, "examples", name)
    if not os.path.isdir(d):
        pytest.skip("reference tree not present")
    for src in sources:
        r = subprocess.run(["g++", "-fsyntax-only", "-w", *flags, f"-I{INC}", f"-I{REF}/examples/common", f"-I{d}",
                            os.path.join(d, src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]


def test_reference_path_tracer_links_against_the_library(tmp_path):
    """The drop-in claim on the real caller: examples/path_tracer/main.cc + its loader, the example's own flags, our header
    and library -- compiled AND linked (running it needs a GPU and an OBJ scene; objrender is run in test_gpu_dropin)."""
    d = os.path.join(REF, "examples", "path_tracer")
    if not os.path.isdir(d) or not os.path.exists(os.path.join(LIB, "libnanort_b200.so")):
        pytest.skip("reference tree or library not present")
    out = tmp_path / "path_tracer"
    r = subprocess.run(["g++", "-O1", "-w", "-std=c++11", "-DNANORT_USE_CPP11_FEATURE", f"-I{INC}", f"-I{d}",
                        f"-I{REF}/examples/common", os.path.join(d, "main.cc"), os.path.join(d, "tiny_obj_loader.cc"),
                        "-pthread", "-o", str(out), f"-L{LIB}", "-lnanort_b200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_reference_scene_graph_header_compiles_on_top_of_the_facade():
    """examples/nanosg/nanosg.h of the reference (unmodified: custom Prim / Pred / Intersector classes for node boxes,
    nanort::safemin / safemax, ListNodeIntersections, one Traverse per node) against include/nanort.h; the GPU run of the
    same build is tests/test_gpu_dropin.py::test_reference_nanosg_header_on_top_of_the_facade_is_identical."""
    d = os.path.join(REF, "examples", "nanosg")
    if not os.path.isdir(d):
        pytest.skip("reference tree not present")
    r = subprocess.run(["g++", "-fsyntax-only", "-w", "-std=c++11", "-DNANORT_USE_CPP11_FEATURE", "-DNANORT_B200_CONFORMANCE",
                        "-pthread", f"-I{d}", f"-I{INC}", os.path.join(ROOT, "examples", "nanosg_check.cc")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_cylinder_like_intersectors_are_recognised_and_refused(tmp_path):
    """examples/cylinder_primitive names its members like the sphere example (`vertices_`, `radiuss_`); the facade must not
    walk such an accel as spheres: the intersector's `test_cap_` member marks it (detail::is_cylinder_like) and Traverse
    refuses it with a message."""
    src = tmp_path / "t.cc"
    src.write_text('''
#include "nanort.h"
struct SphereLike { const float *vertices_; const float *radiuss_; };
struct CylinderLike { const float *vertices_; const float *radiuss_; const bool test_cap_; };
static_assert(!nanort::detail::is_cylinder_like<SphereLike>::value, "sphere intersector");
static_assert(nanort::detail::is_cylinder_like<CylinderLike>::value, "cylinder intersector");
static_assert(!nanort::detail::is_cylinder_like<nanort::TriangleIntersector<> >::value, "triangle intersector");
int main() { return 0; }
''')
    r = subprocess.run(["g++", "-fsyntax-only", "-std=c++11", f"-I{INC}", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
