"""The facade header: the same C++ source compiled against include/nanort.h (GPU) and against the
reference header (CPU) must print the same Build statistics and per-ray Traverse results."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "bin")


def _run(exe, *args):
    return subprocess.run([os.path.join(BIN, exe), *args], check=True, capture_output=True, text=True,
                          timeout=300).stdout.strip().splitlines()


def test_drop_in_source_gives_reference_output():
    if not os.path.exists(os.path.join(BIN, "drop_in_check_b200")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    got = _run("drop_in_check_b200", "24", "400")
    want = open(os.path.join(ROOT, "tests", "golden", "drop_in_check_ref.txt")).read().strip().splitlines()
    assert len(got) == len(want) == 402
    assert got[0] == want[0], "Build statistics / bounding box line"
    diff = [(g, w) for g, w in zip(got[1:], want[1:]) if g != w]
    assert not diff, diff[:5]
    if os.path.exists(os.path.join(BIN, "drop_in_check_ref")):  # live reference binary, when it travelled
        assert _run("drop_in_check_ref", "24", "400") == want


def test_wavefront_example_runs(tmp_path):
    if not os.path.exists(os.path.join(BIN, "ao_wavefront")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    out = subprocess.run([os.path.join(BIN, "ao_wavefront"), "160", "120", "2"], check=True, capture_output=True,
                         text=True, timeout=300, cwd=tmp_path).stdout
    assert "traced" in out and os.path.getsize(tmp_path / "ao.ppm") > 160 * 120


def test_dump_load_interchange_with_cpu_nanort(tmp_path):
    """Dump/Load in the reference's raw format: (1) a tree dumped by CPU nanort (committed fixture) is loaded
    through the facade and traversed on the GPU; (2) a GPU-built tree dumped through the facade is loaded and
    traversed by CPU nanort (when its binary travelled).  Both must print what CPU nanort prints for its own tree."""
    if not os.path.exists(os.path.join(BIN, "dump_load_b200")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    want = open(os.path.join(ROOT, "tests", "golden", "dump_load_ref.txt")).read().strip().splitlines()
    got = _run("dump_load_b200", "load", os.path.join(ROOT, "tests", "golden", "ref_tree_dump.bin"))
    assert got == want
    out = str(tmp_path / "gpu_tree.bin")
    head = _run("dump_load_b200", "dump", out)
    assert head[0].startswith("dumped ") and os.path.getsize(out) > 50000
    if os.path.exists(os.path.join(BIN, "dump_load_ref")):
        assert _run("dump_load_ref", "load", out) == want


def test_drop_in_conformance_mode_reproduces_tree_and_hits():
    """-DNANORT_B200_CONFORMANCE: the facade's Build must produce CPU nanort's exact arrays (hash of every node
    and of indices_) and its per-ray Traverse the exact hit records -- the golden output of the same source
    compiled against the reference header with -DPRINT_TREE."""
    if not os.path.exists(os.path.join(BIN, "drop_in_check_b200_conf")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    got = _run("drop_in_check_b200_conf", "24", "400")
    want = open(os.path.join(ROOT, "tests", "golden", "drop_in_check_ref_tree.txt")).read().strip().splitlines()
    assert got[1].startswith("tree nodes ") and got == want


def test_concurrent_per_ray_traverse_from_worker_threads():
    """8 host threads calling accel.Traverse per ray, as the reference path tracer's row workers do."""
    if not os.path.exists(os.path.join(BIN, "threads_check")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    out = _run("threads_check")
    assert out[-1].endswith("mismatches 0"), out
