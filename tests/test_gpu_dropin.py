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


def _nanosg_golden():
    import gzip

    with gzip.open(os.path.join(ROOT, "tests", "golden", "nanosg_check_ref.txt.gz"), "rt") as f:
        return f.read().strip().splitlines()


def test_nanosg_drop_in_conformance_is_identical():
    """examples/nanosg_check.cc against include/nanosg.h with -DNANORT_B200_CONFORMANCE must print exactly what the
    same source prints against the reference's examples/nanosg/nanosg.h: scene box, node matrices, and for every
    ray node_id / prim_id / t / u / v / P / transformed Ng as raw bits."""
    if not os.path.exists(os.path.join(BIN, "nanosg_check_b200_conf")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    got = _run("nanosg_check_b200_conf")
    want = _nanosg_golden()
    assert len(want) > 2000 and got == want


def test_reference_nanosg_header_on_top_of_the_facade_is_identical():
    """The reference's OWN scene-graph header (examples/nanosg/nanosg.h, unmodified: NodeBBoxGeometry / NodeBBoxPred /
    NodeBBoxIntersector handed to BVHAccel::Build / ListNodeIntersections, one BVHAccel::Traverse per pierced node)
    compiled against include/nanort.h with -DNANORT_B200_CONFORMANCE: the same output as against the reference's nanort.h."""
    if not os.path.exists(os.path.join(BIN, "nanosg_check_mixed_conf")):
        pytest.skip("examples/bin/nanosg_check_mixed_conf not built (needs /root/reference at build time)")
    got = _run("nanosg_check_mixed_conf")
    want = _nanosg_golden()
    assert got == want


def test_nanosg_drop_in_fast_mode_same_hits():
    """Default (fast) mode: same lines except where two surfaces lie at exactly the same distance."""
    if not os.path.exists(os.path.join(BIN, "nanosg_check_b200")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    got = _run("nanosg_check_b200")
    want = _nanosg_golden()
    assert got[:2] == want[:2] and got[-1] == want[-1]  # scene box, node state, hit count
    assert [g.split(":")[0] for g in got] == [w.split(":")[0] for w in want]  # the same rays hit
    diff = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert len(diff) <= 0.01 * len(want)
    for i in diff:  # only the pick differs, not the distance
        assert got[i].split(" t ")[1].split()[0] == want[i].split(" t ")[1].split()[0]


def test_f64_drop_in_prints_the_reference_output():
    """examples/f64_check.cc (BVHAccel<double>: the reference's regression scenario + 2,000 double rays over a
    height field) against include/nanort.h must print what it prints against the reference header: every t / u / v
    as raw 64-bit patterns.  (The mesh has no coincident surfaces, so no tie can pick another triangle.)"""
    import gzip

    if not os.path.exists(os.path.join(BIN, "f64_check_b200")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    got = _run("f64_check_b200")
    with gzip.open(os.path.join(ROOT, "tests", "golden", "f64_check_ref.txt.gz"), "rt") as f:
        want = f.read().strip().splitlines()
    assert len(want) > 1500 and got == want


def test_reference_objrender_program_renders_the_reference_image(tmp_path):
    """The reference's OWN example program (examples/objrender/main.cc, unmodified; one BVHAccel::Traverse per pixel),
    compiled once against the reference header and once against include/nanort.h (examples/Makefile, target `ref`),
    must write byte-identical images; the run also gives the rays/s of the literal per-ray drop-in."""
    import time

    exe_b, exe_r = os.path.join(BIN, "ref_objrender_b200"), os.path.join(BIN, "ref_objrender_ref")
    obj = os.path.join(BIN, "assets", "cornellbox_suzanne.obj")
    if not (os.path.exists(exe_b) and os.path.exists(exe_r) and os.path.exists(obj)):
        pytest.skip("reference example binaries were not built (authoring container without /root/reference)")
    imgs = {}
    for tag, exe in (("b200", exe_b), ("ref", exe_r)):
        d = tmp_path / tag
        d.mkdir()
        for f in ("cornellbox_suzanne.obj", "cornellbox_suzanne.mtl"):
            os.symlink(os.path.join(BIN, "assets", f), d / f)
        t0 = time.time()
        out = subprocess.run([exe, "cornellbox_suzanne.obj"], cwd=d, check=True, capture_output=True, text=True,
                             timeout=600).stdout
        dt = time.time() - t0
        imgs[tag] = (open(d / "render.exr", "rb").read(), open(d / "render.png", "rb").read())
        render = [ln for ln in out.splitlines() if ln.startswith("Render ")]
        print(f"objrender[{tag}]: wall {dt:.2f} s, {render[-1] if render else ''} "
              f"({512 * 512 / max(float(render[-1].split()[1]), 1e-9) / 1e6 if render else 0:.3f} Mrays/s per-ray Traverse)")
    assert imgs["b200"][0] == imgs["ref"][0], "render.exr differs from the reference program's"
    assert imgs["b200"][1] == imgs["ref"][1], "render.png differs from the reference program's"


def test_reference_particle_primitive_program_renders_the_reference_image(tmp_path):
    """The reference's custom-primitive example (examples/particle_primitive/main.cc, unmodified: its own SphereGeometry,
    SpherePred and SphereIntersector classes handed to BVHAccel::Build / Traverse) compiled against include/nanort.h runs
    its spheres on the device kind NRT_PRIM_SPHERES and must write the image the reference header's build writes."""
    exe_b, exe_r = os.path.join(BIN, "ref_particle_b200"), os.path.join(BIN, "ref_particle_ref")
    if not (os.path.exists(exe_b) and os.path.exists(exe_r)):
        pytest.skip("reference example binaries were not built (authoring container without /root/reference)")
    imgs = {}
    for tag, exe in (("b200", exe_b), ("ref", exe_r)):
        d = tmp_path / tag
        d.mkdir()
        subprocess.run([exe, "2000"], cwd=d, check=True, capture_output=True, text=True, timeout=900)  # argv[1] = number of spheres
        imgs[tag] = open(d / "render.png", "rb").read()
    assert imgs["b200"] == imgs["ref"], "render.png differs from the reference program's"
