"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the parity oracle.

 * `Port`      -> oracle/liborc.so           (C restatement, oracle/nanort_oracle.c)
 * `Reference` -> oracle/_ref/libnanort_ref{,03}.so (the unmodified reference header behind
                  oracle/ref_shim.cc; present when it was built in the authoring container)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under nanort_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

RAY_DTYPE = np.dtype(
    [("org", "<f4", (3,)), ("dir", "<f4", (3,)), ("min_t", "<f4"), ("max_t", "<f4"), ("type", "<u4")]
)
HIT_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("t", "<f4"), ("prim_id", "<u4")])
NODE_DTYPE = np.dtype(
    [("bmin", "<f4", (3,)), ("bmax", "<f4", (3,)), ("flag", "<i4"), ("axis", "<i4"), ("data", "<u4", (2,))]
)
BUILD_OPT_DTYPE = np.dtype(
    [
        ("cost_t_aabb", "<f4"),
        ("min_leaf_primitives", "<u4"),
        ("max_tree_depth", "<u4"),
        ("bin_size", "<u4"),
        ("shallow_depth", "<u4"),
        ("min_primitives_for_parallel_build", "<u4"),
        ("cache_bbox", "u1"),
        ("pad", "u1", (3,)),
    ]
)
TRACE_OPT_DTYPE = np.dtype(
    [("prim_ids_range", "<u4", (2,)), ("skip_prim_id", "<u4"), ("cull_back_face", "u1"), ("pad", "u1", (3,))]
)
assert BUILD_OPT_DTYPE.itemsize == 28 and TRACE_OPT_DTYPE.itemsize == 16

MODE_CPP11 = 1
MODE_FIXBINS = 2  # NOT reference behaviour (all three axes binned); experiments only


def build_options(**kw):
    o = np.zeros(1, BUILD_OPT_DTYPE)
    o["cost_t_aabb"] = 0.2
    o["min_leaf_primitives"] = 4
    o["max_tree_depth"] = 256
    o["bin_size"] = 64
    o["shallow_depth"] = 4
    o["min_primitives_for_parallel_build"] = 8192
    for k, v in kw.items():
        o[k] = v
    return o


def trace_options(**kw):
    o = np.zeros(1, TRACE_OPT_DTYPE)
    o["prim_ids_range"] = (0, 0x7FFFFFFF)
    o["skip_prim_id"] = 0xFFFFFFFF
    for k, v in kw.items():
        o[k] = v
    return o


def make(force=False):
    """Compiles liborc.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(os.path.join(HERE, "liborc.so")) or os.path.exists("/root/reference/nanort.h"):
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Counters(C.Structure):
    _fields_ = [("nodes_popped", C.c_uint64), ("prims_tested", C.c_uint64), ("max_stack", C.c_uint32)]


class Port:
    """The C restatement.  Stateless: trees are plain numpy arrays."""

    def __init__(self):
        path = os.path.join(HERE, "liborc.so")
        if not os.path.exists(path):
            make()
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_build.restype = C.c_size_t
        L.orc_build.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_traverse_batch.restype = C.c_size_t
        L.orc_traverse_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                         C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p]
        L.orc_test_prim.restype = C.c_int
        L.orc_test_prim.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_uint32, C.c_void_p]

    def build(self, verts, faces, opts=None, mode=MODE_CPP11):
        verts = np.ascontiguousarray(verts, np.float32)
        faces = np.ascontiguousarray(faces, np.uint32)
        n = len(faces)
        indices = np.zeros(n, np.uint32)
        stats = np.zeros(3, np.uint32)
        out = C.c_void_p()
        nn = self.lib.orc_build(_p(verts), 12, _p(faces), n, _p(opts) if opts is not None else None, mode,
                                C.byref(out), _p(indices), _p(stats))
        if nn == 0:
            return None
        nodes = np.frombuffer((C.c_char * (nn * 40)).from_address(out.value), NODE_DTYPE).copy()
        self.lib.orc_free(out)
        return nodes, indices, {"max_tree_depth": int(stats[0]), "num_leaf_nodes": int(stats[1]),
                                "num_branch_nodes": int(stats[2])}

    def traverse(self, nodes, indices, verts, faces, rays, topts=None, cpp11=True, threads=1, counters=False):
        nodes = np.ascontiguousarray(nodes)
        indices = np.ascontiguousarray(indices, np.uint32)
        verts = np.ascontiguousarray(verts, np.float32)
        faces = np.ascontiguousarray(faces, np.uint32)
        rays = np.ascontiguousarray(rays)
        assert rays.dtype.itemsize == 36 and nodes.dtype.itemsize == 40
        n = len(rays)
        hits = np.zeros(n, HIT_DTYPE)
        mask = np.zeros(n, np.uint8)
        ctr = Counters()
        self.lib.orc_traverse_batch(_p(nodes), _p(indices), _p(verts), 12, _p(faces), _p(rays), n, _p(hits),
                                    _p(mask), _p(topts) if topts is not None else None, 1 if cpp11 else 0,
                                    threads, C.addressof(ctr) if counters else None)
        if counters:
            return hits, mask, {"nodes_popped": ctr.nodes_popped, "prims_tested": ctr.prims_tested,
                                "max_stack": ctr.max_stack}
        return hits, mask

    def test_prim(self, verts, faces, ray, prim, topts=None, cpp11=True):
        verts = np.ascontiguousarray(verts, np.float32)
        faces = np.ascontiguousarray(faces, np.uint32)
        ray = np.ascontiguousarray(ray).reshape(1)
        out = np.zeros(1, HIT_DTYPE)
        ok = self.lib.orc_test_prim(_p(verts), 12, _p(faces), _p(ray), _p(topts) if topts is not None else None,
                                    1 if cpp11 else 0, int(prim), _p(out))
        return bool(ok), out[0]


class Reference:
    """The unmodified reference header (oracle/_ref).  Raises FileNotFoundError when absent."""

    def __init__(self, cpp11=True):
        name = "libnanort_ref.so" if cpp11 else "libnanort_ref03.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.cpp11 = cpp11
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_build.restype = C.c_void_p
        L.ref_build.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p]
        L.ref_adopt.restype = C.c_void_p
        L.ref_adopt.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_bounding_box.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_num_nodes.restype = C.c_size_t
        L.ref_num_nodes.argtypes = [C.c_void_p]
        L.ref_num_indices.restype = C.c_size_t
        L.ref_num_indices.argtypes = [C.c_void_p]
        L.ref_copy_nodes.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_copy_indices.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_traverse_batch.restype = C.c_size_t
        L.ref_traverse_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int]
        L.ref_traverse_one_f64.restype = C.c_int
        L.ref_traverse_one_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_double,
                                           C.c_double, C.c_void_p, C.c_void_p]
        L.ref_sizes.argtypes = [C.c_void_p]

    @staticmethod
    def available(cpp11=True):
        name = "libnanort_ref.so" if cpp11 else "libnanort_ref03.so"
        return os.path.exists(os.path.join(HERE, "_ref", name))

    def sizes(self):
        s = np.zeros(5, np.uint32)
        self.lib.ref_sizes(_p(s))
        return [int(x) for x in s]

    class Accel:
        def __init__(self, ref, handle, verts, faces):
            self.ref, self.h, self.verts, self.faces = ref, handle, verts, faces  # keep geometry alive

        def __del__(self):
            if self.h:
                self.ref.lib.ref_free(self.h)
                self.h = None

        def stats(self):
            s = np.zeros(3, np.uint32)
            self.ref.lib.ref_stats(self.h, _p(s))
            return {"max_tree_depth": int(s[0]), "num_leaf_nodes": int(s[1]), "num_branch_nodes": int(s[2])}

        def bounding_box(self):
            a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
            self.ref.lib.ref_bounding_box(self.h, _p(a), _p(b))
            return a, b

        def nodes(self):
            n = self.ref.lib.ref_num_nodes(self.h)
            out = np.zeros(n, NODE_DTYPE)
            self.ref.lib.ref_copy_nodes(self.h, _p(out))
            return out

        def indices(self):
            n = self.ref.lib.ref_num_indices(self.h)
            out = np.zeros(n, np.uint32)
            self.ref.lib.ref_copy_indices(self.h, _p(out))
            return out

        def traverse(self, rays, topts=None, threads=1, hits=None, mask=None):
            rays = np.ascontiguousarray(rays)
            assert rays.dtype.itemsize == 36
            n = len(rays)
            if hits is None:
                hits = np.zeros(n, HIT_DTYPE)
            if mask is None:
                mask = np.zeros(n, np.uint8)
            self.ref.lib.ref_traverse_batch(self.h, _p(rays), n, _p(hits), _p(mask),
                                            _p(topts) if topts is not None else None, threads)
            return hits, mask

    def build(self, verts, faces, opts=None):
        verts = np.ascontiguousarray(verts, np.float32)
        faces = np.ascontiguousarray(faces, np.uint32)
        h = self.lib.ref_build(_p(verts), 12, _p(faces), len(faces), _p(opts) if opts is not None else None)
        if not h:
            return None
        return Reference.Accel(self, h, verts, faces)

    def adopt(self, nodes, indices, verts, faces):
        nodes = np.ascontiguousarray(nodes)
        indices = np.ascontiguousarray(indices, np.uint32)
        verts = np.ascontiguousarray(verts, np.float32)
        faces = np.ascontiguousarray(faces, np.uint32)
        h = self.lib.ref_adopt(_p(nodes), len(nodes), _p(indices), len(indices), _p(verts), 12, _p(faces))
        if not h:
            return None
        return Reference.Accel(self, h, verts, faces)

    def traverse_one_f64(self, verts, faces, org, dir, min_t=0.0, max_t=1e30):
        verts = np.ascontiguousarray(verts, np.float64)
        faces = np.ascontiguousarray(faces, np.uint32)
        org = np.ascontiguousarray(org, np.float64)
        dir = np.ascontiguousarray(dir, np.float64)
        out = np.zeros(3, np.float64)
        prim = np.zeros(1, np.uint32)
        r = self.lib.ref_traverse_one_f64(_p(verts), _p(faces), len(faces), _p(org), _p(dir), min_t, max_t,
                                          _p(out), _p(prim))
        return r, out, int(prim[0])
