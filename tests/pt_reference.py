"""numpy restatement of the wavefront path tracer (csrc/wavefront.cuh:PathShadeEpilogue) with the ORACLE doing
every Traverse -- the CPU side of tests/test_gpu_path.py.  Same hash, same random dimensions, float32 math."""
import numpy as np

from nanort_b200 import dist as nd, scenes as S

F = np.float32


def _norm(v):
    l = np.sqrt((v * v).sum(axis=1, dtype=F))
    l = np.where(l > 0, l, F(1))
    return (v / l[:, None]).astype(F), l


def _geo_normal(verts, faces, prim):
    f = faces[prim]
    p0, p1, p2 = verts[f[:, 0]], verts[f[:, 1]], verts[f[:, 2]]
    n = np.cross(p1 - p0, p2 - p0).astype(F)
    nn, l = _norm(n)
    return nn, l


def render(port, nodes, idx, verts, faces, cam, W, H, spp, seed, max_bounces, albedo, emission, light_first, light_n,
           tile=(64, 8)):
    pix, smp = nd.slot_pixels(W, H, tile[0], tile[1], 0, 1, spp)
    ok = pix >= 0
    pix, smp = pix[ok], smp[ok]
    accum = np.zeros((W * H, 3), np.float64)
    rays = S.primary_rays(cam, W, H, spp=1, seed=seed, pixels=np.zeros(0, np.int64))  # dtype only
    # camera rays per (pix, smp)
    jx, jy = S.rand_ps(pix, smp, 0, seed), S.rand_ps(pix, smp, 1, seed)
    cam = np.asarray(cam, F)
    sx = ((pix % W).astype(F) + jx) / F(W) - F(0.5)
    sy = F(0.5) - ((pix // W).astype(F) + jy) / F(H)
    d = (cam[3:6][None] * sx[:, None] + cam[6:9][None] * sy[:, None] + cam[9:12][None]).astype(F)
    d, _ = _norm(d)
    org = np.broadcast_to(cam[0:3], d.shape).astype(F)
    w = np.ones((len(pix), 3), F)
    alive = np.arange(len(pix))
    counts = {"radiance": 0, "shadow": 0, "camera": len(pix)}
    albedo, emission = np.asarray(albedo, F), np.asarray(emission, F)
    for b in range(max_bounces):
        if len(alive) == 0:
            break
        r = np.zeros(len(alive), S.RAY_DTYPE)
        r["org"], r["dir"], r["min_t"], r["max_t"] = org, d, F(1e-3), F(1e30)
        counts["radiance"] += len(r)
        hits, mask = port.traverse(nodes, idx, verts, faces, r, threads=8)
        hit = mask.astype(bool)
        prim = hits["prim_id"]
        is_light = hit & (prim >= light_first) & (prim < light_first + light_n)
        n, _ = _geo_normal(verts, faces, np.where(hit, prim, 0))
        ndotd = (n * d).sum(axis=1, dtype=F)
        if b == 0:
            c = np.maximum(-ndotd, F(0))
            sel = is_light
            np.add.at(accum, pix[alive][sel], (c[sel, None] * emission[None] * w[sel]).astype(np.float64))
        diff = hit & ~is_light
        P = (org + d * hits["t"][:, None]).astype(F)
        n = np.where((ndotd > 0)[:, None], -n, n).astype(F)
        dim = 8 + 8 * b
        pa, sa = pix[alive], smp[alive]
        # NEE
        xi1, xi2 = S.rand_ps(pa, sa, dim + 0, seed), S.rand_ps(pa, sa, dim + 1, seed)
        nf = F(light_n)
        face = np.minimum(np.floor(xi1 * nf).astype(np.int64), light_n - 1)
        xi1 = xi1 * nf - face.astype(F)
        fid = light_first + face
        lf = faces[fid]
        v0, v1, v2 = verts[lf[:, 0]], verts[lf[:, 1]], verts[lf[:, 2]]
        s1 = np.sqrt(xi1)
        c0, c1, c2 = F(1) - s1, s1 * (F(1) - xi2), s1 * xi2
        ln, la2 = _geo_normal(verts, faces, fid)
        area = F(0.5) * la2
        L = (c0[:, None] * v0 + c1[:, None] * v1 + c2[:, None] * v2 - P).astype(F)
        ldir, dist = _norm(L)
        cos_l = np.maximum(-(ldir * ln).sum(axis=1, dtype=F), F(0))
        good = diff & (dist > 1e-6) & (cos_l > 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            pdf = (F(1) / nf) * (F(1) / area) * (dist * dist) / cos_l
            cos_t = np.abs((ldir * n).sum(axis=1, dtype=F))
            k = F(1.0 / np.pi) * cos_l * cos_t / pdf
        gi = np.nonzero(good)[0]
        if len(gi):
            sr = np.zeros(len(gi), S.RAY_DTYPE)
            sr["org"], sr["dir"], sr["min_t"], sr["max_t"] = P[gi], ldir[gi], F(1e-5), dist[gi] - F(1e-5)
            counts["shadow"] += len(sr)
            _, occ = port.traverse(nodes, idx, verts, faces, sr, threads=8)
            vis = occ == 0
            contrib = (k[gi, None] * albedo[None] * emission[None] * w[gi]).astype(np.float64)
            np.add.at(accum, pa[gi][vis], contrib[vis])
        # continuation
        if b + 1 >= max_bounces:
            break
        w = (w * albedo[None]).astype(F)
        keep = diff.copy()
        if b + 1 > 3:
            keep &= S.rand_ps(pa, sa, dim + 4, seed) >= F(0.2)
            w = (w * F(1.0 / 0.8)).astype(F)
        sg = np.where(n[:, 2] >= 0, F(1), F(-1))
        a = F(-1) / (sg + n[:, 2])
        bb = n[:, 0] * n[:, 1] * a
        t1 = np.stack([F(1) + sg * n[:, 0] * n[:, 0] * a, sg * bb, -sg * n[:, 0]], 1).astype(F)
        t2 = np.stack([bb, sg + n[:, 1] * n[:, 1] * a, -n[:, 1]], 1).astype(F)
        u1, u2 = S.rand_ps(pa, sa, dim + 2, seed), S.rand_ps(pa, sa, dim + 3, seed)
        rr = np.sqrt(u1)
        ph = F(6.28318530718) * u2
        hx, hy, hz = rr * np.cos(ph), rr * np.sin(ph), np.sqrt(np.maximum(F(0), F(1) - u1))
        wd = (t1 * hx[:, None] + t2 * hy[:, None] + n * hz[:, None]).astype(F)
        wd, _ = _norm(wd)
        ki = np.nonzero(keep)[0]
        alive, org, d, w = alive[ki], P[ki], wd[ki], w[ki]
    return accum.reshape(H, W, 3), counts
