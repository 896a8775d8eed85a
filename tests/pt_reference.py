"""numpy restatement of the wavefront path tracer (csrc/wavefront.cuh:PathShadeEpilogue, i.e. the shading block
of examples/path_tracer/main.cc:856-976) with the ORACLE doing every Traverse -- the CPU side of
tests/test_gpu_path.py.  Same hash, same random dimensions, float32 math."""
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
    l = np.sqrt((n * n).sum(axis=1, dtype=F))
    nn = n / np.where(l > 0, l, F(1))[:, None]
    return nn.astype(F), l


def render(port, nodes, idx, verts, faces, cam, W, H, spp, seed, max_bounces, mats, mat_ids, emissive, tile=(64, 8)):
    pix, smp = nd.slot_pixels(W, H, tile[0], tile[1], 0, 1, spp)
    ok = pix >= 0
    pix, smp = pix[ok], smp[ok]
    accum = np.zeros((W * H, 3), np.float64)
    jx, jy = S.rand_ps(pix, smp, 0, seed), S.rand_ps(pix, smp, 1, seed)
    cam = np.asarray(cam, F)
    sx = ((pix % W).astype(F) + jx) / F(W) - F(0.5)
    sy = F(0.5) - ((pix // W).astype(F) + jy) / F(H)
    d = (cam[3:6][None] * sx[:, None] + cam[6:9][None] * sy[:, None] + cam[9:12][None]).astype(F)
    d, _ = _norm(d)
    org = np.broadcast_to(cam[0:3], d.shape).astype(F)
    w = np.ones((len(pix), 3), F)
    do_em = np.ones(len(pix), bool)
    counts = {"radiance": 0, "shadow": 0, "camera": len(pix)}
    n_em = len(emissive)
    for b in range(max_bounces):
        if len(pix) == 0:
            break
        r = np.zeros(len(pix), S.RAY_DTYPE)
        r["org"], r["dir"], r["min_t"], r["max_t"] = org, d, F(1e-3), F(1e30)
        counts["radiance"] += len(r)
        hits, mask = port.traverse(nodes, idx, verts, faces, r, threads=8)
        hit = mask.astype(bool)
        prim = np.where(hit, hits["prim_id"], 0)
        n, _ = _geo_normal(verts, faces, prim)
        on = n.copy()
        ndotd = (n * d).sum(axis=1, dtype=F)
        n = np.where((ndotd > 0)[:, None], -n, n).astype(F)
        m = mats[mat_ids[prim]]
        inside = np.where(ndotd < 0, F(-1), F(1))
        with np.errstate(divide="ignore", invalid="ignore"):
            n1 = np.where(inside < 0, F(1) / m["ior"], m["ior"]).astype(F)
            n2 = (F(1) / n1).astype(F)
            r0s = ((n1 - n2) / (n1 + n2)).astype(F)
        r0 = r0s * r0s
        hdn = F(1) - (-(d * n).sum(axis=1, dtype=F))
        fres = (r0 + (F(1) - r0) * (hdn * hdn * hdn * hdn * hdn)).astype(F)
        th = F(1.0 / 3.0)
        avg = lambda a: (th * a[:, 0] + th * a[:, 1] + th * a[:, 2]).astype(F)
        rhoS = avg(m["specular"]) * fres
        rhoD = avg(m["diffuse"]) * (F(1) - fres) * (F(1) - m["dissolve"])
        rhoR = avg(m["transmittance"]) * (F(1) - fres) * m["dissolve"]
        rhoE = avg(m["emission"])
        total = (rhoS + rhoD + rhoR + rhoE).astype(F)
        act = hit & ~(total < F(0.0001))
        with np.errstate(divide="ignore", invalid="ignore"):
            rhoS, rhoD, rhoR = rhoS / total, rhoD / total, rhoR / total
        dim = 8 + 8 * b
        pick = S.rand_ps(pix, smp, dim + 5, seed)
        P = (org + d * hits["t"][:, None]).astype(F)
        is_s = act & (pick < rhoS)
        is_d = act & ~is_s & (pick < rhoS + rhoD)
        is_r = act & ~is_s & ~is_d & (pick < rhoD + rhoS + rhoR)
        is_e = act & ~is_s & ~is_d & ~is_r
        out = np.zeros_like(d)
        # glossy
        k = F(2) * (d * n).sum(axis=1, dtype=F)
        out[is_s] = (d - k[:, None] * n)[is_s]
        # diffuse + NEE
        if n_em > 0 and is_d.any():
            xi1, xi2 = S.rand_ps(pix, smp, dim + 0, seed), S.rand_ps(pix, smp, dim + 1, seed)
            nf = F(n_em)
            face = np.minimum(np.floor(xi1 * nf).astype(np.int64), n_em - 1)
            xi1 = xi1 * nf - face.astype(F)
            fid = emissive[face]
            lm = mats[mat_ids[fid]]
            lf = faces[fid]
            v0, v1, v2 = verts[lf[:, 0]], verts[lf[:, 1]], verts[lf[:, 2]]
            s1 = np.sqrt(xi1)
            c0, c1, c2 = F(1) - s1, s1 * (F(1) - xi2), s1 * xi2
            ln, la2 = _geo_normal(verts, faces, fid)
            area = F(0.5) * la2
            L = (c0[:, None] * v0 + c1[:, None] * v1 + c2[:, None] * v2 - P).astype(F)
            dist = np.sqrt((L * L).sum(axis=1, dtype=F))
            ldir = (L * (F(1) / np.where(dist > 0, dist, F(1)))[:, None]).astype(F)
            cos_l = np.maximum(-(ldir * ln).sum(axis=1, dtype=F), F(0))
            good = is_d & (dist > 1e-6) & (area > 0) & (cos_l > 0)
            with np.errstate(divide="ignore", invalid="ignore"):
                pdf = (F(1) / nf) * (F(1) / area) * (dist * dist) / cos_l
                cos_t = np.abs((ldir * n).sum(axis=1, dtype=F))
                kk = F(1.0 / np.pi) * cos_l * cos_t / pdf
            gi = np.nonzero(good)[0]
            if len(gi):
                sr = np.zeros(len(gi), S.RAY_DTYPE)
                sr["org"], sr["dir"], sr["min_t"], sr["max_t"] = P[gi], ldir[gi], F(1e-5), dist[gi] - F(1e-5)
                counts["shadow"] += len(sr)
                _, occ = port.traverse(nodes, idx, verts, faces, sr, threads=8)
                vis = occ == 0
                contrib = (kk[gi, None] * m["diffuse"][gi] * lm["emission"][gi] * w[gi]).astype(np.float64)
                np.add.at(accum, pix[gi][vis], contrib[vis])
        sg = np.where(n[:, 2] >= 0, F(1), F(-1))
        a = F(-1) / (sg + n[:, 2])
        bb = n[:, 0] * n[:, 1] * a
        t1 = np.stack([F(1) + sg * n[:, 0] * n[:, 0] * a, sg * bb, -sg * n[:, 0]], 1).astype(F)
        t2 = np.stack([bb, sg + n[:, 1] * n[:, 1] * a, -n[:, 1]], 1).astype(F)
        u1, u2 = S.rand_ps(pix, smp, dim + 2, seed), S.rand_ps(pix, smp, dim + 3, seed)
        rr = np.sqrt(u1)
        ph = F(6.28318530718) * u2
        hx, hy, hz = rr * np.cos(ph), rr * np.sin(ph), np.sqrt(np.maximum(F(0), F(1) - u1))
        wd = (t1 * hx[:, None] + t2 * hy[:, None] + n * hz[:, None]).astype(F)
        out[is_d] = wd[is_d]
        # refraction
        rn = (-inside[:, None] * on).astype(F)
        ndi = (rn * d).sum(axis=1, dtype=F)
        kr = F(1) - n1 * n1 * (F(1) - ndi * ndi)
        c = n1 * ndi + np.sqrt(np.maximum(kr, F(0)))
        refr = (n1[:, None] * d - c[:, None] * rn).astype(F)
        refr[kr < 0] = 0
        out[is_r] = refr[is_r]
        # emission
        sel = is_e & do_em
        ce = np.maximum(-(on * d).sum(axis=1, dtype=F), F(0))
        np.add.at(accum, pix[sel], (ce[sel, None] * m["emission"][sel] * w[sel]).astype(np.float64))
        # weights / flags
        w = w.copy()
        w[is_s] = (w * m["specular"])[is_s]
        w[is_d] = (w * m["diffuse"])[is_d]
        w[is_r] = (w * m["transmittance"])[is_r]
        do_em = np.where(is_d, False, np.where(is_s | is_r, True, do_em))
        keep = is_s | is_d | is_r
        if b + 1 >= max_bounces:
            break
        if b + 1 > 3:
            keep &= S.rand_ps(pix, smp, dim + 4, seed) >= F(0.2)
            w = (w * F(1.0 / 0.8)).astype(F)
        ki = np.nonzero(keep)[0]
        pix, smp, org, d, w, do_em = pix[ki], smp[ki], P[ki], out[ki].astype(F), w[ki].astype(F), do_em[ki]
    return accum.reshape(H, W, 3), counts
