"""oracle/wgsl_numpy.py -- TEST INFRASTRUCTURE (imported by tests/ only; never by the product).

A SECOND, independent restatement of the reference shader src/shaders/raycast.comp.wgsl, written from the WGSL text (not from
oracle/wxo_render.c) in vectorised numpy float32, and working on the reference's OWN GPU data model: the three R32Uint atlas
textures, the five mask buffers as u32 words and the origins array, exactly what vdb.atlas() / vdb.masks() / vdb.origins()
upload (src/vdb/vdb345.rs:108-264, src/render/wgpu_context.rs:119-159).  Differences in structure from the C oracle are
deliberate, so that a transcription error in either shows up as a disagreement (tests/test_oracle_cross_check.py):

  * all rays of a frame march together under an `active` mask (the C oracle loops over pixels);
  * every lookup starts at the root: the parent cache of get_vdb_leaf_from_leaf (:360-396) is never used, which is
    legitimate because a lookup is a pure function of the position -- and is thereby tested;
  * texels are fetched from the cubic atlas with atlas_origin_from_idx (:516-519), bits from u32 words (:422-425).

Arithmetic: numpy float32 element-wise operations are IEEE binary32 with one rounding per operation and no contraction,
in the operation order of the WGSL text.  `parity unpinned` applies here as to the C oracle: neither was checked against an
execution of the reference (no rustc / wgpu / Vulkan ICD in this environment).
"""
from __future__ import annotations

import numpy as np

F = np.float32
MAX_STEPS = 1000  # HDDA_MAX_RAY_STEPS (:82)
SCALE = (F(1.0), F(8.0), F(128.0), F(4096.0))  # :83

K_D, K_A, REFLECTIVITY, WALL_I = F(0.7), F(0.3), F(0.9), F(0.1)  # :145-148
BASE_COLOR = np.array([0.4, 0.2, 0.2], F)  # :149
AMBIENT_COLOR = np.array([0.4, 0.4, 0.3], F)  # :150


class Scene:
    """The bind groups 2 and 3 of the shader (:28-58)."""

    def __init__(self, node5s, node4s, node3s, kids5, vals5, kids4, vals4, vals3, origins):
        # textureLoad(tex, vec3(x, y, z)) == tex[x, y, z]  (wgpu_context.rs:119-150 uploads cube[z][y][x] row by row)
        self.node5s, self.node4s, self.node3s = node5s, node4s, node3s
        self.kids5, self.vals5, self.kids4, self.vals4, self.vals3 = kids5, vals5, kids4, vals4, vals3
        self.origins = np.asarray(origins, np.int32).reshape(-1, 4)[:, :3]  # array<vec3<i32>>, stride 16

    # ---- get_vdb_leaf_from_nothing -> node5 -> node4 -> node3 (:398-494), for N positions at once --------------------
    def lookup(self, pos):
        """pos: (N, 3) int32.  Returns dist (u32), num_parents (u32), leaf index (parents[2].idx; valid where num_parents == 3)."""
        n = pos.shape[0]
        dist = np.ones(n, np.uint32)  # "VdbLeaf(vec3(0.0), 1u, 0u, ...)" when no root matches (:411)
        nump = np.zeros(n, np.uint32)
        leaf = np.zeros(n, np.uint32)
        node5_global = (pos >> 12) << 12  # global_to_node (:497-500)
        idx5 = np.full(n, -1, np.int64)
        for k in range(self.origins.shape[0] - 1, -1, -1):  # the first matching origin wins (:402-409)
            idx5[(node5_global == self.origins[k]).all(1)] = k
        a = np.flatnonzero(idx5 >= 0)
        if a.size == 0:
            return dist, nump, leaf
        # node5 (:415-444)
        p5, i5 = pos[a], idx5[a]
        child5 = (p5 & 4095) >> 7  # global_to_local, local_to_child_node
        off5 = (child5[:, 0] << 10) | (child5[:, 1] << 5) | child5[:, 2]  # child_to_offset(…, 5u, 10u)
        in_kid5 = (self.kids5[i5, off5 >> 5] >> (off5 & 31).astype(np.uint32)) & 1
        in_val5 = (self.vals5[i5, off5 >> 5] >> (off5 & 31).astype(np.uint32)) & 1
        dim5 = self.node5s.shape[1] >> 5  # textureDimensions(node5s).y >> 5u
        t5 = child5 + 32 * np.stack([i5 % dim5, (i5 // dim5) % dim5, i5 // (dim5 * dim5)], 1)  # atlas_origin_from_idx
        node4_idx = self.node5s[t5[:, 0], t5[:, 1], t5[:, 2]]
        nump[a] = 1
        dist[a] = np.where(in_val5 == 1, 0, node4_idx).astype(np.uint32)  # in_val first (:431-433), then !in_kid (:435-437)
        go = (in_val5 == 0) & (in_kid5 == 1)
        a, i4 = a[go], node4_idx[go].astype(np.int64)
        if a.size == 0:
            return dist, nump, leaf
        # node4 (:446-475)
        p4 = pos[a]
        child4 = (p4 & 127) >> 3
        off4 = (child4[:, 0] << 8) | (child4[:, 1] << 4) | child4[:, 2]
        in_kid4 = (self.kids4[i4, off4 >> 5] >> (off4 & 31).astype(np.uint32)) & 1
        in_val4 = (self.vals4[i4, off4 >> 5] >> (off4 & 31).astype(np.uint32)) & 1
        dim4 = self.node4s.shape[0] >> 4
        t4 = child4 + 16 * np.stack([i4 % dim4, (i4 // dim4) % dim4, i4 // (dim4 * dim4)], 1)
        node3_idx = self.node4s[t4[:, 0], t4[:, 1], t4[:, 2]]
        nump[a] = 2
        dist[a] = np.where(in_val4 == 1, 0, node3_idx).astype(np.uint32)
        go = (in_val4 == 0) & (in_kid4 == 1)
        a, i3 = a[go], node3_idx[go].astype(np.int64)
        if a.size == 0:
            return dist, nump, leaf
        # node3 (:477-494)
        loc3 = pos[a] & 7
        off3 = (loc3[:, 0] << 6) | (loc3[:, 1] << 3) | loc3[:, 2]
        in_val3 = (self.vals3[i3, off3 >> 5] >> (off3 & 31).astype(np.uint32)) & 1
        dim3 = self.node3s.shape[0] >> 3
        t3 = loc3 + 8 * np.stack([i3 % dim3, (i3 // dim3) % dim3, i3 // (dim3 * dim3)], 1)
        voxel = self.node3s[t3[:, 0], t3[:, 1], t3[:, 2]]
        nump[a] = 3
        leaf[a] = i3.astype(np.uint32)
        dist[a] = np.where(in_val3 == 1, 0, voxel).astype(np.uint32)
        return dist, nump, leaf


def sign11(v):  # :70-76
    return np.where(v < F(0.0), F(-1.0), F(1.0)).astype(F)


def normalize(v):  # v / length(v), length = sqrt(x*x + y*y + z*z)
    d = v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]
    return v / np.sqrt(d)[:, None]


def dot(a, b):
    p = a * b
    return p[:, 0] + p[:, 1] + p[:, 2]


def mix(a, b, t):  # e1 * (1 - e3) + e2 * e3
    return a * (F(1.0) - t) + b * t


def _wall_flat(N):  # :296-306, :328-338
    Np = np.maximum(F(0.0), N)
    Nn = -np.minimum(F(0.0), N)
    c = lambda r, g, b: np.array([r, g, b], F)
    out = c(WALL_I, 0, 0) * Np[:, 0:1]
    out = out + c(0, WALL_I, 0) * Np[:, 1:2]
    out = out + c(0, 0, WALL_I) * Np[:, 2:3]
    out = out + c(WALL_I, WALL_I, 0) * Nn[:, 0:1]
    out = out + c(0, WALL_I, WALL_I) * Nn[:, 1:2]
    out = out + c(WALL_I, 0, WALL_I) * Nn[:, 2:3]
    return out


def _sun_lit(sc, st, hit, step, N, sel):
    """BASE_COLOR + I * sun (x 0.05 when the shadow ray finds an occluder) for the rays `sel` (:199-208, :280-290, :317-325)."""
    sun_dir, sun_rgb, sun_a = st["sun_dir"], st["sun_rgb"], st["sun_a"]
    I = (sun_a * K_D) * dot(np.broadcast_to(-sun_dir, N.shape), N)
    I = np.maximum(F(0.0), I)
    occluded = np.zeros(N.shape[0], bool)
    lit = np.flatnonzero(I != F(0.0))
    if lit.size:
        src = hit["p"][sel][lit] - (F(4e-2) * step[lit]) * hit["mask"][sel][lit].astype(F)
        sh = _hdda_many(sc, src, np.broadcast_to(-sun_dir, src.shape).astype(F))["state"]
        occluded[lit] = sh == 0
    full = BASE_COLOR + I[:, None] * sun_rgb
    dim = BASE_COLOR + (I[:, None] * sun_rgb) * F(0.05)
    return np.where(occluded[:, None], dim, full).astype(F)


def _hdda_many(sc, srcs, dirs):
    with np.errstate(all="ignore"):
        return _hdda_core(sc, np.asarray(srcs, F).copy(), np.asarray(dirs, F))


def _hdda_core(sc, p, dirs):
    """hdda_ray (:84-126) for N rays with origins p (N, 3) (modified in place) and directions dirs (N, 3).
    Returns the HDDAout fields (:128-142): state, p, mask (N, 3 bool), i, and of `leaf`: num_parents, parents[2].idx."""
    n = dirs.shape[0]
    step = sign11(dirs)
    step01 = np.maximum(F(0.0), step)
    idir = (F(1.0) / dirs).astype(F)
    mask = np.zeros((n, 3), bool)
    state = np.full(n, 2, np.uint32)
    it = np.full(n, MAX_STEPS, np.uint32)
    nump = np.zeros(n, np.uint32)
    leaf = np.zeros(n, np.uint32)
    active = np.arange(n)
    for i in range(MAX_STEPS):
        if active.size == 0:
            break
        pa = p[active]
        dist, npar, lf = sc.lookup(np.floor(pa).astype(np.int32))
        nump[active], leaf[active] = npar, lf
        hit = dist == 0  # :95-97
        oob = ~hit & (np.abs(pa) > F(4096.0)).any(1)  # :100-102, tested after the hit
        done = hit | oob
        state[active[hit]] = 0
        state[active[oob]] = 1
        it[active[done]] = i
        keep = ~done
        active, pa, dist, npar = active[keep], pa[keep], dist[keep], npar[keep]
        if active.size == 0:
            break
        size = dist.astype(F) * np.choose(3 - npar.astype(np.int64), SCALE).astype(F)  # :104-111: 3 parents -> 1, 2 -> 8, 1 -> 128, 0 -> 4096
        s3 = size[:, None]
        modv = pa - s3 * np.floor(pa / s3)  # modulo_vec3f (:78-80)
        tmax = idir[active] * (s3 * step01[active] - modv)  # :113
        t = np.minimum(np.minimum(tmax[:, 0], tmax[:, 1]), tmax[:, 2])
        pa = pa + t[:, None] * dirs[active]
        m = (tmax <= tmax[:, [1, 2, 0]]) & (tmax <= tmax[:, [2, 0, 1]])
        mask[active] = m
        pa = pa + (F(4e-4) * step[active]) * m.astype(F)
        p[active] = pa
    return {"state": state, "p": p, "mask": mask, "i": it, "num_parents": nump, "leaf": leaf}


def _sub(hit, sel):
    return {k: v[sel] for k, v in hit.items()}


def _reflect_ray1(sc, st, src, dirs):  # :311-342
    hit = _hdda_many(sc, src, dirs)
    step = sign11(dirs)
    out = dirs.copy()  # state 2: vec3(dir)
    h = np.flatnonzero(hit["state"] == 0)
    if h.size:
        N = normalize((-step[h]) * hit["mask"][h].astype(F))
        out[h] = _sun_lit(sc, st, hit, step[h], N, h)
    o = np.flatnonzero(hit["state"] == 1)
    if o.size:
        out[o] = _wall_flat(normalize((-step[o]) * hit["mask"][o].astype(F)))
    return out


def _reflect_ray2(sc, st, src, dirs):  # :267-309
    hit = _hdda_many(sc, src, dirs)
    step = sign11(dirs)
    out = dirs.copy()
    h = np.flatnonzero(hit["state"] == 0)
    if h.size:
        mf = hit["mask"][h].astype(F)
        N = normalize((-step[h]) * mf)
        rdir = normalize(dirs[h] - (F(2.0) * N) * dot(dirs[h], N)[:, None])
        rsrc = hit["p"][h] - (F(4e-2) * step[h]) * mf
        rcol = _reflect_ray1(sc, st, rsrc, rdir)
        mcol = _sun_lit(sc, st, hit, step[h], N, h)
        out[h] = mix(mcol, rcol, REFLECTIVITY)
    o = np.flatnonzero(hit["state"] == 1)
    if o.size:
        out[o] = _wall_flat(normalize((-step[o]) * hit["mask"][o].astype(F)))
    return out


def _fmod(x, y):  # WGSL `%` on f32: x - y * trunc(x / y)
    return x - y * np.trunc(x / y)


def ray_trace(sc: Scene, st: dict, eye, dirs):
    """ray_trace (:152-265) for N primary rays.  Returns (rgb float32 (N,3), primary HDDAout)."""
    with np.errstate(all="ignore"):
        n = dirs.shape[0]
        hit = _hdda_core(sc, np.broadcast_to(np.asarray(eye, F), (n, 3)).astype(F).copy(), dirs)
        step = sign11(dirs)
        mode = st["render_mode"]
        col = dirs.copy()  # :264
        mf = hit["mask"].astype(F)
        h = np.flatnonzero(hit["state"] == 0)
        if h.size:
            fp = np.floor(hit["p"][h])
            grid = np.zeros((h.size, 3), F)
            g3 = (_fmod(fp, F(8.0)) == F(0.0)).any(1) & (st["show_345"][0] == 1)
            g4 = (_fmod(fp, F(128.0)) == F(0.0)).any(1) & (st["show_345"][1] == 1)
            g5 = (_fmod(fp, F(4096.0)) == F(0.0)).any(1) & (st["show_345"][2] == 1)
            grid[g3] = np.array([-0.1, 0.5, 0.3], F)  # lowest priority first (:157-166 is an if / else-if chain)
            grid[g4] = np.array([0.6, -0.2, -0.2], F)
            grid[g5] = np.array([-0.3, -0.3, 1.0], F)
            m = mf[h]
            if mode == 1:
                c = (grid + F(0.1)) + m * np.array([0.4, 0.4, 0.4], F)
            elif mode == 2:
                t = hit["i"][h].astype(F) / F(200.0)
                c = grid + mix(np.array([0.72, 1.0, 0.99], F), np.array([1.0, 0.0, 0.0], F), t[:, None])
            elif mode == 3:
                N = normalize((-step[h]) * m)
                LN = np.maximum(F(0.0), st["sun_a"] * dot(np.broadcast_to(-st["sun_dir"], N.shape), N))
                I_d = ((K_D * st["sun_rgb"]) * BASE_COLOR) * LN[:, None]
                I_a = (K_A * AMBIENT_COLOR) * BASE_COLOR
                occl = np.zeros(h.size, bool)
                lit = np.flatnonzero(LN != F(0.0))
                if lit.size:
                    src = hit["p"][h][lit] - (F(4e-2) * step[h][lit]) * m[lit]
                    occl[lit] = _hdda_many(sc, src, np.broadcast_to(-st["sun_dir"], src.shape).astype(F))["state"] == 0
                c = np.where(occl[:, None], np.broadcast_to(I_a, I_d.shape), I_a + I_d).astype(F)
            elif mode == 4:
                N = normalize((-step[h]) * m)
                mcol = _sun_lit(sc, st, hit, step[h], N, h)
                rdir = normalize(dirs[h] - (F(2.0) * N) * dot(dirs[h], N)[:, None])
                rsrc = hit["p"][h] - (F(4e-2) * step[h]) * m
                rcol = _reflect_ray2(sc, st, rsrc, rdir)
                c = mix(mcol, rcol, REFLECTIVITY)
            else:  # 0 and `default`
                w = m * np.array([0.2, 0.2, 0.3], F)
                c = grid + ((w[:, 0] + w[:, 1]) + w[:, 2])[:, None]  # dot(…, vec3(1.0))
            col[h] = c
        o = np.flatnonzero(hit["state"] == 1)
        if o.size:
            m = mf[o]
            if mode == 2:
                t = hit["i"][o].astype(F) / F(200.0)
                w = m * np.array([0.04, 0.08, 0.12], F)
                c = mix(np.array([0.72, 1.0, 0.99], F), np.array([1.0, 0.0, 0.0], F), t[:, None]) + ((w[:, 0] + w[:, 1]) + w[:, 2])[:, None]
            elif mode == 4:
                N = normalize((-step[o]) * m)
                Np, Nn = np.maximum(F(0.0), N), -np.minimum(F(0.0), N)
                t = (hit["p"][o][:, 1] / F(4096.0))[:, None]
                cc = lambda r, g, b: np.array([r, g, b], F)
                lo = WALL_I * F(0.1)
                c = mix(cc(WALL_I, 0, 0), cc(lo, 0, 0), t) * Np[:, 0:1]
                c = c + cc(0, WALL_I, 0) * Np[:, 1:2]
                c = c + mix(cc(0, 0, WALL_I), cc(0, 0, lo), t) * Np[:, 2:3]
                c = c + mix(cc(WALL_I, WALL_I, 0), cc(lo, lo, 0), t) * Nn[:, 0:1]
                c = c + cc(0, WALL_I, WALL_I) * Nn[:, 1:2]
                c = c + mix(cc(WALL_I, 0, WALL_I), cc(lo, 0, lo), t) * Nn[:, 2:3]
            else:
                w = m * np.array([0.01, 0.02, 0.03], F)
                c = np.zeros((o.size, 3), F) + ((w[:, 0] + w[:, 1]) + w[:, 2])[:, None]
            col[o] = c
        return col.astype(F), hit


def cp_main(sc: Scene, state_bytes: bytes, width: int, height: int):
    """cp_main (:60-68) over the dispatch of wgpu_context.rs:281 ((W/8) x (H/4) workgroups of 8 x 4).
    Returns rgba uint8 (H, W, 4) -- texels outside the dispatch stay 0 -- and the primary rays' HDDAout as (H, W, ...) arrays."""
    s = np.frombuffer(state_bytes, np.float32)
    u32 = np.frombuffer(state_bytes, np.uint32)
    st = {"render_mode": int(u32[48]), "show_345": [int(v) for v in u32[52:55]], "sun_dir": s[56:59].astype(F),
          "sun_rgb": s[60:63].astype(F), "sun_a": F(s[63])}
    eye, u, mv, wp = s[32:35], s[36:39], s[40:43], s[44:47]
    dw, dh = (width // 8) * 8, (height // 4) * 4
    ys, xs = np.meshgrid(np.arange(dh), np.arange(dw), indexing="ij")
    px = xs.ravel().astype(F) + F(0.001)
    py = ys.ravel().astype(F) + F(0.001)
    with np.errstate(all="ignore"):
        d = (px[:, None] * u.astype(F) + py[:, None] * mv.astype(F)) + wp.astype(F)
        dirs = normalize(d.astype(F)).astype(F)
    rgb, hit = ray_trace(sc, st, eye.astype(F), dirs)
    # rgba8unorm store: clamp to [0, 1], x 255, round half to even; NaN -> 0
    with np.errstate(all="ignore"):
        c = np.where(np.isnan(rgb), F(0.0), np.minimum(np.maximum(rgb, F(0.0)), F(1.0)))
        q = np.rint(c * F(255.0)).astype(np.uint8)
    rgba = np.zeros((height, width, 4), np.uint8)
    rgba[:dh, :dw, :3] = q.reshape(dh, dw, 3)
    rgba[:dh, :dw, 3] = 255
    out = {k: v.reshape(dh, dw, *v.shape[1:]) for k, v in hit.items()}
    return rgba, out
