"""py_oracle.py — TEST INFRASTRUCTURE: a second, independent CPU restatement of the reference's path-tracing hot path
(Assets/Scripts/Tracer/RayCompute.compute:10-24 + Assets/Scripts/Tracer/RayCommon.hlsl), written again from the HLSL in
numpy float32 scalar arithmetic, pure-Python loops, for a handful of pixels.

Purpose: the C++ oracle (rt_oracle.cpp) defines "correct" for the CUDA kernels, and the reference ships nothing to pin it
against.  Two restatements written separately and agreeing bit for bit on the same pixels make a transcription slip in
either one visible (tests/test_oracle.py::test_cpp_oracle_equals_the_python_restatement).  Same arithmetic contract as
rt_oracle_math.h: every operation one IEEE binary32 op in source order, NaN-ignoring min/max, the pinned polynomial
log / exp / sin / cos.  Only tests import this file.
"""
from __future__ import annotations

import numpy as np

F = np.float32
INF = F(np.inf)
NAN = F(np.nan)
ZERO, ONE, TWO, HALF = F(0.0), F(1.0), F(2.0), F(0.5)


def _u(x):      # float32 -> uint32 bits
    return int(np.array(x, dtype=np.float32).view(np.uint32))


def _f(bits):   # uint32 bits -> float32
    return np.array(bits & 0xFFFFFFFF, dtype=np.uint32).view(np.float32)[()]


def f2uint_rz(f):   # float -> uint as the GPU converts: truncation, NaN / negative -> 0, saturating
    if not (f > 0):
        return 0
    return 0xFFFFFFFF if f >= 4294967296.0 else int(f)


def fmin(a, b):     # NaN-ignoring, -0 below +0 (the contract of DESIGN.md section 2)
    if a != a: return b
    if b != b: return a
    if a == b: return a if np.signbit(a) else b
    return a if a < b else b


def fmax(a, b):
    if a != a: return b
    if b != b: return a
    if a == b: return b if np.signbit(a) else a
    return a if a > b else b


# ---- pinned transcendental routines (same published schemes as rt_oracle_math.h, re-derived here) -----------------------

def log_rt(x):
    ix = _u(x)
    if x != x:
        return x
    if (ix & 0x7FFFFFFF) == 0:
        return -INF
    if ix & 0x80000000:
        return NAN
    if ix == 0x7F800000:
        return x
    k = 0
    if ix < 0x00800000:
        x = x * F(33554432.0); ix = _u(x); k = -25
    k += (ix >> 23) - 127
    m = _f((ix & 0x007FFFFF) | 0x3F800000)
    if m > F(1.41421354):
        m = m * HALF; k += 1
    f = m - ONE
    s = f / (TWO + f)
    z = s * s
    w = z * z
    t1 = w * (F(0.40000972152) + w * F(0.24279078841))
    t2 = z * (F(0.66666662693) + w * F(0.28498786688))
    R = t2 + t1
    hfsq = (HALF * f) * f
    dk = F(k)
    return dk * F(6.9313812256e-01) - ((hfsq - (s * (hfsq + R) + dk * F(9.0580006145e-06))) - f)


def exp_rt(x):
    if x != x:
        return x
    if x > F(88.72283935546875):
        return INF
    if x < F(-103.972076416015625):
        return ZERO
    fk = x * F(1.4426950216e+00) + (F(-0.5) if x < ZERO else F(0.5))
    k = int(fk)                                   # truncation toward zero
    t = F(k)
    hi = x - t * F(6.9314575195e-01)
    lo = t * F(1.4286067653e-06)
    r = hi - lo
    rr = r * r
    c = r - rr * (F(1.6666625440e-1) + rr * F(-2.7667332906e-3))
    y = ONE - ((lo - (r * c) / (TWO - c)) - hi)
    k1 = int(k / 2)                               # C integer division truncates toward zero
    k2 = k - k1
    return (y * _f((k1 + 127) << 23)) * _f((k2 + 127) << 23)


def _reduce(ax):
    n = int(ax * F(0.636619772367581343) + HALF)
    fn = F(n)
    r = ((ax - fn * F(1.5703125)) - fn * F(4.837512969970703125e-4)) - fn * F(7.549789948768648e-8)
    return n & 3, r


def _sin_poly(r):
    z = r * r
    return ((((F(-1.9515295891e-4) * z + F(8.3321608736e-3)) * z) - F(1.6666654611e-1)) * z) * r + r


def _cos_poly(r):
    z = r * r
    y = (((F(2.443315711809948e-5) * z - F(1.388731625493765e-3)) * z) + F(4.166664568298827e-2)) * (z * z)
    return (y - HALF * z) + ONE


def sin_rt(x):
    ax = abs(x)
    if not (ax <= F(100000.0)):
        return NAN
    q, r = _reduce(ax)
    v = _cos_poly(r) if (q & 1) else _sin_poly(r)
    if q & 2:
        v = -v
    return -v if x < ZERO else v


def cos_rt(x):
    ax = abs(x)
    if not (ax <= F(100000.0)):
        return NAN
    q, r = _reduce(ax)
    v = _sin_poly(r) if (q & 1) else _cos_poly(r)
    if q == 1 or q == 2:
        v = -v
    return v


def pow_rt(x, y):
    return exp_rt(y * log_rt(x))


# ---- small vector helpers (tuples of float32; component-wise, source order) -----------------------------------------------

def v3(x, y, z): return (F(x), F(y), F(z))
def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def mulv(a, b): return (a[0] * b[0], a[1] * b[1], a[2] * b[2])
def muls(a, s): return (a[0] * s, a[1] * s, a[2] * s)
def smul(s, a): return (s * a[0], s * a[1], s * a[2])
def neg(a): return (-a[0], -a[1], -a[2])
def dot(a, b): return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]
def cross(a, b): return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
def normalize(a): return muls(a, ONE / np.sqrt(dot(a, a)))
def lerp(a, b, t): return add(a, smul(t, sub(b, a)))
def sign(x): return F(1.0 if x > ZERO else 0.0) - F(1.0 if x < ZERO else 0.0)
def saturate(x): return fmin(fmax(x, ZERO), ONE)


def smoothstep(a, b, x):
    t = saturate((x - a) / (b - a))
    return t * t * (F(3.0) - TWO * t)


def mul_mat(m, v, w):
    """rows 0..2 of (column-major 4x4 m) * (v, w), each row summed left to right with all four products"""
    w = F(w)
    return tuple(((m[r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r] * w for r in range(3))


class PyShader:
    """One pixel at a time; uniforms / buffers as plain attributes (names as in RayCommon.hlsl:5-26,115-121)."""

    PI = F(3.1415)

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.state = 0

    # RayCommon.hlsl:127-138
    def next_random(self):
        self.state = (self.state * 747796405 + 2891336453) & 0xFFFFFFFF
        s = self.state
        result = (((s >> ((s >> 28) + 4)) ^ s) * 277803737) & 0xFFFFFFFF
        return (result >> 22) ^ result

    def random_value(self):
        return F(self.next_random()) / F(4294967295.0)       # the literal rounds to 2^32 in binary32

    # :141-157
    def random_normal(self):
        theta = F(2.0) * F(3.1415926) * self.random_value()
        rho = np.sqrt(F(-2.0) * log_rt(self.random_value()))
        return rho * cos_rt(theta)

    def random_direction(self):
        x = self.random_normal(); y = self.random_normal(); z = self.random_normal()
        return normalize((x, y, z))

    # :159-164
    def random_point_in_circle(self):
        angle = self.random_value() * TWO * self.PI
        px, py = cos_rt(angle), sin_rt(angle)
        rad = np.sqrt(self.random_value())
        return px * rad, py * rad

    # :167-183
    def environment(self, d):
        if self.UseSky == 0:
            return v3(0, 0, 0)
        ground, horizon, zenith = v3(0.35, 0.3, 0.35), v3(1, 1, 1), v3(0.08, 0.37, 0.73)
        sky_t = pow_rt(smoothstep(ZERO, F(0.4), d[1]), F(0.35))
        g2s = smoothstep(F(-0.01), ZERO, d[1])
        sky = lerp(horizon, zenith, sky_t)
        s = F(1000.0) * ONE / self.SunFocus
        sun = pow_rt(fmax(ZERO, dot(d, self.dirToSun)), s) * self.SunIntensity
        return add(lerp(ground, sky, g2s), muls(smul(sun, self.SunColour), F(1.0 if g2s >= ONE else 0.0)))

    # :188-215
    def ray_triangle(self, pos, d, tri, cull):
        A, B, C = (tuple(tri[k]) for k in ("posA", "posB", "posC"))
        ab, ac = sub(B, A), sub(C, A)
        n = cross(ab, ac)
        ao = sub(pos, A)
        dao = cross(ao, d)
        det = -dot(d, n)
        inv = ONE / det
        dst = dot(ao, n) * inv
        u = dot(ac, dao) * inv
        v = -dot(ab, dao) * inv
        w = ONE - u - v
        keep = det >= F(1e-8) if cull else abs(det) >= F(1e-8)
        hit = bool(keep and dst > ZERO and u >= ZERO and v >= ZERO and w >= ZERO)
        sm = normalize(add(add(muls(tuple(tri["normA"]), w), muls(tuple(tri["normB"]), u)), muls(tuple(tri["normC"]), v)))
        return hit, det < ZERO, dst, muls(sm, sign(det))

    # :219-231
    @staticmethod
    def ray_box(pos, inv, bmin, bmax):
        tmin = mulv(sub(bmin, pos), inv)
        tmax = mulv(sub(bmax, pos), inv)
        t1 = tuple(fmin(a, b) for a, b in zip(tmin, tmax))
        t2 = tuple(fmax(a, b) for a, b in zip(tmin, tmax))
        near = fmax(fmax(t1[0], t1[1]), t1[2])
        far = fmin(fmin(t2[0], t2[1]), t2[2])
        if far >= near and far > ZERO:
            return near if near > ZERO else ZERO
        return INF

    # :234-287
    def ray_bvh(self, pos, d, inv, ray_length, node_offset, tri_offset, cull):
        best = (ray_length, False, None)                     # dst, isBackface, normal
        stack = [node_offset]
        while stack:
            node = self.Nodes[stack.pop()]
            if node["triangleCount"] > 0:
                for i in range(int(node["triangleCount"])):
                    hit, back, dst, nrm = self.ray_triangle(pos, d, self.Triangles[tri_offset + int(node["startIndex"]) + i], cull)
                    if hit and dst < best[0]:
                        best = (dst, back, nrm)
            else:
                ia = node_offset + int(node["startIndex"]); ib = ia + 1
                ca, cb = self.Nodes[ia], self.Nodes[ib]
                da = self.ray_box(pos, inv, tuple(ca["boundsMin"]), tuple(ca["boundsMax"]))
                db = self.ray_box(pos, inv, tuple(cb["boundsMin"]), tuple(cb["boundsMax"]))
                near_a = da <= db
                dn, df = (da, db) if near_a else (db, da)
                i_near, i_far = (ia, ib) if near_a else (ib, ia)
                if df < best[0]:
                    stack.append(i_far)
                if dn < best[0]:
                    stack.append(i_near)
        return best

    # :289-320 with the sphere's own material (extension)
    @staticmethod
    def ray_sphere(pos, d, centre, radius):
        off = sub(pos, centre)
        a = dot(d, d)
        b = TWO * dot(off, d)
        c = dot(off, off) - radius * radius
        disc = b * b - F(4.0) * a * c
        if disc >= ZERO:
            s = np.sqrt(disc)
            near = fmax(ZERO, (-b - s) / (TWO * a))
            far = (-b + s) / (TWO * a)
            if far >= ZERO:
                inside = bool(near == ZERO)
                dst = far if inside else near
                p = add(pos, muls(d, dst))
                return True, inside, dst, p, muls(normalize(sub(p, centre)), F(-1.0 if inside else 1.0))
        return False, False, INF, None, None

    # :335-374 (+ spheres before the models, where the commented call sits, :341)
    def collide(self, pos, d):
        res = dict(hit=False, back=False, dst=INF, normal=None, pos=None, mat=None)
        for sp in self.Spheres:
            ok, inside, dst, p, n = self.ray_sphere(pos, d, tuple(sp["centre"]), sp["radius"])
            if ok and dst < res["dst"]:
                res = dict(hit=True, back=inside, dst=dst, normal=n, pos=p, mat=sp["material"])
        for i in range(self.modelCount):
            model = self.ModelInfo[i]
            w2l, l2w = model["worldToLocal"], model["localToWorld"]
            lp = mul_mat(w2l, pos, 1.0)
            ld = mul_mat(w2l, d, 0.0)
            linv = tuple(ONE / c for c in ld)
            cull = int(model["material"]["flag"]) != 2
            dst, back, nrm = self.ray_bvh(lp, ld, linv, res["dst"], int(model["nodeOffset"]), int(model["triOffset"]), cull)
            if dst < res["dst"]:
                res = dict(hit=True, back=back, dst=dst, normal=normalize(mul_mat(l2w, nrm, 0.0)), pos=add(pos, muls(d, dst)), mat=model["material"])
        return res

    # :383-437
    @staticmethod
    def reflectance(i, n, ior_a, ior_b):
        ratio = ior_a / ior_b
        cos_in = -dot(i, n)
        sin2 = ratio * ratio * (ONE - cos_in * cos_in)
        if sin2 >= ONE:
            return ONE
        cos_r = np.sqrt(ONE - sin2)
        den_perp = ior_a * cos_in + ior_b * cos_r
        den_par = ior_a * cos_in + ior_b * cos_r                  # as in the reference (:392)
        if fmin(den_perp, den_par) < F(1e-8):
            return ONE
        rp = (ior_a * cos_in - ior_b * cos_r) / den_perp
        rp = rp * rp
        rl = (ior_b * cos_in - ior_a * cos_r) / den_par
        rl = rl * rl
        return (rp + rl) / TWO

    @staticmethod
    def refract(i, n, ior_a, ior_b):
        ratio = ior_a / ior_b
        cos_in = -dot(i, n)
        sin2 = ratio * ratio * (ONE - cos_in * cos_in)
        if sin2 > ONE:
            return v3(0, 0, 0)
        return add(smul(ratio, i), smul(ratio * cos_in - np.sqrt(ONE - sin2), n))

    # :450-466
    @staticmethod
    def material_colour(mat, pos, n, specular):
        col = tuple(mat["diffuseCol"][:3])
        if int(mat["flag"]) == 1:
            cp = (pos[0], pos[2])
            if abs(n[0]) > abs(n[1]):
                cp = (pos[2], pos[1])
            if abs(n[2]) > fmax(abs(n[0]), abs(n[1])):
                cp = (pos[0], pos[1])
            cp = (cp[0] * F(1.5), cp[1] * F(1.5))
            fx, fy = np.floor(cp[0]), np.floor(cp[1])
            cx = fx - TWO * np.floor(fx / TWO)
            cy = fy - TWO * np.floor(fy / TWO)
            if not (cx == cy):
                col = tuple(mat["emissionCol"][:3])
        return lerp(col, tuple(mat["specularCol"][:3]), F(1.0 if specular else 0.0))

    # :479-542
    def trace(self, pos, d):
        eps = F(0.001)
        total = v3(0, 0, 0)
        trans = v3(1, 1, 1)
        for _ in range(0, self.MaxBounceCount + 1):
            hit = self.collide(pos, d)
            if not hit["hit"]:
                if self.UseSky:
                    total = add(total, mulv(trans, self.environment(d)))
                break
            mat, n = hit["mat"], hit["normal"]
            if int(mat["flag"]) == 2:
                if hit["back"]:
                    ab = muls(smul(-hit["dst"], tuple(mat["absorption"][:3])), mat["absorptionStrength"])
                    trans = mulv(trans, tuple(exp_rt(c) for c in ab))
                ior_cur = mat["ior"] if hit["back"] else ONE
                ior_next = ONE if hit["back"] else mat["ior"]
                refl = sub(d, smul(TWO * dot(d, n), n))
                refr = self.refract(d, n, ior_cur, ior_next)
                weight = self.reflectance(d, n, ior_cur, ior_next)
                diffuse = normalize(add(n, self.random_direction()))
                refl = normalize(lerp(diffuse, refl, mat["specularProbability"]))
                refr = normalize(lerp(neg(diffuse), refr, mat["smoothness"]))
                follow = self.random_value() <= weight
                d = refl if follow else refr
                pos = add(hit["pos"], muls(smul(eps, n), sign(dot(n, d))))
            else:
                specular = bool(mat["specularProbability"] >= self.random_value())
                pos = add(hit["pos"], muls(n, eps))
                diffuse = normalize(add(n, self.random_direction()))
                spec_dir = sub(d, smul(TWO * dot(n, d), n))              # HLSL reflect(i, n)
                d = normalize(lerp(diffuse, spec_dir, mat["smoothness"] * F(1.0 if specular else 0.0)))
                emitted = muls(tuple(mat["emissionCol"][:3]), mat["emissionStrength"])
                total = add(total, mulv(emitted, trans))
                trans = mulv(trans, self.material_colour(mat, hit["pos"], n, specular))
            p = fmax(trans[0], fmax(trans[1], trans[2]))
            if self.random_value() >= p:
                break
            trans = muls(trans, ONE / p)
        return total

    # RayCompute.compute:15 + RayCommon.hlsl:545-582
    def pixel(self, x, y):
        W, H = self.Resolution
        uv = (F(x) / (F(W) - ONE), F(y) / (F(H) - ONE))
        cam = self.CamLocalToWorldMatrix
        origin = mul_mat(cam, v3(0, 0, 0), 1.0)
        px, py = f2uint_rz(uv[0] * F(W)), f2uint_rz(uv[1] * F(H))
        self.state = (py * W + px + self.Frame * 719393 + self.renderSeed) & 0xFFFFFFFF
        focus_local = mulv((uv[0] - HALF, uv[1] - HALF, ONE), self.ViewParams)
        focus = mul_mat(cam, focus_local, 1.0)
        right, up = (cam[0], cam[1], cam[2]), (cam[4], cam[5], cam[6])
        total = v3(0, 0, 0)
        for _ in range(self.NumRaysPerPixel):
            jx, jy = self.random_point_in_circle()
            jx, jy = jx * self.DefocusStrength / F(W), jy * self.DefocusStrength / F(W)
            ro = add(add(origin, muls(right, jx)), muls(up, jy))
            kx, ky = self.random_point_in_circle()
            kx, ky = kx * self.DivergeStrength / F(W), ky * self.DivergeStrength / F(W)
            fp = add(add(focus, muls(right, kx)), muls(up, ky))
            total = add(total, self.trace(ro, normalize(sub(fp, ro))))
        return tuple(c / F(self.NumRaysPerPixel) for c in total)
