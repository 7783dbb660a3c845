// ctrb_geom.cuh -- device geometry for the tube-vs-environment distance query.
// Semantics follow what python-fcl gives CollisionChecker.check_collision
// (collision_checker.py:35-95,121-147): each tube segment is a SOLID flat-capped
// cylinder (fcl.Cylinder), a mesh obstacle is a triangle soup (surface), the result
// is the exact Euclidean distance, 0 meaning contact.
#pragma once

#include <cuda_runtime.h>

namespace ctrb {

struct d3 {
    double x, y, z;
};
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ d3 operator*(double s, d3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot(d3 a, d3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ d3 cross(d3 a, d3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float dot(f3 a, f3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }

// squared distance from a point to an fp32 AABB
__device__ __forceinline__ float aabb_dist2(const float lo[3], const float hi[3], f3 p) {
    float dx = fmaxf(fmaxf(lo[0] - p.x, p.x - hi[0]), 0.f);
    float dy = fmaxf(fmaxf(lo[1] - p.y, p.y - hi[1]), 0.f);
    float dz = fmaxf(fmaxf(lo[2] - p.z, p.z - hi[2]), 0.f);
    return fmaf(dx, dx, fmaf(dy, dy, dz * dz));
}

// squared point-triangle distance, closest-feature region walk; T = float or double
template <typename V, typename S>
__device__ __forceinline__ S point_tri_dist2(V p, V a, V b, V c) {
    V ab = b - a, ac = c - a, ap = p - a;
    S d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return dot(ap, ap);
    V bp = p - b;
    S d3_ = dot(ab, bp), d4 = dot(ac, bp);
    if (d3_ >= 0 && d4 <= d3_) return dot(bp, bp);
    S vc = d1 * d4 - d3_ * d2;
    if (vc <= 0 && d1 >= 0 && d3_ <= 0) {
        S v = d1 / (d1 - d3_);
        S q = dot(ap, ap) - v * d1;  // |ap - v ab|^2 = ap.ap - 2 v d1 + v^2 ab.ab, with v = d1/(ab.ab) on the line
        // use the direct form for robustness
        V r = {ap.x - v * ab.x, ap.y - v * ab.y, ap.z - v * ab.z};
        (void)q;
        return dot(r, r);
    }
    V cp = p - c;
    S d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return dot(cp, cp);
    S vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        S w = d2 / (d2 - d6);
        V r = {ap.x - w * ac.x, ap.y - w * ac.y, ap.z - w * ac.z};
        return dot(r, r);
    }
    S va = d3_ * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3_) >= 0 && (d5 - d6) >= 0) {
        S w = (d4 - d3_) / ((d4 - d3_) + (d5 - d6));
        V bc = c - b;
        V r = {bp.x - w * bc.x, bp.y - w * bc.y, bp.z - w * bc.z};
        return dot(r, r);
    }
    // interior: distance to the plane
    V n = {ab.y * ac.z - ab.z * ac.y, ab.z * ac.x - ab.x * ac.z, ab.x * ac.y - ab.y * ac.x};
    S nn = dot(n, n);
    if (nn == 0) return dot(ap, ap);
    S t = dot(ap, n);
    return t * t / nn;
}

// fp32: the vector from p to the closest point of triangle (a, b, c) (same region walk as point_tri_dist2)
__device__ __forceinline__ f3 point_tri_closest(f3 p, f3 a, f3 b, f3 c) {
    const f3 ab = b - a, ac = c - a, ap = p - a;
    const float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return f3{-ap.x, -ap.y, -ap.z};
    const f3 bp = p - b;
    const float d3_ = dot(ab, bp), d4 = dot(ac, bp);
    if (d3_ >= 0 && d4 <= d3_) return f3{-bp.x, -bp.y, -bp.z};
    const float vc = d1 * d4 - d3_ * d2;
    if (vc <= 0 && d1 >= 0 && d3_ <= 0) {
        const float v = d1 / (d1 - d3_);
        return f3{v * ab.x - ap.x, v * ab.y - ap.y, v * ab.z - ap.z};
    }
    const f3 cp = p - c;
    const float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return f3{-cp.x, -cp.y, -cp.z};
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        const float w = d2 / (d2 - d6);
        return f3{w * ac.x - ap.x, w * ac.y - ap.y, w * ac.z - ap.z};
    }
    const float va = d3_ * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3_) >= 0 && (d5 - d6) >= 0) {
        const float w = (d4 - d3_) / ((d4 - d3_) + (d5 - d6));
        const f3 bc = c - b;
        return f3{w * bc.x - bp.x, w * bc.y - bp.y, w * bc.z - bp.z};
    }
    // interior: foot of the perpendicular
    const f3 n = {ab.y * ac.z - ab.z * ac.y, ab.z * ac.x - ab.x * ac.z, ab.x * ac.y - ab.y * ac.x};
    const float nn = dot(n, n);
    if (nn == 0) return f3{-ap.x, -ap.y, -ap.z};
    const float t = -dot(ap, n) / nn;
    return f3{t * n.x, t * n.y, t * n.z};
}

// distance from a point to the solid cylinder {c + t a + u : |t| <= h, u _|_ a, |u| <= r}
__device__ __forceinline__ double point_cyl_dist(d3 x, d3 c, d3 a, double h, double r) {
    d3 y = x - c;
    double t = dot(y, a);
    d3 w = y - t * a;
    double rho = sqrt(dot(w, w));
    double dt = fmax(fabs(t) - h, 0.0), dr = fmax(rho - r, 0.0);
    return sqrt(fma(dt, dt, dr * dr));
}

// Segment vs solid cylinder.  In cylinder coordinates the segment is
//   t(s) = t0 + s tp,  rho(s)^2 = A s^2 + 2 Bq s + Cq,  s in [0,1];
// F(s) = max(|t|-h,0)^2 + max(rho-r,0)^2 is convex and C^1, so F' is monotone and
// its root is bracketed (Illinois regula falsi with a bisection safeguard).
struct SegCyl {
    double t0, tp, A, Bq, Cq, h, r;
    __device__ __forceinline__ double eval(double s, double* g) const {
        double t = fma(s, tp, t0);
        double rho2 = fma(s, fma(s, A, 2.0 * Bq), Cq);
        double rho = sqrt(fmax(rho2, 0.0));
        double dt = fabs(t) - h, dr = rho - r;
        double F = 0.0, gg = 0.0;
        if (dt > 0.0) {
            F = dt * dt;
            gg = 2.0 * dt * (t > 0.0 ? tp : -tp);
        }
        if (dr > 0.0) {
            F = fma(dr, dr, F);
            gg = fma(2.0 * dr, fma(A, s, Bq) / rho, gg);
        }
        *g = gg;
        return F;
    }
};

// returns min over the segment p0->p1 of the SQUARED distance to the cylinder (0 = contact)
__device__ __noinline__ double seg_cyl_dist2(d3 p0, d3 p1, d3 c, d3 a, double h, double r) {
    d3 y0 = p0 - c, e = p1 - p0;
    SegCyl sc;
    sc.t0 = dot(y0, a);
    sc.tp = dot(e, a);
    d3 w0 = y0 - sc.t0 * a, wp = e - sc.tp * a;
    sc.A = dot(wp, wp);
    sc.Bq = dot(w0, wp);
    sc.Cq = dot(w0, w0);
    sc.h = h;
    sc.r = r;
    double g0, g1;
    double F0 = sc.eval(0.0, &g0);
    double F1 = sc.eval(1.0, &g1);
    if (F0 == 0.0 || F1 == 0.0) return 0.0;
    if (g0 >= 0.0) return F0;
    if (g1 <= 0.0) return F1;
    double lo = 0.0, hi = 1.0, glo = g0, ghi = g1, best = fmin(F0, F1);
    int side = 0;
#pragma unroll 1
    for (int it = 0; it < 96; ++it) {
        double s = (it & 3) == 3 ? 0.5 * (lo + hi) : (lo * ghi - hi * glo) / (ghi - glo);
        if (!(s > lo && s < hi)) s = 0.5 * (lo + hi);
        if (!(s > lo && s < hi)) break;
        double g, F = sc.eval(s, &g);
        best = fmin(best, F);
        if (F == 0.0 || g == 0.0) return F;
        if (g > 0.0) {
            hi = s;
            ghi = g;
            if (side == 1) glo *= 0.5;
            side = 1;
        } else {
            lo = s;
            glo = g;
            if (side == -1) ghi *= 0.5;
            side = -1;
        }
        if (hi - lo < 1e-14) break;
        // F is convex and [lo, hi] brackets its minimiser: F(s) - min F <= |F'(s)| (hi - lo)
        if (fabs(g) * (hi - lo) <= 1e-14 * F) break;
    }
    return best;
}

__device__ __forceinline__ bool in_tri_plane(d3 z, d3 t0, d3 t1, d3 t2, d3 n) {
    if (dot(cross(t1 - t0, z - t0), n) < 0.0) return false;
    if (dot(cross(t2 - t1, z - t1), n) < 0.0) return false;
    if (dot(cross(t0 - t2, z - t2), n) < 0.0) return false;
    return true;
}

// Exact distance between the solid cylinder and a triangle; 0 = touching/intersecting.
// The minimum of the convex point-to-cylinder distance over the triangle is on an edge,
// or at the interior point facing the cylinder's support point toward the plane.
// `bound`: callers only need the value if it is below bound; edges are skipped early
// once the plane distance alone exceeds it.
__device__ __noinline__ double cyl_tri_dist(d3 c, d3 a, double h, double r, d3 t0, d3 t1, d3 t2, double bound) {
    d3 n = cross(t1 - t0, t2 - t0);
    double nn = sqrt(dot(n, n));
    double sc = 0.0, E = 0.0;
    d3 so = {0, 0, 0};
    bool has_plane = nn > 0.0;
    if (has_plane) {
        n = (1.0 / nn) * n;
        sc = dot(c - t0, n);
        double an = dot(a, n);
        d3 m = n - an * a;
        double mm = sqrt(dot(m, m));
        if (mm > 0.0) m = (1.0 / mm) * m;
        double sg = (an > 0.0) - (an < 0.0);
        so = (sg * h) * a + r * m;
        E = dot(so, n);
        // the plane distance is a lower bound for the whole triangle
        if (fabs(sc) - E >= bound) return fabs(sc) - E;
    }
    double e2 = seg_cyl_dist2(t0, t1, c, a, h, r);
    if (e2 == 0.0) return 0.0;
    double f2 = seg_cyl_dist2(t1, t2, c, a, h, r);
    if (f2 == 0.0) return 0.0;
    e2 = fmin(e2, f2);
    f2 = seg_cyl_dist2(t2, t0, c, a, h, r);
    if (f2 == 0.0) return 0.0;
    e2 = fmin(e2, f2);
    double e = sqrt(e2);
    if (!has_plane) return e;
    if (fabs(sc) <= E) {
        double lam = (E > 0.0) ? -sc / E : 0.0;
        d3 z = c + lam * so;
        return in_tri_plane(z, t0, t1, t2, n) ? 0.0 : e;
    }
    double sgc = sc > 0.0 ? 1.0 : -1.0;
    d3 y = c - sgc * so;
    double dpl = dot(y - t0, n);
    d3 yp = y - dpl * n;
    if (in_tri_plane(yp, t0, t1, t2, n)) return fmin(fabs(sc) - E, e);
    return e;
}

// ------------------------------------------------------------------ solid cylinder vs solid cylinder (GJK)
// Obstacle fcl.Cylinder (init_collision.py:131-135) against a tube cylinder.  FCL runs libccd's GJK with a
// 1e-6 tolerance on this pair; here the same iteration runs to ~1e-14 in fp64.  The simplex is at most 4
// points of the Minkowski difference; its closest point to the origin comes from the region tests.
struct GjkCyl {
    d3 c, a;
    double h, r;
};

__device__ __forceinline__ d3 cyl_support(const GjkCyl& C, d3 d) {
    const double da = dot(d, C.a);
    d3 w = d - da * C.a;
    const double wn = sqrt(dot(w, w));
    const double sg = da >= 0.0 ? 1.0 : -1.0;
    d3 o = C.c + (sg * C.h) * C.a;
    if (wn > 0.0) o = o + (C.r / wn) * w;
    return o;
}

// closest point to the origin of the simplex W[0..n-1], n <= 3; W is reduced to the supporting face
__device__ __noinline__ int simplex_closest3(d3* W, int n, d3& v) {
    if (n == 1) {
        v = W[0];
        return 1;
    }
    if (n == 2) {
        d3 ab = W[1] - W[0];
        double t = -dot(W[0], ab), den = dot(ab, ab);
        if (t <= 0.0 || den == 0.0) {
            v = W[0];
            return 1;
        }
        if (t >= den) {
            W[0] = W[1];
            v = W[0];
            return 1;
        }
        v = W[0] + (t / den) * ab;
        return 2;
    }
    if (n == 3) {
        const d3 a = W[0], b = W[1], c = W[2];
        const d3 ab = b - a, ac = c - a, ap = -1.0 * a;
        const double d1 = dot(ab, ap), d2 = dot(ac, ap);
        if (d1 <= 0.0 && d2 <= 0.0) { v = a; return 1; }
        const d3 bp = -1.0 * b;
        const double d3v = dot(ab, bp), d4 = dot(ac, bp);
        if (d3v >= 0.0 && d4 <= d3v) { W[0] = b; v = b; return 1; }
        const double vc = d1 * d4 - d3v * d2;
        if (vc <= 0.0 && d1 >= 0.0 && d3v <= 0.0) {
            v = a + (d1 / (d1 - d3v)) * ab;
            return 2;
        }
        const d3 cp = -1.0 * c;
        const double d5 = dot(ab, cp), d6 = dot(ac, cp);
        if (d6 >= 0.0 && d5 <= d6) { W[0] = c; v = c; return 1; }
        const double vb = d5 * d2 - d1 * d6;
        if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
            v = a + (d2 / (d2 - d6)) * ac;
            W[1] = c;
            return 2;
        }
        const double va = d3v * d6 - d5 * d4;
        if (va <= 0.0 && (d4 - d3v) >= 0.0 && (d5 - d6) >= 0.0) {
            v = b + ((d4 - d3v) / ((d4 - d3v) + (d5 - d6))) * (c - b);
            W[0] = b;
            W[1] = c;
            return 2;
        }
        const double den = 1.0 / (va + vb + vc);
        v = a + (vb * den) * ab + (vc * den) * ac;
        return 3;
    }
    return 0;
}

// the same for n <= 4 (no recursion: device stack sizes must be static)
__device__ __noinline__ int simplex_closest(d3* W, int n, d3& v) {
    if (n < 4) return simplex_closest3(W, n, v);
    // tetrahedron: inside iff the four sub-volumes share the sign of the whole volume; else the nearest face
    const d3 e1 = W[1] - W[0], e2 = W[2] - W[0], e3 = W[3] - W[0];
    const double vol = dot(cross(e1, e2), e3);
    const double scale = sqrt(dot(e1, e1) * dot(e2, e2) * dot(e3, e3));
    if (fabs(vol) > 1e-12 * scale) {
        bool inside = true;
#pragma unroll 1
        for (int f = 0; f < 4 && inside; ++f) {
            d3 P[4] = {W[0], W[1], W[2], W[3]};
            P[f] = d3{0.0, 0.0, 0.0};
            if (dot(cross(P[1] - P[0], P[2] - P[0]), P[3] - P[0]) * vol < 0.0) inside = false;
        }
        if (inside) {
            v = d3{0.0, 0.0, 0.0};
            return 4;
        }
    }
    double best = 1.7976931348623157e308;
    d3 bv = {0, 0, 0}, BW[3] = {W[0], W[0], W[0]};
    int bn = 1;
#pragma unroll 1
    for (int f = 0; f < 4; ++f) {
        d3 T3[3];
        int k = 0;
        for (int i = 0; i < 4; ++i)
            if (i != f) T3[k++] = W[i];
        d3 tv;
        const int tn = simplex_closest3(T3, 3, tv);
        const double dd = dot(tv, tv);
        if (dd < best) {
            best = dd;
            bv = tv;
            BW[0] = T3[0];
            BW[1] = T3[1];
            BW[2] = T3[2];
            bn = tn;
        }
    }
    W[0] = BW[0];
    W[1] = BW[1];
    W[2] = BW[2];
    v = bv;
    return bn;
}

__device__ __noinline__ double gjk_cyl_cyl(d3 c1, d3 a1, double h1, double r1, d3 c2, d3 a2, double h2, double r2) {
    const GjkCyl A = {c1, a1, h1, r1}, B = {c2, a2, h2, r2};
    d3 W[4];
    d3 v = c1 - c2;
    if (dot(v, v) == 0.0) return 0.0;
    int n = 0, stall = 0;
    double best_vv = dot(v, v);
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
        const d3 w = cyl_support(A, -1.0 * v) - cyl_support(B, v);
        const double vv = dot(v, v), vw = dot(v, w);
        // |v| decreases monotonically in exact arithmetic; keep the best value and stop once rounding stalls it
        if (vv < best_vv * (1.0 - 1e-15)) stall = 0;
        else ++stall;
        best_vv = fmin(best_vv, vv);
        if (vv - vw <= 1e-14 * vv || stall >= 12) return sqrt(best_vv);
        bool dup = false;
        for (int k = 0; k < n; ++k) {
            const d3 e = w - W[k];
            if (dot(e, e) <= 1e-30 * (1.0 + vv)) dup = true;
        }
        if (dup) return sqrt(best_vv);
        W[n++] = w;
        n = simplex_closest(W, n, v);
        if (n == 4) return 0.0;
        if (dot(v, v) <= 1e-28) return 0.0;
    }
    return sqrt(fmin(best_vv, dot(v, v)));
}

// one rim circle of cylinder 1 (centre o, radius r1, in-plane frame e1, e2) against the solid cylinder 2
__device__ __forceinline__ double rim_point_dist(double phi, d3 o, d3 e1, d3 e2, double r1, d3 c2, d3 a2, double h2,
                                                 double r2) {
    double sn, cs;
    sincos(phi, &sn, &cs);
    return point_cyl_dist(o + (r1 * cs) * e1 + (r1 * sn) * e2, c2, a2, h2, r2);
}

// min over the rim of the distance to the other cylinder: 64 samples, then golden-section refinement of every
// local-minimum bracket that can still undercut the best sample (point_cyl_dist is 1-Lipschitz)
__device__ __noinline__ double rim_min(d3 c1, d3 a1, double h1, double r1, double sg, d3 c2, d3 a2, double h2, double r2) {
    constexpr int NS = 64;
    const double step = 6.283185307179586 / NS, gr = 0.6180339887498949;
    d3 t = fabs(a1.x) < 0.6 ? d3{1.0, 0.0, 0.0} : d3{0.0, 1.0, 0.0};
    d3 e1 = cross(a1, t);
    e1 = (1.0 / sqrt(dot(e1, e1))) * e1;
    const d3 e2 = cross(a1, e1);
    const d3 o = c1 + (sg * h1) * a1;
    double best = 1.7976931348623157e308, fprev, ffirst, fcur;
    ffirst = rim_point_dist(0.0, o, e1, e2, r1, c2, a2, h2, r2);
    if (ffirst == 0.0) return 0.0;
    best = ffirst;
#pragma unroll 1
    for (int i = 1; i < NS; ++i) {
        const double f = rim_point_dist(i * step, o, e1, e2, r1, c2, a2, h2, r2);
        if (f == 0.0) return 0.0;
        best = fmin(best, f);
    }
    double res = best;
    fprev = rim_point_dist(-step, o, e1, e2, r1, c2, a2, h2, r2);
    fcur = ffirst;
#pragma unroll 1
    for (int i = 0; i < NS; ++i) {
        const double fnext = rim_point_dist((i + 1) * step, o, e1, e2, r1, c2, a2, h2, r2);
        if (!(fcur > fprev || fcur > fnext || fcur - r1 * step > res)) {
            double lo = (i - 1) * step, hi = (i + 1) * step;
            double x1 = hi - gr * (hi - lo), x2 = lo + gr * (hi - lo);
            double f1 = rim_point_dist(x1, o, e1, e2, r1, c2, a2, h2, r2);
            double f2 = rim_point_dist(x2, o, e1, e2, r1, c2, a2, h2, r2);
#pragma unroll 1
            for (int it = 0; it < 64; ++it) {
                if (f1 <= f2) {
                    hi = x2; x2 = x1; f2 = f1;
                    x1 = hi - gr * (hi - lo);
                    f1 = rim_point_dist(x1, o, e1, e2, r1, c2, a2, h2, r2);
                } else {
                    lo = x1; x1 = x2; f1 = f2;
                    x2 = lo + gr * (hi - lo);
                    f2 = rim_point_dist(x2, o, e1, e2, r1, c2, a2, h2, r2);
                }
            }
            res = fmin(res, fmin(f1, f2));
        }
        fprev = fcur;
        fcur = fnext;
    }
    return res;
}

// Exact solid cylinder vs solid cylinder.  GJK on two curved bodies converges like 1/k^2 (1e-6 after 256
// iterations), so it is the intersection test only.  Disjoint cylinders: the closest pair has a common normal;
// each closest point is the support point of its cylinder along it, i.e. a rim point unless the normal is
// perpendicular or parallel to the axis, and those support sets (generator, cap disk) still meet a rim of one
// of the bodies except when two generators are closest at interior points = axis distance minus both radii.
__device__ __noinline__ double cyl_cyl_dist(d3 c1, d3 a1, double h1, double r1, d3 c2, d3 a2, double h2, double r2) {
    if (gjk_cyl_cyl(c1, a1, h1, r1, c2, a2, h2, r2) == 0.0) return 0.0;
    const d3 w = c1 - c2;
    const double b = dot(a1, a2), d = dot(a1, w), e = dot(a2, w), den = 1.0 - b * b;
    double s = 0.0, t = 0.0;
    if (den > 1e-12) {
        s = (b * e - d) / den;
        t = (e - b * d) / den;
        if (fabs(s) < h1 && fabs(t) < h2) {
            const d3 y = w + s * a1 - t * a2;
            return fmax(sqrt(dot(y, y)) - r1 - r2, 0.0);
        }
    }
    s = fmin(fmax(s, -h1), h1);
    t = fmin(fmax(t, -h2), h2);
    if (point_cyl_dist(c1 + s * a1, c2, a2, h2, r2) == 0.0 || point_cyl_dist(c2 + t * a2, c1, a1, h1, r1) == 0.0) return 0.0;
    double best = rim_min(c1, a1, h1, r1, 1.0, c2, a2, h2, r2);
    best = fmin(best, rim_min(c1, a1, h1, r1, -1.0, c2, a2, h2, r2));
    best = fmin(best, rim_min(c2, a2, h2, r2, 1.0, c1, a1, h1, r1));
    best = fmin(best, rim_min(c2, a2, h2, r2, -1.0, c1, a1, h1, r1));
    return best;
}

// distance from a point to a solid box
__device__ __forceinline__ double point_box_dist(d3 x, const double* center, const double* rot, const double* dims) {
    d3 y = {x.x - center[0], x.y - center[1], x.z - center[2]};
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double l = y.x * rot[0 * 3 + k] + y.y * rot[1 * 3 + k] + y.z * rot[2 * 3 + k];
        double d = fabs(l) - 0.5 * dims[k];
        if (d > 0.0) s = fma(d, d, s);
    }
    return sqrt(s);
}

}  // namespace ctrb
