/*
 * ctr_oracle.c -- CPU restatement of the CTR DaPP hot path (kinematic model,
 * index rounding, eta / follow-the-leader, chain-vs-environment distance, cost).
 *
 * TEST INFRASTRUCTURE ONLY -- see ctr_oracle.h for who may load this and for
 * the parity status of each part.  Reference citations are file:line in
 * ConorMesser/ctr-design-and-path-plan.
 *
 * Build: see oracle/Makefile  (gcc -O2 -pthread -shared -fPIC).
 * -ffast-math must NOT be used: the restatement relies on IEEE semantics
 * (exact theta == 0 branch, round-half-even).
 */
#include "ctr_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------ */
/* small dense helpers                                                  */
/* ------------------------------------------------------------------ */
static void mat4_mul(const double* a, const double* b, double* out) {
    double t[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            t[i * 4 + j] = s;
        }
    memcpy(out, t, sizeof t);
}

static void mat4_eye(double* g) {
    memset(g, 0, 16 * sizeof(double));
    g[0] = g[5] = g[10] = g[15] = 1.0;
}

/* SE(3) inverse (R^T, -R^T p): what numpy.linalg.inv returns for an SE(3)
 * matrix up to LU rounding (kinematic.py:124,191). */
static void se3_inv(const double* g, double* out) {
    double t[16];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) t[i * 4 + j] = g[j * 4 + i];
    for (int i = 0; i < 3; ++i)
        t[i * 4 + 3] = -(t[i * 4 + 0] * g[3] + t[i * 4 + 1] * g[7] + t[i * 4 + 2] * g[11]);
    t[12] = t[13] = t[14] = 0.0;
    t[15] = 1.0;
    memcpy(out, t, sizeof t);
}

static void mat6_vec(const double* m, const double* v, double* out) {
    double t[6];
    for (int i = 0; i < 6; ++i) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += m[i * 6 + k] * v[k];
        t[i] = s;
    }
    memcpy(out, t, sizeof t);
}

/* ------------------------------------------------------------------ */
/* matrix_utils.py                                                      */
/* ------------------------------------------------------------------ */

/* matrix_utils.py:115-135 */
void orc_tilde(const double v[3], double out[9]) {
    memset(out, 0, 9 * sizeof(double));
    out[1] = -v[2];
    out[2] = v[1];
    out[3] = v[2];
    out[5] = -v[0];
    out[6] = -v[1];
    out[7] = v[0];
}

/* matrix_utils.py:138-154 */
void orc_hat(const double s[6], double out[16]) {
    double sk[9];
    orc_tilde(s, sk);
    memset(out, 0, 16 * sizeof(double));
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) out[i * 4 + j] = sk[i * 3 + j];
        out[i * 4 + 3] = s[3 + i];
    }
}

static void adjoint_blocks(const double g[16], double R[9], double pR[9]) {
    double p[3] = {g[3], g[7], g[11]}, pt[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = g[i * 4 + j];
    orc_tilde(p, pt);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += pt[i * 3 + k] * R[k * 3 + j];
            pR[i * 3 + j] = s;
        }
}

/* matrix_utils.py:8-33  Ad(g) = [[R,0],[p~R,R]] */
void orc_big_adjoint(const double g[16], double out[36]) {
    double R[9], pR[9];
    adjoint_blocks(g, R, pR);
    memset(out, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            out[i * 6 + j] = R[i * 3 + j];
            out[(i + 3) * 6 + j] = pR[i * 3 + j];
            out[(i + 3) * 6 + j + 3] = R[i * 3 + j];
        }
}

/* matrix_utils.py:36-57  coAd(g) = [[R,p~R],[0,R]] */
void orc_big_coadjoint(const double g[16], double out[36]) {
    double R[9], pR[9];
    adjoint_blocks(g, R, pR);
    memset(out, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            out[i * 6 + j] = R[i * 3 + j];
            out[i * 6 + j + 3] = pR[i * 3 + j];
            out[(i + 3) * 6 + j + 3] = R[i * 3 + j];
        }
}

/* matrix_utils.py:60-85 */
void orc_little_adjoint(const double s[6], double out[36]) {
    double w[9], v[9];
    orc_tilde(s, w);
    orc_tilde(s + 3, v);
    memset(out, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            out[i * 6 + j] = w[i * 3 + j];
            out[(i + 3) * 6 + j] = v[i * 3 + j];
            out[(i + 3) * 6 + j + 3] = w[i * 3 + j];
        }
}

/* matrix_utils.py:88-112 */
void orc_little_coadjoint(const double s[6], double out[36]) {
    double w[9], v[9];
    orc_tilde(s, w);
    orc_tilde(s + 3, v);
    memset(out, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            out[i * 6 + j] = w[i * 3 + j];
            out[i * 6 + j + 3] = v[i * 3 + j];
            out[(i + 3) * 6 + j + 3] = w[i * 3 + j];
        }
}

/* matrix_utils.py:157-177 -- literal: I + G + a G^2 + c G^3 with
 * matrix_power(G,2) = G@G and matrix_power(G,3) = (G@G)@G, exact theta == 0
 * branch returning I + G. */
void orc_exponential_map(double theta, const double gh[16], double out[16]) {
    double eye[16];
    mat4_eye(eye);
    if (theta == 0.0) {
        for (int i = 0; i < 16; ++i) out[i] = eye[i] + gh[i];
        return;
    }
    double g2[16], g3[16];
    mat4_mul(gh, gh, g2);
    mat4_mul(g2, gh, g3);
    double a = (1.0 - cos(theta)) / (theta * theta);
    double c = (theta - sin(theta)) / (theta * theta * theta);
    for (int i = 0; i < 16; ++i) out[i] = ((eye[i] + gh[i]) + a * g2[i]) + c * g3[i];
}

/* ------------------------------------------------------------------ */
/* strain_bases.py                                                      */
/* ------------------------------------------------------------------ */

/* strain_bases.py:7-224; base is 6 x dof row-major with row stride ORC_MAX_DOF */
int orc_strain_base(int kind, double x, int dof, double* base) {
    memset(base, 0, 6 * ORC_MAX_DOF * sizeof(double));
#define B(r, c) base[(r) * ORC_MAX_DOF + (c)]
    switch (kind) {
        case ORC_BASE_CONSTANT: /* :7-24 */
            B(1, 0) = 1.0;
            break;
        case ORC_BASE_LINEAR: /* :27-51 */
            B(1, 0) = 1.0;
            if (dof > 2) B(1, 1) = x;
            B(2, dof - 1) = x;
            break;
        case ORC_BASE_QUADRATIC: /* :54-78 */
            B(1, 0) = 1.0;
            if (dof > 2) B(1, 1) = x * x;
            B(2, dof - 1) = x * x;
            break;
        case ORC_BASE_PURE_HELIX: /* :81-103 */
            B(1, 0) = sin(x / 10);
            B(2, 1) = cos(x / 10);
            break;
        case ORC_BASE_LINEAR_HELIX: /* :106-128 */
            B(1, 0) = sin(x / 10);
            B(2, 1) = x * cos(x / 10);
            break;
        case ORC_BASE_TORSION_HELIX: /* :131-152 */
            B(0, 0) = 1.0;
            B(1, 1) = 1.0;
            B(2, 2) = 1.0;
            break;
        case ORC_BASE_TORSION_LINEAR_HELIX: /* :155-176 */
            B(0, 0) = 1.0;
            B(1, 1) = 1.0;
            B(2, 2) = x;
            break;
        case ORC_BASE_FULL: { /* :179-224 */
            int bs;
            if (dof % 2 == 1) {
                bs = 1;
                B(0, 0) = 1.0;
            } else {
                bs = 2;
                B(0, 0) = 1.0;
                B(0, 1) = x;
            }
            B(1, bs) = 1.0;
            B(1, bs + 1) = x;
            if (dof <= 6) {
                B(2, bs + 2) = 1.0;
                B(2, bs + 3) = x;
            } else {
                B(1, bs + 2) = x * x;
                B(2, bs + 3) = 1.0;
                B(2, bs + 4) = x;
                B(2, bs + 5) = x * x;
            }
            break;
        }
        default:
            return -1;
    }
#undef B
    return 0;
}

/* ksi = base(x) @ q + strain_bias, strain_bias = [0,0,0,1,0,0] (kinematic.py:49,107-110) */
void orc_ksi(const orc_model* m, int n, double x, double ksi[6]) {
    double base[6 * ORC_MAX_DOF];
    orc_strain_base(m->base_kind[n], x, m->q_dof[n], base);
    for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int c = 0; c < m->q_dof[n]; ++c) s += base[r * ORC_MAX_DOF + c] * m->q[n][c];
        ksi[r] = s;
    }
    ksi[3] += 1.0;
}

/* ------------------------------------------------------------------ */
/* model.py / kinematic.py                                              */
/* ------------------------------------------------------------------ */

/* model.py:37-38: round(length/delta_x)+1 with Python's round-half-even */
int orc_num_discrete_points(double length, double delta_x) { return (int)nearbyint(length / delta_x) + 1; }

static double sign_of(double x) { return (x > 0.0) - (x < 0.0); }

/* kinematic.py:201-274 */
void orc_calculate_indices(int n_tubes, const double* prev_ins, const double* next_ins, const double* max_len,
                           double dx, int* prev_idx, int* new_idx) {
    double length_sum = 0.0;
    for (int n = 0; n < n_tubes; ++n) {
        double min_length = length_sum; /* :224-228 */
        length_sum = max_len[n];
        double max_length = max_len[n];
        double p = prev_ins[n], q = next_ins[n];
        if (p < min_length) p = min_length;
        else if (p > max_length) p = max_length;
        if (q < min_length) q = min_length;
        else if (q > max_length) q = max_length;
        double delta = q - p;
        int pi;
        double di;
        if (fabs(delta) <= 0.5 * dx) { /* :255-265 */
            di = sign_of(delta);
            if (nearbyint(p / dx) == nearbyint(q / dx) &&
                (ceil(p / dx) == ceil(q / dx) || floor(p / dx) == floor(q / dx))) {
                if (di < 0) pi = (int)ceil(p / dx);
                else pi = (int)floor(p / dx);
            } else {
                pi = (int)nearbyint(p / dx);
            }
        } else { /* :266-268 */
            di = nearbyint(delta / dx);
            pi = (int)nearbyint(p / dx);
        }
        prev_idx[n] = pi;
        new_idx[n] = pi + (int)di;
    }
}

/* kinematic.py:82-130 */
int orc_kin_solve_g(const orc_model* m, const int* indices, const double* thetas, int full, double* g_out,
                    int* counts) {
    double g_prev[16];
    mat4_eye(g_prev);
    double* out = g_out;
    for (int n = 0; n < m->tube_num; ++n) {
        /* :95-97 theta rotation about local x through the same exponential_map */
        double xs[6] = {thetas[n], 0, 0, 0, 0, 0}, th[16], te[16];
        orc_hat(xs, th);
        orc_exponential_map(thetas[n], th, te);
        mat4_mul(g_prev, te, g_prev);
        double* tube0 = out;
        memcpy(out, g_prev, sizeof g_prev);
        out += 16;
        int cnt = 1;
        int start = full ? 1 : indices[n] + 1; /* :100-103 */
        for (int i = start; i < m->npts[n]; ++i) { /* :106-118 */
            double s = (i - 0.5) * m->delta_x;
            double ksi[6], kh[16], gh[16], gn[16];
            orc_ksi(m, n, s, ksi);
            double nk = sqrt(ksi[0] * ksi[0] + ksi[1] * ksi[1] + ksi[2] * ksi[2]);
            orc_hat(ksi, kh);
            for (int k = 0; k < 16; ++k) gh[k] = m->delta_x * kh[k];
            orc_exponential_map(m->delta_x * nk, gh, gn);
            mat4_mul(g_prev, gn, g_prev);
            memcpy(out, g_prev, sizeof g_prev);
            out += 16;
            ++cnt;
        }
        if (full && indices[n] != 0) { /* :120-126 re-anchor node idx to the previous tip */
            double inv[16], tr[16];
            se3_inv(tube0 + 16 * indices[n], inv);
            mat4_mul(tube0, inv, tr);
            for (int k = 0; k < cnt; ++k) mat4_mul(tr, tube0 + 16 * k, tube0 + 16 * k);
            memcpy(g_prev, tube0 + 16 * (cnt - 1), sizeof g_prev);
        }
        counts[n] = cnt;
    }
    return 0;
}

/* kinematic.py:132-198 */
int orc_kin_solve_eta(const orc_model* m, const double* velocity, const int* prev_idx, const double* delta_theta,
                      const double* prev_g, const int* prev_counts, double* eta_out, double* ftl_out,
                      int* ftl_counts) {
    double eta_prev[6] = {0, 0, 0, 0, 0, 0};
    double vel_sum = 0.0;
    const double* pg = prev_g;
    double* fo = ftl_out;
    for (int n = 0; n < m->tube_num; ++n) {
        vel_sum += velocity[n];                                     /* :163 */
        double mid = prev_idx[n] * m->delta_x - velocity[n] / 2;     /* :165 */
        double ksi[6];
        orc_ksi(m, n, mid, ksi);                                    /* :167-168 */
        double Ad[36], tr1[6], tr2[6], eta[6];
        orc_big_adjoint(pg, Ad);                                    /* prev_g[n][0] */
        double ex[6] = {delta_theta[n], 0, 0, 0, 0, 0};
        double kv[6];
        for (int k = 0; k < 6; ++k) kv[k] = velocity[n] * ksi[k];
        mat6_vec(Ad, ex, tr1);                                      /* :171 */
        mat6_vec(Ad, kv, tr2);                                      /* :172 */
        for (int k = 0; k < 6; ++k) eta[k] = eta_prev[k] + tr1[k] + tr2[k]; /* :174 */
        int cnt = 0;
        for (int i = prev_idx[n]; i < m->npts[n]; ++i) {            /* :178-193 */
            int k = i - prev_idx[n];
            if (k >= prev_counts[n]) return -2; /* reference would raise IndexError */
            double kh[6], gi[16], Adi[36], loc[6];
            orc_ksi(m, n, m->delta_x * i, kh);
            se3_inv(pg + 16 * k, gi);
            orc_big_adjoint(gi, Adi);
            mat6_vec(Adi, eta, loc);
            for (int c = 0; c < 6; ++c) fo[c] = loc[c] - vel_sum * kh[c];
            fo += 6;
            ++cnt;
        }
        ftl_counts[n] = cnt;
        memcpy(eta_prev, eta, sizeof eta);
        memcpy(eta_out + 6 * n, eta, sizeof eta);
        pg += 16 * prev_counts[n];
    }
    return 0;
}

/* kinematic.py:53-80 */
int orc_kin_solve_integrate(const orc_model* m, const double* delta_theta, const double* delta_ins,
                            const double* this_theta, const double* this_ins, const double* prev_g,
                            const int* prev_counts, int invert_insert, int need_g_out, double* g_out,
                            int* g_counts, double* eta_out, int* new_idx, double* true_ins, double* ftl_out,
                            int* ftl_counts) {
    int T = m->tube_num;
    double ti[ORC_MAX_TUBES], di[ORC_MAX_TUBES], pi[ORC_MAX_TUBES], vel[ORC_MAX_TUBES];
    int prev_idx[ORC_MAX_TUBES];
    for (int n = 0; n < T; ++n) {
        ti[n] = invert_insert ? m->length[n] - this_ins[n] : this_ins[n]; /* :57-59 */
        di[n] = invert_insert ? -delta_ins[n] : delta_ins[n];
        pi[n] = ti[n] - di[n]; /* :61 */
    }
    orc_calculate_indices(T, pi, ti, m->length, m->delta_x, prev_idx, new_idx); /* :63 */
    if (need_g_out) orc_kin_solve_g(m, new_idx, this_theta, 0, g_out, g_counts);
    else
        for (int n = 0; n < T; ++n) g_counts[n] = 0;
    for (int n = 0; n < T; ++n) vel[n] = (prev_idx[n] - new_idx[n]) * m->delta_x; /* :71 */
    int rc = orc_kin_solve_eta(m, vel, prev_idx, delta_theta, prev_g, prev_counts, eta_out, ftl_out, ftl_counts);
    for (int n = 0; n < T; ++n) /* :74-78 */
        true_ins[n] = invert_insert ? m->length[n] - new_idx[n] * m->delta_x : new_idx[n] * m->delta_x;
    return rc;
}

/* ------------------------------------------------------------------ */
/* geometry: exact distances                                            */
/* ------------------------------------------------------------------ */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void sub3(const double* a, const double* b, double* o) {
    o[0] = a[0] - b[0];
    o[1] = a[1] - b[1];
    o[2] = a[2] - b[2];
}
static void cross3(const double* a, const double* b, double* o) {
    double t0 = a[1] * b[2] - a[2] * b[1], t1 = a[2] * b[0] - a[0] * b[2], t2 = a[0] * b[1] - a[1] * b[0];
    o[0] = t0;
    o[1] = t1;
    o[2] = t2;
}

/* Distance from a point to the solid cylinder {c + t a + u : |t|<=h, u _|_ a, |u|<=r}
 * (fcl.Cylinder(radius, length) with h = length/2, collision_checker.py:133). 0 inside. */
double orc_point_cylinder(const double x[3], const double c[3], const double a[3], double h, double r) {
    double y[3];
    sub3(x, c, y);
    double t = dot3(y, a);
    double w[3] = {y[0] - t * a[0], y[1] - t * a[1], y[2] - t * a[2]};
    double rho = sqrt(dot3(w, w));
    double dt = fabs(t) - h, dr = rho - r;
    if (dt < 0) dt = 0;
    if (dr < 0) dr = 0;
    return sqrt(dt * dt + dr * dr);
}

/* F(s) = d_C(p0 + s e)^2 and its derivative; F is convex and C^1 in s. */
static double segcyl_F(double s, const double* y0, const double* e, const double* a, double h, double r,
                       double* dF) {
    double y[3] = {y0[0] + s * e[0], y0[1] + s * e[1], y0[2] + s * e[2]};
    double t = dot3(y, a), tp = dot3(e, a);
    double w[3] = {y[0] - t * a[0], y[1] - t * a[1], y[2] - t * a[2]};
    double wp[3] = {e[0] - tp * a[0], e[1] - tp * a[1], e[2] - tp * a[2]};
    double rho = sqrt(dot3(w, w));
    double dt = fabs(t) - h, dr = rho - r;
    double F = 0.0, g = 0.0;
    if (dt > 0) {
        F += dt * dt;
        g += 2.0 * dt * (t > 0 ? tp : -tp);
    }
    if (dr > 0) {
        F += dr * dr;
        g += 2.0 * dr * dot3(w, wp) / rho;
    }
    *dF = g;
    return F;
}

/* min over s in [0,1] of the distance from p0 + s (p1-p0) to the solid cylinder.
 * d_C is convex, so F' is monotone: bracket its root (Illinois + bisection safeguard). */
double orc_segment_cylinder(const double p0[3], const double p1[3], const double c[3], const double a[3],
                            double h, double r) {
    double y0[3], e[3];
    sub3(p0, c, y0);
    sub3(p1, p0, e);
    double g0, g1;
    double F0 = segcyl_F(0.0, y0, e, a, h, r, &g0);
    double F1 = segcyl_F(1.0, y0, e, a, h, r, &g1);
    if (F0 == 0.0 || F1 == 0.0) return 0.0;
    if (g0 >= 0.0) return sqrt(F0);
    if (g1 <= 0.0) return sqrt(F1);
    double lo = 0.0, hi = 1.0, glo = g0, ghi = g1, best = F0 < F1 ? F0 : F1;
    int side = 0;
    for (int it = 0; it < 200; ++it) {
        double s;
        if (it & 1) s = 0.5 * (lo + hi);
        else s = (lo * ghi - hi * glo) / (ghi - glo);
        if (!(s > lo && s < hi)) s = 0.5 * (lo + hi);
        if (!(s > lo && s < hi)) break;
        double g, F = segcyl_F(s, y0, e, a, h, r, &g);
        if (F < best) best = F;
        if (F == 0.0 || g == 0.0) return sqrt(F);
        if (g > 0) {
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
        if (hi - lo < 1e-15) break;
    }
    return sqrt(best);
}

static int point_in_tri_plane(const double* z, const double* t0, const double* t1, const double* t2,
                              const double* n) {
    const double* v[3] = {t0, t1, t2};
    for (int i = 0; i < 3; ++i) {
        double ed[3], pz[3], cr[3];
        sub3(v[(i + 1) % 3], v[i], ed);
        sub3(z, v[i], pz);
        cross3(ed, pz, cr);
        if (dot3(cr, n) < 0.0) return 0;
    }
    return 1;
}

/* Exact Euclidean distance between a solid flat-capped cylinder and a triangle
 * (a surface patch).  0 when they touch or intersect.  The minimum of the
 * convex function d_C over the triangle is attained either on an edge
 * (orc_segment_cylinder) or at an interior point whose nearest cylinder point
 * is the support point of the cylinder toward the triangle's plane. */
double orc_cylinder_triangle(const double c[3], const double a[3], double h, double r, const double t0[3],
                             const double t1[3], const double t2[3]) {
    double e = orc_segment_cylinder(t0, t1, c, a, h, r);
    if (e == 0.0) return 0.0;
    double e2 = orc_segment_cylinder(t1, t2, c, a, h, r);
    if (e2 == 0.0) return 0.0;
    if (e2 < e) e = e2;
    e2 = orc_segment_cylinder(t2, t0, c, a, h, r);
    if (e2 == 0.0) return 0.0;
    if (e2 < e) e = e2;

    double u[3], v[3], n[3];
    sub3(t1, t0, u);
    sub3(t2, t0, v);
    cross3(u, v, n);
    double nn = sqrt(dot3(n, n));
    if (nn == 0.0) return e; /* degenerate triangle: its edges are the whole set */
    n[0] /= nn;
    n[1] /= nn;
    n[2] /= nn;
    double ct[3];
    sub3(c, t0, ct);
    double sc = dot3(ct, n); /* signed plane offset of the cylinder centre */
    double an = dot3(a, n);
    double m[3] = {n[0] - an * a[0], n[1] - an * a[1], n[2] - an * a[2]}; /* radial part of n */
    double mm = sqrt(dot3(m, m));
    if (mm > 0.0) {
        m[0] /= mm;
        m[1] /= mm;
        m[2] /= mm;
    }
    /* support offset of the cylinder in direction +n:  sgn(a.n) h a + r m */
    double sg = (an > 0) - (an < 0);
    double so[3] = {sg * h * a[0] + r * m[0], sg * h * a[1] + r * m[1], sg * h * a[2] + r * m[2]};
    double E = dot3(so, n); /* half extent of the cylinder along n (>= 0) */
    if (fabs(sc) <= E) {
        /* the plane cuts the cylinder: z = a point of (cylinder /\ plane) on the chord through c */
        double lam = (E > 0.0) ? -sc / E : 0.0;
        double z[3] = {c[0] + lam * so[0], c[1] + lam * so[1], c[2] + lam * so[2]};
        if (point_in_tri_plane(z, t0, t1, t2, n)) return 0.0;
        return e;
    }
    double sgc = sc > 0 ? 1.0 : -1.0;
    double y[3] = {c[0] - sgc * so[0], c[1] - sgc * so[1], c[2] - sgc * so[2]}; /* nearest point to the plane */
    double yt[3];
    sub3(y, t0, yt);
    double dpl = dot3(yt, n);
    double yp[3] = {y[0] - dpl * n[0], y[1] - dpl * n[1], y[2] - dpl * n[2]};
    if (point_in_tri_plane(yp, t0, t1, t2, n)) {
        double d = fabs(sc) - E;
        return d < e ? d : e;
    }
    return e;
}

/* exact point-triangle distance (closest-point region walk, Ericson RTCD 5.1.5) */
double orc_point_triangle(const double p[3], const double a[3], const double b[3], const double c[3]) {
    double ab[3], ac[3], ap[3], bp[3], cp[3], q[3];
    sub3(b, a, ab);
    sub3(c, a, ac);
    sub3(p, a, ap);
    double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0 && d2 <= 0) return sqrt(dot3(ap, ap));
    sub3(p, b, bp);
    double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0 && d4 <= d3) return sqrt(dot3(bp, bp));
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        double v = d1 / (d1 - d3);
        for (int i = 0; i < 3; ++i) q[i] = ap[i] - v * ab[i];
        return sqrt(dot3(q, q));
    }
    sub3(p, c, cp);
    double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0 && d5 <= d6) return sqrt(dot3(cp, cp));
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        double w = d2 / (d2 - d6);
        for (int i = 0; i < 3; ++i) q[i] = ap[i] - w * ac[i];
        return sqrt(dot3(q, q));
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int i = 0; i < 3; ++i) q[i] = bp[i] - w * (c[i] - b[i]);
        return sqrt(dot3(q, q));
    }
    double n[3];
    cross3(ab, ac, n);
    double nn = dot3(n, n);
    if (nn == 0.0) return sqrt(dot3(ap, ap));
    return fabs(dot3(ap, n)) / sqrt(nn);
}

/* distance from a point to a solid box (0 inside); box axes = rows of rot^T, dims = full lengths */
double orc_point_box(const double x[3], const orc_prim* b) {
    double y[3];
    sub3(x, b->center, y);
    double s = 0.0;
    for (int k = 0; k < 3; ++k) {
        /* local coordinate along box axis k = column k of rot */
        double l = y[0] * b->rot[0 * 3 + k] + y[1] * b->rot[1 * 3 + k] + y[2] * b->rot[2 * 3 + k];
        double d = fabs(l) - 0.5 * b->dims[k];
        if (d > 0) s += d * d;
    }
    return sqrt(s);
}

static void prim_axis(const orc_prim* p, double a[3]) { /* local z column of rot */
    a[0] = p->rot[2];
    a[1] = p->rot[5];
    a[2] = p->rot[8];
}

static void box_corner(const orc_prim* b, int k, double out[3]) {
    double s[3] = {(k & 1) ? 0.5 : -0.5, (k & 2) ? 0.5 : -0.5, (k & 4) ? 0.5 : -0.5};
    for (int i = 0; i < 3; ++i) {
        out[i] = b->center[i];
        for (int ax = 0; ax < 3; ++ax) out[i] += b->rot[i * 3 + ax] * s[ax] * b->dims[ax];
    }
}

/* cylinder vs solid box: min over the 12 surface triangles, 0 if the centre is inside */
static double cylinder_box(const double c[3], const double a[3], double h, double r, const orc_prim* b) {
    if (orc_point_box(c, b) == 0.0) return 0.0;
    static const int F[12][3] = {{0, 1, 3}, {0, 3, 2}, {4, 7, 5}, {4, 6, 7}, {0, 5, 1}, {0, 4, 5},
                                 {2, 3, 7}, {2, 7, 6}, {0, 2, 6}, {0, 6, 4}, {1, 5, 7}, {1, 7, 3}};
    double corner[8][3];
    for (int k = 0; k < 8; ++k) box_corner(b, k, corner[k]);
    double best = DBL_MAX;
    for (int f = 0; f < 12; ++f) {
        double d = orc_cylinder_triangle(c, a, h, r, corner[F[f][0]], corner[F[f][1]], corner[F[f][2]]);
        if (d < best) best = d;
        if (best == 0.0) break;
    }
    return best;
}

/* ------------------------------------------------------------------ GJK distance between two solid cylinders
 * (obstacle fcl.Cylinder vs tube fcl.Cylinder, init_collision.py:131-135).  FCL runs libccd's GJK with a 1e-6
 * tolerance here; this restatement iterates the same algorithm to ~1e-13.  The Minkowski difference of the two
 * cylinders is searched through its support map; the closest point of the current simplex (<= 4 points) to the
 * origin is found by the usual region tests. */
typedef struct { double c[3], a[3], h, r; } gjk_cyl;

static void cyl_support(const gjk_cyl* C, const double d[3], double out[3]) {
    double da = dot3(d, C->a);
    double w[3] = {d[0] - da * C->a[0], d[1] - da * C->a[1], d[2] - da * C->a[2]};
    double wn = sqrt(dot3(w, w));
    double sg = da >= 0 ? 1.0 : -1.0;
    for (int k = 0; k < 3; ++k) out[k] = C->c[k] + sg * C->h * C->a[k] + (wn > 0 ? C->r * w[k] / wn : 0.0);
}

/* closest point to the origin on the simplex W[0..n-1]; reduces W to the supporting face; returns n' */
static int simplex_closest(double W[4][3], int n, double v[3]) {
    if (n == 1) {
        memcpy(v, W[0], 3 * sizeof(double));
        return 1;
    }
    if (n == 2) {
        double ab[3];
        sub3(W[1], W[0], ab);
        double t = -dot3(W[0], ab), den = dot3(ab, ab);
        if (t <= 0 || den == 0) {
            memcpy(v, W[0], 3 * sizeof(double));
            return 1;
        }
        if (t >= den) {
            memcpy(W[0], W[1], 3 * sizeof(double));
            memcpy(v, W[0], 3 * sizeof(double));
            return 1;
        }
        t /= den;
        for (int k = 0; k < 3; ++k) v[k] = W[0][k] + t * ab[k];
        return 2;
    }
    if (n == 3) {
        /* closest point of triangle to the origin (Ericson 5.1.5 with p = 0), keeping the supporting vertices */
        double *a = W[0], *b = W[1], *c = W[2];
        double ab[3], ac[3], ap[3] = {-a[0], -a[1], -a[2]};
        sub3(b, a, ab);
        sub3(c, a, ac);
        double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
        if (d1 <= 0 && d2 <= 0) { memcpy(v, a, 24); return 1; }
        double bp[3] = {-b[0], -b[1], -b[2]};
        double d3v = dot3(ab, bp), d4 = dot3(ac, bp);
        if (d3v >= 0 && d4 <= d3v) { memcpy(W[0], b, 24); memcpy(v, W[0], 24); return 1; }
        double vc = d1 * d4 - d3v * d2;
        if (vc <= 0 && d1 >= 0 && d3v <= 0) {
            double t = d1 / (d1 - d3v);
            for (int k = 0; k < 3; ++k) v[k] = a[k] + t * ab[k];
            return 2; /* W[0]=a, W[1]=b */
        }
        double cp[3] = {-c[0], -c[1], -c[2]};
        double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
        if (d6 >= 0 && d5 <= d6) { memcpy(W[0], c, 24); memcpy(v, W[0], 24); return 1; }
        double vb = d5 * d2 - d1 * d6;
        if (vb <= 0 && d2 >= 0 && d6 <= 0) {
            double t = d2 / (d2 - d6);
            for (int k = 0; k < 3; ++k) v[k] = a[k] + t * ac[k];
            memcpy(W[1], c, 24);
            return 2;
        }
        double va = d3v * d6 - d5 * d4;
        if (va <= 0 && (d4 - d3v) >= 0 && (d5 - d6) >= 0) {
            double t = (d4 - d3v) / ((d4 - d3v) + (d5 - d6));
            for (int k = 0; k < 3; ++k) v[k] = b[k] + t * (c[k] - b[k]);
            memcpy(W[0], b, 24);
            memcpy(W[1], c, 24);
            return 2;
        }
        double den = 1.0 / (va + vb + vc), s = vb * den, t = vc * den;
        for (int k = 0; k < 3; ++k) v[k] = a[k] + s * ab[k] + t * ac[k];
        return 3;
    }
    /* tetrahedron: the origin is inside iff the four sub-volumes (origin in place of a vertex) all have the
     * sign of the whole volume; otherwise the closest point lies on one of the four faces */
    {
        static const int F[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
        double e1[3], e2[3], e3[3], cr[3];
        sub3(W[1], W[0], e1);
        sub3(W[2], W[0], e2);
        sub3(W[3], W[0], e3);
        cross3(e1, e2, cr);
        double vol = dot3(cr, e3);
        double scale = sqrt(dot3(e1, e1) * dot3(e2, e2) * dot3(e3, e3));
        if (fabs(vol) > 1e-12 * scale) {
            int inside = 1;
            for (int f = 0; f < 4 && inside; ++f) {
                /* volume of the tetrahedron with vertex f replaced by the origin, same vertex order */
                double P[4][3];
                memcpy(P, W, sizeof P);
                P[f][0] = P[f][1] = P[f][2] = 0.0;
                double a1[3], a2[3], a3[3], c2[3];
                sub3(P[1], P[0], a1);
                sub3(P[2], P[0], a2);
                sub3(P[3], P[0], a3);
                cross3(a1, a2, c2);
                if (dot3(c2, a3) * vol < 0) inside = 0;
            }
            if (inside) {
                v[0] = v[1] = v[2] = 0.0;
                return 4;
            }
        }
        double best = DBL_MAX, bv[3] = {0, 0, 0}, BW[4][3];
        int bn = 0;
        memset(BW, 0, sizeof BW);
        for (int f = 0; f < 4; ++f) {
            double T3[4][3], tv[3];
            memcpy(T3[0], W[F[f][0]], 24);
            memcpy(T3[1], W[F[f][1]], 24);
            memcpy(T3[2], W[F[f][2]], 24);
            int tn = simplex_closest(T3, 3, tv);
            double dd = dot3(tv, tv);
            if (dd < best) {
                best = dd;
                memcpy(bv, tv, 24);
                memcpy(BW, T3, sizeof BW);
                bn = tn;
            }
        }
        memcpy(W, BW, sizeof BW);
        memcpy(v, bv, 24);
        return bn;
    }
}

static int g_gjk_iters = 0, g_gjk_exit = 0, g_gjk_cap = 256;
void orc_gjk_set_cap(int c) { g_gjk_cap = c; }
int orc_gjk_last_iters(void) { return g_gjk_iters; }
int orc_gjk_last_exit(void) { return g_gjk_exit; }

static double gjk_cylinder_cylinder(const double c1[3], const double a1[3], double h1, double r1, const double c2[3],
                                   const double a2[3], double h2, double r2) {
    gjk_cyl A, B;
    memcpy(A.c, c1, 24); memcpy(A.a, a1, 24); A.h = h1; A.r = r1;
    memcpy(B.c, c2, 24); memcpy(B.a, a2, 24); B.h = h2; B.r = r2;
    double W[4][3], v[3];
    sub3(A.c, B.c, v);
    if (dot3(v, v) == 0) return 0.0;
    int n = 0, stall = 0;
    double best_vv = dot3(v, v);
    for (int it = 0; it < g_gjk_cap; ++it) {
        double dneg[3] = {-v[0], -v[1], -v[2]}, sa[3], sb[3], w[3];
        cyl_support(&A, dneg, sa);
        cyl_support(&B, v, sb);
        sub3(sa, sb, w);
        double vv = dot3(v, v), vw = dot3(v, w);
        g_gjk_iters = it;
        /* |v| decreases monotonically in exact arithmetic; keep the best value and stop once rounding stalls it */
        if (vv < best_vv * (1 - 1e-15)) stall = 0;
        else ++stall;
        if (vv < best_vv) best_vv = vv;
        if (vv - vw <= 1e-14 * vv) { g_gjk_exit = 1; return sqrt(best_vv); } /* v is the closest point of A - B */
        if (stall >= 12) { g_gjk_exit = 2; return sqrt(best_vv); }
        int dup = 0;
        for (int k = 0; k < n; ++k) {
            double e[3];
            sub3(w, W[k], e);
            if (dot3(e, e) <= 1e-30 * (1 + vv)) dup = 1;
        }
        if (dup) { g_gjk_exit = 2; return sqrt(best_vv); }
        memcpy(W[n++], w, 24);
        n = simplex_closest(W, n, v);
        if (n == 4) return 0.0;
        if (dot3(v, v) <= 1e-28) return 0.0;
    }
    g_gjk_exit = 3;
    {
        double vv = dot3(v, v);
        if (vv < best_vv) best_vv = vv;
    }
    return sqrt(best_vv);
}


/* minimum over one rim circle of cylinder 1 (centre c1 + sg h1 a1, radius r1) of the distance to the solid
 * cylinder 2: 64 samples, then a golden-section refinement of every local-minimum bracket that can still
 * undercut the best sample (the point-cylinder distance is 1-Lipschitz). */
static double rim_point_dist(double phi, const double o[3], const double e1[3], const double e2[3], double r1,
                             const double c2[3], const double a2[3], double h2, double r2) {
    double cs = cos(phi), sn = sin(phi), x[3];
    for (int k = 0; k < 3; ++k) x[k] = o[k] + r1 * (cs * e1[k] + sn * e2[k]);
    return orc_point_cylinder(x, c2, a2, h2, r2);
}

static double rim_min(const double c1[3], const double a1[3], double h1, double r1, double sg, const double c2[3],
                      const double a2[3], double h2, double r2) {
    enum { NS = 64 };
    const double step = 6.283185307179586 / NS, gr = 0.6180339887498949;
    double t[3] = {0, 0, 0}, e1[3], e2[3], o[3], f[NS];
    t[fabs(a1[0]) < 0.6 ? 0 : 1] = 1.0;
    cross3(a1, t, e1);
    double en = sqrt(dot3(e1, e1));
    for (int k = 0; k < 3; ++k) e1[k] /= en;
    cross3(a1, e1, e2);
    for (int k = 0; k < 3; ++k) o[k] = c1[k] + sg * h1 * a1[k];
    double best = DBL_MAX;
    for (int i = 0; i < NS; ++i) {
        f[i] = rim_point_dist(i * step, o, e1, e2, r1, c2, a2, h2, r2);
        if (f[i] == 0.0) return 0.0;
        if (f[i] < best) best = f[i];
    }
    double res = best;
    for (int i = 0; i < NS; ++i) {
        double fp = f[(i + NS - 1) % NS], fn = f[(i + 1) % NS];
        if (f[i] > fp || f[i] > fn || f[i] - r1 * step > res) continue;
        double lo = (i - 1) * step, hi = (i + 1) * step;
        double x1 = hi - gr * (hi - lo), x2 = lo + gr * (hi - lo);
        double f1 = rim_point_dist(x1, o, e1, e2, r1, c2, a2, h2, r2);
        double f2 = rim_point_dist(x2, o, e1, e2, r1, c2, a2, h2, r2);
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
        double m = f1 < f2 ? f1 : f2;
        if (m < res) res = m;
    }
    return res;
}

/* Exact distance between two solid flat-capped cylinders (obstacle fcl.Cylinder vs tube cylinder).  GJK on two
 * curved bodies converges only like 1/k^2 (1e-6 after 256 iterations), so it serves as the intersection test
 * alone.  For disjoint cylinders the closest pair (x, y) has a common normal n; x is the support point of
 * cylinder 1 along n, which is a rim point unless n is perpendicular or parallel to the axis -- and in those
 * cases the support set (a generator or a cap disk) still meets a rim of one of the two bodies, except when two
 * generators are closest at interior points; that case is the axis-segment distance minus both radii.  Hence
 *   d = |axis1 - axis2| - r1 - r2   if the closest points of the two axis lines are interior to both segments,
 *   d = min over the four rims of min_phi dist(rim(phi), other cylinder)   otherwise. */
double orc_cylinder_cylinder(const double c1[3], const double a1[3], double h1, double r1, const double c2[3],
                             const double a2[3], double h2, double r2) {
    if (gjk_cylinder_cylinder(c1, a1, h1, r1, c2, a2, h2, r2) == 0.0) return 0.0;
    double w[3];
    sub3(c1, c2, w);
    const double b = dot3(a1, a2), d = dot3(a1, w), e = dot3(a2, w), den = 1.0 - b * b;
    double s = 0.0, t = 0.0;
    if (den > 1e-12) {
        s = (b * e - d) / den;
        t = (e - b * d) / den;
        if (fabs(s) < h1 && fabs(t) < h2) {
            double y[3] = {w[0] + s * a1[0] - t * a2[0], w[1] + s * a1[1] - t * a2[1], w[2] + s * a1[2] - t * a2[2]};
            double dd = sqrt(dot3(y, y)) - r1 - r2;
            return dd > 0 ? dd : 0.0;
        }
    }
    /* the axis point of each cylinder nearest the other axis line, clamped: inside the other body => contact */
    s = s > h1 ? h1 : (s < -h1 ? -h1 : s);
    t = t > h2 ? h2 : (t < -h2 ? -h2 : t);
    double p1[3] = {c1[0] + s * a1[0], c1[1] + s * a1[1], c1[2] + s * a1[2]};
    double p2[3] = {c2[0] + t * a2[0], c2[1] + t * a2[1], c2[2] + t * a2[2]};
    if (orc_point_cylinder(p1, c2, a2, h2, r2) == 0.0 || orc_point_cylinder(p2, c1, a1, h1, r1) == 0.0) return 0.0;
    double best = rim_min(c1, a1, h1, r1, 1.0, c2, a2, h2, r2);
    double m = rim_min(c1, a1, h1, r1, -1.0, c2, a2, h2, r2);
    if (m < best) best = m;
    m = rim_min(c2, a2, h2, r2, 1.0, c1, a1, h1, r1);
    if (m < best) best = m;
    m = rim_min(c2, a2, h2, r2, -1.0, c1, a1, h1, r1);
    if (m < best) best = m;
    return best;
}

/* ------------------------------------------------------------------ */
/* environment with a plain median-split AABB tree (CPU baseline speed) */
/* ------------------------------------------------------------------ */
typedef struct {
    double lo[3], hi[3];
    int left, right; /* children, or -1 */
    int first, count;
} orc_node;

typedef struct {
    int ntri;
    double* tris; /* reordered copy [ntri][9] */
    orc_node* nodes;
    int nnodes;
} orc_bvh;

typedef struct orc_env {
    orc_bvh obs, goal_mesh;
    int nprim;
    orc_prim* prims;
    orc_prim goal;
} orc_env;

typedef struct {
    double key;
    int id;
} orc_key;

static int key_cmp(const void* a, const void* b) {
    double x = ((const orc_key*)a)->key, y = ((const orc_key*)b)->key;
    if (x < y) return -1;
    if (x > y) return 1;
    return ((const orc_key*)a)->id - ((const orc_key*)b)->id;
}

static int build_rec(orc_bvh* b, int first, int count, double* cent, int* order) {
    int id = b->nnodes++;
    orc_node* nd = &b->nodes[id];
    for (int k = 0; k < 3; ++k) {
        nd->lo[k] = DBL_MAX;
        nd->hi[k] = -DBL_MAX;
    }
    for (int i = first; i < first + count; ++i)
        for (int v = 0; v < 3; ++v)
            for (int k = 0; k < 3; ++k) {
                double x = b->tris[(size_t)order[i] * 9 + v * 3 + k];
                if (x < nd->lo[k]) nd->lo[k] = x;
                if (x > nd->hi[k]) nd->hi[k] = x;
            }
    nd->first = first;
    nd->count = count;
    nd->left = nd->right = -1;
    if (count <= 4) return id;
    int ax = 0;
    double ext[3] = {nd->hi[0] - nd->lo[0], nd->hi[1] - nd->lo[1], nd->hi[2] - nd->lo[2]};
    if (ext[1] > ext[ax]) ax = 1;
    if (ext[2] > ext[ax]) ax = 2;
    /* median split by centroid along the longest axis */
    orc_key* keys = (orc_key*)malloc((size_t)count * sizeof(orc_key));
    for (int i = 0; i < count; ++i) {
        keys[i].id = order[first + i];
        keys[i].key = cent[(size_t)order[first + i] * 3 + ax];
    }
    qsort(keys, (size_t)count, sizeof(orc_key), key_cmp);
    for (int i = 0; i < count; ++i) order[first + i] = keys[i].id;
    free(keys);
    int l = build_rec(b, first, count / 2, cent, order);
    int r = build_rec(b, first + count / 2, count - count / 2, cent, order);
    b->nodes[id].left = l;
    b->nodes[id].right = r;
    return id;
}

static void bvh_build(orc_bvh* b, const double* tris, int ntri) {
    memset(b, 0, sizeof *b);
    b->ntri = ntri;
    if (ntri <= 0) return;
    b->tris = (double*)malloc((size_t)ntri * 9 * sizeof(double));
    memcpy(b->tris, tris, (size_t)ntri * 9 * sizeof(double));
    b->nodes = (orc_node*)malloc((size_t)(2 * ntri + 1) * sizeof(orc_node));
    double* cent = (double*)malloc((size_t)ntri * 3 * sizeof(double));
    int* order = (int*)malloc((size_t)ntri * sizeof(int));
    for (int i = 0; i < ntri; ++i) {
        order[i] = i;
        for (int k = 0; k < 3; ++k)
            cent[(size_t)i * 3 + k] = (tris[(size_t)i * 9 + k] + tris[(size_t)i * 9 + 3 + k] + tris[(size_t)i * 9 + 6 + k]) / 3.0;
    }
    build_rec(b, 0, ntri, cent, order);
    double* re = (double*)malloc((size_t)ntri * 9 * sizeof(double));
    for (int i = 0; i < ntri; ++i) memcpy(re + (size_t)i * 9, b->tris + (size_t)order[i] * 9, 9 * sizeof(double));
    free(b->tris);
    b->tris = re;
    free(cent);
    free(order);
}

static void bvh_free(orc_bvh* b) {
    free(b->tris);
    free(b->nodes);
}

static double aabb_point_dist(const orc_node* nd, const double* x) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) {
        double d = 0.0;
        if (x[k] < nd->lo[k]) d = nd->lo[k] - x[k];
        else if (x[k] > nd->hi[k]) d = x[k] - nd->hi[k];
        s += d * d;
    }
    return sqrt(s);
}

/* branch and bound: min over triangles of cylinder_triangle, pruned by
 * (point-AABB distance - bounding radius) >= best.  Exactness does not depend on the tree. */
static void bvh_cyl_min(const orc_bvh* b, int id, const double* c, const double* a, double h, double r,
                        double rb, double* best) {
    const orc_node* nd = &b->nodes[id];
    if (nd->left < 0) {
        for (int i = nd->first; i < nd->first + nd->count; ++i) {
            const double* t = b->tris + (size_t)i * 9;
            double d = orc_cylinder_triangle(c, a, h, r, t, t + 3, t + 6);
            if (d < *best) *best = d;
            if (*best == 0.0) return;
        }
        return;
    }
    double dl = aabb_point_dist(&b->nodes[nd->left], c) - rb;
    double dr = aabb_point_dist(&b->nodes[nd->right], c) - rb;
    int first = nd->left, second = nd->right;
    if (dr < dl) {
        double t = dl;
        dl = dr;
        dr = t;
        first = nd->right;
        second = nd->left;
    }
    if (dl < *best) bvh_cyl_min(b, first, c, a, h, r, rb, best);
    if (*best == 0.0) return;
    if (dr < *best) bvh_cyl_min(b, second, c, a, h, r, rb, best);
}

static void bvh_point_min(const orc_bvh* b, int id, const double* x, double* best) {
    const orc_node* nd = &b->nodes[id];
    if (nd->left < 0) {
        for (int i = nd->first; i < nd->first + nd->count; ++i) {
            const double* t = b->tris + (size_t)i * 9;
            double d = orc_point_triangle(x, t, t + 3, t + 6);
            if (d < *best) *best = d;
        }
        return;
    }
    double dl = aabb_point_dist(&b->nodes[nd->left], x);
    double dr = aabb_point_dist(&b->nodes[nd->right], x);
    int first = nd->left, second = nd->right;
    if (dr < dl) {
        double t = dl;
        dl = dr;
        dr = t;
        first = nd->right;
        second = nd->left;
    }
    if (dl < *best) bvh_point_min(b, first, x, best);
    if (dr < *best) bvh_point_min(b, second, x, best);
}

orc_env* orc_env_create(const double* tris, int ntri, const orc_prim* prims, int nprim, const orc_prim* goal,
                        const double* goal_tris, int goal_ntri) {
    orc_env* e = (orc_env*)calloc(1, sizeof *e);
    bvh_build(&e->obs, tris, ntri);
    bvh_build(&e->goal_mesh, goal_tris, goal_ntri);
    e->nprim = nprim;
    if (nprim > 0) {
        e->prims = (orc_prim*)malloc((size_t)nprim * sizeof(orc_prim));
        memcpy(e->prims, prims, (size_t)nprim * sizeof(orc_prim));
    }
    if (goal) e->goal = *goal;
    else e->goal.shape = ORC_SHAPE_NONE;
    return e;
}

void orc_env_destroy(orc_env* e) {
    if (!e) return;
    bvh_free(&e->obs);
    bvh_free(&e->goal_mesh);
    free(e->prims);
    free(e);
}

/* collision_checker.py:35-70 with python-fcl semantics:
 * obstacle_min = min over (cylinder, obstacle) pairs of the exact distance,
 * exactly -1.0 if any pair touches; DBL_MAX with no obstacles;
 * goal_dist = distance(Sphere(rad[-1]) at the tip, goal), -1.0 if touching.
 * brute != 0 ignores the tree (used to prove tree-independence in tests). */
int orc_env_check_collision(const orc_env* env, const double* pos, const int* counts, int ntubes,
                            const double* rad, int brute, double out[2]) {
    double best = DBL_MAX;
    const double* p = pos;
    const double* tip = NULL;
    for (int n = 0; n < ntubes && best > 0.0; ++n) {
        for (int i = 0; i + 1 < counts[n] && best > 0.0; ++i) { /* collision_checker.py:121-147 */
            const double* f = p + 3 * i;
            const double* t = p + 3 * (i + 1);
            double vec[3], c[3], a[3];
            sub3(t, f, vec);
            double len = sqrt(dot3(vec, vec));
            for (int k = 0; k < 3; ++k) {
                c[k] = f[k] + vec[k] / 2;
                a[k] = vec[k] / len;
            }
            double h = 0.5 * len, r = rad[n], rb = sqrt(h * h + r * r);
            if (env->obs.ntri > 0) {
                if (brute) {
                    for (int k = 0; k < env->obs.ntri && best > 0.0; ++k) {
                        const double* tr = env->obs.tris + (size_t)k * 9;
                        double d = orc_cylinder_triangle(c, a, h, r, tr, tr + 3, tr + 6);
                        if (d < best) best = d;
                    }
                } else {
                    bvh_cyl_min(&env->obs, 0, c, a, h, r, rb, &best);
                }
            }
            for (int k = 0; k < env->nprim && best > 0.0; ++k) {
                const orc_prim* o = &env->prims[k];
                double d = DBL_MAX;
                if (o->shape == ORC_SHAPE_SPHERE) {
                    d = orc_point_cylinder(o->center, c, a, h, r) - o->dims[0];
                    if (d < 0) d = 0;
                } else if (o->shape == ORC_SHAPE_BOX) {
                    d = cylinder_box(c, a, h, r, o);
                } else if (o->shape == ORC_SHAPE_CYLINDER) {
                    double oa[3];
                    prim_axis(o, oa);
                    d = orc_cylinder_cylinder(c, a, h, r, o->center, oa, 0.5 * o->dims[1], o->dims[0]);
                } else {
                    return -3;
                }
                if (d < best) best = d;
            }
        }
        p += 3 * counts[n];
    }
    out[0] = (best <= 0.0) ? -1.0 : best;
    { /* curve[-1][-1]: last node of the last tube, whether or not the obstacle loop ended early */
        int total = 0;
        for (int n = 0; n < ntubes; ++n) total += counts[n];
        if (total > 0 && counts[ntubes - 1] > 0) tip = pos + 3 * (size_t)(total - 1);
    }

    /* collision_checker.py:62-68 */
    double gd = DBL_MAX;
    double tr = rad[ntubes - 1];
    const orc_prim* g = &env->goal;
    if (tip) {
        if (g->shape == ORC_SHAPE_SPHERE) {
            double d[3];
            sub3(tip, g->center, d);
            gd = sqrt(dot3(d, d)) - g->dims[0] - tr;
        } else if (g->shape == ORC_SHAPE_BOX) {
            gd = orc_point_box(tip, g) - tr;
        } else if (g->shape == ORC_SHAPE_CYLINDER) {
            double a[3];
            prim_axis(g, a);
            gd = orc_point_cylinder(tip, g->center, a, 0.5 * g->dims[1], g->dims[0]) - tr;
        } else if (g->shape == ORC_SHAPE_MESH && env->goal_mesh.ntri > 0) {
            double b = DBL_MAX;
            if (brute) {
                for (int k = 0; k < env->goal_mesh.ntri; ++k) {
                    const double* t = env->goal_mesh.tris + (size_t)k * 9;
                    double d = orc_point_triangle(tip, t, t + 3, t + 6);
                    if (d < b) b = d;
                }
            } else {
                bvh_point_min(&env->goal_mesh, 0, tip, &b);
            }
            gd = b - tr;
        }
        if (gd <= 0.0) gd = -1.0;
    }
    out[1] = gd;
    return 0;
}

int orc_check_collision(const double* pos, const int* counts, int ntubes, const double* rad, const double* tris,
                        int ntri, const orc_prim* prims, int nprim, const orc_prim* goal,
                        const double* goal_tris, int goal_ntri, double out[2]) {
    orc_env* e = orc_env_create(tris, ntri, prims, nprim, goal, goal_tris, goal_ntri);
    int rc = orc_env_check_collision(e, pos, counts, ntubes, rad, 1, out);
    orc_env_destroy(e);
    return rc;
}

/* ------------------------------------------------------------------ */
/* cost                                                                 */
/* ------------------------------------------------------------------ */

/* obstacles_and_goal.py:39-45,87-88 own cost (generation 0): w*goal + (1/obs)^2;
 * goal_dist < 0 is clamped to 0 by the caller at rrt.py:120-123. */
double orc_soapwg_own_cost(double goal_weight, double obs_min, double goal_dist) {
    if (goal_dist < 0) goal_dist = 0;
    double inv = 1.0 / obs_min;
    return goal_dist * goal_weight + inv * inv;
}

/* follow_the_leader.py:89-116 (translation_only: :173-189) */
double orc_ftl_magnitude(const double f[6], int translation_only) {
    double on = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    double vn = sqrt(f[3] * f[3] + f[4] * f[4] + f[5] * f[5]);
    if (translation_only) return vn;
    return on == 0.0 ? vn : on;
}

/* ------------------------------------------------------------------ */
/* batch drivers (bench cpu_baseline; parity at size)                   */
/* ------------------------------------------------------------------ */
static int total_nodes(const orc_model* m) {
    int s = 0;
    for (int n = 0; n < m->tube_num; ++n) s += m->npts[n];
    return s;
}

/* pthread parallel-for (this image's gcc has no libgomp spec, so no -fopenmp) */
typedef struct {
    const orc_model* m;
    const orc_env* env;
    int B;
    const int* idx;
    const double* theta;
    const double* rad;
    double goal_weight;
    double *obs_min, *goal_dist, *cost, *tip_out;
    int next; /* atomic work counter */
    int chunk;
    int rc;
} orc_job;

static void* worker_solve_g(void* arg) {
    orc_job* j = (orc_job*)arg;
    int T = j->m->tube_num, tot = total_nodes(j->m);
    double* g = (double*)malloc((size_t)tot * 16 * sizeof(double));
    int counts[ORC_MAX_TUBES];
    for (;;) {
        int b0 = __atomic_fetch_add(&j->next, j->chunk, __ATOMIC_RELAXED);
        if (b0 >= j->B) break;
        int b1 = b0 + j->chunk < j->B ? b0 + j->chunk : j->B;
        for (int b = b0; b < b1; ++b) {
            orc_kin_solve_g(j->m, j->idx + (size_t)b * T, j->theta + (size_t)b * T, 0, g, counts);
            int last = 0;
            for (int n = 0; n < T; ++n) last += counts[n];
            memcpy(j->tip_out + (size_t)b * 16, g + (size_t)(last - 1) * 16, 16 * sizeof(double));
        }
    }
    free(g);
    return NULL;
}

static void* worker_eval(void* arg) {
    orc_job* j = (orc_job*)arg;
    int T = j->m->tube_num, tot = total_nodes(j->m);
    double* g = (double*)malloc((size_t)tot * 16 * sizeof(double));
    double* pos = (double*)malloc((size_t)tot * 3 * sizeof(double));
    int counts[ORC_MAX_TUBES];
    for (;;) {
        int b0 = __atomic_fetch_add(&j->next, j->chunk, __ATOMIC_RELAXED);
        if (b0 >= j->B) break;
        int b1 = b0 + j->chunk < j->B ? b0 + j->chunk : j->B;
        for (int b = b0; b < b1; ++b) {
            orc_kin_solve_g(j->m, j->idx + (size_t)b * T, j->theta + (size_t)b * T, 0, g, counts);
            int nn = 0;
            for (int n = 0; n < T; ++n) nn += counts[n];
            for (int k = 0; k < nn; ++k) {
                pos[3 * k + 0] = g[16 * k + 3];
                pos[3 * k + 1] = g[16 * k + 7];
                pos[3 * k + 2] = g[16 * k + 11];
            }
            double out[2];
            int rc = orc_env_check_collision(j->env, pos, counts, T, j->rad, 0, out);
            if (rc) __atomic_store_n(&j->rc, rc, __ATOMIC_RELAXED);
            j->obs_min[b] = out[0];
            j->goal_dist[b] = out[1];
            /* rrt.py:116-129: colliding configs are discarded by the planner (cost NaN here) */
            j->cost[b] = out[0] < 0 ? NAN : orc_soapwg_own_cost(j->goal_weight, out[0], out[1]);
        }
    }
    free(g);
    free(pos);
    return NULL;
}

static void run_job(orc_job* j, void* (*fn)(void*), int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, fn, j);
    fn(j);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
}

int orc_kin_solve_g_batch(const orc_model* m, int B, const int* idx, const double* theta, double* tip_out,
                          int nthreads) {
    orc_job j;
    memset(&j, 0, sizeof j);
    j.m = m;
    j.B = B;
    j.idx = idx;
    j.theta = theta;
    j.tip_out = tip_out;
    j.chunk = 64;
    run_job(&j, worker_solve_g, nthreads);
    return 0;
}

int orc_env_eval_batch(const orc_model* m, const orc_env* env, int B, const int* idx, const double* theta,
                       const double* rad, double goal_weight, double* obs_min, double* goal_dist, double* cost,
                       int nthreads) {
    orc_job j;
    memset(&j, 0, sizeof j);
    j.m = m;
    j.env = env;
    j.B = B;
    j.idx = idx;
    j.theta = theta;
    j.rad = rad;
    j.goal_weight = goal_weight;
    j.obs_min = obs_min;
    j.goal_dist = goal_dist;
    j.cost = cost;
    j.chunk = 16;
    run_job(&j, worker_eval, nthreads);
    return j.rc;
}

int orc_eval_batch(const orc_model* m, int B, const int* idx, const double* theta, const double* rad,
                   const double* tris, int ntri, const orc_prim* prims, int nprim, const orc_prim* goal,
                   const double* goal_tris, int goal_ntri, double goal_weight, double* obs_min,
                   double* goal_dist, double* cost, int nthreads) {
    orc_env* e = orc_env_create(tris, ntri, prims, nprim, goal, goal_tris, goal_ntri);
    int rc = orc_env_eval_batch(m, e, B, idx, theta, rad, goal_weight, obs_min, goal_dist, cost, nthreads);
    orc_env_destroy(e);
    return rc;
}
