// ctrb_static.cu -- the static strain-based model (ctrdapp/model/static.py, static_bases.py
// 'last_strain_base', degree 2) as one warp per configuration on sm_100a.
//
// Static.solve_g (static.py:71-135) = MINPACK hybrd on a 15- (2 tubes) or 30-unknown
// (3 tubes) equilibrium + Magnus integration of the section curves.  Here:
//
//  * residual (static.py:228-454).  calculate_integral (:456-491) only ever samples its
//    integrand at the two Gauss abscissae, by linear interpolation between the bracketing
//    nodes, and every tube!=section twist is a pure x-torsion polynomial, so the relative
//    frames gg21/gg31/gg32 are Rx(theta0 + closed-form integral) and sg = closed-form
//    integral of the basis (2-point Gauss per step is exact for these cubics).  The residual
//    therefore needs 4 node evaluations per section instead of a march over every node, and
//    all wrenches reduce to 3-vector moments (the force rows cancel exactly: v = e_x).
//  * root finder: MINPACK hybrd itself (forward-difference Jacobian, QR, dogleg, Broyden
//    rank-1 updates, SciPy's defaults), so that iterates, evaluation counts and the
//    converged flag follow the reference's.  Work arrays live in shared memory; lanes own
//    Jacobian columns (fdjac1: one perturbed residual per lane), matrix columns (qrfac, qform)
//    and matrix rows (r1mpyq); the O(n^2) scalar recurrences (dogleg, r1updt) run on lane 0.
//  * curves: integrate_g (:165-198) with its absolute-index quirk, as a warp-parallel
//    inclusive scan of the per-node Magnus exponentials (SE(3) product is associative).
#include <float.h>
#include <math_constants.h>

#include <algorithm>
#include <cstring>

#include "ctrb_se3.cuh"

namespace ctrb {

constexpr int kSWarps = 3;    // warps (= configurations) per block: 5 blocks (15 warps) per SM by smem and registers
constexpr int kNP = 31;       // leading dimension of the n x n work matrices (odd: conflict-free column access)
constexpr int kNV = 32;       // padded vector length
constexpr int kMaxNq = kStaticMaxNq;

struct StCfg {
    int i11, i21, i31, i22, i32, i33;
    int L[3];  // nodes of sections 1..3
    double th[3];
    // calculate_integral's two abscissae per section: bracketing nodes and interpolation terms
    int lo[3][2], hi[3][2];
    double dxs[3][2], dxn[3][2], len[3];
};

// ------------------------------------------------------------------ small helpers
struct V3 {
    double x, y, z;
};
__device__ __forceinline__ V3 rotx(double s, double c, V3 v) { return {v.x, c * v.y - s * v.z, s * v.y + c * v.z}; }
__device__ __forceinline__ V3 rotxT(double s, double c, V3 v) { return {v.x, c * v.y + s * v.z, -s * v.y + c * v.z}; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double poly3(const double* q, double x) { return q[0] + q[1] * x + q[2] * (x * x); }
// closed-form integral of q . (1, x, x^2) from 0 to X
__device__ __forceinline__ double ipoly3(const double* q, double X) {
    return q[0] * X + q[1] * (0.5 * X * X) + q[2] * (X * X * X * (1.0 / 3.0));
}

// one out-of-line copy each of the fp64 sincos and of the strain-base switch: the solve kernel is
// instruction-cache bound if these are inlined at every use (ncu: stall_no_instruction 17 per issue)
__device__ __noinline__ void sincos_ni(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __noinline__ void strain_omega_ni(const ModelConst& mc, int n, double x, double w[3]) { strain_omega(mc, n, x, w); }

// rows 0..2 of a strain base as a 3 x dof matrix (strain_bases.py:7-224)
__device__ __noinline__ void strain_base_rows(int kind, int dof, double x, double B[3][CTRB_MAX_DOF]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < CTRB_MAX_DOF; ++k) B[r][k] = 0.0;
    switch (kind) {
        case CTRB_BASE_CONSTANT: B[1][0] = 1.0; break;
        case CTRB_BASE_LINEAR:
            B[1][0] = 1.0;
            if (dof > 2) B[1][1] = x;
            B[2][dof - 1] = x;
            break;
        case CTRB_BASE_QUADRATIC:
            B[1][0] = 1.0;
            if (dof > 2) B[1][1] = x * x;
            B[2][dof - 1] = x * x;
            break;
        case CTRB_BASE_PURE_HELIX: { double sn, cs; sincos_ni(x / 10, &sn, &cs); B[1][0] = sn; B[2][1] = cs; break; }
        case CTRB_BASE_LINEAR_HELIX: { double sn, cs; sincos_ni(x / 10, &sn, &cs); B[1][0] = sn; B[2][1] = x * cs; break; }
        case CTRB_BASE_TORSION_HELIX: B[0][0] = 1.0; B[1][1] = 1.0; B[2][2] = 1.0; break;
        case CTRB_BASE_TORSION_LINEAR_HELIX: B[0][0] = 1.0; B[1][1] = 1.0; B[2][2] = x; break;
        case CTRB_BASE_FULL: {
            int bs = (dof & 1) ? 1 : 2;
            B[0][0] = 1.0;
            if (!(dof & 1)) B[0][1] = x;
            B[1][bs] = 1.0;
            B[1][bs + 1] = x;
            if (dof <= 6) {
                B[2][bs + 2] = 1.0;
                B[2][bs + 3] = x;
            } else {
                B[1][bs + 2] = x * x;
                B[2][bs + 3] = 1.0;
                B[2][bs + 4] = x;
                B[2][bs + 5] = x * x;
            }
            break;
        }
        default: break;
    }
}

// ------------------------------------------------------------------ per-configuration setup
// static.py:81-95 section indices; :456-491 quadrature brackets
__device__ __noinline__ void st_setup(const StModel& sm, const int* idx, const double* theta, StCfg& c) {
    const int T = sm.mc.T;
    const int* N = sm.mc.npts;
    int i11 = idx[0], i22 = idx[T - 2], i33 = idx[T - 1];
    int i21 = i22 - (N[0] - i11 - 1);
    int i32 = i33 - (N[T - 2] - i22 - 1);
    int i31 = i32 - (N[0] - i11 - 1);
    if (T == 2) i11 = i21 = i31 = 0;
    c.i11 = i11; c.i21 = i21; c.i31 = i31; c.i22 = i22; c.i32 = i32; c.i33 = i33;
    c.L[0] = T == 3 ? N[0] - i11 : 1;
    c.L[1] = N[T - 2] - i22;
    c.L[2] = N[T - 1] - i33;
    for (int n = 0; n < 3; ++n) c.th[n] = n < T ? theta[n] : 0.0;
    const double cs[2] = {(1.0 / sqrt(3.0) + 1.0) / 2.0, (-1.0 / sqrt(3.0) + 1.0) / 2.0};  // static.py:601-603
    for (int s = 0; s < 3; ++s) {
        const int np = c.L[s];
        const double length = (np - 1) * sm.mc.dx;
        c.len[s] = length;
        for (int a = 0; a < 2; ++a) {
            c.lo[s][a] = c.hi[s][a] = 0;
            c.dxs[s][a] = 1.0;
            c.dxn[s][a] = 0.0;
            if (!(length > 0.0)) continue;
            const double step = length / (np - 1);  // np.linspace(0, length, num_points)
            const double xn = cs[a] * length;
            // searchsorted(x_vals, xn, side='left') clipped to [1, np-1]
            int hi = (int)ceil(xn / step);
            if (hi < 0) hi = 0;
            if (hi > np - 1) hi = np - 1;
            while (hi < np - 1 && ((hi == np - 1) ? length : hi * step) < xn) ++hi;
            while (hi > 0 && (((hi - 1) == np - 1) ? length : (hi - 1) * step) >= xn) --hi;
            if (hi < 1) hi = 1;
            const int lo = hi - 1;
            const double xlo = lo * step, xhi = (hi == np - 1) ? length : hi * step;
            c.lo[s][a] = lo;
            c.hi[s][a] = hi;
            c.dxs[s][a] = xhi - xlo;
            c.dxn[s][a] = xn - xlo;
        }
    }
}

// ------------------------------------------------------------------ integrands at one node
// section 1 (static.py:261-345), three tubes; y = [intc1 (9), intg21 (3), intg31 (3)]
__device__ __noinline__ void node_sec1(const StModel& sm, const StCfg& c, const double* q, int i, double y[15]) {
    const double dx = sm.mc.dx, X = i * dx;
    const double* q11 = q + sm.qoff[0];
    const double* q21 = q + sm.qoff[1];
    const double* q31 = q + sm.qoff[2];
    V3 w1 = {poly3(q11, X), poly3(q11 + 3, X), poly3(q11 + 6, X)};
    const double tg21 = poly3(q21, X), tg31 = poly3(q31, X);
    double s21, c21, s31, c31;
    sincos_ni(c.th[1] + ipoly3(q21, X), &s21, &c21);  // gg21 = Rx(theta1 + int tau21)
    sincos_ni(c.th[2] + ipoly3(q31, X), &s31, &c31);
    V3 w21 = rotxT(s21, c21, w1);
    w21.x += tg21;
    V3 w31 = rotxT(s31, c31, w21);
    w31.x += tg31;
    double ws[3];
    strain_omega_ni(sm.mc, 0, (i + c.i11) * dx, ws);
    V3 m11 = {sm.Kr[0][0] * (w1.x - ws[0]), sm.Kr[0][1] * (w1.y - ws[1]), sm.Kr[0][2] * (w1.z - ws[2])};
    strain_omega_ni(sm.mc, 1, (i + c.i21) * dx, ws);
    V3 m21 = {sm.Kr[1][0] * (w21.x - ws[0]), sm.Kr[1][1] * (w21.y - ws[1]), sm.Kr[1][2] * (w21.z - ws[2])};
    strain_omega_ni(sm.mc, 2, (i + c.i31) * dx, ws);
    V3 m31 = {sm.Kr[2][0] * (w31.x - ws[0]), sm.Kr[2][1] * (w31.y - ws[1]), sm.Kr[2][2] * (w31.z - ws[2])};
    V3 r31m = rotx(s31, c31, m31);                 // coAd(gg31) fi_31 (moment part)
    V3 u = {m21.x + r31m.x, m21.y + r31m.y, m21.z + r31m.z};
    V3 ru = rotx(s21, c21, u);                      // coAd21 (fi_21 + coAd31 fi_31)
    V3 wsum = {m11.x + ru.x, m11.y + ru.y, m11.z + ru.z};
    V3 z21 = cross3(w1, ru);                        // coad(ksi_1) (...)
    V3 z31 = cross3(w1, rotx(s21, c21, r31m));
    const double P[3] = {1.0, X, X * X};
    const double S[3] = {X, 0.5 * X * X, X * X * X * (1.0 / 3.0)};  // sg (row 0) = int of the basis
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        y[k] = P[k] * wsum.x;
        y[3 + k] = P[k] * wsum.y;
        y[6 + k] = P[k] * wsum.z;
        y[9 + k] = P[k] * (m21.x + m31.x) - S[k] * z21.x;
        y[12 + k] = P[k] * m31.x - S[k] * z31.x;
    }
}

// section 2 (static.py:347-436); y = [intc2 (9), intg32 (3), intg312 (3)]
__device__ __noinline__ void node_sec2(const StModel& sm, const StCfg& c, const double* q, int i, double s31e, double c31e,
                          const double S31e[3], double y[15]) {
    const int T = sm.mc.T;
    const double dx = sm.mc.dx, X = i * dx;
    const double* q22 = q + sm.qoff[3];
    const double* q32 = q + sm.qoff[4];
    V3 w2 = {poly3(q22, X), poly3(q22 + 3, X), poly3(q22 + 6, X)};
    const double t32 = poly3(q32, X);
    double s32, c32;
    sincos_ni((T == 2 ? c.th[1] : 0.0) + ipoly3(q32, X), &s32, &c32);
    V3 w32 = rotxT(s32, c32, rotxT(s31e, c31e, w2));
    w32.x += t32;
    double ws[3];
    strain_omega_ni(sm.mc, T - 2, (i + c.i22) * dx, ws);
    V3 m22 = {sm.Kr[T - 2][0] * (w2.x - ws[0]), sm.Kr[T - 2][1] * (w2.y - ws[1]), sm.Kr[T - 2][2] * (w2.z - ws[2])};
    strain_omega_ni(sm.mc, T - 1, (i + c.i32) * dx, ws);
    V3 m32 = {sm.Kr[T - 1][0] * (w32.x - ws[0]), sm.Kr[T - 1][1] * (w32.y - ws[1]), sm.Kr[T - 1][2] * (w32.z - ws[2])};
    V3 a = rotx(s31e, c31e, rotx(s32, c32, m32));   // coAd31 coAd32 fi_32
    V3 wsum = {m22.x + a.x, m22.y + a.y, m22.z + a.z};
    V3 z = cross3(w2, a);
    const double P[3] = {1.0, X, X * X};
    const double S[3] = {X, 0.5 * X * X, X * X * X * (1.0 / 3.0)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        y[k] = P[k] * wsum.x;
        y[3 + k] = P[k] * wsum.y;
        y[6 + k] = P[k] * wsum.z;
        y[9 + k] = P[k] * m32.x - S[k] * z.x;
        y[12 + k] = S31e[k] * z.x;
    }
}

// section 3 (static.py:438-452); y = intc3 (dof of the last tube)
__device__ __noinline__ void node_sec3(const StModel& sm, const StCfg& c, const double* q, int i, double y[CTRB_MAX_DOF]) {
    const int T = sm.mc.T, dof = sm.mc.dof[T - 1];
    const double* q33 = q + sm.qoff[5];
    double B[3][CTRB_MAX_DOF];
    strain_base_rows(sm.mc.kind[T - 1], dof, (i + c.i33) * sm.mc.dx, B);
    double m[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < dof; ++k) {
            a += B[r][k] * q33[k];
            b += B[r][k] * sm.mc.q[T - 1][k];
        }
        m[r] = sm.Kr[T - 1][r] * (a - b);
    }
    for (int k = 0; k < dof; ++k) y[k] = B[0][k] * m[0] + B[1][k] * m[1] + B[2][k] * m[2];
}

// static_equilibrium (static.py:228-259): F(q), n = sm.nq
__device__ __noinline__ void st_residual(const StModel& sm, const StCfg& c, const double* q, double* F) {
    const int T = sm.mc.T;
    double s31e = 0.0, c31e = 1.0, S31e[3] = {0.0, 0.0, 0.0};
    double F31[3] = {0.0, 0.0, 0.0};
    if (T == 3) {
        double acc[15];
#pragma unroll
        for (int k = 0; k < 15; ++k) acc[k] = 0.0;
        if (c.len[0] > 0.0) {
#pragma unroll 1
            for (int a = 0; a < 2; ++a) {
                double y2[2][15];
#pragma unroll 1
                for (int e = 0; e < 2; ++e) node_sec1(sm, c, q, e ? c.hi[0][a] : c.lo[0][a], y2[e]);
                const double inv = 1.0 / c.dxs[0][a];
#pragma unroll 1
                for (int k = 0; k < 15; ++k) acc[k] += 0.5 * (((y2[1][k] - y2[0][k]) * inv) * c.dxn[0][a] + y2[0][k]);
            }
        }
        const double h = c.len[0] / 2;
        for (int k = 0; k < 9; ++k) F[sm.qoff[0] + k] = h * acc[k];
        for (int k = 0; k < 3; ++k) F[sm.qoff[1] + k] = h * acc[9 + k];
        for (int k = 0; k < 3; ++k) F31[k] = h * acc[12 + k];
        // final relative frame / basis integral of tube 3 over section 1 (equilibrium_three's return)
        const double Xe = (c.L[0] - 1) * sm.mc.dx;
        sincos_ni(c.th[2] + ipoly3(q + sm.qoff[2], Xe), &s31e, &c31e);
        S31e[0] = Xe;
        S31e[1] = 0.5 * Xe * Xe;
        S31e[2] = Xe * Xe * Xe * (1.0 / 3.0);
    }
    {
        double acc[15];
#pragma unroll
        for (int k = 0; k < 15; ++k) acc[k] = 0.0;
        if (c.len[1] > 0.0) {
#pragma unroll 1
            for (int a = 0; a < 2; ++a) {
                double y2[2][15];
#pragma unroll 1
                for (int e = 0; e < 2; ++e) node_sec2(sm, c, q, e ? c.hi[1][a] : c.lo[1][a], s31e, c31e, S31e, y2[e]);
                const double inv = 1.0 / c.dxs[1][a];
#pragma unroll 1
                for (int k = 0; k < 15; ++k) acc[k] += 0.5 * (((y2[1][k] - y2[0][k]) * inv) * c.dxn[1][a] + y2[0][k]);
            }
        }
        const double h = c.len[1] / 2;
        for (int k = 0; k < 9; ++k) F[sm.qoff[3] + k] = h * acc[k];
        for (int k = 0; k < 3; ++k) F[sm.qoff[4] + k] = h * acc[9 + k];
        if (T == 3)
            for (int k = 0; k < 3; ++k) F[sm.qoff[2] + k] = F31[k] - h * acc[12 + k];  // intg_31 - intg_312
    }
    {
        const int dof = sm.ndof[5];
        double acc[CTRB_MAX_DOF];
        for (int k = 0; k < CTRB_MAX_DOF; ++k) acc[k] = 0.0;
        if (c.len[2] > 0.0) {
#pragma unroll 1
            for (int a = 0; a < 2; ++a) {
                double y2[2][CTRB_MAX_DOF];
#pragma unroll 1
                for (int e = 0; e < 2; ++e) node_sec3(sm, c, q, e ? c.hi[2][a] : c.lo[2][a], y2[e]);
                const double inv = 1.0 / c.dxs[2][a];
#pragma unroll 1
                for (int k = 0; k < dof; ++k) acc[k] += 0.5 * (((y2[1][k] - y2[0][k]) * inv) * c.dxn[2][a] + y2[0][k]);
            }
        }
        const double h = c.len[2] / 2;
        for (int k = 0; k < dof; ++k) F[sm.qoff[5] + k] = h * acc[k];
    }
}

// ------------------------------------------------------------------ MINPACK hybrd, warp-cooperative
// All arrays are shared memory of this warp.  Matrices are column-major with leading dimension kNP.
struct HybrdMem {
    double fjac[kNP * kMaxNq];
    double r[kMaxNq * (kMaxNq + 1) / 2];
    double x[kNV], fvec[kNV], qtf[kNV], wa1[kNV], wa2[kNV], wa3[kNV], wa4[kNV];
    double ybuf0[12][16];     // node integrands of a residual (3 sections x 2 abscissae x {lo, hi})
    StCfg cfg;
    // During the forward-difference Jacobian two residuals are evaluated side by side.  R, wa3 and wa4 are
    // dead then (R is rebuilt from the QR right after), so the second set of node vectors and the two
    // perturbed unknown vectors alias them.
    __device__ double (*ybuf(int g))[16] { return g == 0 ? ybuf0 : reinterpret_cast<double(*)[16]>(r); }
    __device__ double* qbuf(int g) { return g == 0 ? wa3 : wa4; }
};

__device__ __noinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

__device__ __noinline__ double st_enorm_full(int n, const double* x) {
    // MINPACK enorm with its three scaled accumulators (only reached outside the intermediate range)
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    const double agiant = rgiant / (double)n;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const double xabs = fabs(x[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 += xabs * xabs;
        } else if (xabs <= rdwarf) {
            if (xabs > x3max) {
                const double rr = x3max / xabs;
                s3 = 1 + s3 * rr * rr;
                x3max = xabs;
            } else if (xabs != 0) {
                const double rr = xabs / x3max;
                s3 += rr * rr;
            }
        } else {
            if (xabs > x1max) {
                const double rr = x1max / xabs;
                s1 = 1 + s1 * rr * rr;
                x1max = xabs;
            } else {
                const double rr = xabs / x1max;
                s1 += rr * rr;
            }
        }
    }
    if (s1 != 0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0) {
        if (s2 >= x3max) return sqrt(s2 * (1 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

// MINPACK enorm of a length-n vector held one element per lane (v = 0 in lanes >= n).  In the
// intermediate range the routine is sqrt(sum x^2); the warp tree-sum only changes the order of the adds.
__device__ __noinline__ double enorm_w(int n, double v, const double* x_smem_or_null) {
    const double xabs = fabs(v);
    const bool odd = (xabs >= 1.304e19 / (double)n) || (xabs <= 3.834e-20 && xabs != 0.0);
    if (__any_sync(0xffffffffu, odd) && x_smem_or_null) return st_enorm_full(n, x_smem_or_null);
    return sqrt(warp_sum(xabs * xabs));
}

__device__ __forceinline__ int rowoff(int n, int j) { return j * n - (j * (j - 1)) / 2; }  // packed upper-triangular rows

// Cooperative residual, phase 1: work item t in 0..11 = (section, abscissa, bracket end) evaluates one node
// integrand vector into y (shared memory).  q is a shared-memory vector of the unknowns.
__device__ __noinline__ void res_node(const StModel& sm, const StCfg& c, const double* q, int t, double* y) {
    const int T = sm.mc.T;
    const int sec = t >> 2, a = (t >> 1) & 1, e = t & 1;
    if (c.len[sec] > 0.0 && (sec != 0 || T == 3)) {
        const int node = e ? c.hi[sec][a] : c.lo[sec][a];
        if (sec == 0) {
            node_sec1(sm, c, q, node, y);
        } else if (sec == 1) {
            double s31e = 0.0, c31e = 1.0, S31e[3] = {0.0, 0.0, 0.0};
            if (T == 3) {  // final relative frame / basis integral of tube 3 over section 1 (equilibrium_three's return)
                const double Xe = (c.L[0] - 1) * sm.mc.dx;
                sincos_ni(c.th[2] + ipoly3(q + sm.qoff[2], Xe), &s31e, &c31e);
                S31e[0] = Xe;
                S31e[1] = 0.5 * Xe * Xe;
                S31e[2] = Xe * Xe * Xe * (1.0 / 3.0);
            }
            node_sec2(sm, c, q, node, s31e, c31e, S31e, y);
        } else {
            node_sec3(sm, c, q, node, y);
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < 16; ++k) y[k] = 0.0;
    }
}

// phase 2: component k of F from the 12 node vectors (calculate_integral, static.py:456-491)
__device__ __noinline__ double res_assemble(const StModel& sm, const StCfg& c, const double (*yb)[16], int comp) {
    int blk = 0;  // order (1,1),(2,1),(3,1),(2,2),(3,2),(3,3)
#pragma unroll
    for (int b = 1; b < 6; ++b)
        if (comp >= sm.qoff[b] && sm.ndof[b] > 0) blk = b;
    const int k = comp - sm.qoff[blk];
    const int sec = blk < 3 ? 0 : (blk < 5 ? 1 : 2);
    const int yo = (blk == 0 || blk == 3 || blk == 5) ? 0 : ((blk == 1 || blk == 4) ? 9 : 12);
    double acc = 0.0;
    if (c.len[sec] > 0.0) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const double ylo = yb[sec * 4 + a * 2 + 0][yo + k], yhi = yb[sec * 4 + a * 2 + 1][yo + k];
            acc += 0.5 * (((yhi - ylo) / c.dxs[sec][a]) * c.dxn[sec][a] + ylo);
        }
    }
    double val = (c.len[sec] / 2) * acc;
    if (blk == 2) {  // intg_31 - intg_312: the section-2 share lives at offset 12 of the section-2 vectors
        double acc2 = 0.0;
        if (c.len[1] > 0.0) {
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const double ylo = yb[4 + a * 2 + 0][12 + k], yhi = yb[4 + a * 2 + 1][12 + k];
                acc2 += 0.5 * (((yhi - ylo) / c.dxs[1][a]) * c.dxn[1][a] + ylo);
            }
        }
        val -= (c.len[1] / 2) * acc2;
    }
    return val;
}

// one residual evaluation F(q) by the whole warp; q and F are shared-memory vectors
__device__ __forceinline__ void st_residual_w(const StModel& sm, HybrdMem& M, const double* q, double* F, int lane) {
    if (lane < 12) res_node(sm, M.cfg, q, lane, M.ybuf0[lane]);
    __syncwarp();
    if (lane < sm.nq) F[lane] = res_assemble(sm, M.cfg, M.ybuf0, lane);
    __syncwarp();
}

// dogleg (MINPACK): Gauss-Newton step by back substitution on the packed R, scaled gradient, and their
// convex combination inside the trust region.  One vector element per lane; returns this lane's p.
__device__ __noinline__ double dogleg_w(int n, const double* r, double diag_l, const double* qtb, double delta, double* tmp,
                                        int lane) {
    const double epsmch = DBL_EPSILON;
    double xg = 0.0;  // Gauss-Newton component of this lane
#pragma unroll 1
    for (int j = n - 1; j >= 0; --j) {
        const int ro = rowoff(n, j);
        const double term = (lane > j && lane < n) ? r[ro + lane - j] * xg : 0.0;
        const double sum = warp_sum(term);
        double temp = r[ro];
        if (temp == 0.0) {
            const double colv = (lane <= j) ? fabs(r[rowoff(n, lane) + j - lane]) : 0.0;
            temp = epsmch * warp_max(colv);
            if (temp == 0.0) temp = epsmch;
        }
        const double xj = (qtb[j] - sum) / temp;
        if (lane == j) xg = xj;
    }
    const bool act = lane < n;
    const double qnorm = enorm_w(n, act ? diag_l * xg : 0.0, nullptr);
    if (qnorm <= delta) return xg;
    // scaled gradient direction: wa1_i = (sum_{j<=i} r(j,i) qtb_j) / diag_i
    double g = 0.0;
    if (act) {
#pragma unroll 1
        for (int j = 0; j <= lane; ++j) g += r[rowoff(n, j) + lane - j] * qtb[j];
        g = g / diag_l;
    }
    const double gnorm = enorm_w(n, g, nullptr);
    double sgnorm = 0.0, alpha = delta / qnorm;
    if (gnorm != 0.0) {
        g = act ? (g / gnorm) / diag_l : 0.0;
        if (act) tmp[lane] = g;
        __syncwarp();
        double w2 = 0.0;
        if (act) {
            const int ro = rowoff(n, lane);
#pragma unroll 1
            for (int i = lane; i < n; ++i) w2 += r[ro + i - lane] * tmp[i];
        }
        double temp = enorm_w(n, w2, nullptr);
        sgnorm = (gnorm / temp) / temp;
        alpha = 0.0;
        if (sgnorm < delta) {
            const double bnorm = enorm_w(n, act ? qtb[lane] : 0.0, nullptr);
            temp = (bnorm / gnorm) * (bnorm / qnorm) * (sgnorm / delta);
            const double dq = delta / qnorm, sd = sgnorm / delta;
            temp = temp - dq * sd * sd + sqrt((temp - dq) * (temp - dq) + (1 - dq * dq) * (1 - sd * sd));
            alpha = (dq * (1 - sd * sd)) / temp;
        }
        __syncwarp();
    }
    const double temp = (1 - alpha) * fmin(sgnorm, delta);
    return temp * g + alpha * xg;
}

__device__ __forceinline__ void givens(double a, double b, double* c_, double* s_, double* tau) {
    // the rotation MINPACK builds to annihilate b against a (r1updt): returns cos, sin and the stored tau
    if (fabs(a) < fabs(b)) {
        const double cotan = a / b;
        *s_ = 0.5 / sqrt(0.25 + 0.25 * cotan * cotan);
        *c_ = *s_ * cotan;
        *tau = 1;
        if (fabs(*c_) * DBL_MAX > 1) *tau = 1 / *c_;
    } else {
        const double tn = b / a;
        *c_ = 0.5 / sqrt(0.25 + 0.25 * tn * tn);
        *s_ = *c_ * tn;
        *tau = *s_;
    }
}

// r1updt: QR factors of R + u v^T.  Lane i owns w_i and element i of every packed row segment.
// On return v (smem) and w (smem) hold the rotations for r1mpyq.
__device__ __noinline__ void r1updt_w(int n, double* s, double u_l, double* v, double* w, int lane) {
    double wl = 0.0;
    if (lane == n - 1) wl = s[rowoff(n, n - 1)];
    double vn = v[n - 1];
#pragma unroll 1
    for (int j = n - 2; j >= 0; --j) {
        if (lane == j) wl = 0.0;
        const double vj = v[j];
        __syncwarp();
        if (vj != 0.0) {
            double c_, s_, tau;
            givens(vn, vj, &c_, &s_, &tau);
            vn = s_ * vj + c_ * vn;
            if (lane == 0) v[j] = tau;
            if (lane >= j && lane < n) {
                const int l = rowoff(n, j) + lane - j;
                const double sl = s[l];
                s[l] = c_ * sl - s_ * wl;
                wl = s_ * sl + c_ * wl;
            }
        }
    }
    if (lane == 0) v[n - 1] = vn;
    if (lane < n) wl += vn * u_l;
#pragma unroll 1
    for (int j = 0; j < n - 1; ++j) {
        const double wj = __shfl_sync(0xffffffffu, wl, j);
        if (wj != 0.0) {
            const int ro = rowoff(n, j);
            double c_, s_, tau;
            givens(s[ro], wj, &c_, &s_, &tau);
            __syncwarp();
            if (lane >= j && lane < n) {
                const int l = ro + lane - j;
                const double sl = s[l];
                s[l] = c_ * sl + s_ * wl;
                wl = -s_ * sl + c_ * wl;
            }
            if (lane == j) wl = tau;
            __syncwarp();
        }
    }
    if (lane == n - 1) s[rowoff(n, n - 1)] = wl;
    if (lane < n) w[lane] = wl;
    __syncwarp();
}

// apply the 2(n-1) stored rotations to ONE row a[0..n-1] with stride lda (r1mpyq, one row per lane)
__device__ __noinline__ void st_r1mpyq_row(int n, double* a, int lda, const double* v, const double* w) {
#pragma unroll 1
    for (int j = n - 2; j >= 0; --j) {
        double c_, s_;
        if (fabs(v[j]) > 1) {
            c_ = 1 / v[j];
            s_ = sqrt(1 - c_ * c_);
        } else {
            s_ = v[j];
            c_ = sqrt(1 - s_ * s_);
        }
        const double temp = c_ * a[j * lda] - s_ * a[(n - 1) * lda];
        a[(n - 1) * lda] = s_ * a[j * lda] + c_ * a[(n - 1) * lda];
        a[j * lda] = temp;
    }
#pragma unroll 1
    for (int j = 0; j < n - 1; ++j) {
        double c_, s_;
        if (fabs(w[j]) > 1) {
            c_ = 1 / w[j];
            s_ = sqrt(1 - c_ * c_);
        } else {
            s_ = w[j];
            c_ = sqrt(1 - s_ * s_);
        }
        const double temp = c_ * a[j * lda] + s_ * a[(n - 1) * lda];
        a[(n - 1) * lda] = -s_ * a[j * lda] + c_ * a[(n - 1) * lda];
        a[j * lda] = temp;
    }
}

// MINPACK hybrd, warp-cooperative.  Every scalar of the iteration (delta, norms, counters, ratio) is
// computed redundantly and identically by all lanes; vectors are one element per lane.
// returns info (1 = converged, 2 = maxfev, 3 = xtol too small, 4/5 = slow progress)
__device__ __noinline__ int st_hybrd(const StModel& sm, HybrdMem& M, int lane, double xtol, int maxfev, double factor,
                                     int* nfev_out) {
    const int n = sm.nq;
    const bool act = lane < n;
    const double epsmch = DBL_EPSILON, p1 = 0.1, p5 = 0.5, p001 = 1e-3, p0001 = 1e-4;
    double* fjac = M.fjac;
    int info = 0, nfev;
    st_residual_w(sm, M, M.x, M.fvec, lane);
    nfev = 1;
    double fnorm = enorm_w(n, act ? M.fvec[lane] : 0.0, M.fvec);
    int iter = 1, ncsuc = 0, ncfail = 0, nslow1 = 0, nslow2 = 0;
    double delta = 0, xnorm = 0;
    double diag_l = 1.0;
#pragma unroll 1
    for (;;) {  // outer loop: fresh forward-difference Jacobian
        bool jeval = true;
        // fdjac1 (dense): columns two at a time -- lanes 0..11 evaluate the node integrands of x + h_j e_j,
        // lanes 12..23 those of x + h_(j+1) e_(j+1); then lanes 0..n-1 assemble both columns
        {
            const double eps = sqrt(epsmch);
#pragma unroll 1
            for (int j0 = 0; j0 < n; j0 += 2) {
                if (act) {
                    const double xv = M.x[lane];
                    double h = eps * fabs(xv);
                    if (h == 0) h = eps;
                    M.qbuf(0)[lane] = (lane == j0) ? xv + h : xv;
                    M.qbuf(1)[lane] = (lane == j0 + 1) ? xv + h : xv;
                }
                __syncwarp();
                if (lane < 24) {
                    const int g = lane >= 12 ? 1 : 0;
                    if (j0 + g < n) res_node(sm, M.cfg, M.qbuf(g), lane - 12 * g, M.ybuf(g)[lane - 12 * g]);
                }
                __syncwarp();
                if (act) {
#pragma unroll 1
                    for (int g = 0; g < 2; ++g) {
                        const int j = j0 + g;
                        if (j >= n) break;
                        const double xv = M.x[j];
                        double h = eps * fabs(xv);
                        if (h == 0) h = eps;
                        const double Fp = res_assemble(sm, M.cfg, M.ybuf(g), lane);
                        fjac[lane + j * kNP] = (Fp - M.fvec[lane]) / h;
                    }
                }
                __syncwarp();
            }
        }
        nfev += n;
        __syncwarp();
        // qrfac: Householder QR without pivoting; lane k owns column k.  acnorm -> wa2, rdiag -> wa1
        double acnorm_l = 0.0;
        if (act) {
            double s2 = 0.0;
#pragma unroll 1
            for (int i = 0; i < n; ++i) s2 += fjac[i + lane * kNP] * fjac[i + lane * kNP];
            acnorm_l = sqrt(s2);
        }
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            double ajnorm = enorm_w(n, (lane >= j && act) ? fjac[lane + j * kNP] : 0.0, nullptr);  // lane = row
            if (ajnorm != 0) {
                if (fjac[j + j * kNP] < 0) ajnorm = -ajnorm;
                __syncwarp();
                if (lane >= j && act) fjac[lane + j * kNP] /= ajnorm;
                __syncwarp();
                if (lane == 0) fjac[j + j * kNP] += 1;
                __syncwarp();
                if (lane > j && act) {  // lane = column k
                    double sum = 0;
#pragma unroll 1
                    for (int i = j; i < n; ++i) sum += fjac[i + j * kNP] * fjac[i + lane * kNP];
                    const double temp = sum / fjac[j + j * kNP];
#pragma unroll 1
                    for (int i = j; i < n; ++i) fjac[i + lane * kNP] -= temp * fjac[i + j * kNP];
                }
            }
            if (lane == 0) M.wa1[j] = -ajnorm;
            __syncwarp();
        }
        if (iter == 1) {
            diag_l = acnorm_l == 0 ? 1.0 : acnorm_l;
            xnorm = enorm_w(n, act ? diag_l * M.x[lane] : 0.0, nullptr);
            delta = factor * xnorm;
            if (delta == 0) delta = factor;
        }
        // qtf = Q^T fvec through the stored Householder vectors
        double qtf_l = act ? M.fvec[lane] : 0.0;
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const double fjj = fjac[j + j * kNP];
            if (fjj != 0) {
                const double fij = (lane >= j && act) ? fjac[lane + j * kNP] : 0.0;
                const double sum = warp_sum(fij * qtf_l);
                qtf_l += fij * (-sum / fjj);
            }
        }
        if (act) {
            M.qtf[lane] = qtf_l;
            // copy R: strict upper part of column `lane` from fjac, diagonal from rdiag
#pragma unroll 1
            for (int i = 0; i < lane; ++i) M.r[rowoff(n, i) + lane - i] = fjac[i + lane * kNP];
            M.r[rowoff(n, lane)] = M.wa1[lane];
        }
        __syncwarp();
        // qform: accumulate Q in fjac; lane = column of the trailing update
#pragma unroll 1
        for (int j = 1; j < n; ++j)
            if (lane < j) fjac[lane + j * kNP] = 0;
        __syncwarp();
#pragma unroll 1
        for (int l = 0; l < n; ++l) {
            const int k = n - l - 1;
            if (lane >= k && act) M.wa3[lane] = fjac[lane + k * kNP];
            __syncwarp();
            if (lane >= k && act) fjac[lane + k * kNP] = (lane == k) ? 1.0 : 0.0;
            __syncwarp();
            if (M.wa3[k] != 0) {
                if (lane >= k && act) {
                    double sum = 0;
#pragma unroll 1
                    for (int i = k; i < n; ++i) sum += fjac[i + lane * kNP] * M.wa3[i];
                    const double temp = sum / M.wa3[k];
#pragma unroll 1
                    for (int i = k; i < n; ++i) fjac[i + lane * kNP] -= temp * M.wa3[i];
                }
            }
            __syncwarp();
        }
        diag_l = fmax(diag_l, acnorm_l);
#pragma unroll 1
        for (;;) {  // inner loop
            // direction p (this lane's component), trial point, scaled step
            double p_l = -dogleg_w(n, M.r, diag_l, M.qtf, delta, M.wa3, lane);
            if (!act) p_l = 0.0;
            if (act) {
                M.wa1[lane] = p_l;
                M.wa2[lane] = M.x[lane] + p_l;
            }
            const double pnorm = enorm_w(n, diag_l * p_l, nullptr);
            if (iter == 1) delta = fmin(delta, pnorm);
            __syncwarp();
            st_residual_w(sm, M, M.wa2, M.wa4, lane);
            nfev++;
            const double fnorm1 = enorm_w(n, act ? M.wa4[lane] : 0.0, M.wa4);
            double actred = -1;
            if (fnorm1 < fnorm) actred = 1 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
            // predicted reduction: || qtf + R p ||
            double w3 = 0.0;
            if (act) {
                const int ro = rowoff(n, lane);
                double sum = 0;
#pragma unroll 1
                for (int j = lane; j < n; ++j) sum += M.r[ro + j - lane] * M.wa1[j];
                w3 = M.qtf[lane] + sum;
            }
            const double temp = enorm_w(n, w3, nullptr);
            double prered = 0;
            if (temp < fnorm) prered = 1 - (temp / fnorm) * (temp / fnorm);
            double ratio = 0;
            if (prered > 0) ratio = actred / prered;
            if (ratio < p1) {
                ncsuc = 0;
                ncfail++;
                delta = p5 * delta;
            } else {
                ncfail = 0;
                ncsuc++;
                if (ratio >= p5 || ncsuc > 1) delta = fmax(delta, pnorm / p5);
                if (fabs(ratio - 1) <= p1) delta = pnorm / p5;
            }
            if (ratio >= p0001) {  // successful iteration
                if (act) {
                    M.x[lane] = M.wa2[lane];
                    M.fvec[lane] = M.wa4[lane];
                }
                xnorm = enorm_w(n, act ? diag_l * M.wa2[lane] : 0.0, nullptr);
                fnorm = fnorm1;
                iter++;
            }
            nslow1++;
            if (actred >= p001) nslow1 = 0;
            if (jeval) nslow2++;
            if (actred >= p1) nslow2 = 0;
            if (delta <= xtol * xnorm || fnorm == 0) info = 1;
            if (info == 0) {
                if (nfev >= maxfev) info = 2;
                if (p1 * fmax(p1 * delta, pnorm) <= epsmch * xnorm) info = 3;
                if (nslow2 == 5) info = 4;
                if (nslow1 == 10) info = 5;
            }
            __syncwarp();
            if (info != 0) {
                *nfev_out = nfev;
                return info;
            }
            if (ncfail == 2) break;  // recompute the Jacobian
            // rank-one (Broyden) update of R, Q and qtf
            double u_l = 0.0;
            if (act) {
                double sum = 0;
#pragma unroll 1
                for (int i = 0; i < n; ++i) sum += fjac[i + lane * kNP] * M.wa4[i];
                M.wa2[lane] = (sum - w3) / pnorm;                 // v
                u_l = diag_l * ((diag_l * p_l) / pnorm);          // u
                if (ratio >= p0001) M.qtf[lane] = sum;
            }
            __syncwarp();
            r1updt_w(n, M.r, u_l, M.wa2, M.wa3, lane);
            if (act) st_r1mpyq_row(n, fjac + lane, kNP, M.wa2, M.wa3);
            if (lane == 31) st_r1mpyq_row(n, M.qtf, 1, M.wa2, M.wa3);
            __syncwarp();
            jeval = false;
        }
    }
}

// ------------------------------------------------------------------ curve integration (integrate_g)
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

// half_step_integral for a tube==section basis (with bias): the Magnus exponential of one step.
// wA, wB = omega at the two Gauss points; v = e_x at both.
__device__ __noinline__ SE3 magnus_step(const double wA[3], const double wB[3], double dx) {
    const double cf = (sqrt(3.0) * dx * dx) / 12;
    // Gamma = dx/2 (xi1^ + xi2^) + cf [xi2^, xi1^];  [.,.] = (w2 x w1, w2 x v1 - w1 x v2), v1 = v2 = e_x
    V3 a = {wA[0], wA[1], wA[2]}, b = {wB[0], wB[1], wB[2]};
    V3 cw = cross3(b, a);
    V3 d = {b.x - a.x, b.y - a.y, b.z - a.z};
    V3 cv = cross3(d, V3{1.0, 0.0, 0.0});
    double w[3] = {(dx / 2) * (a.x + b.x) + cf * cw.x, (dx / 2) * (a.y + b.y) + cf * cw.y, (dx / 2) * (a.z + b.z) + cf * cw.z};
    double v[3] = {(dx / 2) * 2.0 + cf * cv.x, cf * cv.y, cf * cv.z};
    return se3_exp(w, v, 1.0);
}

// omega of the (tube==section) basis at absolute section coordinate `sidx` (static_bases.py:61-81)
__device__ __noinline__ void sect_omega(const StModel& sm, const StCfg& c, const double* qs, int sect, double sidx,
                                           double w[3]) {
    const double dx = sm.mc.dx;
    if (sect == 3) {
        const int T = sm.mc.T, dof = sm.mc.dof[T - 1];
        double B[3][CTRB_MAX_DOF];
        strain_base_rows(sm.mc.kind[T - 1], dof, (sidx + c.i33) * dx, B);  // _easy_last_basis adds idx33 (again)
        for (int r = 0; r < 3; ++r) {
            double a = 0.0;
            for (int k = 0; k < dof; ++k) a += B[r][k] * qs[k];
            w[r] = a;
        }
    } else {
        const double x = sidx * dx;
        w[0] = poly3(qs, x);
        w[1] = poly3(qs + 3, x);
        w[2] = poly3(qs + 6, x);
    }
}

// Writes `count` nodes of one section curve: node 0 = g0, node k = g0 * E_1 ... E_k with
// E_k the Magnus step at absolute index start + k.  Returns the last node (all lanes).
__device__ __noinline__ SE3 integrate_section(const StModel& sm, const StCfg& c, const double* qs, int sect, int start, int count,
                                 SE3 g0, double* out, int lane) {
    const double dx = sm.mc.dx;
    const double cg1 = 0.5 - sqrt(3.0) / 6, cg2 = 0.5 + sqrt(3.0) / 6;  // static.py:597-599
    if (lane == 0 && out) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            out[r * 4 + 0] = g0.r[r * 3 + 0];
            out[r * 4 + 1] = g0.r[r * 3 + 1];
            out[r * 4 + 2] = g0.r[r * 3 + 2];
            out[r * 4 + 3] = g0.p[r];
        }
        out[12] = 0; out[13] = 0; out[14] = 0; out[15] = 1;
    }
    SE3 carry = g0;
    for (int base = 1; base < count; base += 32) {
        const int k = base + lane;
        SE3 e;
        se3_identity(e);
        if (k < count) {
            const double si = (double)(start + k);
            double wA[3], wB[3];
            sect_omega(sm, c, qs, sect, si + cg1 - 1, wA);
            sect_omega(sm, c, qs, sect, si + cg2 - 1, wB);
            e = magnus_step(wA, wB, dx);
        }
        // inclusive scan of SE(3) products across lanes (Hillis-Steele)
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            SE3 o;
#pragma unroll
            for (int t = 0; t < 9; ++t) o.r[t] = shfl_up_d(e.r[t], off);
#pragma unroll
            for (int t = 0; t < 3; ++t) o.p[t] = shfl_up_d(e.p[t], off);
            if (lane >= off) e = se3_mul(o, e);
        }
        SE3 g = se3_mul(carry, e);
        if (k < count && out) {
            double* dst = out + (size_t)k * 16;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                dst[r * 4 + 0] = g.r[r * 3 + 0];
                dst[r * 4 + 1] = g.r[r * 3 + 1];
                dst[r * 4 + 2] = g.r[r * 3 + 2];
                dst[r * 4 + 3] = g.p[r];
            }
            dst[12] = 0; dst[13] = 0; dst[14] = 0; dst[15] = 1;
        }
        const int last = min(31, count - 1 - base);
#pragma unroll
        for (int t = 0; t < 9; ++t) carry.r[t] = shfl_d(g.r[t], last);
#pragma unroll
        for (int t = 0; t < 3; ++t) carry.p[t] = shfl_d(g.p[t], last);
    }
    return carry;
}

__device__ __forceinline__ SE3 se3_rx(double angle) {
    SE3 g;
    se3_identity(g);
    double s, c;
    sincos_ni(angle, &s, &c);
    g.r[4] = c; g.r[5] = -s; g.r[7] = s; g.r[8] = c;
    return g;
}

// integral of q.(1,x,x^2) over the ABSOLUTE interval integrate_g walks (static.py:192-195)
__device__ __forceinline__ double ipoly3_abs(const double* q, double a, double b) { return ipoly3(q, b) - ipoly3(q, a); }

// ------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(kSWarps * 32, 5)
    static_solve_kernel(StModel sm, long long B, const int32_t* __restrict__ idx, const double* __restrict__ theta,
                        const double* __restrict__ x0, int x0_per_config, double xtol, const long long* __restrict__ offsets,
                        double* __restrict__ g_out, double* __restrict__ sol_out, int32_t* __restrict__ nfev_out,
                        int32_t* __restrict__ info_out, unsigned long long* __restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    HybrdMem& M = reinterpret_cast<HybrdMem*>(smem_raw)[warp];
    const int T = sm.mc.T, n = sm.nq;
    for (;;) {
        // dynamic schedule: the evaluation count varies from ~30 to 200(n+1) between configurations
        unsigned long long tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1ULL);
        const long long b = (long long)__shfl_sync(0xffffffffu, tk, 0);
        if (b >= B) break;
        if (lane == 0) st_setup(sm, idx + b * T, theta + b * T, M.cfg);
        if (lane < n) M.x[lane] = x0 ? x0[(x0_per_config ? b * n : 0) + lane] : sm.x0[lane];
        __syncwarp();
        const StCfg& c = M.cfg;
        int nfev = 0;
        const int info = st_hybrd(sm, M, lane, xtol, 200 * (n + 1), 100.0, &nfev);
        __syncwarp();
        if (sol_out && lane < n) sol_out[b * n + lane] = M.x[lane];
        if (lane == 0) {
            if (nfev_out) nfev_out[b] = nfev;
            if (info_out) info_out[b] = info;
        }
        if (g_out) {
            const double* q = M.x;
            const double dx = sm.mc.dx;
            SE3 g11_last, g21_last, g31_last;
            se3_identity(g11_last);
            if (T == 3) {
                g11_last = integrate_section(sm, c, q + sm.qoff[0], 1, c.i11, c.L[0], se3_rx(c.th[0]),
                                             g_out + offsets[b * T + 0] * 16, lane);
                g21_last = se3_rx(c.th[1] + ipoly3_abs(q + sm.qoff[1], c.i21 * dx, (c.i21 + c.L[0] - 1) * dx));
                g31_last = se3_rx(c.th[2] + ipoly3_abs(q + sm.qoff[2], c.i31 * dx, (c.i31 + c.L[0] - 1) * dx));
            } else {
                g21_last = se3_rx(c.th[0]);  // static.py:114-118
                g31_last = se3_rx(c.th[1]);
            }
            SE3 g22_last = integrate_section(sm, c, q + sm.qoff[3], 2, c.i22, c.L[1], se3_mul(g11_last, g21_last),
                                             g_out + offsets[b * T + (T - 2)] * 16, lane);
            SE3 g32_last = se3_rx(ipoly3_abs(q + sm.qoff[4], c.i32 * dx, (c.i32 + c.L[1] - 1) * dx));
            integrate_section(sm, c, q + sm.qoff[5], 3, c.i33, c.L[2], se3_mul(se3_mul(g22_last, g31_last), g32_last),
                              g_out + offsets[b * T + (T - 1)] * 16, lane);
        }
        __syncwarp();
    }
}

// residual only (parity tests): one thread per (config) evaluates F(q)
__global__ void static_residual_kernel(StModel sm, long long B, const int32_t* __restrict__ idx,
                                       const double* __restrict__ theta, const double* __restrict__ q,
                                       double* __restrict__ F) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    StCfg c;
    st_setup(sm, idx + b * sm.mc.T, theta + b * sm.mc.T, c);
    double ql[kMaxNq], Fl[kMaxNq];
    for (int k = 0; k < sm.nq; ++k) ql[k] = q[b * sm.nq + k];
    st_residual(sm, c, ql, Fl);
    for (int k = 0; k < sm.nq; ++k) F[b * sm.nq + k] = Fl[k];
}

// ------------------------------------------------------------------ host side
int static_model_init(const ctrb_model_desc* d, const ModelConst& mc, StModel* out) {
    StModel sm;
    memset(&sm, 0, sizeof sm);
    sm.mc = mc;
    const int T = mc.T;
    if (T != 2 && T != 3) {
        set_error("%d is not supported by the static model", T);  // static.py:77-78
        return CTRB_ERR_INVALID;
    }
    if (d->static_degree != 2 || d->static_basis_type != 0) {
        set_error("static model: only basis_type 'last_strain_base' with degree 2 is supported (the reference's other "
                  "bases do not run, SURVEY App. B#2; _simple_basis hard-codes column offsets 3 and 6)");
        return CTRB_ERR_UNSUPPORTED;
    }
    const double pi = 3.141592653589793;
    for (int n = 0; n < T; ++n) {  // static.py:561-585
        const double o = d->radius_outer[n], i = d->radius_inner[n];
        const double polar = pi / 4 * (o * o * o * o - i * i * i * i);
        const double mom = polar * 2;
        sm.Kr[n][0] = 23.1e6 * mom;
        sm.Kr[n][1] = 60e6 * polar;
        sm.Kr[n][2] = 60e6 * polar;
    }
    const int da = 3;
    int nd[6] = {T == 3 ? 3 * da : 0, T == 3 ? da : 0, T == 3 ? da : 0, 3 * da, da, mc.dof[T - 1]};
    int o = 0;
    for (int k = 0; k < 6; ++k) {
        sm.ndof[k] = nd[k];
        sm.qoff[k] = o;
        o += nd[k];
    }
    sm.nq = o;
    if (o > kMaxNq) {
        set_error("static model: %d unknowns exceed the supported %d", o, kMaxNq);
        return CTRB_ERR_UNSUPPORTED;
    }
    const int first = o - nd[5];  // static_bases.py:142-168
    for (int k = 0; k < first; ++k) sm.x0[k] = (k % da == 0) ? 0.01 : 0.0;
    for (int k = 0; k < nd[5]; ++k) sm.x0[first + k] = mc.q[T - 1][k];
    *out = sm;
    return CTRB_OK;
}

int launch_static_solve(ctrb_ctx* ctx, const StModel& sm, long long B, const int32_t* idx, const double* theta,
                        const double* x0, int x0_per_config, double xtol, const long long* offsets, double* g_out,
                        double* sol_out, int32_t* nfev_out, int32_t* info_out) {
    if (B <= 0) return CTRB_OK;
    const size_t smem = sizeof(HybrdMem) * kSWarps;
    CTRB_CUDA(cudaFuncSetAttribute(static_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks_needed = (B + kSWarps - 1) / kSWarps;
    int per_sm = 1;
    CTRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, static_solve_kernel, kSWarps * 32, smem));
    int grid = (int)std::min<long long>(blocks_needed, (long long)ctx->sm_count * std::max(per_sm, 1));
    CTRB_CUDA(cudaMemsetAsync(ctx->d_counters + 5, 0, sizeof(unsigned long long), ctx->stream));
    {
        KernelTimer kt(ctx, CTRB_KERNEL_STATIC_SOLVE);
        static_solve_kernel<<<grid, kSWarps * 32, smem, ctx->stream>>>(sm, B, idx, theta, x0, x0_per_config,
                                                                      xtol > 0 ? xtol : 1.49012e-8, offsets, g_out, sol_out,
                                                                      nfev_out, info_out, ctx->d_counters + 5);
    }
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

int launch_static_residual(ctrb_ctx* ctx, const StModel& sm, long long B, const int32_t* idx, const double* theta,
                           const double* q, double* F) {
    if (B <= 0) return CTRB_OK;
    static_residual_kernel<<<(unsigned)((B + 63) / 64), 64, 0, ctx->stream>>>(sm, B, idx, theta, q, F);
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

}  // namespace ctrb
