// ctrb_static_common.cuh -- device code shared by the two static-model solve kernels:
//   ctrb_static.cu       MINPACK hybrd with its QR factors (the literal algorithm; fallback + degenerate sections)
//   ctrb_static_fast.cu  the same iteration carried on the explicit Jacobian and its inverse (Sherman-Morrison)
// per-configuration setup (static.py:81-95), the closed-form node integrands of the residual
// (static.py:228-454), calculate_integral's assembly (:456-491), warp norms and the curve integration
// (integrate_g, :165-198).
#pragma once
#include <float.h>
#include <math_constants.h>

#include "ctrb_se3.cuh"

namespace ctrb {
namespace {  // internal linkage: two translation units include this header

constexpr int kMaxNq = kStaticMaxNq;
// unroll factors of two loops, for the compile-time sweeps (tools/build_variant.sh)
#ifndef CTRB_FDJ_NORM_UNROLL
#define CTRB_FDJ_NORM_UNROLL 2
#endif
constexpr int kFdjNormUnroll = CTRB_FDJ_NORM_UNROLL;  // column-norm loop of warp_fdjac
#ifndef CTRB_NODE_OUT_UNROLL
#define CTRB_NODE_OUT_UNROLL 3
#endif
constexpr int kNodeOutUnroll = CTRB_NODE_OUT_UNROLL;  // node_sec12's loop over the three powers of x
#ifndef CTRB_WARP_SUM_UNROLL
#define CTRB_WARP_SUM_UNROLL 5
#endif
constexpr int kWarpSumUnroll = CTRB_WARP_SUM_UNROLL;  // the butterfly of warp_sum
#ifndef CTRB_ASM_UNROLL
#define CTRB_ASM_UNROLL 2
#endif
constexpr int kAsmUnroll = CTRB_ASM_UNROLL;  // asm_part's loop over the two Gauss abscissae

// The model constants as every device function below reads them: a block-wide shared-memory copy that each kernel
// fills first (load_model).  Through the kernel parameter (a generic pointer once it crosses an out-of-line call)
// every field access was a generic LD behind two R2UR descriptor moves: 22 % of node_sec12's instructions.
__shared__ StModel g_sm;
#define CTRB_USE_SM      \
    (void)sm_;           \
    const StModel& sm = g_sm
// Pointers into the warp's shared-memory workspace lose their address space when they cross an out-of-line call
// inside a struct; the hint turns the generic LD / ST (+ descriptor moves) back into LDS / STS.
#define CTRB_IN_SHARED(p) __builtin_assume(__isShared(p))
__device__ __forceinline__ void load_model(const StModel& src) {
    const int* s = reinterpret_cast<const int*>(&src);
    int* d = reinterpret_cast<int*>(&g_sm);
    for (int i = threadIdx.x; i < (int)(sizeof(StModel) / sizeof(int)); i += blockDim.x) d[i] = s[i];
    __syncthreads();
}

struct StCfg {
    int i11, i21, i31, i22, i32, i33;
    int L[3];  // nodes of sections 1..3
    double th[3];
    // calculate_integral's two abscissae per section: bracketing nodes and interpolation terms
    int lo[3][2], hi[3][2];
    double wq[3][2], len[3];  // (x_new - x_lo) / (x_hi - x_lo) of interp1d at the two abscissae; section length
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

// one out-of-line copy of the fp64 sincos: the solve kernels are instruction-cache bound if it is inlined at
// every use (ncu: stall_no_instruction is the top stall); results come back in registers
struct SinCos {
    double s, c;
};
__device__ __noinline__ SinCos sincos_ni(double a) {
    SinCos r;
    sincos(a, &r.s, &r.c);
    return r;
}
// two independent angles in one call: the two dependency chains interleave (the kernels are latency-bound)
struct SinCos2 {
    double s0, c0, s1, c1;
};
#ifndef CTRB_SINCOS_PAIR
#define CTRB_SINCOS_PAIR 1
#endif
#if CTRB_SINCOS_PAIR
__device__ __noinline__ SinCos2 sincos2_ni(double a, double b) {
    SinCos2 r;
    sincos(a, &r.s0, &r.c0);
    sincos(b, &r.s1, &r.c1);
    return r;
}
#else
__device__ __forceinline__ SinCos2 sincos2_ni(double a, double b) {  // two calls of the one out-of-line copy: half the code
    const SinCos x = sincos_ni(a), y = sincos_ni(b);
    return SinCos2{x.s, x.c, y.s, y.c};
}
#endif

// one non-zero per column of a strain base (strain_bases.py:7-224): its row (0..2) and its value at x
__host__ __device__ __forceinline__ void strain_base_col(int kind, int dof, int k, double x, int* row, double* val) {
    int r = 1;
    double v = 1.0;
    switch (kind) {
        case CTRB_BASE_CONSTANT: break;
        case CTRB_BASE_LINEAR:
            if (k == dof - 1) { r = 2; v = x; } else if (k == 1) { v = x; }
            break;
        case CTRB_BASE_QUADRATIC:
            if (k == dof - 1) { r = 2; v = x * x; } else if (k == 1) { v = x * x; }
            break;
        case CTRB_BASE_PURE_HELIX:
            if (k == 0) { v = sin(x / 10); } else { r = 2; v = cos(x / 10); }
            break;
        case CTRB_BASE_LINEAR_HELIX:
            if (k == 0) { v = sin(x / 10); } else { r = 2; v = x * cos(x / 10); }
            break;
        case CTRB_BASE_TORSION_HELIX: r = k; break;
        case CTRB_BASE_TORSION_LINEAR_HELIX:
            r = k;
            if (k == 2) v = x;
            break;
        case CTRB_BASE_FULL: {
            const int bs = (dof & 1) ? 1 : 2;  // torsion columns
            const int nb = (dof - bs) / 2;     // columns per bending direction (2 or 3)
            if (k < bs) {
                r = 0;
                v = k == 0 ? 1.0 : x;
            } else {
                const int kk = k - bs;
                r = kk < nb ? 1 : 2;
                const int pw = kk < nb ? kk : kk - nb;
                v = pw == 0 ? 1.0 : (pw == 1 ? x : x * x);
            }
            break;
        }
        default: break;
    }
    *row = r;
    *val = v;
}

// ------------------------------------------------------------------ per-configuration setup
// static.py:456-491 quadrature brackets of a section of np nodes (runs once per design and np: static_tables_kernel)
__device__ __noinline__ void st_quad(double dx, int np, StQuad& e) {
    const double cs[2] = {(1.0 / sqrt(3.0) + 1.0) / 2.0, (-1.0 / sqrt(3.0) + 1.0) / 2.0};  // static.py:601-603
    const double length = (np - 1) * dx;
    e.len = length;
    for (int a = 0; a < 2; ++a) {
        e.lo[a] = e.hi[a] = 0;
        e.wq[a] = 0.0;
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
        e.lo[a] = lo;
        e.hi[a] = hi;
        e.wq[a] = (xn - xlo) / (xhi - xlo);
    }
}

// static.py:81-95 section indices; the quadrature brackets come from the per-design table (the formula, with its
// fifteen divisions, was 700 instructions streamed through the instruction cache once per configuration)
__device__ __noinline__ void st_setup(const StModel& sm_, const int* idx, const double* theta, StCfg& c) {
    CTRB_USE_SM;
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
#pragma unroll 1
    for (int s = 0; s < 3; ++s) {
        const int np = c.L[s];
        if (np >= 0 && np <= sm.quad_max) {
            const StQuad e = sm.quad_tab[np];
            c.len[s] = e.len;
            c.lo[s][0] = e.lo[0]; c.lo[s][1] = e.lo[1];
            c.hi[s][0] = e.hi[0]; c.hi[s][1] = e.hi[1];
            c.wq[s][0] = e.wq[0]; c.wq[s][1] = e.wq[1];
        } else {  // indices outside the tubes (rejected by the API's validation before any launch)
            c.len[s] = 0.0;
            c.lo[s][0] = c.lo[s][1] = c.hi[s][0] = c.hi[s][1] = 0;
            c.wq[s][0] = c.wq[s][1] = 0.0;
        }
    }
}

// ------------------------------------------------------------------ integrands at one node
// The unknown vector as the node functions read it: shared-memory vector q with, optionally, component jp
// displaced by hp -- the forward-difference Jacobian's x + h e_j without materialising a second vector
// (q[k] + 0.0 is exact, so unperturbed components are bit-identical to a plain read).
struct QV {
    const double* q;
    int jp;
    double hp;
    __device__ __forceinline__ double operator()(int k) const { return k == jp ? q[k] + hp : q[k]; }
};
__device__ __forceinline__ double poly3v(QV q, int o, double x) { return q(o) + q(o + 1) * x + q(o + 2) * (x * x); }
__device__ __forceinline__ double ipoly3v(QV q, int o, double X) {
    return q(o) * X + q(o + 1) * (0.5 * X * X) + q(o + 2) * (X * X * X * (1.0 / 3.0));
}
// section 1 (static.py:261-345), three tubes; y = [intc1 (9), intg21 (3), intg31 (3)]
// omega of tube n's own pre-curvature at absolute node `node` (get_ksi_star, static.py:618-634): per-design table,
// or the formula when a derived section index runs outside the tube (negative idx21 / idx31 / idx32)
__device__ __noinline__ V3 ksi_star_cold(const ModelConst& mc, int n, double x) {
    double w[3];
    strain_omega(mc, n, x, w);
    return V3{w[0], w[1], w[2]};
}
__device__ __forceinline__ V3 ksi_star_w(const StModel& sm_, int n, int node) {
    CTRB_USE_SM;
    if (node >= 0 && node < sm.mc.npts[n]) {
        const double* ws = sm.ws_tab + 3 * (sm.mc.node_base[n] + node);
        return V3{ws[0], ws[1], ws[2]};
    }
    return ksi_star_cold(sm.mc, n, node * sm.mc.dx);
}

// Sections 1 and 2 (static.py:261-345 and 347-436) through ONE routine, so that the lanes holding section-1 and
// section-2 nodes do not diverge.  Section 1: outer tube with curvature polynomials w, then two nested relative
// x-rotations Rx(aM), Rx(aI) of the middle and inner tubes, each adding its torsion polynomial.  Section 2 is the
// same chain without a middle tube (zero stiffness, no torsion) whose rotation is the relative frame gg31 reached
// at the end of section 1 (equilibrium_three's return value; identity for two tubes).
//   sec 0: y = [intc1 (9), intg21 (3), intg31 (3)]      sec 1: y = [intc2 (9), intg32 (3), intg312 (3)]
template <bool kShared = true>  // kShared: c, q and y live in shared memory (all callers but the scalar parity kernel)
__device__ __noinline__ void node_sec12(const StModel& sm_, const StCfg& c, QV q, int sec, int i, double y[15]) {
    CTRB_USE_SM;
    if (kShared) {
        CTRB_IN_SHARED(&c);
        CTRB_IN_SHARED(q.q);
        CTRB_IN_SHARED(y);
    }
    const int T = sm.mc.T;
    const bool s2 = sec == 1;
    const double dx = sm.mc.dx, X = i * dx;
    const double Xe = T == 3 ? (c.L[0] - 1) * dx : 0.0;
    const int qo = sm.qoff[s2 ? 3 : 0], qm = sm.qoff[s2 ? 2 : 1], qi = sm.qoff[s2 ? 4 : 2];
    const int nO = s2 ? T - 2 : 0, nI = s2 ? T - 1 : 2;
    V3 w1 = {poly3v(q, qo, X), poly3v(q, qo + 3, X), poly3v(q, qo + 6, X)};
    const double tgM = s2 ? 0.0 : poly3v(q, qm, X), tgI = poly3v(q, qi, X);
    const bool mid_identity = s2 && T == 2;
    const double aM = mid_identity ? 0.0 : (s2 ? c.th[2] : c.th[1]) + ipoly3v(q, qm, s2 ? Xe : X);
    const double aI = (s2 ? (T == 2 ? c.th[1] : 0.0) : c.th[2]) + ipoly3v(q, qi, X);
    // M: gg21 = Rx(theta1 + int tau21) | gg31 at the end of section 1;   I: gg31 | gg32
    const SinCos2 rr = sincos2_ni(aM, aI);
    const double sM = rr.s0, cM = rr.c0, sI = rr.s1, cI = rr.c1;
    V3 wM = rotxT(sM, cM, w1);
    wM.x += tgM;
    V3 wI = rotxT(sI, cI, wM);
    wI.x += tgI;
    V3 ws = ksi_star_w(sm, nO, i + (s2 ? c.i22 : c.i11));
    V3 mO = {sm.Kr[nO][0] * (w1.x - ws.x), sm.Kr[nO][1] * (w1.y - ws.y), sm.Kr[nO][2] * (w1.z - ws.z)};
    V3 mM = {0.0, 0.0, 0.0};
    if (!s2) {
        ws = ksi_star_w(sm, 1, i + c.i21);
        mM = V3{sm.Kr[1][0] * (wM.x - ws.x), sm.Kr[1][1] * (wM.y - ws.y), sm.Kr[1][2] * (wM.z - ws.z)};
    }
    ws = ksi_star_w(sm, nI, i + (s2 ? c.i32 : c.i31));
    V3 mI = {sm.Kr[nI][0] * (wI.x - ws.x), sm.Kr[nI][1] * (wI.y - ws.y), sm.Kr[nI][2] * (wI.z - ws.z)};
    V3 rIm = rotx(sI, cI, mI);                      // coAd(gg_I) f_I (moment part)
    V3 u = {mM.x + rIm.x, mM.y + rIm.y, mM.z + rIm.z};
    V3 ru = rotx(sM, cM, u);                        // coAd(gg_M) (f_M + coAd(gg_I) f_I)
    V3 wsum = {mO.x + ru.x, mO.y + ru.y, mO.z + ru.z};
    V3 zM = cross3(w1, ru);                         // coad(ksi_O) (...)
    V3 zI = cross3(w1, rotx(sM, cM, rIm));
    const double P[3] = {1.0, X, X * X};
    const double S[3] = {X, 0.5 * X * X, X * X * X * (1.0 / 3.0)};            // sg (row 0) = int of the basis
    const double Se[3] = {Xe, 0.5 * Xe * Xe, Xe * Xe * Xe * (1.0 / 3.0)};     // sg31 at the end of section 1
#pragma unroll(kNodeOutUnroll)
    for (int k = 0; k < 3; ++k) {
        y[k] = P[k] * wsum.x;
        y[3 + k] = P[k] * wsum.y;
        y[6 + k] = P[k] * wsum.z;
        y[9 + k] = P[k] * (mM.x + mI.x) - S[k] * zM.x;
        y[12 + k] = s2 ? Se[k] * zI.x : P[k] * mI.x - S[k] * zI.x;
    }
}

// section 3 (static.py:438-452); y = intc3 (dof of the last tube): B^T K (B q33 - B q*) with the base's single
// non-zero per column taken from the per-design table
template <bool kShared = true>
__device__ __noinline__ void node_sec3(const StModel& sm_, const StCfg& c, QV q, int i, double y[CTRB_MAX_DOF]) {
    CTRB_USE_SM;
    if (kShared) {
        CTRB_IN_SHARED(&c);
        CTRB_IN_SHARED(q.q);
        CTRB_IN_SHARED(y);
    }
    const int T = sm.mc.T, dof = sm.mc.dof[T - 1];
    const int q33 = sm.qoff[5];
    const double* bt = sm.b3_tab + (CTRB_MAX_DOF + 3) * (i + c.i33);
    // B q33 and B q* are accumulated by the same explicitly rounded operations: with q33 == q* (the initial
    // guess, static_bases.py:164-166) the difference must be exactly zero, as it is in the reference
    double a[3] = {0.0, 0.0, 0.0}, b[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
    for (int k = 0; k < dof; ++k) {
        const double ta = __dmul_rn(bt[k], q(q33 + k)), tb = __dmul_rn(bt[k], sm.mc.q[T - 1][k]);
        const int r = sm.b3_row[k];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
            a[rr] = __dadd_rn(a[rr], r == rr ? ta : 0.0);
            b[rr] = __dadd_rn(b[rr], r == rr ? tb : 0.0);
        }
    }
    const double m0 = sm.Kr[T - 1][0] * (a[0] - b[0]);
    const double m1 = sm.Kr[T - 1][1] * (a[1] - b[1]);
    const double m2 = sm.Kr[T - 1][2] * (a[2] - b[2]);
#pragma unroll 1
    for (int k = 0; k < dof; ++k) {
        const int r = sm.b3_row[k];
        y[k] = bt[k] * (r == 0 ? m0 : (r == 1 ? m1 : m2));
    }
}

__device__ __noinline__ double warp_sum(double v) {
#pragma unroll(kWarpSumUnroll)
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// four sums at once: the same butterfly per value (bit-identical to four warp_sum calls), one dependent chain instead of four
struct D4 {
    double a, b, c, d;
};
__device__ __noinline__ D4 warp_sum4(double a, double b, double c, double d) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
        c += __shfl_xor_sync(0xffffffffu, c, off);
        d += __shfl_xor_sync(0xffffffffu, d, off);
    }
    return D4{a, b, c, d};
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

// Cooperative residual, phase 1: work item t in 0..11 = (section, abscissa, bracket end) evaluates one node
// integrand vector into y (shared memory).  q is a shared-memory vector of the unknowns.
__device__ __noinline__ void res_node(const StModel& sm_, const StCfg& c, QV q, int t, double* y) {
    CTRB_USE_SM;
    CTRB_IN_SHARED(&c);
    CTRB_IN_SHARED(q.q);
    CTRB_IN_SHARED(y);
    const int T = sm.mc.T;
    const int sec = t >> 2, a = (t >> 1) & 1, e = t & 1;
    if (c.len[sec] > 0.0 && (sec != 0 || T == 3)) {
        const int node = e ? c.hi[sec][a] : c.lo[sec][a];
        if (sec < 2) {
            node_sec12(sm, c, q, sec, node, y);
        } else {
            node_sec3(sm, c, q, node, y);
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < 16; ++k) y[k] = 0.0;
    }
}

// phase 2: component k of F from the 12 node vectors (calculate_integral, static.py:456-491)
__device__ __noinline__ double res_assemble(const StModel& sm_, const StCfg& c, const double (*yb)[16], int comp) {
    CTRB_USE_SM;
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
            acc += 0.5 * ((yhi - ylo) * c.wq[sec][a] + ylo);
        }
    }
    double val = (c.len[sec] / 2) * acc;
    if (blk == 2) {  // intg_31 - intg_312: the section-2 share lives at offset 12 of the section-2 vectors
        double acc2 = 0.0;
        if (c.len[1] > 0.0) {
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const double ylo = yb[4 + a * 2 + 0][12 + k], yhi = yb[4 + a * 2 + 1][12 + k];
                acc2 += 0.5 * ((yhi - ylo) * c.wq[1][a] + ylo);
            }
        }
        val -= (c.len[1] / 2) * acc2;
    }
    return val;
}

constexpr int kJLD = 31;  // leading dimension of the column-major n x n matrices (odd: conflict-free row and column access)

// per-lane description of residual row `lane` (which block of static.py:63 it belongs to); `scale` multiplies the row
// (1 everywhere except under the one-step-section rule, see pin_mask)
struct RowMeta {
    int blk, k, sec, yo;
    double scale;
};

__device__ __forceinline__ RowMeta row_meta(const StModel& sm_, int comp) {
    CTRB_USE_SM;
    RowMeta r;
    r.blk = 0;
#pragma unroll
    for (int b = 1; b < 6; ++b)
        if (comp >= sm.qoff[b] && sm.ndof[b] > 0) r.blk = b;
    r.k = comp - sm.qoff[r.blk];
    r.sec = r.blk < 3 ? 0 : (r.blk < 5 ? 1 : 2);
    r.yo = (r.blk == 0 || r.blk == 3 || r.blk == 5) ? 0 : ((r.blk == 1 || r.blk == 4) ? 9 : 12);
    r.scale = 1.0;
    return r;
}

// One-step sections (two nodes).  calculate_integral (static.py:456-491) samples a section's integrand at
// its nodes only, and a one-step section has the nodes x = 0 and x = dx: the curvature polynomial c0 + c1 x + c2 x^2 of
// its tube==section block enters the residual through c0 and c1 dx + c2 dx^2 alone, so the Jacobian columns of c1 and
// c2 are proportional (identical for dx = 1), and so are the residual rows that weigh the integrand by x and x^2: the
// system is consistent and rank-deficient by three per one-step section.  MINPACK's qrfac meets an (exactly or nearly)
// zero diagonal of R there, and what dogleg's back substitution makes of it -- 0/0 in exact arithmetic, rounding noise
// over rounding noise in floating point -- decides where in the null space hybrd lands (the reference's own result changes
// under a 1e-13 perturbation of its initial guess in a fifth of such configurations).  The rule used here is hybrd on
// the reduced system, i.e. what MINPACK's rank-deficient rule gives when the zero is exact: the step component of the
// coincident column is zero, so c2 keeps its initial value, and the duplicated row is carried once with the weight
// sqrt(1 + dx^2) that keeps |F|, J^T F and the column norms what they are in the full system.  Implemented without changing
// the system size: pinned unknown j has column e_j and row e_j^T in the Jacobian and a zero residual component (the
// Broyden updates preserve that structure exactly: u_j = 0, v_j = 0), row j - 1 carries the weight.
// Bit j of the mask = unknown j is pinned.
__device__ __forceinline__ unsigned pin_mask(const StModel& sm_, const StCfg& c) {
    CTRB_USE_SM;
    unsigned m = 0;
    const unsigned x2 = (1u << 2) | (1u << 5) | (1u << 8);  // the x^2 coefficients of the three curvature components
    if (sm.mc.T == 3 && c.L[0] == 2) m |= x2 << sm.qoff[0];
    if (c.L[1] == 2) m |= x2 << sm.qoff[3];
    return m;
}
__device__ __forceinline__ void pin_row_meta(const StModel& sm_, unsigned pin, int lane, RowMeta& rm) {
    CTRB_USE_SM;
    if ((pin >> lane) & 1u) rm.scale = 0.0;
    else if (lane < 31 && ((pin >> (lane + 1)) & 1u)) rm.scale = sqrt(1.0 + sm.mc.dx * sm.mc.dx);
}

// calculate_integral (static.py:456-491) of one integrand component over one section: the four node values are
// yb[base + 2 a + e][off] (a = Gauss abscissa, e = lower / upper bracketing node)
__device__ __forceinline__ double asm_part(const StCfg& c, const double (*yb)[16], int base, int sec, int off) {
    if (!(c.len[sec] > 0.0)) return 0.0;
    double acc = 0.0;
#pragma unroll(kAsmUnroll)
    for (int a = 0; a < 2; ++a) {
        const double ylo = yb[base + a * 2 + 0][off], yhi = yb[base + a * 2 + 1][off];
        acc += 0.5 * ((yhi - ylo) * c.wq[sec][a] + ylo);
    }
    return (c.len[sec] / 2) * acc;
}

// F(q) by the whole warp: 12 node integrands (3 sections x 2 abscissae x 2 bracketing nodes), then one residual
// component per lane.  Returns this lane's component; p1 / p2 are its section shares (row block (3,1) is
// intg_31 - intg_312, static.py:250-252; p2 = 0 elsewhere).
// `nnodes` = 12, or 8 when the last block is carried implicitly (see block5_norm2): its nodes are not needed then.
__device__ __forceinline__ double warp_residual(const StModel& sm_, const StCfg& cfg, double (*y)[16], const double* q,
                                                const RowMeta& rm, bool act, int lane, int nnodes, double* p1, double* p2) {
    CTRB_USE_SM;
    if (lane < nnodes) res_node(sm, cfg, QV{q, -1, 0.0}, lane, y[lane]);
    __syncwarp();
    double a = 0.0, b = 0.0;
    if (act) {
        a = rm.scale * asm_part(cfg, y, rm.sec * 4, rm.sec, rm.yo + rm.k);  // scale: 1.0 (exact) unless the row is pinned / weighted
        if (rm.blk == 2) b = asm_part(cfg, y, 4, 1, 12 + rm.k);
    }
    __syncwarp();
    *p1 = a;
    *p2 = b;
    return rm.blk == 2 ? a - b : a;
}

// column-slot table of the forward-difference Jacobian: slot = (unknown j, section whose 4 nodes are re-evaluated)
struct SlotTable {
    signed char j[48], sec[48], col_blk[32];
    int nslots;
};

// thread 0 of the block fills the table (it only depends on the design); `reduced` leaves the last block out
__device__ __forceinline__ void build_slot_table(const StModel& sm_, SlotTable& st, bool reduced) {
    CTRB_USE_SM;
    int ns = 0;
    for (int b = 0; b < 6; ++b)
        for (int k = 0; k < sm.ndof[b]; ++k) st.col_blk[sm.qoff[b] + k] = (signed char)b;
    // the two-slot columns first, so that a pair never straddles a round of 8 slots
    for (int k = 0; k < sm.ndof[2]; ++k)
        for (int s = 0; s < 2; ++s) {
            st.j[ns] = (signed char)(sm.qoff[2] + k);
            st.sec[ns++] = (signed char)s;
        }
    for (int b = 0; b < (reduced ? 5 : 6); ++b) {
        if (b == 2) continue;
        for (int k = 0; k < sm.ndof[b]; ++k) {
            st.j[ns] = (signed char)(sm.qoff[b] + k);
            st.sec[ns++] = (signed char)(b < 3 ? 0 : (b < 5 ? 1 : 2));
        }
    }
    st.nslots = ns;
}

// The last block (3,3) of the system is linear and decoupled: F_33 = int B^T K B (q_33 - q*) depends on q_33 only and
// nothing else depends on q_33 (static.py:438-452).  From the default initial guess q_33 = q* (static_bases.py:164-166)
// F_33 is exactly zero, the Jacobian is exactly block diagonal, and hybrd never moves q_33: its Gauss-Newton and
// gradient components vanish, Broyden's update vectors have zeros there, the QR factors stay block diagonal.  The
// ONLY place the block enters the iteration is the scaled norm |D x| (trust-region start, the xtol test), through the
// constant  sum_j (D_j q*_j)^2  with D_j = the forward-difference column norms of the block.  So the solvers iterate on
// the first n - dof_33 unknowns and add this constant to |D x|^2: the same iterates with 27 instead of 30 unknowns.
// Returns that constant (all lanes); yb is scratch for 32 node vectors.
__device__ __noinline__ double block5_norm2(const StModel& sm_, const StCfg& cfg, const double* x, double (*yb)[16], int lane) {
    CTRB_USE_SM;
    const int dof = sm.ndof[5], q0 = sm.qoff[5];
    const double eps = sqrt(DBL_EPSILON);
    const int slot = lane >> 2;
    if (slot < dof) {
        const double xv = x[q0 + slot];
        double h = eps * fabs(xv);
        if (h == 0.0) h = eps;
        res_node(sm, cfg, QV{x, q0 + slot, h}, 8 + (lane & 3), yb[lane]);
    }
    __syncwarp();
    double c2 = 0.0;
#pragma unroll 1
    for (int k = 0; k < dof; ++k) {
        const double xv = x[q0 + k];
        double h = eps * fabs(xv);
        if (h == 0.0) h = eps;
        const double col = lane < dof ? asm_part(cfg, yb, k * 4, 2, lane) / h : 0.0;  // (F_33(x + h e_k) - 0) / h, row `lane`
        double d = sqrt(warp_sum(col * col));
        if (d == 0.0) d = 1.0;
        c2 = fma(d * xv, d * xv, c2);
    }
    __syncwarp();
    return c2;
}

// forward-difference Jacobian (MINPACK fdjac1, dense band) using the block structure; f_l / p1b / p2b are this
// lane's residual component at x and its section shares.  Writes J (column-major, leading dimension kJLD); yb is
// scratch for 32 node vectors; returns this lane's column norm.
__device__ __noinline__ double warp_fdjac(const StModel& sm_, const StCfg& cfg, const SlotTable& st, const RowMeta& rm,
                                          const double* x, double* J, double (*yb)[16], double f_l, double p1b, double p2b,
                                          int n, int lane, unsigned pin = 0u) {
    CTRB_USE_SM;
    const bool act = lane < n;
    const double eps = sqrt(DBL_EPSILON);
    // The quotient (F(x + h e_j) - F(x)) / h is taken as a product with 1 / h: one division per COLUMN (by the lane that owns
    // it, broadcast below) instead of one per entry, at most one ulp away from the quotient.  The block structure makes most
    // numerators exactly zero, and a zero numerator sends the IEEE division through its slow path.
    double rh_l = 0.0;
    {
        double h = eps * fabs(act ? x[lane] : 0.0);
        if (h == 0.0) h = eps;
        rh_l = 1.0 / h;
    }
    for (int base = 0; base < st.nslots; base += 8) {
        {
            const int slot = base + (lane >> 2);
            if (slot < st.nslots) {
                const int j = st.j[slot];
                const double xv = x[j];
                double h = eps * fabs(xv);
                if (h == 0.0) h = eps;
                res_node(sm, cfg, QV{x, j, h}, st.sec[slot] * 4 + (lane & 3), yb[lane]);
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int s8 = 0; s8 < 8; ++s8) {
            const int slot = base + s8;
            if (slot >= st.nslots) break;
            const int j = st.j[slot], sec = st.sec[slot];
            const bool pair = st.col_blk[j] == 2;  // (3,1) torsion: sections 1 and 2 both move; slots (s, s + 1)
            if (pair && sec == 1) continue;
            const double rh = __shfl_sync(0xffffffffu, rh_l, j);
            if (act) {
                const bool hit1 = rm.sec == sec || (pair && rm.sec == 1);
                const bool hit2 = rm.blk == 2 && (sec == 1 || pair);
                double val = 0.0;
                if (hit1 || hit2) {
                    const int b1 = (pair && rm.sec == 1) ? (s8 + 1) * 4 : s8 * 4;
                    const double p1 = hit1 ? rm.scale * asm_part(cfg, yb, b1, rm.sec, rm.yo + rm.k) : p1b;
                    double Fp = p1;
                    if (rm.blk == 2) {
                        const int b2 = pair ? (s8 + 1) * 4 : s8 * 4;
                        const double p2 = hit2 ? asm_part(cfg, yb, b2, 1, 12 + rm.k) : p2b;
                        Fp = p1 - p2;
                    }
                    val = (Fp - f_l) * rh;
                }
                if ((pin >> j) & 1u) val = lane == j ? 1.0 : 0.0;  // pinned unknown: column e_j (its row is zero: scale 0)
                J[lane + j * kJLD] = val;
            }
        }
        __syncwarp();
    }
    double s2 = 0.0;
    if (act) {
        const double* m = J + lane * kJLD;
#pragma unroll(kFdjNormUnroll)
        for (int i = 0; i < n; ++i) s2 = fma(m[i], m[i], s2);
    }
    return sqrt(s2);
}

// ------------------------------------------------------------------ curve integration (integrate_g)
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

// half_step_integral for a tube==section basis (with bias): the Magnus exponential of one step.
// wA, wB = omega at the two Gauss points; v = e_x at both.
__device__ __forceinline__ SE3 magnus_step_body(const double wA[3], const double wB[3], double dx) {
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
// out of line for the warp-per-configuration scan (its code size matters there), inline for the thread-per-configuration
// kernel (a call passes the 12 doubles of the result through the stack: 2.92 -> 2.63 ms per 2^22 configurations)
__device__ __noinline__ SE3 magnus_step(const double wA[3], const double wB[3], double dx) { return magnus_step_body(wA, wB, dx); }
__device__ __forceinline__ SE3 magnus_step_inl(const double wA[3], const double wB[3], double dx) { return magnus_step_body(wA, wB, dx); }

// omega of the (tube==section) basis at absolute section coordinate `sidx` (static_bases.py:61-81)
__device__ __noinline__ void sect_omega(const StModel& sm_, const StCfg& c, const double* qs, int sect, double sidx,
                                           double w[3]) {
    CTRB_USE_SM;
    const double dx = sm.mc.dx;
    if (sect == 3) {
        const int T = sm.mc.T, dof = sm.mc.dof[T - 1];
        const double x = (sidx + c.i33) * dx;  // _easy_last_basis adds idx33 (again)
        w[0] = w[1] = w[2] = 0.0;
#pragma unroll 1
        for (int k = 0; k < dof; ++k) {
            int r;
            double v;
            strain_base_col(sm.mc.kind[T - 1], dof, k, x, &r, &v);
            const double t = v * qs[k];
            w[0] += r == 0 ? t : 0.0;
            w[1] += r == 1 ? t : 0.0;
            w[2] += r == 2 ? t : 0.0;
        }
    } else {
        const double x = sidx * dx;
        w[0] = poly3(qs, x);
        w[1] = poly3(qs + 3, x);
        w[2] = poly3(qs + 6, x);
    }
}

// one curve node: a 4x4 row-major matrix (16 doubles), or -- kPos, for the fused solve -> collision path, which only
// reads positions -- its position (3 doubles)
template <bool kPos>
__device__ __forceinline__ void store_node(double* out, size_t k, const SE3& g) {
    if (kPos) {
        double* dst = out + k * 3;
        dst[0] = g.p[0]; dst[1] = g.p[1]; dst[2] = g.p[2];
    } else {
        double* dst = out + k * 16;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            dst[r * 4 + 0] = g.r[r * 3 + 0];
            dst[r * 4 + 1] = g.r[r * 3 + 1];
            dst[r * 4 + 2] = g.r[r * 3 + 2];
            dst[r * 4 + 3] = g.p[r];
        }
        dst[12] = 0; dst[13] = 0; dst[14] = 0; dst[15] = 1;
    }
}

// Writes `count` nodes of one section curve: node 0 = g0, node k = g0 * E_1 ... E_k with
// E_k the Magnus step at absolute index start + k.  Returns the last node (all lanes).
template <bool kPos>
__device__ __noinline__ SE3 integrate_section(const StModel& sm_, const StCfg& c, const double* qs, int sect, int start, int count,
                                 SE3 g0, double* out, int lane) {
    CTRB_USE_SM;
    const double dx = sm.mc.dx;
    const double cg1 = 0.5 - sqrt(3.0) / 6, cg2 = 0.5 + sqrt(3.0) / 6;  // static.py:597-599
    if (lane == 0 && out) store_node<kPos>(out, 0, g0);
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
        // inclusive scan of SE(3) products across lanes (Hillis-Steele); only the first m lanes hold steps, and a round
        // with off >= m changes none of them
        const int m = min(32, count - base);
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            if (off >= m) break;
            SE3 o;
#pragma unroll
            for (int t = 0; t < 9; ++t) o.r[t] = shfl_up_d(e.r[t], off);
#pragma unroll
            for (int t = 0; t < 3; ++t) o.p[t] = shfl_up_d(e.p[t], off);
            if (lane >= off) e = se3_mul(o, e);
        }
        SE3 g = se3_mul(carry, e);
        if (k < count && out) store_node<kPos>(out, (size_t)k, g);
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
    const SinCos r = sincos_ni(angle);
    const double s = r.s, c = r.c;
    g.r[4] = c; g.r[5] = -s; g.r[7] = s; g.r[8] = c;
    return g;
}

// integral of q.(1,x,x^2) over the ABSOLUTE interval integrate_g walks (static.py:192-195)
__device__ __forceinline__ double ipoly3_abs(const double* q, double a, double b) { return ipoly3(q, b) - ipoly3(q, a); }

// Static.solve_g's curve assembly (static.py:110-130) from a solved q (shared-memory vector): the whole warp
// writes the T section curves of configuration b into the packed g_out.
template <bool kPos>
__device__ __noinline__ void st_write_curves(const StModel& sm_, const StCfg& c, const double* q, const long long* offsets,
                                             long long b, double* g_out, int lane) {
    constexpr int kNode = kPos ? 3 : 16;  // doubles per node
    CTRB_USE_SM;
    const int T = sm.mc.T;
    const double dx = sm.mc.dx;
    SE3 g11_last, g21_last, g31_last;
    se3_identity(g11_last);
    if (T == 3) {
        g11_last = integrate_section<kPos>(sm, c, q + sm.qoff[0], 1, c.i11, c.L[0], se3_rx(c.th[0]),
                                           g_out + offsets[b * T + 0] * kNode, lane);
        g21_last = se3_rx(c.th[1] + ipoly3_abs(q + sm.qoff[1], c.i21 * dx, (c.i21 + c.L[0] - 1) * dx));
        g31_last = se3_rx(c.th[2] + ipoly3_abs(q + sm.qoff[2], c.i31 * dx, (c.i31 + c.L[0] - 1) * dx));
    } else {
        g21_last = se3_rx(c.th[0]);  // static.py:114-118
        g31_last = se3_rx(c.th[1]);
    }
    SE3 g22_last = integrate_section<kPos>(sm, c, q + sm.qoff[3], 2, c.i22, c.L[1], se3_mul(g11_last, g21_last),
                                           g_out + offsets[b * T + (T - 2)] * kNode, lane);
    SE3 g32_last = se3_rx(ipoly3_abs(q + sm.qoff[4], c.i32 * dx, (c.i32 + c.L[1] - 1) * dx));
    integrate_section<kPos>(sm, c, q + sm.qoff[5], 3, c.i33, c.L[2], se3_mul(se3_mul(g22_last, g31_last), g32_last),
                            g_out + offsets[b * T + (T - 1)] * kNode, lane);
}

}  // namespace
}  // namespace ctrb
