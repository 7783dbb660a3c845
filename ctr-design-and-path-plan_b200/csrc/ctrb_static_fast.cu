// ctrb_static_fast.cu -- Static.solve_g's root find (static.py:102, scipy.optimize.root 'hybr') carried on the
// explicit Jacobian and its inverse instead of MINPACK's QR factors.  One warp per configuration, sm_100a.
//
// hybrd's iteration only ever uses its factors J = Q R through four products (dogleg, hybrd, More/Garbow/Hillstrom):
//     Gauss-Newton step  R^-1 Q^T f      = J^-1 f
//     scaled gradient    R^T Q^T f / D   = J^T f / D
//     |R g|, |Q^T f + R p|               = |J g|, |f + J p|       (Q is orthogonal)
// and updates them with Broyden's rank-one formula J+ = J + u v^T.  Holding J and H = J^-1 explicitly and
// updating H by Sherman-Morrison,  H+ = H - (H u)(v^T H) / (1 + v^T H u),  gives the same iterates in exact
// arithmetic, and every step becomes a lane-parallel matrix-vector product (lane i owns row i / column i /
// element i) with no serial Givens chain: r1updt + r1mpyq + the back substitution were 60 % of the QR kernel's
// instructions (profiles/README.md).  On 700 random three-tube configurations the CPU restatement of this
// variant follows the literal hybrd as closely as hybrd follows itself under a 1e-13 perturbation of its
// initial guess.
//
// What it cannot do is MINPACK's treatment of a singular R (a zero diagonal is replaced by eps * |column|).  A
// configuration whose Jacobian the Gauss-Jordan inversion finds singular (pivot below 1e-12 of the column norm: in
// practice a forward-difference column that vanished because hybrd's step h = eps |x_j| underflowed against the other
// terms) is appended to a fallback list that static_solve_kernel -- the literal QR algorithm -- solves in the same launch
// sequence; so is one with a zero-length section.  One-step sections, whose x and x^2 curvature columns are proportional
// (SURVEY App. B#5), are solved HERE on the reduced system (pin_mask in ctrb_static_common.cuh) unless the caller asks for
// MINPACK's literal treatment of them (CTRB_OPT_STATIC_ONE_STEP = 1: they then go to the QR kernel).
//
// The forward-difference Jacobian uses the block structure of the residual: perturbing an unknown of block
// (tube, section) only changes the node integrands of that section (and of section 2 for the (3,1) torsion,
// through the relative frame at the end of section 1), so a column needs 4 (or 8) node evaluations, not 12,
// and the untouched entries are exact zeros either way.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ctrb_static_common.cuh"

namespace ctrb {

#ifndef CTRB_FAST_WARPS
#define CTRB_FAST_WARPS 4
#endif
#ifndef CTRB_FAST_BLOCKS
#define CTRB_FAST_BLOCKS 3
#endif
constexpr int kFWarps = CTRB_FAST_WARPS;   // warps (= configurations in flight) per block
constexpr int kLD = kJLD;    // leading dimension, column-major: element (i, j) at i + j * kLD.  Odd, so both
                             // "lane = row" (stride 1) and "lane = column" (stride 31) accesses are conflict-free
constexpr unsigned kFull = 0xffffffffu;
constexpr int kFastAllowBlock = 1, kFastPinOneStep = 2;  // kernel flags
// unroll factors of the n-long loops when n is a template parameter: pairs of columns for the matrix-vector products
// (complete) and for the rank-one update, columns for the Gauss-Jordan step.  Measured on C3 (exposure 2..8, no
// fallbacks), M cfg/s: everything complete 11.19; update 3 -> 11.79; + Gauss-Jordan 3 -> 12.10; (2, 2) 12.07; (1, 1) 11.84;
// matrix-vector products by 5 or 8 instead of complete: 12.2 / 12.1 with the two-tube model slower.  The kernel runs on
// instruction fetch: ncu's stall_no_instruction went from 1.3 to 2.4 per issue with everything unrolled.
// Round 2 (one-step configurations inside): matrix-vector products 3 / 5 / 6 / 7 / 12 / complete: 90.9 / 91.7 / 91.1 / 93.3 /
// 93.7 / 91.9 ms per 2^20 shallow configurations, the same order on the uniform and the well-posed workloads.  The curve is not
// monotone: the kernel's speed follows its code layout as much as its instruction count.
#ifndef CTRB_MV_UNROLL
#define CTRB_MV_UNROLL 3
#endif
// Round 2, with the one-step configurations in this kernel (more Jacobian evaluations per configuration, so more of the
// code is live at once): the same sweep now favours the rolled loops -- rank-one update 3 / 2 / 1: 98.1 / 97.2 / 95.8 ms per
// 2^20 shallow configurations, Gauss-Jordan 3 / 1: 98.1 / 97.4; matrix-vector products 4 / 8 / complete: 104.9 / 98.1 / 98.1.
#ifndef CTRB_R1_UNROLL
#define CTRB_R1_UNROLL 1
#endif
#ifndef CTRB_INV_UNROLL
#define CTRB_INV_UNROLL 1
#endif
#ifndef CTRB_BLOCK_INVERT
#define CTRB_BLOCK_INVERT 1
#endif
#ifndef CTRB_B27_UNROLL
#define CTRB_B27_UNROLL 1   // 3 / 2 / 1: 95.4 / 94.0 / 92.1 ms per 2^20 shallow configurations
#endif
#ifndef CTRB_COPY_UNROLL
#define CTRB_COPY_UNROLL 4
#endif
constexpr int kCopyUnroll = CTRB_COPY_UNROLL;
constexpr int kB27Unroll = CTRB_B27_UNROLL;  // the three j-loops of fast_invert_block27
constexpr bool kBlockInvert = CTRB_BLOCK_INVERT != 0;
constexpr int kMvUnroll = CTRB_MV_UNROLL, kR1Unroll = CTRB_R1_UNROLL, kInvUnroll = CTRB_INV_UNROLL;

struct alignas(16) FastMem {
    double J[kLD * kMaxNq];
    double H[kLD * kMaxNq + 34];  // + slack: lanes 30, 31 read (and discard) one column past the end
    double x[32], f[32], va[32], vb[32], vc[32];
    double y[12][16];             // node integrands of the most recent residual
    StCfg cfg;
    // during the forward-difference Jacobian H is dead (it is rebuilt from J right after) and holds the node
    // integrands of 8 column slots x 4 nodes
    __device__ double (*ybuf())[16] { return reinterpret_cast<double(*)[16]>(H); }
};

// y_i = sum_j M[i][j] v[j]: lane = row i, v broadcast from shared memory (two elements per 128-bit load; every
// vector of FastMem is 16-byte aligned)
// N = the system size when the kernel is instantiated for it (loops unroll completely), 0 = run-time n
template <int N>
__device__ __noinline__ double matvec_row(const double* M, const double* v, int n_rt, int lane) {
    const int n = N ? N : n_rt;
    CTRB_IN_SHARED(M);
    CTRB_IN_SHARED(v);
    const double* m = M + lane;
    const double2* v2 = reinterpret_cast<const double2*>(v);
    double a0 = 0.0, a1 = 0.0;
    int j = 0;
#pragma unroll(N ? kMvUnroll : 3)
    for (; j + 1 < n; j += 2) {
        const double2 vv = v2[j >> 1];
        a0 = fma(m[j * kLD], vv.x, a0);
        a1 = fma(m[(j + 1) * kLD], vv.y, a1);
    }
    if (j < n) a0 = fma(m[j * kLD], v[j], a0);
    return a0 + a1;
}

// y_j = sum_i M[i][j] v[i]: lane = column j
template <int N>
__device__ __noinline__ double matvec_col(const double* M, const double* v, int n_rt, int lane) {
    const int n = N ? N : n_rt;
    CTRB_IN_SHARED(M);
    CTRB_IN_SHARED(v);
    const double* m = M + lane * kLD;
    const double2* v2 = reinterpret_cast<const double2*>(v);
    double a0 = 0.0, a1 = 0.0;
    int i = 0;
#pragma unroll(N ? kMvUnroll : 3)
    for (; i + 1 < n; i += 2) {
        const double2 vv = v2[i >> 1];
        a0 = fma(m[i], vv.x, a0);
        a1 = fma(m[i + 1], vv.y, a1);
    }
    if (i < n) a0 = fma(m[i], v[i], a0);
    return a0 + a1;
}

__device__ __forceinline__ double wnorm(double v) { return sqrt(warp_sum(v * v)); }

// H = J^-1 for the three-tube system from the default initial guess (27 unknowns) through its block structure.  The
// forward-difference Jacobian is block tridiagonal with blocks of 12 (section 1: blocks (1,1), (2,1)), 3 (the (3,1)
// torsion, which also turns the frame section 2 starts from) and 12 (section 2: blocks (2,2), (3,2)) unknowns, with
// exact zeros in the two off-corner 12 x 12 blocks (warp_fdjac writes them as 0.0):
//     J = [A11 A12 0; A21 A22 A23; 0 A32 A33],   J^-1 = diag(P, 0, Q) + W S^-1 Z,   P = A11^-1, Q = A33^-1,
//     W = [P A12; -I; Q A32],  Z = [A21 P, -I, A23 Q],  S = A22 - A21 P A12 - A23 Q A32  (3 x 3).
// P and Q are found together, one per half-warp (rows 0..11 on lanes 0..11, rows 15..26 on lanes 16..27), by the same
// exchange steps as fast_invert: 12 steps on 12 columns instead of 27 on 27.  Returns false when any pivot (or S) is
// small; the caller then runs the general Gauss-Jordan, which decides whether the Jacobian is singular.
__device__ __noinline__ bool fast_invert_block27(FastMem& M, double acnorm_l, int lane) {
    CTRB_IN_SHARED(&M);
    constexpr int n = 27, nb = 12, c2 = 12;  // c2: first row / column of the middle block
    double* Tm = M.H;
    const double* Jm = M.J;
#pragma unroll(kCopyUnroll)
    for (int e = lane; e < kLD * n; e += 32) Tm[e] = Jm[e];
    __syncwarp();
    const int half = lane >> 4, li = lane & 15;
    const bool act = li < nb;
    const int cb = half ? 15 : 0;          // first row / column of this half-warp's diagonal block
    const int row = cb + (act ? li : 0);   // this lane's row (and, where lane = column, its column)
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    bool used = false;
    int mycol = 0, prow = 0;               // block-local: the column this row pivoted; the pivot row of column li
#pragma unroll 1
    for (int k = 0; k < nb; ++k) {
        const double a = act ? Tm[row + (cb + k) * kLD] : 0.0;
        const unsigned long long key = (act && !used) ? (unsigned long long)__double_as_longlong(fabs(a)) : 0ULL;
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned mh = __reduce_max_sync(hmask, hi);
        const unsigned ml = __reduce_max_sync(hmask, hi == mh ? lo : 0u);
        const int r = __ffs(__ballot_sync(kFull, hi == mh && lo == ml && act && !used) & hmask) - 1;  // pivot lane
        const double thr = 1e-12 * __shfl_sync(kFull, acnorm_l, cb + k);
        const double piv = __shfl_sync(kFull, a, r < 0 ? 0 : r);
        const bool ok = r >= 0 && fabs(piv) > thr && fabs(piv) < 1.7e308;
        if (!__all_sync(kFull, ok)) return false;
        const int rrow = cb + (r & 15);
        const double ipiv = 1.0 / piv;
        const double m = (lane == r) ? 0.0 : a * ipiv;
        if (act && lane != r) {
            const double* pr = Tm + rrow + cb * kLD;
            double* tr = Tm + row + cb * kLD;
#pragma unroll(kB27Unroll)
            for (int j = 0; j < nb; ++j) tr[j * kLD] = fma(-m, pr[j * kLD], tr[j * kLD]);
        }
        __syncwarp();
        if (act) Tm[rrow + row * kLD] *= ipiv;  // lane = column of the pivot row
        __syncwarp();
        if (act) Tm[row + (cb + k) * kLD] = (lane == r) ? ipiv : -m;
        __syncwarp();
        if (lane == r) {
            used = true;
            mycol = k;
        }
        if (act && li == k) prow = r & 15;
    }
    {   // unscramble: P[mycol(i)][prow(j)] = T[i][j] inside each diagonal block
        double rv[nb];
#pragma unroll
        for (int j = 0; j < nb; ++j) rv[j] = act ? Tm[row + (cb + j) * kLD] : 0.0;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < nb; ++j) {
            const int pj = __shfl_sync(kFull, prow, (half << 4) + j);
            if (act) Tm[(cb + mycol) + (cb + pj) * kLD] = rv[j];
        }
        __syncwarp();
    }
    // u = this row of P A12 (Q A32), v = this column of A21 P (A23 Q)
    double u0 = 0.0, u1 = 0.0, u2 = 0.0, v0 = 0.0, v1 = 0.0, v2 = 0.0;
    if (act) {
        const double* pr = Tm + row + cb * kLD;           // row of P / Q
        const double* pc = Tm + cb + row * kLD;           // column of P / Q
        const double* a12 = Jm + cb + c2 * kLD;           // A12 / A32: rows cb.., columns 12..14
        const double* a21 = Jm + c2 + cb * kLD;           // A21 / A23: rows 12..14, columns cb..
#pragma unroll(kB27Unroll)
        for (int j = 0; j < nb; ++j) {
            const double p = pr[j * kLD], q = pc[j];
            u0 = fma(p, a12[j], u0);
            u1 = fma(p, a12[j + kLD], u1);
            u2 = fma(p, a12[j + 2 * kLD], u2);
            v0 = fma(a21[j * kLD], q, v0);
            v1 = fma(a21[j * kLD + 1], q, v1);
            v2 = fma(a21[j * kLD + 2], q, v2);
        }
    }
    // S = A22 - A21 (P A12) - A23 (Q A32): the sums run over the rows of both diagonal blocks, i.e. over the lanes
    double S[3][3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const double ab = act ? Jm[(c2 + b) + row * kLD] : 0.0;  // A21[b][i] / A23[b][i]
        S[b][0] = Jm[(c2 + b) + (c2 + 0) * kLD] - warp_sum(ab * u0);
        S[b][1] = Jm[(c2 + b) + (c2 + 1) * kLD] - warp_sum(ab * u1);
        S[b][2] = Jm[(c2 + b) + (c2 + 2) * kLD] - warp_sum(ab * u2);
    }
    // G = S^-1 by the adjugate (every lane; S is the same in all of them)
    const double c00 = S[1][1] * S[2][2] - S[1][2] * S[2][1], c01 = S[1][2] * S[2][0] - S[1][0] * S[2][2],
                 c02 = S[1][0] * S[2][1] - S[1][1] * S[2][0];
    const double det = S[0][0] * c00 + S[0][1] * c01 + S[0][2] * c02;
    double scale = 1.0;
#pragma unroll
    for (int b = 0; b < 3; ++b) scale *= fmax(fabs(S[b][0]), fmax(fabs(S[b][1]), fabs(S[b][2])));
    if (!(fabs(det) > 1e-10 * scale) || !(fabs(det) < 1.7e308)) return false;
    const double idet = 1.0 / det;
    double G[3][3];
    G[0][0] = c00 * idet;
    G[1][0] = c01 * idet;
    G[2][0] = c02 * idet;
    G[0][1] = (S[0][2] * S[2][1] - S[0][1] * S[2][2]) * idet;
    G[1][1] = (S[0][0] * S[2][2] - S[0][2] * S[2][0]) * idet;
    G[2][1] = (S[0][1] * S[2][0] - S[0][0] * S[2][1]) * idet;
    G[0][2] = (S[0][1] * S[1][2] - S[0][2] * S[1][1]) * idet;
    G[1][2] = (S[0][2] * S[1][0] - S[0][0] * S[1][2]) * idet;
    G[2][2] = (S[0][0] * S[1][1] - S[0][1] * S[1][0]) * idet;
    // back to lane = row / column i of the full matrix: rows 15..26 sit one lane higher
    const int src = lane >= 15 ? lane + 1 : lane;
    const bool midl = lane >= c2 && lane < c2 + 3;
    double w0 = __shfl_sync(kFull, u0, src), w1 = __shfl_sync(kFull, u1, src), w2 = __shfl_sync(kFull, u2, src);
    double z0 = __shfl_sync(kFull, v0, src), z1 = __shfl_sync(kFull, v1, src), z2 = __shfl_sync(kFull, v2, src);
    if (midl) {  // W and Z carry -I in the middle block
        w0 = lane == c2 ? -1.0 : 0.0; w1 = lane == c2 + 1 ? -1.0 : 0.0; w2 = lane == c2 + 2 ? -1.0 : 0.0;
        z0 = w0; z1 = w1; z2 = w2;
    }
    M.va[lane] = z0;
    M.vb[lane] = z1;
    M.vc[lane] = z2;
    __syncwarp();
    if (lane < n) {
        const double t0 = fma(w0, G[0][0], fma(w1, G[1][0], w2 * G[2][0]));
        const double t1 = fma(w0, G[0][1], fma(w1, G[1][1], w2 * G[2][1]));
        const double t2 = fma(w0, G[0][2], fma(w1, G[1][2], w2 * G[2][2]));
        double* hr = Tm + lane;
        const bool top = lane < nb, bot = lane >= 15;
#pragma unroll(kB27Unroll)
        for (int j = 0; j < n; ++j) {
            const bool diag = (top && j < nb) || (bot && j >= 15);
            const double d = diag ? hr[j * kLD] : 0.0;
            hr[j * kLD] = fma(t0, M.va[j], fma(t1, M.vb[j], fma(t2, M.vc[j], d)));
        }
    }
    __syncwarp();
    return true;
}

// H = J^-1 by Gauss-Jordan exchange steps with implicit partial pivoting; lane = row.  Returns false when a
// pivot falls below 1e-12 of its column norm (numerically singular Jacobian).
template <int N>
__device__ __noinline__ bool fast_invert(FastMem& M, int n_rt, double acnorm_l, int lane) {
    const int n = N ? N : n_rt;
    CTRB_IN_SHARED(&M);
    const bool act = lane < n;
    double* Tm = M.H;
    for (int e = lane; e < kLD * n; e += 32) Tm[e] = M.J[e];
    __syncwarp();
    bool used = false;
    int mycol = 0, prow = 0;  // lane r: column it pivoted; lane k: pivot row of column k
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const double a = act ? Tm[lane + k * kLD] : 0.0;
        const unsigned long long key = (act && !used) ? (unsigned long long)__double_as_longlong(fabs(a)) : 0ULL;
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned mh = __reduce_max_sync(kFull, hi);
        const unsigned ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
        const int r = __ffs(__ballot_sync(kFull, hi == mh && lo == ml && act && !used)) - 1;
        const double thr = 1e-12 * __shfl_sync(kFull, acnorm_l, k);
        const double piv = __shfl_sync(kFull, a, r < 0 ? 0 : r);
        if (r < 0 || !(fabs(piv) > thr) || !(fabs(piv) < 1.7e308)) return false;
        const double ipiv = 1.0 / piv;
        const double m = (lane == r) ? 0.0 : a * ipiv;
        if (act && lane != r) {
            const double* pr = Tm + r;
            double* tr = Tm + lane;
#pragma unroll(N ? kInvUnroll : 3)
            for (int j = 0; j < n; ++j) tr[j * kLD] = fma(-m, pr[j * kLD], tr[j * kLD]);
        }
        __syncwarp();
        if (act) Tm[r + lane * kLD] *= ipiv;  // lane = column j of the pivot row
        __syncwarp();
        if (act) Tm[lane + k * kLD] = (lane == r) ? ipiv : -m;
        __syncwarp();
        if (lane == r) {
            used = true;
            mycol = k;
        }
        if (lane == k) prow = r;
    }
    // unscramble: H[mycol(i)][prow(j)] = T[i][j]
    double row[kMaxNq];
#pragma unroll
    for (int j = 0; j < kMaxNq; ++j) row[j] = (j < n && act) ? Tm[lane + j * kLD] : 0.0;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kMaxNq; ++j) {
        const int pj = __shfl_sync(kFull, prow, j);
        if (j < n && act) M.H[mycol + pj * kLD] = row[j];
    }
    __syncwarp();
    return true;
}

// returns MINPACK's info (1 = converged, 2 = maxfev, 3 = xtol too small, 4/5 = slow progress), or a negative reason when the
// configuration has to go through the QR kernel (-1 degenerate section, -2 non-finite residual at the initial guess, -3 / -4
// singular Jacobian at the first / a later evaluation, -5 non-finite trial point, -6 singular Broyden update)
template <int N>
__device__ __noinline__ int fast_hybrd(const StModel& sm_, FastMem& M, const SlotTable& st, int lane, double xtol, int maxfev,
                                       double factor, bool reduced, bool allow_block, unsigned pin, int* nfev_out) {
    CTRB_USE_SM;
    CTRB_IN_SHARED(&M);
    CTRB_IN_SHARED(&st);
    // with the default initial guess the decoupled linear block (3,3) is carried implicitly (block5_norm2, tabulated per design)
    const int n = N ? N : (reduced ? sm.nq - sm.ndof[5] : sm.nq);
    const int nnodes = reduced ? 8 : 12;
    const double c33 = reduced ? sm.c33_tab[M.cfg.i33] : 0.0;
    const bool act = lane < n;
    RowMeta rm = row_meta(sm, act ? lane : 0);
    if (pin) pin_row_meta(sm, pin, lane, rm);  // one-step sections (pin_mask): pinned rows vanish, their twins carry the weight
    const double epsmch = DBL_EPSILON, p1c = 0.1, p5 = 0.5, p001 = 1e-3, p0001 = 1e-4;
    double x_l = act ? M.x[lane] : 0.0;
    double pa, pb;
    double f_l = warp_residual(sm, M.cfg, M.y, M.x, rm, act, lane, nnodes, &pa, &pb);
    int nfev = 1, info = 0;
    double fnorm = wnorm(f_l);
    if (!(fnorm < 1.7e308)) return -2;  // non-finite residual at the initial guess
    double rfnorm = 1.0 / fnorm;
    int iter = 1, ncsuc = 0, ncfail = 0, nslow1 = 0, nslow2 = 0;
    double delta = 0.0, xnorm = 0.0, diag_l = 1.0;
#pragma unroll 1
    for (;;) {
        bool jeval = true;
        const double acnorm_l = warp_fdjac(sm, M.cfg, st, rm, M.x, M.J, M.ybuf(), f_l, pa, pb, n, lane, pin);
        nfev += sm.nq;  // MINPACK's count: one evaluation per unknown
        // three tubes from the default initial guess: block-structured inverse; anything it declines (a small pivot)
        // goes through the general Gauss-Jordan, which alone decides what is handed to the QR kernel
        const bool blocked = kBlockInvert && allow_block && reduced && n == 27 && sm.mc.T == 3 && fast_invert_block27(M, acnorm_l, lane);
        if (!blocked && !fast_invert<N>(M, n, acnorm_l, lane)) return iter == 1 && nfev <= sm.nq + 1 ? -3 : -4;  // singular Jacobian (first / later)
        if (iter == 1) {
            diag_l = acnorm_l == 0.0 ? 1.0 : acnorm_l;
            xnorm = sqrt(warp_sum(act ? (diag_l * x_l) * (diag_l * x_l) : 0.0) + c33);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        diag_l = fmax(diag_l, acnorm_l);
        const double rdiag_l = 1.0 / diag_l;  // divisions by per-lane / per-iteration quantities become multiplications
        if (act) M.f[lane] = f_l;
        __syncwarp();
#pragma unroll 1
        for (;;) {
            // ---- dogleg (MINPACK): Gauss-Newton step H f, scaled gradient J^T f / D, their trust-region blend
            const double xg = act ? matvec_row<N>(M.H, M.f, n, lane) : 0.0;
            const double qnorm = wnorm(diag_l * xg);
            double step = xg;
            if (!(qnorm <= delta)) {
                double g = act ? matvec_col<N>(M.J, M.f, n, lane) * rdiag_l : 0.0;
                const double gnorm = wnorm(g);
                double sgnorm = 0.0, alpha = delta / qnorm;
                if (gnorm != 0.0) {
                    g = act ? (g * (1.0 / gnorm)) * rdiag_l : 0.0;
                    if (act) M.va[lane] = g;
                    __syncwarp();
                    const double w2 = act ? matvec_row<N>(M.J, M.va, n, lane) : 0.0;
                    double temp = wnorm(w2);
                    sgnorm = (gnorm / temp) / temp;
                    alpha = 0.0;
                    if (sgnorm < delta) {
                        const double bnorm = fnorm;  // |Q^T f| = |f|
                        temp = (bnorm / gnorm) * (bnorm / qnorm) * (sgnorm / delta);
                        const double dq = delta / qnorm, sd = sgnorm / delta;
                        temp = temp - dq * sd * sd + sqrt((temp - dq) * (temp - dq) + (1 - dq * dq) * (1 - sd * sd));
                        alpha = (dq * (1 - sd * sd)) / temp;
                    }
                    __syncwarp();
                }
                const double temp = (1 - alpha) * fmin(sgnorm, delta);
                step = temp * g + alpha * xg;
            }
            const double p_l = act ? -step : 0.0;
            const double xt_l = x_l + p_l;
            // inside the trust region p = -(H f), so |D p| is |D (H f)| = qnorm to the bit: one warp reduction less
            // (measured: 92.0 -> 89.3 ms per 2^20 shallow configurations)
            const double pnorm = (qnorm <= delta) ? qnorm : wnorm(diag_l * p_l);
            if (iter == 1) delta = fmin(delta, pnorm);
            if (act) {
                M.va[lane] = p_l;
                M.vb[lane] = xt_l;
            }
            __syncwarp();
            double qa, qb;
            const double fn_l = warp_residual(sm, M.cfg, M.y, M.vb, rm, act, lane, nnodes, &qa, &qb);
            // Everything the rest of the iteration reduces over the warp is known here: |f+|, |f + J p|, |D x+| and v.(H u) of the
            // Sherman-Morrison denominator (the Broyden vectors only need f+, J p and |D p|).  One four-value butterfly instead of
            // four chains of five dependent shuffle-add steps; the two products with H move ahead of the acceptance test (wasted
            // in the last iteration and when the Jacobian is re-evaluated).  Same sums, same order: bit-identical results (checked on
            // 2^20 configurations), 89.3 -> 88.2 ms.
            nfev++;
            const double w3 = act ? f_l + matvec_row<N>(M.J, M.va, n, lane) : 0.0;  // f + J p
            // Broyden vectors (hybrd: wa2 = (Q^T f+ - (Q^T f + R p)) / pnorm, wa1 = D ((D p) / pnorm))
            const double ipn = 1.0 / pnorm;
            const double u_l = act ? (fn_l - w3) * ipn : 0.0;
            const double v_l = act ? diag_l * ((diag_l * p_l) * ipn) : 0.0;
            if (act) {
                M.vb[lane] = u_l;
                M.vc[lane] = v_l;
            }
            __syncwarp();
            const double Hu = act ? matvec_row<N>(M.H, M.vb, n, lane) : 0.0;
            const double vH = act ? matvec_col<N>(M.H, M.vc, n, lane) : 0.0;
            const D4 red = warp_sum4(fn_l * fn_l, w3 * w3, act ? (diag_l * xt_l) * (diag_l * xt_l) : 0.0, v_l * Hu);
            const double fnorm1 = sqrt(red.a);
            if (!(fnorm1 < 1.7e308) || !(pnorm < 1.7e308)) return -5;  // non-finite: let the QR kernel decide
            double actred = -1.0;
            // hybrd only compares these quantities with constants: the two quotients by |f| become products with 1 / |f| (kept
            // from the last accepted iterate), and ratio = actred / prered is compared as actred against c * prered: two
            // divisions less per iteration; same decisions (results bit-identical on 2 x 2^20 configurations), 88.1 -> 87.4 ms
            if (fnorm1 < fnorm) actred = 1 - (fnorm1 * rfnorm) * (fnorm1 * rfnorm);
            const double temp = sqrt(red.b);
            double prered = 0.0;
            if (temp < fnorm) prered = 1 - (temp * rfnorm) * (temp * rfnorm);
            const bool pos = prered > 0;
            const bool r_lt_p1 = !pos ? true : actred < p1c * prered;     // ratio < 0.1   (ratio = 0 when prered <= 0)
            const bool r_ge_p5 = pos && actred >= p5 * prered;             // ratio >= 0.5
            const bool r_near1 = pos && fabs(actred - prered) <= p1c * prered;  // |ratio - 1| <= 0.1
            const bool r_ok = pos && actred >= p0001 * prered;             // ratio >= 1e-4
            if (r_lt_p1) {
                ncsuc = 0;
                ncfail++;
                delta = p5 * delta;
            } else {
                ncfail = 0;
                ncsuc++;
                if (r_ge_p5 || ncsuc > 1) delta = fmax(delta, pnorm / p5);
                if (r_near1) delta = pnorm / p5;
            }
            if (r_ok) {  // successful iteration
                rfnorm = 1.0 / fnorm1;
                x_l = xt_l;
                f_l = fn_l;
                pa = qa;
                pb = qb;
                if (act) {
                    M.x[lane] = x_l;
                    M.f[lane] = f_l;
                }
                xnorm = sqrt(red.c + c33);
                fnorm = fnorm1;
                iter++;
            }
            nslow1++;
            if (actred >= p001) nslow1 = 0;
            if (jeval) nslow2++;
            if (actred >= p1c) nslow2 = 0;
            if (delta <= xtol * xnorm || fnorm == 0) info = 1;
            if (info == 0) {
                if (nfev >= maxfev) info = 2;
                if (p1c * fmax(p1c * delta, pnorm) <= epsmch * xnorm) info = 3;
                if (nslow2 == 5) info = 4;
                if (nslow1 == 10) info = 5;
            }
            __syncwarp();
            if (info != 0) {
                *nfev_out = nfev;
                return info;
            }
            if (ncfail == 2) break;  // recompute the Jacobian
            // ---- rank-one update of J and, by Sherman-Morrison, of H
            const double den = 1.0 + red.d;
            if (!(fabs(den) > 1e-8) || !(fabs(den) < 1e100)) return -6;  // J+ is (numerically) singular
            __syncwarp();
            if (act) M.vb[lane] = vH;
            __syncwarp();
            if (act) {
                const double hs = -Hu * (1.0 / den);
                double* jr = M.J + lane;
                double* hr = M.H + lane;
                const double2* vc2 = reinterpret_cast<const double2*>(M.vc);
                const double2* vb2 = reinterpret_cast<const double2*>(M.vb);
                int j = 0;
#pragma unroll(N ? kR1Unroll : 3)
                for (; j + 1 < n; j += 2) {
                    const double2 c2 = vc2[j >> 1], b2 = vb2[j >> 1];
                    jr[j * kLD] = fma(u_l, c2.x, jr[j * kLD]);
                    hr[j * kLD] = fma(hs, b2.x, hr[j * kLD]);
                    jr[(j + 1) * kLD] = fma(u_l, c2.y, jr[(j + 1) * kLD]);
                    hr[(j + 1) * kLD] = fma(hs, b2.y, hr[(j + 1) * kLD]);
                }
                if (j < n) {
                    jr[j * kLD] = fma(u_l, M.vc[j], jr[j * kLD]);
                    hr[j * kLD] = fma(hs, M.vb[j], hr[j * kLD]);
                }
            }
            __syncwarp();
            jeval = false;
        }
    }
}

template <int N>
__global__ void __launch_bounds__(kFWarps * 32, CTRB_FAST_BLOCKS)
    static_fast_kernel(const __grid_constant__ StModel sm_, long long B, const int32_t* __restrict__ idx, const double* __restrict__ theta,
                       const double* __restrict__ x0, int x0_per_config, double xtol, const long long* __restrict__ offsets,
                       double* __restrict__ g_out, double* __restrict__ sol_out, int32_t* __restrict__ nfev_out,
                       int32_t* __restrict__ info_out, long long* __restrict__ fb_list,
                       unsigned long long* __restrict__ fb_count, unsigned long long* __restrict__ ticket, int flags,
                       unsigned long long* __restrict__ fb_reasons) {
    load_model(sm_);
    const StModel& sm = g_sm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SlotTable st;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FastMem& M = reinterpret_cast<FastMem*>(smem_raw)[warp];
    const int T = sm.mc.T, n = sm.nq;
    const bool reduced = x0 == nullptr;  // default initial guess: q_33 = q*
    if (threadIdx.x == 0) build_slot_table(sm, st, reduced);
    __syncthreads();
    for (;;) {
        unsigned long long tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1ULL);
        const long long b = (long long)__shfl_sync(kFull, tk, 0);
        if (b >= B) break;
        if (lane == 0) st_setup(sm, idx + b * T, theta + b * T, M.cfg);
        M.x[lane] = lane < n ? (x0 ? x0[(x0_per_config ? b * n : 0) + lane] : sm.x0[lane]) : 0.0;
        M.f[lane] = 0.0;
        M.va[lane] = 0.0;
        M.vb[lane] = 0.0;
        M.vc[lane] = 0.0;
        __syncwarp();
        const StCfg& c = M.cfg;
        // zero-length sections make whole residual blocks vanish (static.py:478-479): singular by construction, left to the
        // QR kernel.  A one-step section 1 or 2 gives the x and x^2 curvature coefficients proportional columns (SURVEY
        // App. B#5): solved here on the reduced system (pin_mask), or handed to the QR kernel when the caller asked for
        // MINPACK's own treatment of them (CTRB_OPT_STATIC_ONE_STEP = 1)
        const int lmin = (flags & kFastPinOneStep) ? 2 : 3;
        const bool degenerate = (T == 3 && c.L[0] < lmin) || c.L[1] < lmin || c.L[2] < 2;
        const unsigned pin = (flags & kFastPinOneStep) ? pin_mask(sm, c) : 0u;
        int nfev = 0;
        const int info = degenerate ? -1
                                    : fast_hybrd<N>(sm, M, st, lane, xtol, 200 * (n + 1), 100.0, reduced,
                                                    (flags & kFastAllowBlock) != 0, pin, &nfev);
        __syncwarp();
        if (info < 0) {
            if (lane == 0) {
                fb_list[atomicAdd(fb_count, 1ULL)] = b;
                atomicAdd(fb_reasons + (-info - 1), 1ULL);
            }
            continue;
        }
        if (sol_out && lane < n) sol_out[b * n + lane] = M.x[lane];
        if (lane == 0) {
            if (nfev_out) nfev_out[b] = nfev;
            if (info_out) info_out[b] = info;
        }
        __syncwarp();
    }
}

template <int N>
static int launch_static_fast_n(ctrb_ctx* ctx, const StModel& sm, long long B, const int32_t* idx, const double* theta,
                                const double* x0, int x0_per_config, double xtol, const long long* offsets, double* g_out,
                                double* sol_out, int32_t* nfev_out, int32_t* info_out, long long* fb_list,
                                unsigned long long* fb_count, unsigned long long* ticket) {
    // CTRB_OPT_FAST_BLOCK_INVERT = 0: the general Gauss-Jordan inverse for every configuration (parity test of the
    // block-structured one); CTRB_OPT_STATIC_ONE_STEP = 1: one-step sections go to the QR kernel
    const int flags = (ctx->opt[CTRB_OPT_FAST_BLOCK_INVERT] ? kFastAllowBlock : 0) |
                      (ctx->opt[CTRB_OPT_STATIC_ONE_STEP] == 0 ? kFastPinOneStep : 0);
    const size_t smem = sizeof(FastMem) * kFWarps;
    CTRB_CUDA(cudaFuncSetAttribute(static_fast_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CTRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, static_fast_kernel<N>, kFWarps * 32, smem));
    const long long blocks_needed = (B + kFWarps - 1) / kFWarps;
    const int grid = (int)std::min<long long>(blocks_needed, (long long)ctx->sm_count * std::max(per_sm, 1));
    static_fast_kernel<N><<<grid, kFWarps * 32, smem, ctx->stream>>>(sm, B, idx, theta, x0, x0_per_config, xtol, offsets, g_out,
                                                                    sol_out, nfev_out, info_out, fb_list, fb_count, ticket, flags, ctx->d_counters + 16);
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

// The kernel is instantiated for the system sizes of the reference's configurations (three tubes: 30 unknowns, 27 from
// the default initial guess; two tubes: 15 / 12) so that its n-long loops unroll completely; any other size takes the
// run-time-n instantiation.
int launch_static_fast(ctrb_ctx* ctx, const StModel& sm, long long B, const int32_t* idx, const double* theta,
                       const double* x0, int x0_per_config, double xtol, const long long* offsets, double* g_out,
                       double* sol_out, int32_t* nfev_out, int32_t* info_out, long long* fb_list,
                       unsigned long long* fb_count, unsigned long long* ticket) {
    const int n = x0 == nullptr ? sm.nq - sm.ndof[5] : sm.nq;  // as fast_hybrd derives it
    const bool generic = ctx->opt[CTRB_OPT_FAST_GENERIC_N] != 0;  // run-time-n instantiation (parity test)
#define CTRB_FAST_ARGS ctx, sm, B, idx, theta, x0, x0_per_config, xtol, offsets, g_out, sol_out, nfev_out, info_out, fb_list, fb_count, ticket
    if (!generic) {
        switch (n) {
            case 27: return launch_static_fast_n<27>(CTRB_FAST_ARGS);
            case 30: return launch_static_fast_n<30>(CTRB_FAST_ARGS);
            case 12: return launch_static_fast_n<12>(CTRB_FAST_ARGS);
            case 15: return launch_static_fast_n<15>(CTRB_FAST_ARGS);
            default: break;
        }
    }
    return launch_static_fast_n<0>(CTRB_FAST_ARGS);
#undef CTRB_FAST_ARGS
}

}  // namespace ctrb
