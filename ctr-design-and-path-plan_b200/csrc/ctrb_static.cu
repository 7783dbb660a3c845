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
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ctrb_static_common.cuh"

namespace ctrb {

#ifndef CTRB_QR_WARPS
#define CTRB_QR_WARPS 3
#endif
#ifndef CTRB_QR_BLOCKS
#define CTRB_QR_BLOCKS 5
#endif
#ifndef CTRB_QR_UNROLL
// unroll factor of the dot-product / update loops of qrfac, qform, dogleg and the Broyden vectors (one multiply-add per
// 10 instructions when rolled).  Measured on C3 shallow, QR kernel M cfg/s: 1 -> 2.55, 2 -> 2.69, 4 -> 2.55, 8 -> 2.28:
// beyond 2 the larger code costs more instruction-cache misses than the saved loop overhead buys.
#define CTRB_QR_UNROLL 2
#endif
// The implicit last block's constant: from the per-design table (as the fast kernel does) or evaluated per configuration.
// Same value either way; the table is 2.6 % SLOWER here (2.64 vs 2.71 M cfg/s): with this kernel's iteration at ~48 KB of
// code against a 32 KB instruction cache its speed follows the code layout more than the instruction count.
#ifndef CTRB_QR_C33_TABLE
#define CTRB_QR_C33_TABLE 0
#endif
#ifndef CTRB_QR_ITER_UNROLL
#define CTRB_QR_ITER_UNROLL 2   // the same for the loops inside hybrd's iteration (dogleg, predicted reduction, Broyden vectors)
#endif
constexpr int kQrUnroll = CTRB_QR_UNROLL, kQrIterUnroll = CTRB_QR_ITER_UNROLL;
constexpr int kSWarps = CTRB_QR_WARPS;    // warps (= configurations) per block: 5 blocks (15 warps) per SM by smem and registers
constexpr int kNP = 31;       // leading dimension of the n x n work matrices (odd: conflict-free column access)
constexpr int kNV = 32;       // padded vector length

// static_equilibrium (static.py:228-259): F(q), n = sm.nq
__device__ __noinline__ void st_residual(const StModel& sm_, const StCfg& c, const double* qp, double* F) {
    CTRB_USE_SM;
    const int T = sm.mc.T;
    const QV q = {qp, -1, 0.0};
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
                for (int e = 0; e < 2; ++e) node_sec12<false>(sm, c, q, 0, e ? c.hi[0][a] : c.lo[0][a], y2[e]);
#pragma unroll 1
                for (int k = 0; k < 15; ++k) acc[k] += 0.5 * ((y2[1][k] - y2[0][k]) * c.wq[0][a] + y2[0][k]);
            }
        }
        const double h = c.len[0] / 2;
        for (int k = 0; k < 9; ++k) F[sm.qoff[0] + k] = h * acc[k];
        for (int k = 0; k < 3; ++k) F[sm.qoff[1] + k] = h * acc[9 + k];
        for (int k = 0; k < 3; ++k) F31[k] = h * acc[12 + k];
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
                for (int e = 0; e < 2; ++e) node_sec12<false>(sm, c, q, 1, e ? c.hi[1][a] : c.lo[1][a], y2[e]);
#pragma unroll 1
                for (int k = 0; k < 15; ++k) acc[k] += 0.5 * ((y2[1][k] - y2[0][k]) * c.wq[1][a] + y2[0][k]);
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
                for (int e = 0; e < 2; ++e) node_sec3<false>(sm, c, q, e ? c.hi[2][a] : c.lo[2][a], y2[e]);
#pragma unroll 1
                for (int k = 0; k < dof; ++k) acc[k] += 0.5 * ((y2[1][k] - y2[0][k]) * c.wq[2][a] + y2[0][k]);
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
    double wa1[kNV], wa2[kNV], wa3[kNV], wa4[kNV];
    double x[kNV], fvec[kNV], qtf[kNV];
    double ybuf0[12][16];     // node integrands of a residual (3 sections x 2 abscissae x {lo, hi})
    StCfg cfg;
    // During the forward-difference Jacobian R and wa1..wa4 are dead (R is rebuilt from the QR right after):
    // the 32 node vectors of a round of column slots alias them (465 + 128 >= 512 doubles).
    __device__ double (*ybuf32())[16] { return reinterpret_cast<double(*)[16]>(r); }
};
static_assert(kMaxNq * (kMaxNq + 1) / 2 + 4 * kNV >= 32 * 16, "fdjac scratch does not fit in R + wa1..wa4");

__device__ __forceinline__ int rowoff(int n, int j) { return j * n - (j * (j - 1)) / 2; }  // packed upper-triangular rows

// one residual evaluation F(q) by the whole warp; q and F are shared-memory vectors
// dogleg (MINPACK): Gauss-Newton step by back substitution on the packed R, scaled gradient, and their
// convex combination inside the trust region.  One vector element per lane; returns this lane's p.
__device__ __noinline__ double dogleg_w(int n, const double* r, double diag_l, const double* qtb, double delta, double* tmp,
                                        int lane) {
    const double epsmch = DBL_EPSILON;
    // Gauss-Newton step R x = qtb by column-oriented back substitution: lane i carries b_i; once x_j is known every
    // lane i < j subtracts R[i][j] x_j.  1 / R[j][j] is formed by all lanes at once (a zero diagonal is replaced by
    // eps * max |column|, as dogleg does).
    double rd = 1.0, bl = 0.0;
    if (lane < n) {
        double temp = r[rowoff(n, lane)];
        if (temp == 0.0) {
#pragma unroll 1
            for (int i = 0; i <= lane; ++i) temp = fmax(temp, fabs(r[rowoff(n, i) + lane - i]));
            temp = epsmch * temp;
            if (temp == 0.0) temp = epsmch;
        }
        rd = 1.0 / temp;
        bl = qtb[lane];
    }
    double xg = 0.0;  // Gauss-Newton component of this lane
#pragma unroll 1
    for (int j = n - 1; j >= 0; --j) {
        const double xj = __shfl_sync(0xffffffffu, bl * rd, j);
        if (lane == j) xg = xj;
        if (lane < j) bl = fma(-r[rowoff(n, lane) + j - lane], xj, bl);
    }
    const bool act = lane < n;
    const double qnorm = enorm_w(n, act ? diag_l * xg : 0.0, nullptr);
    if (qnorm <= delta) return xg;
    // scaled gradient direction: wa1_i = (sum_{j<=i} r(j,i) qtb_j) / diag_i
    double g = 0.0;
    if (act) {
#pragma unroll kQrIterUnroll
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
#pragma unroll kQrIterUnroll
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

// The plane rotation (c, s) = (a, b) / hypot(a, b); identity when both vanish.  MINPACK stores its rotations as
// one number (tau) and rebuilds cos / sin with a square root wherever they are applied; here they are kept as pairs.
// b == 0 is MINPACK's "if (w[j] != 0)" guard: no rotation.  (a / |a|, 0) with a < 0 would be a rotation by pi, which r1updt's
// "s != 0" shortcut skips on R while r1mpyq applies it to Q: the factors then disagree.  It cannot happen by rounding, but it
// does for the decoupled unit rows of pinned unknowns (pin_mask), whose w_j is exactly zero and whose R_jj is -1.
__device__ __forceinline__ void rot_of(double a, double b, double* c_, double* s_) {
    const double h2 = fma(a, a, b * b);
    const bool rot = b != 0.0 && h2 > 0.0;
    const double ri = rot ? rsqrt(h2) : 0.0;
    *c_ = rot ? a * ri : 1.0;
    *s_ = b * ri;
}

// r1updt: QR factors of R + u v^T (R packed by rows in s).  Lane l owns w_l and element l of every row.
// First sweep (rotations that fold v into v_{n-1} e_{n-1}): |v_{n-1}| after step j is sqrt(sum_{k>=j} v_k^2), so all
// n-1 rotations come from one suffix scan; second sweep (Hessenberg -> triangular) is sequential in j.
// Outputs the rotations as c1/s1 (first sweep) and c2/s2 (second sweep), j = 0 .. n-2, for r1mpyq.
__device__ __noinline__ void r1updt_w(int n, double* s, double u_l, const double* v, double* c1, double* s1, double* c2,
                                      double* s2, int lane) {
    const bool act = lane < n;
    const double vl = act ? v[lane] : 0.0;
    double S = vl * vl;  // suffix sums of squares
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double t = __shfl_down_sync(0xffffffffu, S, off);
        if (lane + off < 32) S += t;
    }
    const double vlast = __shfl_sync(0xffffffffu, vl, n - 1);
    const double Snext = __shfl_down_sync(0xffffffffu, S, 1);
    const double S0 = __shfl_sync(0xffffffffu, S, 0);
    __syncwarp();
    if (lane < n - 1) {
        double c_ = 1.0, s_ = 0.0;
        if (vl != 0.0) {
            const double ri = rsqrt(S);
            const double vn = (lane == n - 2) ? vlast : sqrt(Snext);  // v_{n-1} before this step
            c_ = vn * ri;
            s_ = vl * ri;
        }
        c1[lane] = c_;
        s1[lane] = s_;
    }
    const double vn_final = (S0 - vlast * vlast > 0.0) ? sqrt(S0) : vlast;
    __syncwarp();
    double wl = 0.0;
    if (lane == n - 1) wl = s[rowoff(n, n - 1)];
#pragma unroll 1
    for (int j = n - 2; j >= 0; --j) {
        const double s_ = s1[j];
        if (s_ != 0.0 && lane >= j && act) {
            const double c_ = c1[j];
            const int l = rowoff(n, j) + lane - j;
            const double sl = s[l];
            s[l] = c_ * sl - s_ * wl;
            wl = s_ * sl + c_ * wl;
        }
    }
    if (act) wl = fma(vn_final, u_l, wl);
    __syncwarp();
    // Second sweep (upper Hessenberg -> triangular): rotation j needs R[j][j] and w_j, both held by lane j.  Every lane
    // forms "its" rotation (no divergence), lane j's is broadcast by shuffle, and the row-(j+1) element is fetched
    // before the rotation is applied so that the shared-memory latency overlaps the dependent arithmetic.
    {
        int ro = 0;  // rowoff(n, 0)
        double sl = act ? s[lane] : 0.0;
#pragma unroll 1
        for (int j = 0; j < n - 1; ++j) {
            double c_, s_;
            rot_of(sl, wl, &c_, &s_);
            c_ = __shfl_sync(0xffffffffu, c_, j);
            s_ = __shfl_sync(0xffffffffu, s_, j);
            const int ro_next = ro + (n - j);
            const double sl_next = (lane > j && act) ? s[ro_next + lane - (j + 1)] : 0.0;
            if (lane == j) {
                c2[j] = c_;
                s2[j] = s_;
            }
            if (s_ != 0.0 && lane >= j && act) {
                s[ro + lane - j] = c_ * sl + s_ * wl;
                wl = -s_ * sl + c_ * wl;
            }
            sl = sl_next;
            ro = ro_next;
        }
    }
    if (lane == n - 1) s[rowoff(n, n - 1)] = wl;
    __syncwarp();
}

// apply the 2(n-1) rotations to ONE row a[0..n-1] with stride lda (r1mpyq, one row per lane)
__device__ __noinline__ void st_r1mpyq_row(int n, double* a, int lda, const double* c1, const double* s1, const double* c2,
                                           const double* s2) {
    double an = a[(n - 1) * lda];
#pragma unroll 2
    for (int j = n - 2; j >= 0; --j) {
        const double c_ = c1[j], s_ = s1[j], aj = a[j * lda];
        a[j * lda] = c_ * aj - s_ * an;
        an = s_ * aj + c_ * an;
    }
#pragma unroll 2
    for (int j = 0; j < n - 1; ++j) {
        const double c_ = c2[j], s_ = s2[j], aj = a[j * lda];
        a[j * lda] = c_ * aj + s_ * an;
        an = -s_ * aj + c_ * an;
    }
    a[(n - 1) * lda] = an;
}

// MINPACK hybrd, warp-cooperative.  Every scalar of the iteration (delta, norms, counters, ratio) is
// computed redundantly and identically by all lanes; vectors are one element per lane.
// returns info (1 = converged, 2 = maxfev, 3 = xtol too small, 4/5 = slow progress)
__device__ __noinline__ int st_hybrd(const StModel& sm_, HybrdMem& M, const SlotTable& st, int lane, double xtol, int maxfev,
                                     double factor, bool reduced, unsigned pin, int* nfev_out) {
    CTRB_USE_SM;
    // with the default initial guess the decoupled linear block (3,3) is carried implicitly (block5_norm2, tabulated per design)
    const int n = reduced ? sm.nq - sm.ndof[5] : sm.nq;
    const int nnodes = reduced ? 8 : 12;
    // Three tubes from the default initial guess: the forward-difference Jacobian is block tridiagonal (12 | 3 | 12
    // unknowns, exact zeros in the two corner blocks, see fast_invert_block27), so the Householder vectors of the first
    // 12 columns end at row 14.  The loops over them stop there: the skipped terms are products with exact zeros.
    const bool tri = reduced && n == 27 && sm.mc.T == 3;
#if CTRB_QR_C33_TABLE
    const double c33 = reduced ? sm.c33_tab[M.cfg.i33] : 0.0;
#else
    const double c33 = reduced ? block5_norm2(sm, M.cfg, M.x, M.ybuf32(), lane) : 0.0;
#endif
    const bool act = lane < n;
    const double epsmch = DBL_EPSILON, p1 = 0.1, p5 = 0.5, p001 = 1e-3, p0001 = 1e-4;
    double* fjac = M.fjac;
    int info = 0, nfev;
    RowMeta rm = row_meta(sm, act ? lane : 0);
    if (pin) pin_row_meta(sm, pin, lane, rm);  // one-step sections on the reduced system, as in the fast kernel (pin_mask)
    double pa, pb;  // section shares of this lane's residual component at the accepted x (for the sparse fdjac)
    {
        const double f0 = warp_residual(sm, M.cfg, M.ybuf0, M.x, rm, act, lane, nnodes, &pa, &pb);
        if (act) M.fvec[lane] = f0;
        __syncwarp();
    }
    nfev = 1;
    double fnorm = enorm_w(n, act ? M.fvec[lane] : 0.0, M.fvec);
    int iter = 1, ncsuc = 0, ncfail = 0, nslow1 = 0, nslow2 = 0;
    double delta = 0, xnorm = 0;
    double diag_l = 1.0;
#pragma unroll 1
    for (;;) {  // outer loop: fresh forward-difference Jacobian
        bool jeval = true;
        // fdjac1 (dense band) through the block structure of the residual: 4 (8) node evaluations per column
        warp_fdjac(sm, M.cfg, st, rm, M.x, fjac, M.ybuf32(), act ? M.fvec[lane] : 0.0, pa, pb, n, lane, pin);
        nfev += sm.nq;  // MINPACK's count: one evaluation per unknown
        __syncwarp();
        // qrfac: Householder QR without pivoting; lane k owns column k.  acnorm -> wa2, rdiag -> wa1
        double acnorm_l = 0.0;
        if (act) {
            double s2 = 0.0;
#pragma unroll kQrUnroll
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
                    const int iend = (tri && j < 12) ? 15 : n;  // rows 15.. of this Householder vector are exact zeros
                    double sum = 0;
#pragma unroll kQrUnroll
                    for (int i = j; i < iend; ++i) sum += fjac[i + j * kNP] * fjac[i + lane * kNP];
                    const double temp = sum / fjac[j + j * kNP];
#pragma unroll kQrUnroll
                    for (int i = j; i < iend; ++i) fjac[i + lane * kNP] -= temp * fjac[i + j * kNP];
                }
            }
            if (lane == 0) M.wa1[j] = -ajnorm;
            __syncwarp();
        }
        if (iter == 1) {
            diag_l = acnorm_l == 0 ? 1.0 : acnorm_l;
            xnorm = sqrt(warp_sum(act ? (diag_l * M.x[lane]) * (diag_l * M.x[lane]) : 0.0) + c33);
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
                    const int iend = (tri && k < 12) ? 15 : n;  // wa3 = Householder vector k: exact zeros from row 15 on
                    double sum = 0;
#pragma unroll kQrUnroll
                    for (int i = k; i < iend; ++i) sum += fjac[i + lane * kNP] * M.wa3[i];
                    const double temp = sum / M.wa3[k];
#pragma unroll kQrUnroll
                    for (int i = k; i < iend; ++i) fjac[i + lane * kNP] -= temp * M.wa3[i];
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
            double qa, qb;
            {
                const double f1 = warp_residual(sm, M.cfg, M.ybuf0, M.wa2, rm, act, lane, nnodes, &qa, &qb);
                if (act) M.wa4[lane] = f1;
                __syncwarp();
            }
            nfev++;
            const double fnorm1 = enorm_w(n, act ? M.wa4[lane] : 0.0, M.wa4);
            double actred = -1;
            if (fnorm1 < fnorm) actred = 1 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
            // predicted reduction: || qtf + R p ||
            double w3 = 0.0;
            if (act) {
                const int ro = rowoff(n, lane);
                double sum = 0;
#pragma unroll kQrIterUnroll
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
                pa = qa;
                pb = qb;
                xnorm = sqrt(warp_sum(act ? (diag_l * M.wa2[lane]) * (diag_l * M.wa2[lane]) : 0.0) + c33);
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
#pragma unroll kQrIterUnroll
                for (int i = 0; i < n; ++i) sum += fjac[i + lane * kNP] * M.wa4[i];
                M.wa2[lane] = (sum - w3) / pnorm;                 // v
                u_l = diag_l * ((diag_l * p_l) / pnorm);          // u
                if (ratio >= p0001) M.qtf[lane] = sum;
            }
            __syncwarp();
            r1updt_w(n, M.r, u_l, M.wa2, M.wa1, M.wa2, M.wa3, M.wa4, lane);  // v is read into registers before c1/s1 land
            // rows of Q on lanes 0..n-1, qtf (one more row under the same rotations) on lane 31
            if (act || lane == 31) st_r1mpyq_row(n, act ? fjac + lane : M.qtf, act ? kNP : 1, M.wa1, M.wa2, M.wa3, M.wa4);
            __syncwarp();
            jeval = false;
        }
    }
}

// ------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(kSWarps * 32, CTRB_QR_BLOCKS)
    static_solve_kernel(const __grid_constant__ StModel sm_, long long B, const int32_t* __restrict__ idx, const double* __restrict__ theta,
                        const double* __restrict__ x0, int x0_per_config, double xtol, const long long* __restrict__ offsets,
                        double* __restrict__ g_out, double* __restrict__ sol_out, int32_t* __restrict__ nfev_out,
                        int32_t* __restrict__ info_out, unsigned long long* __restrict__ ticket,
                        const long long* __restrict__ list, const unsigned long long* __restrict__ list_count,
                        unsigned long long* __restrict__ list_total, int pin_one_step) {
    if (list && list_total && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(list_total, *list_count);
    load_model(sm_);
    const StModel& sm = g_sm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SlotTable st;
    const bool reduced = x0 == nullptr;  // default initial guess: q_33 = q*
    if (threadIdx.x == 0) build_slot_table(sm, st, reduced);
    __syncthreads();
    // with a list the kernel solves configurations list[0 .. *list_count): the ones the fast kernel handed back
    const long long work = list ? (long long)*list_count : B;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    HybrdMem& M = reinterpret_cast<HybrdMem*>(smem_raw)[warp];
    const int T = sm.mc.T, n = sm.nq;
    for (;;) {
        // dynamic schedule: the evaluation count varies from ~30 to 200(n+1) between configurations
        unsigned long long tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1ULL);
        const long long w = (long long)__shfl_sync(0xffffffffu, tk, 0);
        if (w >= work) break;
        const long long b = list ? list[w] : w;
        if (lane == 0) st_setup(sm, idx + b * T, theta + b * T, M.cfg);
        if (lane < n) M.x[lane] = x0 ? x0[(x0_per_config ? b * n : 0) + lane] : sm.x0[lane];
        __syncwarp();
        int nfev = 0;
        // one-step sections under the product's rule (CTRB_OPT_STATIC_ONE_STEP = 0) keep it here too: a configuration the
        // fast kernel hands over after a singular Jacobian re-evaluation is solved on the same reduced system
        const unsigned pin = pin_one_step ? pin_mask(sm, M.cfg) : 0u;
        const int info = st_hybrd(sm, M, st, lane, xtol, 200 * (n + 1), 100.0, reduced, pin, &nfev);
        __syncwarp();
        if (sol_out && lane < n) sol_out[b * n + lane] = M.x[lane];
        if (lane == 0) {
            if (nfev_out) nfev_out[b] = nfev;
            if (info_out) info_out[b] = info;
        }
        __syncwarp();
    }
}

// Section curves from the solved unknowns (Static.solve_g's integrate_g calls, static.py:110-130): one warp per
// configuration.  A kernel of its own: integrate_g's code (Magnus exponentials, SE(3) scan) is as large as the
// solver's inner loop, and running it at the end of every solve evicted that loop from the instruction cache
// (+46 % on the fast kernel, profiles/README.md).
constexpr int kCWarps = 8;
struct CurveMem {
    StCfg cfg;
    double q[kNV];
};
template <bool kPos>  // kPos: node positions only (3 doubles per node) -- what the fused solve -> collision path reads
__global__ void __launch_bounds__(kCWarps * 32)
    static_curves_kernel(const __grid_constant__ StModel sm_, long long B, const int32_t* __restrict__ idx,
                         const double* __restrict__ theta, const double* __restrict__ sol,
                         const long long* __restrict__ offsets, double* __restrict__ g_out) {
    load_model(sm_);
    const StModel& sm = g_sm;
    __shared__ CurveMem mem[kCWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    CurveMem& M = mem[warp];
    const int T = sm.mc.T, n = sm.nq;
    for (long long b = (long long)blockIdx.x * kCWarps + warp; b < B; b += (long long)gridDim.x * kCWarps) {
        if (lane == 0) st_setup(sm, idx + b * T, theta + b * T, M.cfg);
        if (lane < n) M.q[lane] = sol[b * n + lane];
        __syncwarp();
        st_write_curves<kPos>(sm, M.cfg, M.q, offsets, b, g_out, lane);
        __syncwarp();
    }
}

// The same curves for the fused solve -> collision path, which reads node positions only: ONE THREAD per configuration
// walks its three section curves serially (Magnus step, SE(3) product, 24 B store per node).  The warp-per-configuration
// scan above keeps 5 of 32 lanes busy on a "shallow" section (5 nodes) and spends four scan rounds on it; a lane per
// configuration does each node's work once: ~400 instructions per node and lane, 32 configurations per warp.
// (The 4x4 output of Static.solve_g keeps the scan kernel: its stores are coalesced 128 B rows.)
__device__ __forceinline__ void omega_poly(const double* qs, double x, double w[3]) {
    w[0] = poly3(qs, x);
    w[1] = poly3(qs + 3, x);
    w[2] = poly3(qs + 6, x);
}
__device__ __forceinline__ void omega_last(const StModel& sm_, const double* qs, double x, double w[3]) {
    CTRB_USE_SM;
    const int T = sm.mc.T, dof = sm.mc.dof[T - 1];
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
}
// nodes 0 .. count-1 of one section curve (positions), serially; returns the last node.  kLast: the section of the last
// tube alone (its own strain base, evaluated at the section coordinate + idx33 once more, static_bases.py:77-81)
template <bool kLast>
__device__ __forceinline__ SE3 chain_section_pos(const StModel& sm_, const double* qs, int i33, int start, int count, SE3 g,
                                                 double* __restrict__ out) {
    CTRB_USE_SM;
    const double dx = sm.mc.dx;
    const double cg1 = 0.5 - sqrt(3.0) / 6, cg2 = 0.5 + sqrt(3.0) / 6;  // static.py:597-599
    out[0] = g.p[0]; out[1] = g.p[1]; out[2] = g.p[2];
#pragma unroll 1
    for (int k = 1; k < count; ++k) {
        const double si = (double)(start + k);
        double wA[3], wB[3];
        if (kLast) {
            omega_last(sm, qs, (si + cg1 - 1 + i33) * dx, wA);
            omega_last(sm, qs, (si + cg2 - 1 + i33) * dx, wB);
        } else {
            omega_poly(qs, (si + cg1 - 1) * dx, wA);
            omega_poly(qs, (si + cg2 - 1) * dx, wB);
        }
        g = se3_mul(g, magnus_step_inl(wA, wB, dx));
        out[3 * k + 0] = g.p[0]; out[3 * k + 1] = g.p[1]; out[3 * k + 2] = g.p[2];
    }
    return g;
}
constexpr int kCPThreads = 128;
__global__ void __launch_bounds__(kCPThreads)
    static_curves_pos_kernel(const __grid_constant__ StModel sm_, long long B, const int32_t* __restrict__ idx,
                             const double* __restrict__ theta, const double* __restrict__ sol,
                             const long long* __restrict__ offsets, double* __restrict__ pos_out) {
    load_model(sm_);
    const StModel& sm = g_sm;
    const long long b = (long long)blockIdx.x * kCPThreads + threadIdx.x;
    if (b >= B) return;
    const int T = sm.mc.T, n = sm.nq;
    const int* N = sm.mc.npts;
    const double dx = sm.mc.dx;
    const int32_t* ib = idx + b * T;
    const double* q = sol + b * n;
    int i11 = ib[0];
    const int i22 = ib[T - 2], i33 = ib[T - 1];   // static.py:81-95
    int i21 = i22 - (N[0] - i11 - 1);
    const int i32 = i33 - (N[T - 2] - i22 - 1);
    int i31 = i32 - (N[0] - i11 - 1);
    if (T == 2) i11 = i21 = i31 = 0;
    const int L0 = T == 3 ? N[0] - i11 : 1, L1 = N[T - 2] - i22, L2 = N[T - 1] - i33;
    double qs[9];
    SE3 g11_last, g21_last, g31_last;
    se3_identity(g11_last);
    if (T == 3) {  // static.py:110-113
#pragma unroll
        for (int k = 0; k < 9; ++k) qs[k] = q[sm.qoff[0] + k];
        g11_last = chain_section_pos<false>(sm, qs, 0, i11, L0, se3_rx(theta[b * T + 0]), pos_out + offsets[b * T + 0] * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) qs[k] = q[sm.qoff[1] + k];
        g21_last = se3_rx(theta[b * T + 1] + ipoly3_abs(qs, i21 * dx, (i21 + L0 - 1) * dx));
#pragma unroll
        for (int k = 0; k < 3; ++k) qs[k] = q[sm.qoff[2] + k];
        g31_last = se3_rx(theta[b * T + 2] + ipoly3_abs(qs, i31 * dx, (i31 + L0 - 1) * dx));
    } else {       // static.py:114-118
        g21_last = se3_rx(theta[b * T + 0]);
        g31_last = se3_rx(theta[b * T + 1]);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) qs[k] = q[sm.qoff[3] + k];
    const SE3 g22_last = chain_section_pos<false>(sm, qs, 0, i22, L1, se3_mul(g11_last, g21_last),
                                                  pos_out + offsets[b * T + (T - 2)] * 3);
#pragma unroll
    for (int k = 0; k < 3; ++k) qs[k] = q[sm.qoff[4] + k];
    const SE3 g32_last = se3_rx(ipoly3_abs(qs, i32 * dx, (i32 + L1 - 1) * dx));
#pragma unroll 1
    for (int k = 0; k < CTRB_MAX_DOF; ++k) qs[k] = k < sm.ndof[5] ? q[sm.qoff[5] + k] : 0.0;
    chain_section_pos<true>(sm, qs, i33, i33, L2, se3_mul(se3_mul(g22_last, g31_last), g32_last),
                            pos_out + offsets[b * T + (T - 1)] * 3);
}

// residual only (parity tests): one thread per (config) evaluates F(q)
__global__ void static_residual_kernel(StModel sm_, long long B, const int32_t* __restrict__ idx,
                                       const double* __restrict__ theta, const double* __restrict__ q,
                                       double* __restrict__ F) {
    load_model(sm_);
    const StModel& sm = g_sm;
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    StCfg c;
    st_setup(sm, idx + b * sm.mc.T, theta + b * sm.mc.T, c);
    double ql[kMaxNq], Fl[kMaxNq];
    for (int k = 0; k < sm.nq; ++k) ql[k] = q[b * sm.nq + k];
    st_residual(sm, c, ql, Fl);
    for (int k = 0; k < sm.nq; ++k) F[b * sm.nq + k] = Fl[k];
}

// per-design tables of StModel: one thread per node
__global__ void static_tables_kernel(StModel sm, double* __restrict__ ws, double* __restrict__ b3, StQuad* __restrict__ quad,
                                     int quad_n) {
    if ((int)(blockIdx.x * blockDim.x + threadIdx.x) < quad_n) {
        StQuad e;
        st_quad(sm.mc.dx, blockIdx.x * blockDim.x + threadIdx.x, e);
        quad[blockIdx.x * blockDim.x + threadIdx.x] = e;
    }
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int T = sm.mc.T;
    if (j < sm.mc.total_nodes) {
        int n = 0;
        while (n + 1 < T && j >= sm.mc.node_base[n + 1]) ++n;
        double w[3];
        strain_omega(sm.mc, n, (j - sm.mc.node_base[n]) * sm.mc.dx, w);  // get_ksi_star: index * delta_x
        ws[3 * j + 0] = w[0];
        ws[3 * j + 1] = w[1];
        ws[3 * j + 2] = w[2];
    }
    if (j < sm.mc.npts[T - 1]) {
        const int dof = sm.mc.dof[T - 1];
        double* o = b3 + (CTRB_MAX_DOF + 3) * j;
        double bq[3] = {0.0, 0.0, 0.0};
        for (int k = 0; k < CTRB_MAX_DOF; ++k) {
            int r = 1;
            double v = 0.0;
            if (k < dof) strain_base_col(sm.mc.kind[T - 1], dof, k, j * sm.mc.dx, &r, &v);
            o[k] = v;
            if (k < dof) bq[r] += v * sm.mc.q[T - 1][k];
        }
        o[CTRB_MAX_DOF + 0] = bq[0];
        o[CTRB_MAX_DOF + 1] = bq[1];
        o[CTRB_MAX_DOF + 2] = bq[2];
    }
}

// block5_norm2 for every insertion index of the last tube (one warp each), from the default initial guess
__global__ void static_c33_kernel(StModel sm_, double* __restrict__ c33) {
    load_model(sm_);
    const StModel& sm = g_sm;
    __shared__ StCfg cfg;
    __shared__ double x[32], yb[32][16];
    const int T = sm.mc.T, lane = threadIdx.x, i33 = blockIdx.x;
    if (lane == 0) {
        int idx[3] = {0, 0, 0};
        double theta[3] = {0.0, 0.0, 0.0};
        for (int n = 0; n < T; ++n) idx[n] = sm.mc.npts[n] > 2 ? sm.mc.npts[n] - 3 : 0;  // sections 1, 2 are not read
        idx[T - 1] = i33;
        st_setup(sm, idx, theta, cfg);
    }
    x[lane] = lane < sm.nq ? sm.x0[lane] : 0.0;
    __syncwarp();
    const double v = block5_norm2(sm, cfg, x, yb, lane);
    if (lane == 0) c33[i33] = v;
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
    for (int k = 0; k < CTRB_MAX_DOF; ++k) {
        int r = 1;
        double v = 0.0;
        if (k < mc.dof[T - 1]) strain_base_col(mc.kind[T - 1], mc.dof[T - 1], k, 1.0, &r, &v);
        sm.b3_row[k] = r;
    }
    *out = sm;
    return CTRB_OK;
}

int launch_static_tables(ctrb_ctx* ctx, ctrb_model* m) {
    const int tot = m->mc.total_nodes, T = m->mc.T;
    CTRB_CUDA(cudaMalloc(&m->d_ws_tab, (size_t)tot * 3 * sizeof(double)));
    CTRB_CUDA(cudaMalloc(&m->d_b3_tab, (size_t)m->mc.npts[T - 1] * (CTRB_MAX_DOF + 3) * sizeof(double)));
    int quad_max = 1;
    for (int n = 0; n < T; ++n) quad_max = std::max(quad_max, m->mc.npts[n]);
    CTRB_CUDA(cudaMalloc(&m->d_quad_tab, (size_t)(quad_max + 1) * sizeof(StQuad)));
    m->sm.ws_tab = m->d_ws_tab;
    m->sm.b3_tab = m->d_b3_tab;
    m->sm.quad_tab = m->d_quad_tab;
    m->sm.quad_max = quad_max;
    const int work = std::max(tot, quad_max + 1);
    static_tables_kernel<<<(work + 127) / 128, 128, 0, ctx->stream>>>(m->sm, m->d_ws_tab, m->d_b3_tab, m->d_quad_tab, quad_max + 1);
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    CTRB_CUDA(cudaMalloc(&m->d_c33_tab, (size_t)m->mc.npts[T - 1] * sizeof(double)));
    static_c33_kernel<<<m->mc.npts[T - 1], 32, 0, ctx->stream>>>(m->sm, m->d_c33_tab);  // stream order: after the tables it reads
    m->sm.c33_tab = m->d_c33_tab;
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

int launch_static_fast(ctrb_ctx* ctx, const StModel& sm, long long B, const int32_t* idx, const double* theta,
                       const double* x0, int x0_per_config, double xtol, const long long* offsets, double* g_out,
                       double* sol_out, int32_t* nfev_out, int32_t* info_out, long long* fb_list,
                       unsigned long long* fb_count, unsigned long long* ticket);

// integrate_g of given unknowns (one warp per configuration); pos_only: 3 doubles per node instead of 16
int launch_static_curves(ctrb_ctx* ctx, const StModel& sm, long long B, const int32_t* idx, const double* theta,
                         const double* sol, const long long* offsets, double* g_out, int pos_only) {
    if (B <= 0) return CTRB_OK;
    KernelTimer kc(ctx, CTRB_KERNEL_STATIC_CURVES);
    const long long blocks = (B + kCWarps - 1) / kCWarps;
    const int cgrid = (int)std::min<long long>(blocks, (long long)ctx->sm_count * 16);
    if (pos_only == 1) {
        static_curves_pos_kernel<<<(unsigned)((B + kCPThreads - 1) / kCPThreads), kCPThreads, 0, ctx->stream>>>(sm, B, idx, theta, sol,
                                                                                                              offsets, g_out);
    } else if (pos_only == 2) {  // the scan kernel writing positions (parity test of the lane-per-configuration kernel)
        static_curves_kernel<true><<<cgrid, kCWarps * 32, 0, ctx->stream>>>(sm, B, idx, theta, sol, offsets, g_out);
    } else {
        static_curves_kernel<false><<<cgrid, kCWarps * 32, 0, ctx->stream>>>(sm, B, idx, theta, sol, offsets, g_out);
    }
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

// Static.solve_g for B configurations: static_fast_kernel carries hybrd's iteration on the
// explicit Jacobian and its inverse and hands back the configurations it cannot treat -- zero-length sections, a Jacobian
// that turns out numerically singular, and, with CTRB_OPT_STATIC_ONE_STEP = 1, one-step sections --
// which static_solve_kernel then solves with MINPACK's own QR machinery; static_curves_kernel integrates the section
// curves.  CTRB_OPT_STATIC_SOLVER = 1 sends everything through the QR kernel.
int launch_static_solve(ctrb_ctx* ctx, const StModel& sm, long long B, const int32_t* idx, const double* theta,
                        const double* x0, int x0_per_config, double xtol, const long long* offsets, double* g_out,
                        double* sol_out, int32_t* nfev_out, int32_t* info_out, int pos_only) {
    if (B <= 0) return CTRB_OK;
    const bool qr_only = ctx->opt[CTRB_OPT_STATIC_SOLVER] == 1;
    const size_t smem = sizeof(HybrdMem) * kSWarps;
    CTRB_CUDA(cudaFuncSetAttribute(static_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks_needed = (B + kSWarps - 1) / kSWarps;
    int per_sm = 1;
    CTRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, static_solve_kernel, kSWarps * 32, smem));
    int grid = (int)std::min<long long>(blocks_needed, (long long)ctx->sm_count * std::max(per_sm, 1));
    // d_counters[8] = QR ticket, [9] = fast ticket, [10] = fallback count ([0..7] belong to the collision kernels)
    CTRB_CUDA(cudaMemsetAsync(ctx->d_counters + 8, 0, 3 * sizeof(unsigned long long), ctx->stream));
    if (g_out && !sol_out) {  // the curves kernel reads the solutions back
        void* p = nullptr;
        int rc = ctx_scratch(ctx, 13, (size_t)B * sm.nq * sizeof(double), &p);
        if (rc) return rc;
        sol_out = (double*)p;
    }
    long long* fb_list = nullptr;
    if (!qr_only) {
        void* p = nullptr;
        int rc = ctx_scratch(ctx, 12, (size_t)B * sizeof(long long), &p);
        if (rc) return rc;
        fb_list = (long long*)p;
    }
    KernelTimer kt(ctx, CTRB_KERNEL_STATIC_SOLVE);
    if (!qr_only) {
        KernelTimer kf(ctx, CTRB_KERNEL_STATIC_FAST);
        int rc = launch_static_fast(ctx, sm, B, idx, theta, x0, x0_per_config, xtol > 0 ? xtol : 1.49012e-8, offsets, g_out,
                                    sol_out, nfev_out, info_out, fb_list, ctx->d_counters + 10, ctx->d_counters + 9);
        if (rc) return rc;
    }
    {
        KernelTimer kq(ctx, CTRB_KERNEL_STATIC_QR);
        static_solve_kernel<<<grid, kSWarps * 32, smem, ctx->stream>>>(sm, B, idx, theta, x0, x0_per_config,
                                                                      xtol > 0 ? xtol : 1.49012e-8, offsets, g_out, sol_out,
                                                                      nfev_out, info_out, ctx->d_counters + 8, fb_list,
                                                                      ctx->d_counters + 10, ctx->d_counters + 11,
                                                                      ctx->opt[CTRB_OPT_STATIC_ONE_STEP] == 0);
        ctx->launches++;
        CTRB_CUDA(cudaGetLastError());
    }
    if (g_out) return launch_static_curves(ctx, sm, B, idx, theta, sol_out, offsets, g_out, pos_only);
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
