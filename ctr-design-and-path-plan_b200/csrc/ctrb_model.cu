// ctrb_model.cu -- kinematic strain-based model kernels (sm_100a).
//
//   kin_table_kernel        per-design table of cumulative steps C_n[j] (exp chain of kinematic.py:106-118)
//   kin_gcurve_kernel       Kinematic.solve_g for a batch: G_j = A_n * C_n[j], A_n = G_start * C_n[idx]^-1
//   calc_indices_kernel     calculate_indices (kinematic.py:201-274)
//   node_count/scan kernels ragged output offsets
//   kin_eta_ftl_kernel      Kinematic.solve_eta (kinematic.py:132-198)
//
// The strain field depends on (design q, tube, node) only -- never on the (insertion, rotation)
// tuple -- so the se(3) exponentials and their running product are evaluated once per model
// handle (kin_table_kernel) and every configuration reduces to one SE(3) compose per emitted
// node.  The g-curve kernel is therefore bound by HBM writes (128 B / node in fp64).
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

#include "ctrb_se3.cuh"

namespace ctrb {

// ------------------------------------------------------------------ table build
__global__ void kin_table_kernel(ModelConst mc, double* tab, double* tab_inv, float* tabR32) {
    int n = threadIdx.x;
    if (n >= mc.T) return;
    const int tot = mc.total_nodes;
    SE3 c;
    se3_identity(c);
    const double bias_v[3] = {1.0, 0.0, 0.0};  // strain_bias = [0,0,0,1,0,0], kinematic.py:49
    for (int j = 0; j < mc.npts[n]; ++j) {
        if (j > 0) {
            double w[3];
            strain_omega(mc, n, (j - 0.5) * mc.dx, w);  // centered_s, kinematic.py:107
            SE3 e = se3_exp(w, bias_v, mc.dx);
            c = se3_mul(c, e);
        }
        SE3 ci = se3_inv(c);
        int col = mc.node_base[n] + j;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            tab[k * tot + col] = c.r[k];
            tab_inv[k * tot + col] = ci.r[k];
            tabR32[k * tot + col] = (float)c.r[k];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            tab[(9 + k) * tot + col] = c.p[k];
            tab_inv[(9 + k) * tot + col] = ci.p[k];
        }
    }
}

// ------------------------------------------------------------------ g-curve kernel
// One warp owns 32 configurations at a time.
//  phase 1 (lane = configuration): the short sequential chain over tubes,
//     A_n = G_tip(n-1) * Rx(theta_n) * C_n[idx_n]^-1, staged in shared memory;
//  phase 2 (lane = (node slot, matrix row)): 8 nodes per warp instruction, each lane
//     produces one 4-wide row of G_j = A_n * C_n[j] and stores it with one 256-bit
//     (fp64) or 128-bit (fp32) instruction: a warp store covers 1 KiB / 512 B contiguous.
constexpr int kGWarps = 4;

template <typename OutT>
struct RowStore;
template <>
struct RowStore<double> {
    static __device__ __forceinline__ void st(double* p, double a, double b, double c, double d) {
        asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
    }
};
template <>
struct RowStore<float> {
    static __device__ __forceinline__ void st(float* p, float a, float b, float c, float d) {
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
    }
};

template <typename OutT>
__global__ void __launch_bounds__(kGWarps * 32)
    kin_gcurve_kernel(ModelConst mc, ModelTables tb, long long B, const int32_t* __restrict__ idx,
                      const double* __restrict__ theta, int full, const long long* __restrict__ offsets,
                      OutT* __restrict__ out, int tab_in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = mc.T, tot = mc.total_nodes;
    // layout: [A: kGWarps*32*T*12 doubles][output offsets: kGWarps*32*T int64][start idx: kGWarps*32*T ints]
    //         [table (optional): fp64 12*tot doubles | fp32 position rows 3*tot doubles + rotation 9*tot floats]
    double* sA = reinterpret_cast<double*>(smem_raw);
    long long* sOff = reinterpret_cast<long long*>(sA + kGWarps * 32 * T * 12);
    int* sIdx = reinterpret_cast<int*>(sOff + kGWarps * 32 * T);
    const double* tab = tb.tab;
    const float* tabR = tb.tabR32;
    const double* tabP = tb.tab + 9 * tot;  // position rows, the only fp64 part the fp32 path reads
    if (tab_in_smem) {
        double* st = reinterpret_cast<double*>(sOff + kGWarps * 32 * T) + (kGWarps * 32 * T + 1) / 2;
        if (sizeof(OutT) == 8) {
            for (int i = threadIdx.x; i < 12 * tot; i += blockDim.x) st[i] = tb.tab[i];
            tab = st;
        } else {  // the fp32 path reads the position rows (9..11) in fp64 and the rotation rows in fp32
            for (int i = threadIdx.x; i < 3 * tot; i += blockDim.x) st[i] = tb.tab[9 * tot + i];
            tabP = st;
            float* sf = reinterpret_cast<float*>(st + 3 * tot);
            for (int i = threadIdx.x; i < 9 * tot; i += blockDim.x) sf[i] = tb.tabR32[i];
            tabR = sf;
        }
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* wA = sA + (size_t)warp * 32 * T * 12;
    long long* wOff = sOff + warp * 32 * T;
    int* wIdx = sIdx + warp * 32 * T;
    const long long ngroups = (B + 31) / 32;
    const int slot = lane >> 2, row = lane & 3;

    for (long long grp = (long long)blockIdx.x * kGWarps + warp; grp < ngroups; grp += (long long)gridDim.x * kGWarps) {
        const long long b0 = grp * 32;
        // ---- phase 1: lane = configuration
        {
            long long b = b0 + lane;
            if (b < B) {
                SE3 g;
                se3_identity(g);
                for (int n = 0; n < T; ++n) {
                    int id = idx[b * T + n];
                    se3_rot_x(g, theta[b * T + n]);                              // kinematic.py:95-97
                    SE3 a = se3_mul(g, load_tab(tb.tab_inv, tot, mc.node_base[n] + id));
                    double* dst = wA + (lane * T + n) * 12;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {  // row-major rows [r0 r1 r2 p]
                        dst[k * 4 + 0] = a.r[k * 3 + 0];
                        dst[k * 4 + 1] = a.r[k * 3 + 1];
                        dst[k * 4 + 2] = a.r[k * 3 + 2];
                        dst[k * 4 + 3] = a.p[k];
                    }
                    wIdx[lane * T + n] = id;
                    wOff[lane * T + n] = offsets[b * T + n];
                    g = se3_mul(a, load_tab(tb.tab, tot, mc.node_base[n] + mc.npts[n] - 1));  // tube tip
                }
            }
        }
        __syncwarp();
        // ---- phase 2
        const int ncfg = (int)min((long long)32, B - b0);
        if (sizeof(OutT) == 8) {
            // fp64: lane = (8 node slots, 4 matrix rows); one 256-bit store per lane, 1 KiB per warp instruction
            for (int c = 0; c < ncfg; ++c) {
                for (int n = 0; n < T; ++n) {
                    const int id = wIdx[c * T + n];
                    const int start = full ? 0 : id;                               // kinematic.py:100-103
                    const int cnt = mc.npts[n] - start;
                    const long long obase = wOff[c * T + n];
                    double a0 = 0, a1 = 0, a2 = 0, ap = 0;
                    if (row < 3) {
                        const double* ar = wA + (c * T + n) * 12 + row * 4;
                        a0 = ar[0];
                        a1 = ar[1];
                        a2 = ar[2];
                        ap = ar[3];
                    }
                    const int colbase = mc.node_base[n] + start;
                    for (int k = slot; k < cnt; k += 8) {
                        const int col = colbase + k;
                        double* dst = reinterpret_cast<double*>(out) + ((obase + k) * 16 + row * 4);
                        if (row == 3) {
                            RowStore<double>::st(dst, 0.0, 0.0, 0.0, 1.0);
                        } else {
                            double o0 = fma(a0, tab[0 * tot + col], fma(a1, tab[3 * tot + col], a2 * tab[6 * tot + col]));
                            double o1 = fma(a0, tab[1 * tot + col], fma(a1, tab[4 * tot + col], a2 * tab[7 * tot + col]));
                            double o2 = fma(a0, tab[2 * tot + col], fma(a1, tab[5 * tot + col], a2 * tab[8 * tot + col]));
                            double o3 = fma(a0, tab[9 * tot + col], fma(a1, tab[10 * tot + col], fma(a2, tab[11 * tot + col], ap)));
                            RowStore<double>::st(dst, o0, o1, o2, o3);
                        }
                    }
                }
            }
        } else {
            // fp32: lane = (16 node slots, 2 row pairs); rows (0,1) or (2,3) = 32 B = one 256-bit store per lane.
            // Rotation in fp32, the position entry in fp64 (its two ~L-sized terms cancel).
            const int slot16 = lane >> 1, half = lane & 1;
            for (int c = 0; c < ncfg; ++c) {
                for (int n = 0; n < T; ++n) {
                    const int id = wIdx[c * T + n];
                    const int start = full ? 0 : id;
                    const int cnt = mc.npts[n] - start;
                    const long long obase = wOff[c * T + n];
                    const double* ar = wA + (c * T + n) * 12 + (half ? 8 : 0);   // row 0 or row 2
                    const double a0 = ar[0], a1 = ar[1], a2 = ar[2], ap = ar[3];
                    const double b0_ = half ? 0.0 : ar[4], b1_ = half ? 0.0 : ar[5], b2_ = half ? 0.0 : ar[6], bp = half ? 0.0 : ar[7];
                    const float f0 = (float)a0, f1 = (float)a1, f2 = (float)a2;
                    const float g0 = (float)b0_, g1 = (float)b1_, g2 = (float)b2_;
                    const int colbase = mc.node_base[n] + start;
                    for (int k = slot16; k < cnt; k += 16) {
                        const int col = colbase + k;
                        const float r0 = tabR[0 * tot + col], r1 = tabR[1 * tot + col], r2 = tabR[2 * tot + col];
                        const float r3 = tabR[3 * tot + col], r4 = tabR[4 * tot + col], r5 = tabR[5 * tot + col];
                        const float r6 = tabR[6 * tot + col], r7 = tabR[7 * tot + col], r8 = tabR[8 * tot + col];
                        const double px = tabP[col], py = tabP[tot + col], pz = tabP[2 * tot + col];
                        float o[8];
                        o[0] = fmaf(f0, r0, fmaf(f1, r3, f2 * r6));
                        o[1] = fmaf(f0, r1, fmaf(f1, r4, f2 * r7));
                        o[2] = fmaf(f0, r2, fmaf(f1, r5, f2 * r8));
                        o[3] = (float)fma(a0, px, fma(a1, py, fma(a2, pz, ap)));
                        if (half) {
                            o[4] = 0.f; o[5] = 0.f; o[6] = 0.f; o[7] = 1.f;
                        } else {
                            o[4] = fmaf(g0, r0, fmaf(g1, r3, g2 * r6));
                            o[5] = fmaf(g0, r1, fmaf(g1, r4, g2 * r7));
                            o[6] = fmaf(g0, r2, fmaf(g1, r5, g2 * r8));
                            o[7] = (float)fma(b0_, px, fma(b1_, py, fma(b2_, pz, bp)));
                        }
                        float* dst = reinterpret_cast<float*>(out) + ((obase + k) * 16 + half * 8);
                        asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(o[0]), "f"(o[1]),
                                     "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                                     : "memory");
                    }
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ calculate_indices (kinematic.py:201-274)
__device__ __forceinline__ void calc_indices_one(double p, double q, double min_len, double max_len, double dx,
                                                 int* prev_idx, int* new_idx) {
    p = p < min_len ? min_len : (p > max_len ? max_len : p);
    q = q < min_len ? min_len : (q > max_len ? max_len : q);
    double delta = q - p;
    int pi;
    double di;
    if (fabs(delta) <= 0.5 * dx) {
        di = (delta > 0.0) - (delta < 0.0);
        double pd = p / dx, qd = q / dx;
        if (rint(pd) == rint(qd) && (ceil(pd) == ceil(qd) || floor(pd) == floor(qd))) {
            pi = di < 0 ? (int)ceil(pd) : (int)floor(pd);
        } else {
            pi = (int)rint(pd);  // Python round(): half to even
        }
    } else {
        di = rint(delta / dx);
        pi = (int)rint(p / dx);
    }
    *prev_idx = pi;
    *new_idx = pi + (int)di;
}

__global__ void calc_indices_kernel(ModelConst mc, long long B, const double* __restrict__ prev_ins,
                                    const double* __restrict__ next_ins, int32_t* __restrict__ prev_idx,
                                    int32_t* __restrict__ new_idx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * mc.T) return;
    int n = (int)(i % mc.T);
    double min_len = n == 0 ? 0.0 : mc.L[n - 1];  // kinematic.py:224-228
    int p, q;
    calc_indices_one(prev_ins[i], next_ins[i], min_len, mc.L[n], mc.dx, &p, &q);
    prev_idx[i] = p;
    new_idx[i] = q;
}

// ------------------------------------------------------------------ ragged offsets: exclusive scan of node counts
constexpr int kScanBlock = 1024;

__global__ void node_count_scan1(ModelConst mc, long long n_items, const int32_t* __restrict__ idx, int full,
                                 long long* __restrict__ offsets, long long* __restrict__ block_sums) {
    __shared__ long long s[kScanBlock];
    long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
    long long v = 0;
    if (i < n_items) {
        int n = (int)(i % mc.T);
        v = full ? mc.npts[n] : mc.npts[n] - idx[i];
    }
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < kScanBlock; off <<= 1) {
        long long t = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n_items) offsets[i] = s[threadIdx.x] - v;  // exclusive within block
    if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = s[threadIdx.x];
}

__global__ void node_count_scan2(long long nblocks, long long* __restrict__ block_sums) {
    // single block: exclusive scan of block sums in place (serial over chunks of kScanBlock)
    __shared__ long long s[kScanBlock];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < nblocks; base += kScanBlock) {
        long long i = base + threadIdx.x;
        long long v = i < nblocks ? block_sums[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < kScanBlock; off <<= 1) {
            long long t = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblocks) block_sums[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) carry += s[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nblocks] = carry;  // grand total
}

__global__ void node_count_scan3(long long n_items, long long* __restrict__ offsets,
                                 const long long* __restrict__ block_sums, long long nblocks) {
    long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
    if (i < n_items) offsets[i] += block_sums[blockIdx.x];
    if (i == 0) offsets[n_items] = block_sums[nblocks];
}

// ------------------------------------------------------------------ eta / follow-the-leader (kinematic.py:132-198)
// One warp per configuration.  eta_n is a T-step recurrence (all lanes redundantly, it is
// ~60 FLOP per tube); the per-node FTL twist is lane-parallel over the parent's nodes:
//   ftl_k = Ad(P_k^-1) eta_n - V_n xi(i dx) = [R^T w - V w_i ; R^T (v - p x w) - V e_x]
#ifndef CTRB_FTL_WARPS
#define CTRB_FTL_WARPS 8
#endif
#ifndef CTRB_FTL_UNROLL2
#define CTRB_FTL_UNROLL2 0
#endif
constexpr int kEWarps = CTRB_FTL_WARPS;

// rows 0..2 of a 4x4 node: three 256-bit loads (sm_100 has them; the node is 32-byte aligned) through the read-only path
__device__ __forceinline__ void ld_row(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ SE3 load_g64(const double* g) {
    SE3 o;
    ld_row(g, o.r[0], o.r[1], o.r[2], o.p[0]);
    ld_row(g + 4, o.r[3], o.r[4], o.r[5], o.p[1]);
    ld_row(g + 8, o.r[6], o.r[7], o.r[8], o.p[2]);
    return o;
}

__device__ __forceinline__ double twist_mag(const double f[6], bool translation_only) {
    double on = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    double vn = sqrt(f[3] * f[3] + f[4] * f[4] + f[5] * f[5]);
    if (translation_only) return vn;        // follow_the_leader.py:173-189
    return on == 0.0 ? vn : on;             // follow_the_leader.py:89-116
}

__global__ void __launch_bounds__(kEWarps * 32)
    kin_eta_ftl_kernel(ModelConst mc, long long B, const int32_t* __restrict__ prev_idx,
                       const int32_t* __restrict__ new_idx, const double* __restrict__ velocity,
                       const double* __restrict__ delta_theta,
                       const double* __restrict__ prev_g, const long long* __restrict__ prev_off,
                       double* __restrict__ eta_out, double* __restrict__ ftl_out, double* __restrict__ ftl_mag,
                       int* __restrict__ err_flag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = mc.T;
    for (long long b = (long long)blockIdx.x * kEWarps + warp; b < B; b += (long long)gridDim.x * kEWarps) {
        // ---- phase 1, lane n = tube n: its share Ad(P_n[0]) (dtheta_n e_x + v_n xi(mid)) of eta (kinematic.py:168-177).
        // The T first-node loads are in flight together (they were one exposed DRAM latency per tube when the tubes were
        // walked one after the other); eta_n = sum_{m <= n} share_m and V_n = sum_{m <= n} v_m follow by a prefix sum over lanes.
        double sh[6] = {0, 0, 0, 0, 0, 0};
        double vsum_l = 0.0;
        long long o0_l = 0;
        int pi_l = 0, cnt_l = 0;
        bool bad = false;
        if (lane < T) {
            const int n = lane;
            pi_l = prev_idx[b * T + n];
            // kinematic.py:71: velocity from the rounded indices, unless the caller passes it (solve_eta's own signature)
            const double v = velocity ? velocity[b * T + n] : (pi_l - new_idx[b * T + n]) * mc.dx;
            vsum_l = v;                                            // :163
            const double mid = pi_l * mc.dx - v / 2;               // :165
            double wm[3];
            strain_omega(mc, n, mid, wm);
            o0_l = prev_off[b * T + n];
            const long long o1 = prev_off[b * T + n + 1];
            cnt_l = mc.npts[n] - pi_l;
            bad = o1 - o0_l < cnt_l || cnt_l < 1;                  // the reference raises IndexError here
            if (!bad) {
                const SE3 p0 = load_g64(prev_g + o0_l * 16);
                // twist in the tube's base frame: dtheta e_x + v xi(mid)  (:171-172)
                const double tw[3] = {delta_theta[b * T + n] + v * wm[0], v * wm[1], v * wm[2]};
                double Rw[3], Rv[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    Rw[i] = p0.r[i * 3] * tw[0] + p0.r[i * 3 + 1] * tw[1] + p0.r[i * 3 + 2] * tw[2];
                    Rv[i] = p0.r[i * 3] * v;
                }
                // Ad(g)[w;v] = [R w ; p x (R w) + R v]
                sh[0] = Rw[0];
                sh[1] = Rw[1];
                sh[2] = Rw[2];
                sh[3] = p0.p[1] * Rw[2] - p0.p[2] * Rw[1] + Rv[0];
                sh[4] = p0.p[2] * Rw[0] - p0.p[0] * Rw[2] + Rv[1];
                sh[5] = p0.p[0] * Rw[1] - p0.p[1] * Rw[0] + Rv[2];
            }
        }
        // the first tube with too short a parent curve ends the configuration there (as the sequential walk did)
        const unsigned badmask = __ballot_sync(0xffffffffu, bad);
        const int t_ok = badmask ? __ffs(badmask) - 1 : T;
        if (badmask && lane == 0) atomicExch(err_flag, 1);
        // inclusive prefix over lanes 0..T-1 in tube order: eta_n accumulates share_0 + ... + share_n exactly as the loop did
#pragma unroll
        for (int m = 1; m < CTRB_MAX_TUBES; ++m) {
            double add[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) add[j] = __shfl_up_sync(0xffffffffu, sh[j], 1);
            const double va = __shfl_up_sync(0xffffffffu, vsum_l, 1);
            // after step m lane n holds the sum of shares max(0, n - m) .. n; adding the neighbour's running sum once per step
            // in this order reproduces ((s0 + s1) + s2) + s3
            if (lane == m && lane < T) {
#pragma unroll
                for (int j = 0; j < 6; ++j) sh[j] = add[j] + sh[j];
                vsum_l = va + vsum_l;
            }
        }
        if (lane < t_ok) {
#pragma unroll
            for (int j = 0; j < 6; ++j) eta_out[(b * T + lane) * 6 + j] = sh[j];
        }
        double sum_mag = 0.0, sum_tmag = 0.0, tip_mag = 0.0, tip_tmag = 0.0;
        long long n_nodes = 0;
        // ---- phase 2: the follow-the-leader twists of every parent node, lanes over nodes
        for (int n = 0; n < t_ok; ++n) {
            double eta[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) eta[j] = __shfl_sync(0xffffffffu, sh[j], n);
            const double vel_sum = __shfl_sync(0xffffffffu, vsum_l, n);
            const long long o0 = __shfl_sync(0xffffffffu, o0_l, n);
            const int pi = __shfl_sync(0xffffffffu, pi_l, n), cnt = __shfl_sync(0xffffffffu, cnt_l, n);
            // one parent node per lane and trip (two with CTRB_FTL_UNROLL2: both nodes' loads are in flight before either is used)
            auto node = [&](const SE3& g, int k) {                 // :178-193
                double wi[3];
                strain_omega(mc, n, mc.dx * (pi + k), wi);
                double u[3] = {eta[3] - (g.p[1] * eta[2] - g.p[2] * eta[1]),
                               eta[4] - (g.p[2] * eta[0] - g.p[0] * eta[2]),
                               eta[5] - (g.p[0] * eta[1] - g.p[1] * eta[0])};
                double f[6];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    f[j] = g.r[0 * 3 + j] * eta[0] + g.r[1 * 3 + j] * eta[1] + g.r[2 * 3 + j] * eta[2] - vel_sum * wi[j];
                    f[3 + j] = g.r[0 * 3 + j] * u[0] + g.r[1 * 3 + j] * u[1] + g.r[2 * 3 + j] * u[2];
                }
                f[3] -= vel_sum;
                if (ftl_out) {
                    double2* dst = reinterpret_cast<double2*>(ftl_out + (o0 + k) * 6);
                    dst[0] = make_double2(f[0], f[1]);
                    dst[1] = make_double2(f[2], f[3]);
                    dst[2] = make_double2(f[4], f[5]);
                }
                double m = twist_mag(f, false), tm = twist_mag(f, true);
                sum_mag += m;
                sum_tmag += tm;
                if (n == T - 1 && k == cnt - 1) {
                    tip_mag = m;
                    tip_tmag = tm;
                }
            };
#if CTRB_FTL_UNROLL2
            for (int k = lane; k < cnt; k += 64) {
                const bool two = k + 32 < cnt;
                const SE3 g = load_g64(prev_g + (o0 + k) * 16);
                const SE3 g2 = load_g64(prev_g + (o0 + (two ? k + 32 : k)) * 16);
                node(g, k);
                if (two) node(g2, k + 32);
            }
#else
            for (int k = lane; k < cnt; k += 32) node(load_g64(prev_g + (o0 + k) * 16), k);
#endif
            n_nodes += cnt;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sum_mag += __shfl_xor_sync(0xffffffffu, sum_mag, off);
            sum_tmag += __shfl_xor_sync(0xffffffffu, sum_tmag, off);
            tip_mag += __shfl_xor_sync(0xffffffffu, tip_mag, off);
            tip_tmag += __shfl_xor_sync(0xffffffffu, tip_tmag, off);
        }
        if (lane == 0 && ftl_mag) {
            double inv = n_nodes > 0 ? 1.0 / (double)n_nodes : 0.0;
            ftl_mag[b * 4 + 0] = tip_mag;
            ftl_mag[b * 4 + 1] = sum_mag * inv;
            ftl_mag[b * 4 + 2] = tip_tmag;
            ftl_mag[b * 4 + 3] = sum_tmag * inv;
        }
    }
}

// ------------------------------------------------------------------ FMA peak probe (compute-roofline denominator)
template <typename F>
__global__ void __launch_bounds__(256) fma_peak_kernel(F* out, int iters, F seed) {
    F a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + (F)(threadIdx.x + k);
    const F m = (F)0.999999, c = (F)1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = a[k] * m + c;
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = a[k] * m + c;
    }
    F s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == (F)-1) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the chain live
}

int launch_fma_peak(ctrb_ctx* ctx, int precision, double* tflops) {
    void* d = nullptr;
    int rc = ctx_scratch(ctx, 0, 1 << 20, &d);
    if (rc) return rc;
    const int grid = ctx->sm_count * 8, iters = precision == CTRB_F64 ? 4096 : 16384;
    cudaEvent_t e0, e1;
    CTRB_CUDA(cudaEventCreate(&e0));
    CTRB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CTRB_CUDA(cudaEventRecord(e0, ctx->stream));
        if (precision == CTRB_F64) fma_peak_kernel<double><<<grid, 256, 0, ctx->stream>>>((double*)d, iters, 1.0);
        else fma_peak_kernel<float><<<grid, 256, 0, ctx->stream>>>((float*)d, iters, 1.0f);
        CTRB_CUDA(cudaEventRecord(e1, ctx->stream));
        CTRB_CUDA(cudaEventSynchronize(e1));
        float ms;
        CTRB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
        ctx->launches++;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double flops = 2.0 * 16.0 * (double)iters * 256.0 * (double)grid;
    *tflops = flops / (best * 1e-3) / 1e12;
    return CTRB_OK;
}

// ------------------------------------------------------------------ host launchers (called from ctrb_api.cu)
int launch_table(ctrb_ctx* ctx, const ctrb_model* m) {
    kin_table_kernel<<<1, 32, 0, ctx->stream>>>(m->mc, m->d_tab, m->d_tab_inv, m->d_tabR32);
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

int launch_gcurve(ctrb_ctx* ctx, const ctrb_model* m, long long B, const int32_t* idx, const double* theta, int full,
                  int precision, const long long* offsets, void* out) {
    if (B <= 0) return CTRB_OK;
    const int T = m->mc.T, tot = m->mc.total_nodes;
    size_t base = (size_t)kGWarps * 32 * T * (12 * sizeof(double) + sizeof(long long)) +
                  (((size_t)kGWarps * 32 * T + 1) / 2) * sizeof(double);
    size_t with_tab = base + (precision == CTRB_F32 ? (size_t)3 * tot * sizeof(double) + (size_t)9 * tot * sizeof(float)
                                                    : (size_t)12 * tot * sizeof(double));
    int tab_in_smem = with_tab <= 100 * 1024;
    size_t smem = tab_in_smem ? with_tab : base;
    long long ngroups = (B + 31) / 32;
    long long blocks_needed = (ngroups + kGWarps - 1) / kGWarps;
    // persistent grid of resident blocks only (a fixed 4 per SM left the fp32 kernel, whose shared memory then allowed 3,
    // with a quarter of its blocks in a second wave: 4.28 TB/s).  Measured on C2 (10^6 configurations), blocks per SM ->
    // TB/s fp64 / fp32: 2 -> 5.76 / 5.55, 3 -> 5.91 / 6.12, 4 -> 5.71 / 5.63, 5+ (fp32) -> 5.63: twelve write streams
    // per SM are enough to cover the store latency and more only scatter the DRAM pages.  CTRB_OPT_GCURVE_BLOCKS overrides.
    int per_sm = 1;
    if (precision == CTRB_F64) {
        CTRB_CUDA(cudaFuncSetAttribute(kin_gcurve_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CTRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kin_gcurve_kernel<double>, kGWarps * 32, smem));
    } else {
        CTRB_CUDA(cudaFuncSetAttribute(kin_gcurve_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CTRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kin_gcurve_kernel<float>, kGWarps * 32, smem));
    }
    const int max_blocks = ctx->opt[CTRB_OPT_GCURVE_BLOCKS] > 0 ? ctx->opt[CTRB_OPT_GCURVE_BLOCKS] : 3;
    per_sm = std::max(1, std::min(per_sm, max_blocks));
    int grid = (int)std::min<long long>(blocks_needed, (long long)ctx->sm_count * per_sm);
    KernelTimer kt(ctx, CTRB_KERNEL_GCURVE);
    if (precision == CTRB_F64) {
        kin_gcurve_kernel<double><<<grid, kGWarps * 32, smem, ctx->stream>>>(m->mc, m->tables(), B, idx, theta, full, offsets,
                                                                            (double*)out, tab_in_smem);
    } else {
        kin_gcurve_kernel<float><<<grid, kGWarps * 32, smem, ctx->stream>>>(m->mc, m->tables(), B, idx, theta, full, offsets,
                                                                           (float*)out, tab_in_smem);
    }
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

int launch_calc_indices(ctrb_ctx* ctx, const ctrb_model* m, long long B, const double* prev_ins, const double* next_ins,
                        int32_t* prev_idx, int32_t* new_idx) {
    if (B <= 0) return CTRB_OK;
    long long n = B * m->mc.T;
    calc_indices_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(m->mc, B, prev_ins, next_ins, prev_idx, new_idx);
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

int launch_offsets(ctrb_ctx* ctx, const ctrb_model* m, long long B, const int32_t* idx, int full, long long* offsets) {
    long long n = B * m->mc.T;
    if (n <= 0) {
        CTRB_CUDA(cudaMemsetAsync(offsets, 0, sizeof(long long), ctx->stream));
        return CTRB_OK;
    }
    long long nblocks = (n + kScanBlock - 1) / kScanBlock;
    void* bs = nullptr;
    int rc = ctx_scratch(ctx, 0, (size_t)(nblocks + 1) * sizeof(long long), &bs);
    if (rc) return rc;
    node_count_scan1<<<(unsigned)nblocks, kScanBlock, 0, ctx->stream>>>(m->mc, n, idx, full, offsets, (long long*)bs);
    node_count_scan2<<<1, kScanBlock, 0, ctx->stream>>>(nblocks, (long long*)bs);
    node_count_scan3<<<(unsigned)nblocks, kScanBlock, 0, ctx->stream>>>(n, offsets, (const long long*)bs, nblocks);
    ctx->launches += 3;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

int launch_eta_ftl(ctrb_ctx* ctx, const ctrb_model* m, long long B, const int32_t* prev_idx, const int32_t* new_idx,
                   const double* velocity, const double* delta_theta, const double* prev_g, const long long* prev_off, double* eta_out,
                   double* ftl_out, double* ftl_mag, int* err_flag) {
    if (B <= 0) return CTRB_OK;
    long long blocks_needed = (B + kEWarps - 1) / kEWarps;
    int grid = (int)std::min<long long>(blocks_needed, (long long)ctx->sm_count * 8);
    KernelTimer kt(ctx, CTRB_KERNEL_ETA_FTL);
    kin_eta_ftl_kernel<<<grid, kEWarps * 32, 0, ctx->stream>>>(m->mc, B, prev_idx, new_idx, velocity, delta_theta, prev_g, prev_off,
                                                              eta_out, ftl_out, ftl_mag, err_flag);
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

}  // namespace ctrb
