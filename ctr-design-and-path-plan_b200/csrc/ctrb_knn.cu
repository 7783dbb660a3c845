// ctrb_knn.cu -- nearest-neighbour queries of the planner's node store (ctrdapp/solve/dynamic_tree.py:22-120,
// which delegates to pympnn.KDTree_topologies: k-d trees over a product of R^1 and S^1 factors).
//
// The tree holds at most iteration_number + 2 points (<= a few thousand 2T-dimensional points), so a batched
// brute-force scan beats a k-d tree on this hardware: one warp per query, lanes stride over the points, and
// the k <= 32 best candidates live one per lane in registers (sorted; an insertion is one ballot + one shuffle).
// Metric: sqrt( sum_i (w_i * delta_i)^2 ) with delta_i = |a - b| on linear dimensions (topology 1) and
// min(|a - b|, 2 pi - |a - b|) on circular ones (topology 2).  Ties are broken by the lower point index, so
// results do not depend on the scan order.
#include <math_constants.h>

#include "ctrb_internal.cuh"

namespace ctrb {

constexpr int kKnnWarps = 8;

__global__ void __launch_bounds__(kKnnWarps * 32)
    knn_kernel(const __grid_constant__ KnnParams P, const double* __restrict__ pts, const double* __restrict__ qry,
               int32_t* __restrict__ idx_out, double* __restrict__ dist_out) {
    const int lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * kKnnWarps + (threadIdx.x >> 5);
    if (q >= P.Q) return;
    double x[kKnnMaxDim];
#pragma unroll
    for (int d = 0; d < kKnnMaxDim; ++d) x[d] = d < P.dim ? qry[q * P.dim + d] : 0.0;
    // lane j holds the j-th best (distance^2, index) so far
    double best_d = CUDART_INF;
    long long best_i = 0x7fffffffffffffffLL;
    const unsigned kmask = P.k >= 32 ? 0xffffffffu : ((1u << P.k) - 1u);
    for (long long base = 0; base < P.N; base += 32) {
        const long long i = base + lane;
        double d2 = CUDART_INF;
        if (i < P.N) {
            d2 = 0.0;
#pragma unroll
            for (int d = 0; d < kKnnMaxDim; ++d) {
                if (d < P.dim) {
                    double a = fabs(pts[i * P.dim + d] - x[d]);
                    if (P.topo[d] == 2) a = fmin(a, fabs(2.0 * CUDART_PI - a));
                    a *= P.w[d];
                    d2 = fma(a, a, d2);
                }
            }
        }
        const double kth_d = __shfl_sync(0xffffffffu, best_d, P.k - 1);
        const long long kth_i = __shfl_sync(0xffffffffu, best_i, P.k - 1);
        unsigned cand = __ballot_sync(0xffffffffu, i < P.N && (d2 < kth_d || (d2 == kth_d && i < kth_i)));
        while (cand) {
            const int src = __ffs(cand) - 1;
            cand &= cand - 1;
            const double nd = __shfl_sync(0xffffffffu, d2, src);
            const long long ni = base + src;
            // position of the newcomer = number of kept entries that sort before it
            const bool before = best_d < nd || (best_d == nd && best_i < ni);
            const int pos = __popc(__ballot_sync(0xffffffffu, before) & kmask);
            const double up_d = __shfl_up_sync(0xffffffffu, best_d, 1);
            const long long up_i = __shfl_up_sync(0xffffffffu, best_i, 1);
            if (pos < P.k) {
                if (lane == pos) {
                    best_d = nd;
                    best_i = ni;
                } else if (lane > pos) {
                    best_d = up_d;
                    best_i = up_i;
                }
            }
        }
    }
    if (lane < P.k) {
        const bool valid = best_d < CUDART_INF;
        idx_out[q * P.k + lane] = valid ? (int32_t)best_i : -1;
        dist_out[q * P.k + lane] = valid ? sqrt(best_d) : CUDART_INF;
    }
}

int launch_knn(ctrb_ctx* ctx, const KnnParams& P, const double* pts, const double* qry, int32_t* idx_out, double* dist_out) {
    if (P.Q <= 0) return CTRB_OK;
    const unsigned grid = (unsigned)((P.Q + kKnnWarps - 1) / kKnnWarps);
    knn_kernel<<<grid, kKnnWarps * 32, 0, ctx->stream>>>(P, pts, qry, idx_out, dist_out);
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

}  // namespace ctrb
