// ctrb_internal.cuh -- shared declarations of libctrb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "ctrb200.h"

namespace ctrb {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
#define CTRB_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            ctrb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                            cudaGetErrorString(e__));                                        \
            return CTRB_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)
#define CTRB_REQUIRE(cond, ...)             \
    do {                                    \
        if (!(cond)) {                      \
            ctrb::set_error(__VA_ARGS__);   \
            return CTRB_ERR_INVALID;        \
        }                                   \
    } while (0)

// ---------------------------------------------------------------- model constants (kernel parameter, by value)
struct ModelConst {
    int T;
    int kind[CTRB_MAX_TUBES];
    int dof[CTRB_MAX_TUBES];
    int npts[CTRB_MAX_TUBES];       // num_discrete_points, model.py:37-38
    int node_base[CTRB_MAX_TUBES];  // first table column of tube n
    int total_nodes;
    double q[CTRB_MAX_TUBES][CTRB_MAX_DOF];
    double L[CTRB_MAX_TUBES];
    double dx;
};

// Table of cumulative transforms C_n[j] = E_n[1] ... E_n[j] (C_n[0] = I), where
// E_n[i] = exp(dx * hat(B_n((i-1/2)dx) q_n + bias)) is the configuration-independent
// per-node step of kinematic.py:106-118.  SoA: tab[k * total_nodes + node_base[n] + j],
// k = 0..8 rotation row-major, 9..11 position.  tab_inv holds C_n[j]^-1.
struct ModelTables {
    const double* tab;
    const double* tab_inv;
    const float* tabR32;  // rotation part of tab in fp32 (9 rows)
};

constexpr int kStaticMaxNq = 30;
// static model constants (kernel parameter, by value); see ctrb_static.cu
struct StModel {
    ModelConst mc;
    double Kr[CTRB_MAX_TUBES][3];  // rotational stiffness diag (static.py:561-585): G*2I, E*I, E*I
    int nq;
    int ndof[6];                   // order (1,1),(2,1),(3,1),(2,2),(3,2),(3,3)  (static.py:63)
    int qoff[6];
    double x0[kStaticMaxNq];       // get_init_vals (static_bases.py:142-168)
    // per-design tables (device memory, built once at ctrb_model_create by static_tables_kernel):
    const double* ws_tab;          // [total_nodes][3]: omega of the tube's own pre-curvature ksi* at node j (get_ksi_star)
    const double* b3_tab;          // [npts_last][CTRB_MAX_DOF + 3]: last tube's strain base at node j, one value per
                                   // column (every base has one non-zero per column), then (B q*)[0:3]
    int b3_row[CTRB_MAX_DOF];      // the row of that non-zero
    const struct StQuad* quad_tab; // [max npts + 1]: calculate_integral's brackets for a section of np nodes (st_setup)
    int quad_max;                  // largest np in the table
    const double* c33_tab;         // [npts_last]: block5_norm2 (the implicit last block's share of |D x|^2 from the default
                                   // initial guess) by insertion index of the last tube -- it depends on nothing else
};

// calculate_integral (static.py:456-491) for a section of np nodes: the two abscissae's bracketing nodes, interp1d's
// (x_new - x_lo) / (x_hi - x_lo), and the section length.  Depends on (np, delta_x) only: tabulated per design.
struct StQuad {
    int lo[2], hi[2];
    double wq[2], len;
};

}  // namespace ctrb

struct ctrb_ctx {
    int device;
    cudaStream_t stream;
    cudaStream_t own_stream;
    cudaStream_t stream2;  // second stream for host-memory chunk pipelining
    cudaEvent_t ev[2];
    cudaEvent_t ev_chunk[2];  // host-memory pipeline: chunk inputs have landed (recorded on stream2)
    cudaEvent_t ev_done[2];   //                       chunk kernels have finished (recorded on stream)
    cudaEvent_t kev[8][2];  // per-kernel start/stop of the most recent launch (ctrb_kernel_id)
    int64_t launches;
    // grow-only device scratch
    void* scratch[32];        // [1..8] Stage buffers, [9..15] per-call scratch, [17..24] second set of Stage buffers
    size_t scratch_bytes[32];
    unsigned long long* d_counters;  // [32]: [16..21] why the fast kernel handed configurations over (ctrb_static_last_fallback_reasons); [0..7] collision work counters + ticket, [8..10] static-solve tickets + fallback count, [11] fallbacks summed over the chunks of one call
    int sm_count;
    int opt[16];              // ctrb_option values (ctrb_ctx_set_option); defaults set by ctrb_ctx_create
};

struct ctrb_model {
    ctrb_model_desc desc;
    ctrb::ModelConst mc;
    int device;
    double* d_tab;
    double* d_tab_inv;
    float* d_tabR32;
    int is_static;
    ctrb::StModel sm;
    double* d_ws_tab;
    double* d_b3_tab;
    ctrb::StQuad* d_quad_tab;
    double* d_c33_tab;
    ctrb::ModelTables tables() const { return ctrb::ModelTables{d_tab, d_tab_inv, d_tabR32}; }
};

namespace ctrb {

// ---------------------------------------------------------------- BVH (host build, device traversal)
// 64-byte node: both children's fp32 AABBs (rounded outward) + two child links.
// link >= 0: inner node index;  link < 0: leaf, ~link = (first_tri << 3) | (count - 1), count <= 8.
struct alignas(16) BvhNode {
    float lo0[3], hi0[3];
    float lo1[3], hi1[3];
    int32_t c0, c1;
    int32_t pad[2];
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

struct HostBvh {
    std::vector<BvhNode> nodes;     // BFS order, root = 0 (empty when ntri == 0)
    std::vector<double> tri64;      // [ntri][9] in leaf order
    std::vector<float> tri32;       // [ntri][12]: v0.xyz,0, v1.xyz,0, v2.xyz,0 (rounded to nearest)
    int ntri = 0, nleaves = 0, depth = 0;
    int root_leaf_link = 0;         // used when the whole mesh is one leaf (nodes has a single dummy)
    float slack = 0.f;              // absolute fp32 error budget for every fp32 distance bound
};
void build_bvh(const double* tris, int ntri, HostBvh* out);

struct DevBvh {
    const BvhNode* nodes;
    const float4* tri32;
    const double* tri64;
    int nnodes;
    int ntri;
    float slack;
};

struct DevPrim {
    int shape;
    double center[3];
    double rot[9];
    double dims[3];
};

constexpr int kMaxPrims = 16;

struct EnvConst {
    DevBvh obs;
    DevBvh goal_mesh;
    int nprim;
    DevPrim prims[kMaxPrims];
    DevPrim goal;  // shape NONE when absent
};

}  // namespace ctrb

struct ctrb_env {
    int device;
    ctrb::EnvConst ec;
    ctrb::EnvConst* d_ec;  // device copy (too big for comfortable by-value passing with 16 prims)
    void* d_nodes[2];
    void* d_tri32[2];
    void* d_tri64[2];
    int32_t stats[4];
};

namespace ctrb {
// nearest-neighbour query parameters (ctrb_knn.cu)
constexpr int kKnnMaxDim = 2 * CTRB_MAX_TUBES;
struct KnnParams {
    int dim, k;
    long long N, Q;
    double w[kKnnMaxDim];
    int topo[kKnnMaxDim];
};
int launch_knn(ctrb_ctx* ctx, const KnnParams& P, const double* pts, const double* qry, int32_t* idx_out, double* dist_out);

// records the bracketing events of one of the four hot kernels on the launching stream
struct KernelTimer {
    ctrb_ctx* c;
    int id;
    KernelTimer(ctrb_ctx* ctx, int kid) : c(ctx), id(kid) { cudaEventRecord(c->kev[id][0], c->stream); }
    ~KernelTimer() { cudaEventRecord(c->kev[id][1], c->stream); }
};
// scratch helper: returns device pointer of at least `bytes` in slot `slot`
int ctx_scratch(ctrb_ctx* ctx, int slot, size_t bytes, void** out);
}  // namespace ctrb
