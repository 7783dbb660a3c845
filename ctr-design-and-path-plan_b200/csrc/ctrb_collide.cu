// ctrb_collide.cu -- tube-vs-environment distance / collision and the fused
// g-curve + collision + cost kernels (sm_100a).
//
// Replaces CollisionChecker.check_collision (collision_checker.py:35-95) and, in the fused
// variant, the solve_g -> check_collision -> heuristic chain of rrt.py:110-129.
//
// Every pair of consecutive nodes is a solid flat-capped cylinder (collision_checker.py:121-147).
// Two schedules of the same exact query:
//   chain_eval_group_kernel  a warp takes a group of configurations (lane = configuration for the
//                            positions and results, lane = cylinder for the BVH walks); the default
//   chain_eval_kernel        one warp per configuration; for configurations that do not fit the
//                            grouped kernel's node pool / stack, and as its cross-check
// In both, lanes walk a BVH over the triangle soup for one cylinder at a time and share, per
// configuration, one upper bound on the minimum distance (shared-memory atomicMin on the float
// bits) and one contact flag that ends the whole query (python-fcl's defaultDistanceCallback
// stops at the first contact); surviving (cylinder, triangle) pairs are queued and decided by the
// warp in lock-step.  In the fused variants p_j = A_n C_n[j].p comes straight from the model
// table: the g-curve never touches HBM.
//
// Culling is fp32 (boxes rounded outward, `slack` absorbs every fp32 rounding), the
// decisive cylinder<->triangle distance is fp64 and exact (ctrb_geom.cuh).  No tensor
// cores: there is no dense contraction anywhere on this path.
#include <float.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ctrb_geom.cuh"
#include "ctrb_se3.cuh"

namespace ctrb {

#ifndef CTRB_EVAL_WARPS
#define CTRB_EVAL_WARPS 8
#endif
#ifndef CTRB_EVAL_MINBLOCKS
#define CTRB_EVAL_MINBLOCKS 2
#endif
#ifndef CTRB_TOP_NODES
#define CTRB_TOP_NODES 256
#endif
constexpr int kCWarps = CTRB_EVAL_WARPS;      // warps (= configurations in flight) per block
constexpr int kTopNodes = CTRB_TOP_NODES;     // BVH nodes staged in shared memory (BFS prefix = top levels)
constexpr int kStack = 48;

constexpr int kQueue = 96;  // deferred exact tests per warp (drained 32 at a time, all lanes in lock-step)
#ifndef CTRB_STEPS_PER_ROUND
#define CTRB_STEPS_PER_ROUND 8
#endif
constexpr int kStepsPerRound = CTRB_STEPS_PER_ROUND;  // BVH node steps a lane takes between two warp-wide votes
struct WarpCtl {
    int bound_bits;   // float bits of an upper bound on the config's min distance (>= 0 => int-ordered)
    int collided;
    int ticket;
    int qcount;
    int q_tri[kQueue];    // triangle (leaf order) of a candidate pair
    int q_item[kQueue];   // its cylinder
    float q_lb[kQueue];   // fp32 lower bound on their distance
};

struct EvalParams {
    ModelConst mc;
    ModelTables tb;
    const EnvConst* env;
    long long B;
    const int32_t* idx;
    const double* theta;
    // non-fused input: node k of a tube at g + (offsets[..] + k) * g_stride, its position component j at
    // [g_poff + j * g_pstep] (4x4 row-major nodes: 16, 3, 4; position-only nodes: 3, 0, 1)
    const double* g;
    const long long* offsets;
    int g_stride, g_poff, g_pstep;
    int tube_num;
    double rad[CTRB_MAX_TUBES];
    // outputs
    double* obstacle_min;
    double* goal_dist;
    double* cost;
    uint8_t* collide;
    ctrb_cost_desc cd;
    unsigned long long* counters;  // [0..3] work counters, [4] config ticket
    int max_nodes;                 // shared-memory rows per warp
    int bvh_depth;                 // of the obstacle mesh's BVH (host-side knowledge: picks the traversal kernel)
    int chunk;                     // grouped kernel: configurations a warp takes per ticket (<= 32)
    int pool;                      // grouped kernel: node positions per warp
};

__device__ __forceinline__ BvhNode load_node(const DevBvh& bvh, const BvhNode* s_top, int ntop, int ni) {
    if (ni < ntop) return s_top[ni];
    BvhNode n;
    const float4* p = reinterpret_cast<const float4*>(bvh.nodes + ni);
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    n.lo0[0] = a.x; n.lo0[1] = a.y; n.lo0[2] = a.z; n.hi0[0] = a.w;
    n.hi0[1] = b.x; n.hi0[2] = b.y; n.lo1[0] = b.z; n.lo1[1] = b.w;
    n.lo1[2] = c.x; n.hi1[0] = c.y; n.hi1[1] = c.z; n.hi1[2] = c.w;
    n.c0 = __float_as_int(d.x);
    n.c1 = __float_as_int(d.y);
    return n;
}

struct Work {
    unsigned int nodes, pre, exact;
};

// exact nearest distance from a point to a mesh (goal mesh): fp32 culling, fp64 leaves
__device__ __noinline__ double point_vs_mesh(const DevBvh& bvh, d3 x, Work& wk) {
    const f3 xf = {(float)x.x, (float)x.y, (float)x.z};
    const float slack = bvh.slack;
    double best = DBL_MAX;
    float bound = FLT_MAX;
    int stack[kStack];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const int ni = stack[--sp];
        const BvhNode nd = load_node(bvh, nullptr, 0, ni);
        wk.nodes++;
        float reach = bound + slack;
        float reach2 = reach < 1e18f ? reach * reach * 1.000001f : FLT_MAX;
        float d0 = aabb_dist2(nd.lo0, nd.hi0, xf);
        float d1 = aabb_dist2(nd.lo1, nd.hi1, xf);
        int l0 = nd.c0, l1 = nd.c1;
        if (d1 < d0) {
            float t = d0; d0 = d1; d1 = t;
            int ti = l0; l0 = l1; l1 = ti;
        }
        if (l1 >= 0 && d1 <= reach2 && sp < kStack) stack[sp++] = l1;
        if (l0 >= 0 && d0 <= reach2 && sp < kStack) stack[sp++] = l0;
        for (int side = 0; side < 2; ++side) {
            const int link = side == 0 ? l0 : l1;
            const float dd = side == 0 ? d0 : d1;
            if (link >= 0 || dd > reach2) continue;
            const int code = ~link;
            const int first = code >> 3, cnt = (code & 7) + 1;
            for (int t = first; t < first + cnt; ++t) {
                const double* tq = bvh.tri64 + (size_t)t * 9;
                d3 t0 = {__ldg(tq + 0), __ldg(tq + 1), __ldg(tq + 2)};
                d3 t1 = {__ldg(tq + 3), __ldg(tq + 4), __ldg(tq + 5)};
                d3 t2 = {__ldg(tq + 6), __ldg(tq + 7), __ldg(tq + 8)};
                double d = sqrt(point_tri_dist2<d3, double>(x, t0, t1, t2));
                wk.exact++;
                if (d < best) {
                    best = d;
                    bound = __double2float_ru(d);
                }
            }
        }
    }
    return best;
}

__device__ void box_corner(const DevPrim& b, int k, d3& out) {
    double s[3] = {(k & 1) ? 0.5 : -0.5, (k & 2) ? 0.5 : -0.5, (k & 4) ? 0.5 : -0.5};
    double o[3];
    for (int i = 0; i < 3; ++i) {
        o[i] = b.center[i];
        for (int ax = 0; ax < 3; ++ax) o[i] += b.rot[i * 3 + ax] * s[ax] * b.dims[ax];
    }
    out = {o[0], o[1], o[2]};
}

// solid cylinder vs primitive obstacle (fcl.Sphere / fcl.Cylinder / fcl.Box, init_collision.py:129-139)
__device__ __noinline__ double cylinder_vs_prim(const DevPrim& o, d3 c, d3 a, double h, double r, double bound) {
    if (o.shape == CTRB_SHAPE_SPHERE) {
        double d = point_cyl_dist(d3{o.center[0], o.center[1], o.center[2]}, c, a, h, r) - o.dims[0];
        return d < 0.0 ? 0.0 : d;
    }
    if (o.shape == CTRB_SHAPE_BOX) {
        if (point_box_dist(c, o.center, o.rot, o.dims) == 0.0) return 0.0;
        const int F[12][3] = {{0, 1, 3}, {0, 3, 2}, {4, 7, 5}, {4, 6, 7}, {0, 5, 1}, {0, 4, 5},
                              {2, 3, 7}, {2, 7, 6}, {0, 2, 6}, {0, 6, 4}, {1, 5, 7}, {1, 7, 3}};
        double best = DBL_MAX;
        for (int f = 0; f < 12; ++f) {
            d3 p0, p1, p2;
            box_corner(o, F[f][0], p0);
            box_corner(o, F[f][1], p1);
            box_corner(o, F[f][2], p2);
            double d = cyl_tri_dist(c, a, h, r, p0, p1, p2, fmin(best, bound));
            if (d < best) best = d;
            if (best == 0.0) break;
        }
        return best;
    }
    if (o.shape == CTRB_SHAPE_CYLINDER) {  // axis = local z column of rotate_align('direction')
        return cyl_cyl_dist(c, a, h, r, d3{o.center[0], o.center[1], o.center[2]}, d3{o.rot[2], o.rot[5], o.rot[8]},
                            0.5 * o.dims[1], o.dims[0]);
    }
    return DBL_MAX;
}

template <bool FUSED>
__global__ void __launch_bounds__(kCWarps * 32, CTRB_EVAL_MINBLOCKS) chain_eval_kernel(const __grid_constant__ EvalParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BvhNode* s_top = reinterpret_cast<BvhNode*>(smem_raw);
    double* s_pos_all = reinterpret_cast<double*>(smem_raw + (size_t)kTopNodes * sizeof(BvhNode));
    WarpCtl* s_ctl_all = reinterpret_cast<WarpCtl*>(s_pos_all + (size_t)kCWarps * P.max_nodes * 3);

    const EnvConst& env = *P.env;
    const int ntop = min(env.obs.nnodes, kTopNodes);
    for (int i = threadIdx.x; i < ntop * 4; i += blockDim.x)
        reinterpret_cast<float4*>(s_top)[i] = __ldg(reinterpret_cast<const float4*>(env.obs.nodes) + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* s_pos = s_pos_all + (size_t)warp * P.max_nodes * 3;
    WarpCtl* ctl = s_ctl_all + warp;
    const int T = FUSED ? P.mc.T : P.tube_num;
    Work wk = {0, 0, 0};
    unsigned int nconf = 0;

    for (;;) {
        // dynamic schedule: configurations differ wildly in cost (first-contact exit vs full distance)
        long long b;
        {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(P.counters + 4, 1ULL);
            b = (long long)__shfl_sync(0xffffffffu, t, 0);
        }
        if (b >= P.B) break;
        nconf++;

        // ---- stage A: node positions of this configuration into shared memory
        int cnt[CTRB_MAX_TUBES], first[CTRB_MAX_TUBES];
        int total = 0;
        if (FUSED) {
            const int tot = P.mc.total_nodes;
            SE3 g;
            se3_identity(g);
            for (int n = 0; n < T; ++n) {
                const int id = P.idx[b * T + n];
                se3_rot_x(g, P.theta[b * T + n]);
                SE3 a = se3_mul(g, load_tab(P.tb.tab_inv, tot, P.mc.node_base[n] + id));
                cnt[n] = P.mc.npts[n] - id;  // kinematic.py:100-103 (full=False)
                first[n] = total;
                const int colbase = P.mc.node_base[n] + id;
                for (int k = lane; k < cnt[n]; k += 32) {
                    const int col = colbase + k;
                    double cx = __ldg(P.tb.tab + 9 * tot + col), cy = __ldg(P.tb.tab + 10 * tot + col),
                           cz = __ldg(P.tb.tab + 11 * tot + col);
                    double* dst = s_pos + (size_t)(total + k) * 3;
                    dst[0] = fma(a.r[0], cx, fma(a.r[1], cy, fma(a.r[2], cz, a.p[0])));
                    dst[1] = fma(a.r[3], cx, fma(a.r[4], cy, fma(a.r[5], cz, a.p[1])));
                    dst[2] = fma(a.r[6], cx, fma(a.r[7], cy, fma(a.r[8], cz, a.p[2])));
                }
                total += cnt[n];
                // tube tip frame = A * C[N-1]
                g = se3_mul(a, load_tab(P.tb.tab, tot, P.mc.node_base[n] + P.mc.npts[n] - 1));
            }
        } else {
            for (int n = 0; n < T; ++n) {
                const long long o0 = P.offsets[b * T + n], o1 = P.offsets[b * T + n + 1];
                cnt[n] = (int)(o1 - o0);
                if (total + cnt[n] > P.max_nodes) cnt[n] = max(0, P.max_nodes - total);  // guarded on the host
                first[n] = total;
                for (int k = lane; k < cnt[n]; k += 32) {
                    const double* gg = P.g + (o0 + k) * P.g_stride + P.g_poff;
                    double* dst = s_pos + (size_t)(total + k) * 3;
                    dst[0] = __ldg(gg);
                    dst[1] = __ldg(gg + P.g_pstep);
                    dst[2] = __ldg(gg + 2 * P.g_pstep);
                }
                total += cnt[n];
            }
        }
        if (lane == 0) {
            ctl->bound_bits = __float_as_int(FLT_MAX);
            ctl->collided = 0;
            ctl->ticket = 0;
            ctl->qcount = 0;
        }
        __syncwarp();

        // ---- stage B: cylinders.  item k of tube n joins nodes first[n]+k and first[n]+k+1
        int ncyl = 0;
        int cyl_first[CTRB_MAX_TUBES];
        for (int n = 0; n < T; ++n) {
            cyl_first[n] = ncyl;
            ncyl += max(cnt[n] - 1, 0);
        }
        double best = DBL_MAX;
        const bool have_obstacles = env.obs.ntri > 0 || env.nprim > 0;
        if (have_obstacles) {
            // Every lane walks the BVH for one cylinder at a time (private stack, fp32 culling against the bound all
            // lanes share).  The decisive fp64 cylinder<->triangle tests are NOT done inside the walk, where each lane
            // would run the ~1.5 k-instruction routine alone while the others wait (ncu: 3.6 of 32 lanes active): a
            // surviving (cylinder, triangle) pair goes into a shared-memory queue and the whole warp drains the queue
            // in lock-step, one pair per lane.
            volatile int* vbound = &ctl->bound_bits;
            volatile int* vcoll = &ctl->collided;
            const DevBvh& bvh = env.obs;
            const float slack = bvh.slack;
            auto cylinder_of = [&](int item, d3& c, d3& a, double& h, double& r) {
                int n = 0;
#pragma unroll
                for (int q = 1; q < CTRB_MAX_TUBES; ++q)
                    if (q < T && item >= cyl_first[q]) n = q;
                const double* pf = s_pos + (size_t)(first[n] + (item - cyl_first[n])) * 3;
                d3 f = {pf[0], pf[1], pf[2]}, t = {pf[3], pf[4], pf[5]};
                d3 vec = t - f;
                const double len = sqrt(dot(vec, vec));    // collision_checker.py:130-131
                c = f + 0.5 * vec;                          // mid_point
                a = (1.0 / len) * vec;                      // unit_vec -> cylinder axis
                h = 0.5 * len;
                r = P.rad[n];
            };
            auto exact_pair = [&](int item, int t, float bound) {
                d3 c, a;
                double h, r;
                cylinder_of(item, c, a, h, r);
                const double* tq = bvh.tri64 + (size_t)t * 9;
                d3 t0 = {__ldg(tq + 0), __ldg(tq + 1), __ldg(tq + 2)};
                d3 t1 = {__ldg(tq + 3), __ldg(tq + 4), __ldg(tq + 5)};
                d3 t2 = {__ldg(tq + 6), __ldg(tq + 7), __ldg(tq + 8)};
                const double d = cyl_tri_dist(c, a, h, r, t0, t1, t2, (double)bound);
                wk.exact++;
                if (d <= 0.0) {
                    *vcoll = 1;
                } else if (d < best) {
                    best = d;
                    atomicMin(&ctl->bound_bits, __float_as_int(__double2float_ru(d)));
                }
            };
            // every decision that guards a __syncwarp is taken on lane 0's view (lanes of a warp do not run in
            // lock-step: a flag read lane by lane can differ, and lanes would wait at different barriers)
            auto drain = [&]() {
                __syncwarp();
                const int qc = min(__shfl_sync(0xffffffffu, *(volatile int*)&ctl->qcount, 0), kQueue);
                for (int base = 0; base < qc; base += 32) {
                    if (__shfl_sync(0xffffffffu, *vcoll, 0)) break;
                    const int k = base + lane;
                    const float bound = __int_as_float(*vbound);
                    if (k < qc && ctl->q_lb[k] <= bound) exact_pair(ctl->q_item[k], ctl->q_tri[k], bound);
                    __syncwarp();
                }
                __syncwarp();
                if (lane == 0) ctl->qcount = 0;
                __syncwarp();
            };
            int stack[kStack];
            int sp = 0, item = 0;
            bool exhausted = false;  // this lane has seen the end of the cylinder ticket
            f3 cf = {0.f, 0.f, 0.f};
            float rb = 0.f, rin = 0.f;
            for (;;) {
                if (sp == 0 && !exhausted && !*vcoll) {  // next cylinder for this lane
                    item = atomicAdd(&ctl->ticket, 1);
                    exhausted = item >= ncyl;
                    if (!exhausted) {
                        d3 c, a;
                        double h, r;
                        cylinder_of(item, c, a, h, r);
                        for (int k = 0; k < env.nprim && !*vcoll; ++k) {
                            const double d = cylinder_vs_prim(env.prims[k], c, a, h, r, best);
                            if (d <= 0.0) *vcoll = 1;
                            else if (d < best) {
                                best = d;
                                atomicMin(&ctl->bound_bits, __float_as_int(__double2float_ru(d)));
                            }
                        }
                        if (bvh.ntri > 0 && !*vcoll) {
                            cf = f3{(float)c.x, (float)c.y, (float)c.z};
                            rb = __fsqrt_ru(__fmaf_ru((float)h, (float)h, (float)r * (float)r)) * 1.000001f;
                            rin = fminf((float)h, (float)r);
                            stack[sp++] = 0;
                        }
                    }
                }
                if (!__any_sync(0xffffffffu, sp > 0 || (!exhausted && !*vcoll))) break;
#pragma unroll 1
                for (int rep = 0; rep < kStepsPerRound && sp > 0; ++rep) {  // a few node steps between two warp votes
                    if (*vcoll) {
                        sp = 0;
                    } else {
                        const int ni = stack[--sp];
                        const BvhNode nd = load_node(bvh, s_top, ntop, ni);
                        wk.nodes++;
                        float bound = __int_as_float(*vbound);
                        const float reach = bound + rb + slack;
                        const float reach2 = reach * reach * 1.000001f;
                        float d0 = aabb_dist2(nd.lo0, nd.hi0, cf);
                        float d1 = aabb_dist2(nd.lo1, nd.hi1, cf);
                        int l0 = nd.c0, l1 = nd.c1;
                        if (d1 < d0) {  // near child first
                            float t = d0; d0 = d1; d1 = t;
                            int ti = l0; l0 = l1; l1 = ti;
                        }
                        if (l1 >= 0 && d1 <= reach2 && sp < kStack) stack[sp++] = l1;
                        if (l0 >= 0 && d0 <= reach2 && sp < kStack) stack[sp++] = l0;
#pragma unroll 1
                        for (int side = 0; side < 2; ++side) {
                            const int link = side == 0 ? l0 : l1;
                            const float dd = side == 0 ? d0 : d1;
                            if (link >= 0 || dd > reach2) continue;
                            const int code = ~link;
                            const int tfirst = code >> 3, tcnt = (code & 7) + 1;
#pragma unroll 1
                            for (int t = tfirst; t < tfirst + tcnt && !*vcoll; ++t) {
                                const float4* tp = bvh.tri32 + (size_t)t * 3;
                                float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
                                const float dpt = sqrtf(point_tri_dist2<f3, float>(cf, f3{v0.x, v0.y, v0.z}, f3{v1.x, v1.y, v1.z},
                                                                                    f3{v2.x, v2.y, v2.z}));
                                wk.pre++;
                                if (dpt + slack < rin) {  // a ball of radius min(h, r) about the centre lies inside the cylinder
                                    *vcoll = 1;
                                    break;
                                }
                                // the centre belongs to the cylinder: d(cyl, tri) <= dpt
                                atomicMin(&ctl->bound_bits, __float_as_int(dpt + slack));
                                bound = __int_as_float(*vbound);
                                const float lb = dpt - rb - slack;  // d(cyl, tri) >= dpt - rb
                                if (lb > bound) continue;
                                const int slot = atomicAdd(&ctl->qcount, 1);
                                if (slot < kQueue) {
                                    ctl->q_tri[slot] = t;
                                    ctl->q_item[slot] = item;
                                    ctl->q_lb[slot] = lb;
                                } else {
                                    exact_pair(item, t, bound);  // queue full: decide here (rare)
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (__shfl_sync(0xffffffffu, *(volatile int*)&ctl->qcount, 0) >= 32) drain();
            }
            drain();
        }
        __syncwarp();
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, off));
        const bool collided = *(volatile int*)&ctl->collided != 0;

        // ---- goal: Sphere(rad[-1]) at the last node of the last tube (collision_checker.py:62-68)
        if (lane == 0) {
            double gd = DBL_MAX;
            if (total > 0 && env.goal.shape != CTRB_SHAPE_NONE) {
                // curve[-1][-1]: last node of the last tube that has nodes
                const double* pt = s_pos + (size_t)(total - 1) * 3;
                d3 tip = {pt[0], pt[1], pt[2]};
                const double tr = P.rad[T - 1];
                const DevPrim& gl = env.goal;
                if (gl.shape == CTRB_SHAPE_SPHERE) {
                    d3 dd = tip - d3{gl.center[0], gl.center[1], gl.center[2]};
                    gd = sqrt(dot(dd, dd)) - gl.dims[0] - tr;
                } else if (gl.shape == CTRB_SHAPE_BOX) {
                    gd = point_box_dist(tip, gl.center, gl.rot, gl.dims) - tr;
                } else if (gl.shape == CTRB_SHAPE_CYLINDER) {
                    d3 ax = {gl.rot[2], gl.rot[5], gl.rot[8]};
                    gd = point_cyl_dist(tip, d3{gl.center[0], gl.center[1], gl.center[2]}, ax, 0.5 * gl.dims[1], gl.dims[0]) - tr;
                } else if (gl.shape == CTRB_SHAPE_MESH && env.goal_mesh.ntri > 0) {
                    gd = point_vs_mesh(env.goal_mesh, tip, wk) - tr;
                }
                if (gd <= 0.0) gd = -1.0;
            }
            const double om = collided ? -1.0 : (have_obstacles ? best : DBL_MAX);
            P.obstacle_min[b] = om;
            P.goal_dist[b] = gd;
            if (P.collide) P.collide[b] = collided ? 1 : 0;
            if (P.cost) {
                // rrt.py:116-129: a colliding configuration is discarded; at the goal the distance clamps to 0
                double cst = CUDART_NAN;
                if (!collided) {
                    const double g0 = gd < 0.0 ? 0.0 : gd;
                    if (P.cd.cost_type == CTRB_COST_SOAPWG) {
                        const double inv = 1.0 / om;  // obstacles_and_goal.py:39-45,87-88
                        cst = g0 * P.cd.goal_weight + inv * inv;
                    } else if (P.cd.cost_type == CTRB_COST_ONLY_GOAL) {
                        cst = g0;                      // only_goal_distance.py:20-24
                    } else {
                        cst = 0.0;                     // FTL family: the twist term comes from ctrb_eta_ftl_batch
                    }
                }
                P.cost[b] = cst;
            }
        }
        __syncwarp();
    }
    // work counters
    for (int off = 16; off > 0; off >>= 1) {
        wk.nodes += __shfl_xor_sync(0xffffffffu, wk.nodes, off);
        wk.pre += __shfl_xor_sync(0xffffffffu, wk.pre, off);
        wk.exact += __shfl_xor_sync(0xffffffffu, wk.exact, off);
    }
    if (lane == 0) {
        atomicAdd(P.counters + 0, (unsigned long long)wk.nodes);
        atomicAdd(P.counters + 1, (unsigned long long)wk.pre);
        atomicAdd(P.counters + 2, (unsigned long long)wk.exact);
        atomicAdd(P.counters + 3, (unsigned long long)nconf);
    }
}

// ------------------------------------------------------------------ grouped variant
// chain_eval_kernel above gives a whole warp to one configuration, which leaves most lanes idle when a
// configuration has few cylinders (ncu, C3 "shallow": 8 of 32 lanes active in the walk, 1.8 in the leaf tests).
// chain_eval_group_kernel lets a warp take as many configurations as fit its node pool (up to 32):
//   lane = configuration for the SE(3) chain, the node positions and the epilogue (coalesced result stores);
//   lane = work item for the BVH walks: the cylinders of the whole group are one item list behind a shared-memory
//   ticket; bounds, contact flags and minima are per configuration slot.
// The walk is "while-while": a lane pops inner nodes until a LEAF link surfaces (leaf links are pushed like inner
// nodes), then all lanes that hold
// a leaf run the fp32 point-triangle prefilter together; surviving (cylinder, triangle) pairs go to the warp's queue
// and are decided by the exact fp64 test 32 at a time.
#ifndef CTRB_GROUP_WARPS
#define CTRB_GROUP_WARPS 8
#endif
#ifndef CTRB_GROUP_BLOCKS
#define CTRB_GROUP_BLOCKS 2
#endif
#ifndef CTRB_GROUP_QUEUE
#define CTRB_GROUP_QUEUE 64
#endif
constexpr int kGWarps = CTRB_GROUP_WARPS;
// Node positions (fp64 xyz) per warp: the pool is sized at launch, kPoolMin or kPoolMax, whichever the longest
// configuration needs.  Shared memory is taken from the SM's L1, and this kernel lives on L1 hits (BVH nodes,
// triangles): 2 x 8 warps with 256-node pools and 128-entry queues (197 KB of shared memory) ran at 83 M configs/s,
// with 208-node pools and 64-entry queues (162 KB) at 92 M; more resident warps (20-24) made it slower (49-58 M).
constexpr int kPoolMin = 208, kPoolMax = 256;
constexpr int kGQueue = CTRB_GROUP_QUEUE;  // deferred exact tests per warp
constexpr int kGStack = 24;    // traversal stack entries per lane, in SHARED memory (a BVH deeper than kGStack - 2 takes
                               // chain_eval_kernel): with 32 busy lanes per warp a local-memory stack thrashed the ~40 KB
                               // of L1 that the shared-memory carve-out leaves (ncu: L1 hit rate 50 %, long-scoreboard 8.4)
#ifndef CTRB_GROUP_MIN
#define CTRB_GROUP_MIN 6
#endif
// No BVH levels are staged in shared memory here: the staged-or-global branch in load_node diverges between lanes that
// are at different depths, and the staging takes L1 capacity away from the nodes below (measured: 45.6 -> 63 M configs/s
// when the 128-node staging was removed).
constexpr int kGroupMin = CTRB_GROUP_MIN;   // fewer configurations than this in a group: one pass per configuration
#ifndef CTRB_GROUP_INNER
#define CTRB_GROUP_INNER 2
#endif
#ifndef CTRB_GROUP_REPS
#define CTRB_GROUP_REPS 8
#endif

// ---- long cylinders.  Non-physical equilibria of the static model (one-step sections) put up to tens
// of metres between two consecutive nodes.  A ball about the centre of such a cylinder covers the whole mesh: every
// triangle became a candidate pair and 0.9 % of a static batch took as long as the other 99.1 %.  Cylinders with a
// half-length above kLongHalf are therefore set aside during the group's walk and handled afterwards, when the
// configuration's bound is known, by their AXIS: the axis is clipped to the mesh's bounding box grown by bound + r (points
// beyond it cannot beat the bound), boxes are culled by a slab test against the grown box, and the prefilter is the
// fp32 segment-triangle distance d: d - r <= d(cylinder, triangle) <= d.  The decisive test stays the exact fp64 one on
// the whole cylinder.
constexpr float kLongHalf = 1.5f;
constexpr float kFarHalf = 25.f;  // ... and above this one they also get a first look BEFORE the other cylinders
constexpr unsigned int kLongFlag = 0x80000000u;  // item code bit: this cylinder is in the long list
constexpr int kLongMax = 32;  // long cylinders set aside per group; any beyond that go the ordinary (slow) way

// does the segment c +- h a pass through the box [lo - grow, hi + grow]?  (slab test, fp32, conservative through `grow`)
__device__ __forceinline__ bool axis_hits_box(const float lo[3], const float hi[3], f3 c, f3 a, float h, float grow) {
    const float o[3] = {c.x - h * a.x, c.y - h * a.y, c.z - h * a.z};
    const float d[3] = {2.f * h * a.x, 2.f * h * a.y, 2.f * h * a.z};
    float tmin = 0.f, tmax = 1.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float dk = copysignf(fmaxf(fabsf(d[k]), 1e-20f), d[k]);
        const float inv = 1.f / dk;
        const float t0 = (lo[k] - grow - o[k]) * inv, t1 = (hi[k] + grow - o[k]) * inv;
        tmin = fmaxf(tmin, fminf(t0, t1));
        tmax = fminf(tmax, fmaxf(t0, t1));
    }
    return tmin <= tmax * 1.00001f + 1e-6f;
}

// squared distance between segments p1 + s d1 (s in [0,1]) and p2..q2; *s_out = the parameter on the first one
__device__ __forceinline__ float seg_seg_dist2_f(f3 p1, f3 d1, f3 p2, f3 q2, float* s_out) {
    const f3 d2 = q2 - p2, r = p1 - p2;
    const float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), c = dot(d1, r);
    float s = 0.f, t = 0.f;
    if (a > 1e-20f && e > 1e-20f) {
        const float b = dot(d1, d2), den = a * e - b * b;
        s = den > 0.f ? fminf(fmaxf((b * f - c * e) / den, 0.f), 1.f) : 0.f;
        t = (b * s + f) / e;
        if (t < 0.f) {
            t = 0.f;
            s = fminf(fmaxf(-c / a, 0.f), 1.f);
        } else if (t > 1.f) {
            t = 1.f;
            s = fminf(fmaxf((b - c) / a, 0.f), 1.f);
        }
    } else if (a > 1e-20f) {
        s = fminf(fmaxf(-c / a, 0.f), 1.f);
    } else if (e > 1e-20f) {
        t = fminf(fmaxf(f / e, 0.f), 1.f);
    }
    const f3 w = {r.x + s * d1.x - t * d2.x, r.y + s * d1.y - t * d2.y, r.z + s * d1.z - t * d2.z};
    *s_out = s;
    return dot(w, w);
}

// fp32 distance between the segment c +- h a and triangle (v0, v1, v2), everything taken relative to c;
// *s_out = where on the segment (0 .. 1) the closest point lies
__device__ __noinline__ float seg_tri_dist_f(f3 c, f3 a, float h, f3 v0, f3 v1, f3 v2, float* s_out) {
    const f3 A = v0 - c, B = v1 - c, C = v2 - c;
    const f3 p = {-h * a.x, -h * a.y, -h * a.z}, q = {h * a.x, h * a.y, h * a.z};
    const f3 d = {2.f * h * a.x, 2.f * h * a.y, 2.f * h * a.z};
    float s = 0.f, sb;
    float best = point_tri_dist2<f3, float>(p, A, B, C);
    float t = point_tri_dist2<f3, float>(q, A, B, C);
    if (t < best) { best = t; s = 1.f; }
    t = seg_seg_dist2_f(p, d, A, B, &sb);
    if (t < best) { best = t; s = sb; }
    t = seg_seg_dist2_f(p, d, B, C, &sb);
    if (t < best) { best = t; s = sb; }
    t = seg_seg_dist2_f(p, d, C, A, &sb);
    if (t < best) { best = t; s = sb; }
    // the segment may pierce the triangle
    const f3 ab = B - A, ac = C - A;
    const f3 n = {ab.y * ac.z - ab.z * ac.y, ab.z * ac.x - ab.x * ac.z, ab.x * ac.y - ab.y * ac.x};
    const float dp = dot(n, p - A), dq = dot(n, q - A);
    if ((dp < 0.f) != (dq < 0.f) && dp != dq) {
        const float u = dp / (dp - dq);
        const f3 x = {p.x + u * d.x, p.y + u * d.y, p.z + u * d.z};
        if (point_tri_dist2<f3, float>(x, A, B, C) <= 1e-12f * fmaxf(dot(x, x), 1.f)) {
            best = 0.f;
            s = u;
        }
    }
    *s_out = s;
    return sqrtf(best);
}

struct LongItem {
    f3 c, a;       // clipped axis segment: centre, unit direction
    float h, r;    // its half-length; the cylinder's radius
    int active;
};
// Clip the axis of cylinder (c, a, h) to the mesh's bounding box grown by `grow` (fp64: the cylinder may be kilometres
// long).  Points of the axis outside that box are farther than `grow` from every triangle.
__device__ __noinline__ LongItem clip_long(const DevBvh& bvh, d3 c, d3 a, double h, double r, double grow) {
    LongItem L;
    L.active = 0;
    L.r = (float)r;
    const BvhNode root = bvh.nodes[0];
    double t0 = -h, t1 = h;
    const double cc[3] = {c.x, c.y, c.z}, aa[3] = {a.x, a.y, a.z};
    for (int k = 0; k < 3; ++k) {
        const double lo = (double)fminf(root.lo0[k], root.lo1[k]) - grow, hi = (double)fmaxf(root.hi0[k], root.hi1[k]) + grow;
        if (fabs(aa[k]) < 1e-300) {
            if (cc[k] < lo || cc[k] > hi) return L;
        } else {
            const double u0 = (lo - cc[k]) / aa[k], u1 = (hi - cc[k]) / aa[k];
            t0 = fmax(t0, fmin(u0, u1));
            t1 = fmin(t1, fmax(u0, u1));
        }
    }
    if (!(t0 <= t1)) return L;
    const double mid = 0.5 * (t0 + t1);
    L.c = f3{(float)(c.x + mid * a.x), (float)(c.y + mid * a.y), (float)(c.z + mid * a.z)};
    L.a = f3{(float)a.x, (float)a.y, (float)a.z};
    L.h = (float)(0.5 * (t1 - t0)) * 1.000001f + 1e-6f;
    L.active = 1;
    return L;
}

struct GroupMem {
    double* pos;                     // [pool][3]  (behind the GroupMem array in dynamic shared memory)
    unsigned int* items;             // [pool]     node (10 bits) | slot (5 bits) << 10 | tube (2 bits) << 15
    unsigned long long best_bits[32];
    int bound_bits[32];
    int collided[32];
    int q_tri[kGQueue];
    unsigned int q_item[kGQueue];
    float q_lb[kGQueue];
    int stack[kGStack][32];          // [entry][lane]: conflict-free
    unsigned int long_items[kLongMax];  // long cylinders set aside during the walk
    int nlong;
    int ticket;
    int qcount;
    int pad_;
};

template <bool FUSED>
__global__ void __launch_bounds__(kGWarps * 32, CTRB_GROUP_BLOCKS) chain_eval_group_kernel(const __grid_constant__ EvalParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GroupMem* s_all = reinterpret_cast<GroupMem*>(smem_raw);
    const EnvConst& env = *P.env;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GroupMem& W = s_all[warp];
    const int pool = P.pool;
    if (lane == 0) {
        double* pos0 = reinterpret_cast<double*>(s_all + kGWarps);
        W.pos = pos0 + (size_t)warp * pool * 3;
        W.items = reinterpret_cast<unsigned int*>(pos0 + (size_t)kGWarps * pool * 3) + (size_t)warp * pool;
    }
    __syncwarp();
    int* const stk = &W.stack[0][lane];  // entry k of this lane's stack = stk[k * 32]
    const int T = FUSED ? P.mc.T : P.tube_num;
    const bool have_obstacles = env.obs.ntri > 0 || env.nprim > 0;
    const DevBvh& bvh = env.obs;
    const float slack = bvh.slack;
    volatile int* vbound = W.bound_bits;
    volatile int* vcoll = W.collided;
    Work wk = {0, 0, 0};
    unsigned int nconf = 0;

    // the solid cylinder of an item: nodes `node`, `node + 1` of the pool (collision_checker.py:121-147)
    auto cylinder_of = [&](unsigned code, d3& c, d3& a, double& h, double& r) {
        const double* pf = W.pos + (size_t)(code & 1023u) * 3;
        d3 f = {pf[0], pf[1], pf[2]}, t = {pf[3], pf[4], pf[5]};
        d3 vec = t - f;
        const double len = sqrt(dot(vec, vec));
        c = f + 0.5 * vec;
        a = (1.0 / len) * vec;
        h = 0.5 * len;
        r = P.rad[(code >> 15) & 3u];
    };
    auto record = [&](int slot, double d) {
        if (d <= 0.0) {
            vcoll[slot] = 1;
        } else {
            atomicMin(&W.best_bits[slot], (unsigned long long)__double_as_longlong(d));
            atomicMin(&W.bound_bits[slot], __float_as_int(__double2float_ru(d)));
        }
    };
    auto exact_pair = [&](unsigned code, int t, float bound) {
        d3 c, a;
        double h, r;
        cylinder_of(code, c, a, h, r);
        const double* tq = bvh.tri64 + (size_t)t * 9;
        d3 t0 = {__ldg(tq + 0), __ldg(tq + 1), __ldg(tq + 2)};
        d3 t1 = {__ldg(tq + 3), __ldg(tq + 4), __ldg(tq + 5)};
        d3 t2 = {__ldg(tq + 6), __ldg(tq + 7), __ldg(tq + 8)};
        const double d = cyl_tri_dist(c, a, h, r, t0, t1, t2, (double)bound);
        wk.exact++;
        record((code >> 10) & 31u, d);
    };
    // every decision that guards a __syncwarp is taken on lane 0's view (see chain_eval_kernel)
    auto drain = [&]() {
        __syncwarp();
        const int qc = min(__shfl_sync(0xffffffffu, *(volatile int*)&W.qcount, 0), kGQueue);
        for (int base = 0; base < qc; base += 32) {
            const int k = base + lane;
            if (k < qc) {
                const unsigned code = W.q_item[k];
                const int slot = (code >> 10) & 31u;
                const float bound = __int_as_float(vbound[slot]);
                if (!vcoll[slot] && W.q_lb[k] <= bound) exact_pair(code, W.q_tri[k], bound);
            }
            __syncwarp();
        }
        if (lane == 0) W.qcount = 0;
        __syncwarp();
    };

    for (;;) {
        long long b0;
        {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(P.counters + 4, (unsigned long long)P.chunk);
            b0 = (long long)__shfl_sync(0xffffffffu, t, 0);
        }
        if (b0 >= P.B) break;
        const int nchunk = (int)min((long long)P.chunk, P.B - b0);
        int done = 0;
        while (done < nchunk) {
            // ---- lane = configuration: node counts, how many configurations fit the pool
            const int ci = done + lane;
            const bool valid = ci < nchunk;
            const long long b = b0 + ci;
            int cnt[CTRB_MAX_TUBES];
            int tn = 0, ncyl = 0;
#pragma unroll
            for (int n = 0; n < CTRB_MAX_TUBES; ++n) {
                cnt[n] = 0;
                if (valid && n < T) {
                    if (FUSED) cnt[n] = P.mc.npts[n] - P.idx[b * T + n];  // kinematic.py:100-103 (full=False)
                    else cnt[n] = (int)(P.offsets[b * T + n + 1] - P.offsets[b * T + n]);
                    tn += cnt[n];
                    ncyl += max(cnt[n] - 1, 0);
                }
            }
            int nend = tn, iend = ncyl;  // inclusive scans over lanes
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int a = __shfl_up_sync(0xffffffffu, nend, off), c = __shfl_up_sync(0xffffffffu, iend, off);
                if (lane >= off) {
                    nend += a;
                    iend += c;
                }
            }
            const unsigned fit = __ballot_sync(0xffffffffu, valid && nend <= pool);
            int m = __ffs(~fit) - 1;  // leading lanes that fit (the host guarantees one configuration always fits)
            if (m < 0) m = 32;
            m = max(m, 1);
            const bool mine = valid && lane < m;
            const int nstart = nend - tn, istart = iend - ncyl;
            const int total_items = __shfl_sync(0xffffffffu, iend, m - 1);
            // ---- positions into the pool, cylinder items into the list
            if (mine) nconf++;
            if (m >= kGroupMin) {
                // many short configurations: lane = configuration, nodes one after the other
                if (mine) {
                    double* dst = W.pos + (size_t)nstart * 3;
                    unsigned int* it = W.items + istart;
                    int node = nstart;
                    if (FUSED) {
                        const int tot = P.mc.total_nodes;
                        SE3 g;
                        se3_identity(g);
                        for (int n = 0; n < T; ++n) {
                            const int id = P.idx[b * T + n];
                            se3_rot_x(g, P.theta[b * T + n]);
                            const SE3 a = se3_mul(g, load_tab(P.tb.tab_inv, tot, P.mc.node_base[n] + id));
                            const int colbase = P.mc.node_base[n] + id;
#pragma unroll 1
                            for (int k = 0; k < cnt[n]; ++k) {
                                const int col = colbase + k;
                                const double cx = __ldg(P.tb.tab + 9 * tot + col), cy = __ldg(P.tb.tab + 10 * tot + col),
                                             cz = __ldg(P.tb.tab + 11 * tot + col);
                                dst[0] = fma(a.r[0], cx, fma(a.r[1], cy, fma(a.r[2], cz, a.p[0])));
                                dst[1] = fma(a.r[3], cx, fma(a.r[4], cy, fma(a.r[5], cz, a.p[1])));
                                dst[2] = fma(a.r[6], cx, fma(a.r[7], cy, fma(a.r[8], cz, a.p[2])));
                                dst += 3;
                                if (k + 1 < cnt[n]) *it++ = (unsigned)(node + k) | ((unsigned)lane << 10) | ((unsigned)n << 15);
                            }
                            node += cnt[n];
                            g = se3_mul(a, load_tab(P.tb.tab, tot, P.mc.node_base[n] + P.mc.npts[n] - 1));
                        }
                    } else {
                        for (int n = 0; n < T; ++n) {
                            const double* gg = P.g + P.offsets[b * T + n] * P.g_stride + P.g_poff;
#pragma unroll 1
                            for (int k = 0; k < cnt[n]; ++k) {
                                dst[0] = __ldg(gg + (size_t)k * P.g_stride);
                                dst[1] = __ldg(gg + (size_t)k * P.g_stride + P.g_pstep);
                                dst[2] = __ldg(gg + (size_t)k * P.g_stride + 2 * P.g_pstep);
                                dst += 3;
                                if (k + 1 < cnt[n]) *it++ = (unsigned)(node + k) | ((unsigned)lane << 10) | ((unsigned)n << 15);
                            }
                            node += cnt[n];
                        }
                    }
                }
            } else {
                // a few long configurations: one after the other, lanes stride over the nodes of each tube
#pragma unroll 1
                for (int sl = 0; sl < m; ++sl) {
                    const long long bs = b0 + done + sl;
                    int node = __shfl_sync(0xffffffffu, nstart, sl);
                    int item = __shfl_sync(0xffffffffu, istart, sl);
                    SE3 g;
                    se3_identity(g);
                    for (int n = 0; n < T; ++n) {
                        const int cn = __shfl_sync(0xffffffffu, cnt[n], sl);
                        if (FUSED) {
                            const int tot = P.mc.total_nodes;
                            const int id = P.idx[bs * T + n];
                            se3_rot_x(g, P.theta[bs * T + n]);
                            const SE3 a = se3_mul(g, load_tab(P.tb.tab_inv, tot, P.mc.node_base[n] + id));
                            const int colbase = P.mc.node_base[n] + id;
                            for (int k = lane; k < cn; k += 32) {
                                const int col = colbase + k;
                                const double cx = __ldg(P.tb.tab + 9 * tot + col), cy = __ldg(P.tb.tab + 10 * tot + col),
                                             cz = __ldg(P.tb.tab + 11 * tot + col);
                                double* dst = W.pos + (size_t)(node + k) * 3;
                                dst[0] = fma(a.r[0], cx, fma(a.r[1], cy, fma(a.r[2], cz, a.p[0])));
                                dst[1] = fma(a.r[3], cx, fma(a.r[4], cy, fma(a.r[5], cz, a.p[1])));
                                dst[2] = fma(a.r[6], cx, fma(a.r[7], cy, fma(a.r[8], cz, a.p[2])));
                            }
                            g = se3_mul(a, load_tab(P.tb.tab, tot, P.mc.node_base[n] + P.mc.npts[n] - 1));
                        } else {
                            const double* gg = P.g + P.offsets[bs * T + n] * P.g_stride + P.g_poff;
                            for (int k = lane; k < cn; k += 32) {
                                double* dst = W.pos + (size_t)(node + k) * 3;
                                dst[0] = __ldg(gg + (size_t)k * P.g_stride);
                                dst[1] = __ldg(gg + (size_t)k * P.g_stride + P.g_pstep);
                                dst[2] = __ldg(gg + (size_t)k * P.g_stride + 2 * P.g_pstep);
                            }
                        }
                        for (int k = lane; k + 1 < cn; k += 32)
                            W.items[item + k] = (unsigned)(node + k) | ((unsigned)sl << 10) | ((unsigned)n << 15);
                        node += cn;
                        item += max(cn - 1, 0);
                    }
                }
            }
            W.bound_bits[lane] = __float_as_int(FLT_MAX);
            W.collided[lane] = 0;
            W.best_bits[lane] = (unsigned long long)__double_as_longlong(DBL_MAX);
            if (lane == 0) {
                W.ticket = 0;
                W.qcount = 0;
                W.nlong = 0;
            }
            __syncwarp();
            // ---- long cylinders are taken out of the item list
            if (have_obstacles && bvh.ntri > 0) {
                for (int i = lane; i < total_items; i += 32) {
                    const unsigned code = W.items[i];
                    const double* pf = W.pos + (size_t)(code & 1023u) * 3;
                    const double dx = pf[3] - pf[0], dy = pf[4] - pf[1], dz = pf[5] - pf[2];
                    if (fma(dx, dx, fma(dy, dy, dz * dz)) > 4.0 * (double)kLongHalf * (double)kLongHalf) {
                        const int kl = atomicAdd(&W.nlong, 1);
                        if (kl < kLongMax) {  // (any beyond that stay in the list and go the ordinary way)
                            W.long_items[kl] = code;
                            W.items[i] = code | kLongFlag;
                        }
                    }
                }
            }
            __syncwarp();
            const int nl = min(__shfl_sync(0xffffffffu, *(volatile int*)&W.nlong, 0), kLongMax);
            // phase 0: the long cylinders with a thin tube about their axis only (does it pierce the wall? what is near
            //          it?) -- a contact ends the configuration before anything else is walked, and otherwise the
            //          configuration has a finite bound before its far-away cylinders are looked at;
            //          then the walk of all ordinary cylinders;
            // phase 1: the long cylinders again, tube widened up to the (now tight) bound.
#pragma unroll 1
            for (int phase = 0; phase < 2; ++phase) {
            if (nl > 0) {
                unsigned code = 0;
                int slot = 0, sp = 0;
                LongItem li;
                li.active = 0;
                li.h = 0.f;
                li.r = 0.f;
                li.c = li.a = f3{0.f, 0.f, 0.f};
                if (lane < nl) {
                    code = W.long_items[lane];
                    slot = (code >> 10) & 31u;
                    if (!vcoll[slot]) {
                        d3 c, a;
                        double h, r;
                        cylinder_of(code, c, a, h, r);
                        // phase 0 is for the cylinders long enough to carry the rest of the chain out of the scene
                        if (phase == 0 && h < (double)kFarHalf) h = -1.0;
                        if (h > 0.0 && __int_as_float(vbound[slot]) > 1e30f) {
                            // no bound yet: the cylinder's centre to any one vertex of the mesh is an upper bound
                            const d3 v = {__ldg(bvh.tri64 + 0), __ldg(bvh.tri64 + 1), __ldg(bvh.tri64 + 2)};
                            const d3 dv = v - c;
                            atomicMin(&W.bound_bits[slot], __float_as_int(__double2float_ru(sqrt(dot(dv, dv)))));
                        }
                        // phase 0 looks no farther than 32 r from the axis (tubes of 2 r, 8 r, 32 r); phase 1 as far as the bound requires
                        const double reach = phase == 0 ? 32.0 * r : (double)__int_as_float(vbound[slot]) * 1.000001 + r;
                        if (h > 0.0) li = clip_long(bvh, c, a, h, r, reach + 4.0 * (double)slack);
                        if (li.active) {
                            stk[0] = 0;
                            sp = 1;
                        }
                    }
                }
                const float lslack = fmaxf(slack, 1.6e-5f * fmaxf(fmaxf(fabsf(li.c.x), fabsf(li.c.y)), fabsf(li.c.z))) * 4.f +
                                     li.h * 4e-6f;  // fp32 error of a point far out on the clipped axis
                // The tree is walked with a tube about the axis that starts thin (2 r) and is widened fourfold, from the
                // root again, until it reaches bound + r: the boxes the axis itself runs through come first, so a
                // cylinder that pierces the wall finds its contact in the first pass, and one that does not has a tight
                // bound (from the triangles next to it) before the wide passes.
                float gcap = 2.f * li.r;
                for (;;) {
                    if (sp == 0 && li.active) {
                        if (!vcoll[slot] && gcap < __int_as_float(vbound[slot]) + li.r && (phase == 1 || gcap < 32.f * li.r)) {
                            gcap *= 4.f;
                            stk[0] = 0;
                            sp = 1;
                        } else {
                            li.active = 0;
                        }
                    }
                    if (!__any_sync(0xffffffffu, sp > 0)) break;
#pragma unroll 1
                    for (int rep = 0; rep < 8 && sp > 0; ++rep) {
                        if (vcoll[slot]) {
                            sp = 0;
                            break;
                        }
                        const int e = stk[--sp * 32];
                        if (e >= 0) {
                            const BvhNode nd = load_node(bvh, nullptr, 0, e);
                            wk.nodes++;
                            const float grow = fminf(__int_as_float(vbound[slot]) + li.r, gcap) * 1.00001f + lslack;
                            if (axis_hits_box(nd.lo1, nd.hi1, li.c, li.a, li.h, grow) && sp < kGStack) stk[sp++ * 32] = nd.c1;
                            if (axis_hits_box(nd.lo0, nd.hi0, li.c, li.a, li.h, grow) && sp < kGStack) stk[sp++ * 32] = nd.c0;
                            continue;
                        }
                        const int lc = ~e;
                        const int tfirst = lc >> 3, tcnt = (lc & 7) + 1;
#pragma unroll 1
                        for (int t = tfirst; t < tfirst + tcnt && !vcoll[slot]; ++t) {
                            const float4* tp = bvh.tri32 + (size_t)t * 3;
                            const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
                            float sq;
                            const float dst = seg_tri_dist_f(li.c, li.a, li.h, f3{v0.x, v0.y, v0.z}, f3{v1.x, v1.y, v1.z},
                                                             f3{v2.x, v2.y, v2.z}, &sq);
                            wk.pre++;
                            // a triangle point within r of an interior point of the axis is inside the solid cylinder
                            if (dst + lslack < li.r && sq > 0.001f && sq < 0.999f) {
                                vcoll[slot] = 1;
                                sp = 0;
                                break;
                            }
                            // the axis belongs to the cylinder: d(cyl, tri) <= d(axis, tri); and the cylinder lies inside
                            // the capsule of radius r about its axis: d(cyl, tri) >= d(axis, tri) - r
                            atomicMin(&W.bound_bits[slot], __float_as_int(dst + lslack));
                            const float bound = __int_as_float(vbound[slot]);
                            const float lb = dst - li.r - lslack;
                            if (lb > bound) continue;
                            const int qs = atomicAdd(&W.qcount, 1);
                            if (qs < kGQueue) {
                                W.q_tri[qs] = t;
                                W.q_item[qs] = code;
                                W.q_lb[qs] = lb;
                            } else {
                                exact_pair(code, t, bound);
                            }
                        }
                    }
                    __syncwarp();
                    if (__shfl_sync(0xffffffffu, *(volatile int*)&W.qcount, 0) >= 32) drain();
                }
                drain();
            }
            if (phase == 1) break;
            // ---- lane = work item
            // Many short configurations: one pass over the whole item list.  A few long ones: one pass per configuration,
            // so that a contact ends that configuration's walks at once and the warp moves on together (python-fcl's
            // callback stops at the first contact; most long configurations collide).
            const int npass = !have_obstacles ? 0 : (m >= kGroupMin ? 1 : m);
#pragma unroll 1
            for (int pass = 0; pass < npass; ++pass) {
                int range_end = total_items;
                if (m < kGroupMin) {
                    range_end = __shfl_sync(0xffffffffu, iend, pass);
                    const int range_begin = __shfl_sync(0xffffffffu, istart, pass);
                    if (lane == 0) W.ticket = range_begin;
                    __syncwarp();
                }
                int sp = 0, slot = 0;
                unsigned code = 0;
                bool exhausted = false;
                f3 cf = {0.f, 0.f, 0.f}, af = {0.f, 0.f, 0.f};
                float rb = 0.f, hf = 0.f, rf = 0.f;
                float slk = slack;  // `slack` is sized for the mesh's coordinates; a cylinder far outside them needs its own
                for (;;) {
                    if (sp == 0 && !exhausted) {  // next live cylinder for this lane
#pragma unroll 1
                        for (int tries = 0; tries < 4; ++tries) {
                            const int q = atomicAdd(&W.ticket, 1);
                            if (q >= range_end) {
                                exhausted = true;
                                break;
                            }
                            code = W.items[q];
                            if (code & kLongFlag) continue;  // a long cylinder: handled by its axis, before and after this walk
                            slot = (code >> 10) & 31u;
                            if (vcoll[slot]) {
                                if (m < kGroupMin) {  // this pass is that configuration's alone: nothing left to do
                                    exhausted = true;
                                    break;
                                }
                                continue;
                            }
                            d3 c, a;
                            double h, r;
                            cylinder_of(code, c, a, h, r);
                            for (int k = 0; k < env.nprim && !vcoll[slot]; ++k)
                                record(slot, cylinder_vs_prim(env.prims[k], c, a, h, r,
                                                              __longlong_as_double((long long)W.best_bits[slot])));
                            if (bvh.ntri > 0 && !vcoll[slot]) {
                                cf = f3{(float)c.x, (float)c.y, (float)c.z};
                                af = f3{(float)a.x, (float)a.y, (float)a.z};
                                hf = (float)h;
                                rf = (float)r;
                                rb = __fsqrt_ru(__fmaf_ru(hf, hf, rf * rf)) * 1.000001f;
                                slk = fmaxf(slack, 1.6e-5f * fmaxf(fmaxf(fabsf(cf.x), fabsf(cf.y)), fabsf(cf.z)));  // 64 ulp of 2 |c|
                                stk[0] = 0;
                                sp = 1;
                                break;
                            }
                        }
                    }
                    if (!__any_sync(0xffffffffu, sp > 0 || !exhausted)) break;
#pragma unroll 1
                    for (int rep = 0; rep < CTRB_GROUP_REPS; ++rep) {
                        int leaf = 0;
                        bool has_leaf = false;
#pragma unroll 1
                        for (int mstep = 0; mstep < CTRB_GROUP_INNER && sp > 0 && !has_leaf; ++mstep) {
                            if (vcoll[slot]) {
                                sp = 0;
                                break;
                            }
                            const int e = stk[--sp * 32];
                            if (e < 0) {
                                leaf = e;
                                has_leaf = true;
                            } else {
                                const BvhNode nd = load_node(bvh, nullptr, 0, e);
                                wk.nodes++;
                                const float reach = __int_as_float(vbound[slot]) + rb + slk;
                                const float reach2 = reach * reach * 1.000001f;
                                float d0 = aabb_dist2(nd.lo0, nd.hi0, cf);
                                float d1 = aabb_dist2(nd.lo1, nd.hi1, cf);
                                int l0 = nd.c0, l1 = nd.c1;
                                if (d1 < d0) {  // near child on top of the stack
                                    float t = d0; d0 = d1; d1 = t;
                                    int ti = l0; l0 = l1; l1 = ti;
                                }
                                if (d1 <= reach2 && sp < kGStack) stk[sp++ * 32] = l1;
                                if (d0 <= reach2 && sp < kGStack) stk[sp++ * 32] = l0;
                            }
                        }
                        if (has_leaf) {
                            const int lc = ~leaf;
                            const int tfirst = lc >> 3, tcnt = (lc & 7) + 1;
#pragma unroll 1
                            for (int t = tfirst; t < tfirst + tcnt && !vcoll[slot]; ++t) {
                                const float4* tp = bvh.tri32 + (size_t)t * 3;
                                const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
                                // q = the triangle's closest point to the cylinder's centre, in cylinder coordinates
                                const f3 w = point_tri_closest(cf, f3{v0.x, v0.y, v0.z}, f3{v1.x, v1.y, v1.z}, f3{v2.x, v2.y, v2.z});
                                wk.pre++;
                                const float dpt2 = dot(w, w), tq = dot(w, af);
                                const float dpt = sqrtf(dpt2), rho = sqrtf(fmaxf(dpt2 - tq * tq, 0.f));
                                const float et = fabsf(tq) - hf, er = rho - rf;
                                if (et + slk < 0.f && er + slk < 0.f) {  // q is inside the solid cylinder for certain
                                    vcoll[slot] = 1;
                                    sp = 0;
                                    break;
                                }
                                // q belongs to the triangle: d(cyl, tri) <= d(q, cyl)
                                const float dt_ = fmaxf(et, 0.f), dr_ = fmaxf(er, 0.f);
                                atomicMin(&W.bound_bits[slot], __float_as_int(sqrtf(fmaf(dt_, dt_, dr_ * dr_)) + slk));
                                const float bound = __int_as_float(vbound[slot]);
                                const float lb = dpt - rb - slk;  // d(cyl, tri) >= dpt - rb
                                if (lb > bound) continue;
                                const int qs = atomicAdd(&W.qcount, 1);
                                if (qs < kGQueue) {
                                    W.q_tri[qs] = t;
                                    W.q_item[qs] = code;
                                    W.q_lb[qs] = lb;
                                } else {
                                    exact_pair(code, t, bound);  // queue full: decide here (rare)
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (__shfl_sync(0xffffffffu, *(volatile int*)&W.qcount, 0) >= 32) drain();
                }
                drain();
            }
            __syncwarp();
            }  // phase
            __syncwarp();
            // ---- lane = configuration again: goal distance, cost, results (coalesced across lanes)
            if (mine) {
                const bool collided = vcoll[lane] != 0;
                const double best = __longlong_as_double((long long)W.best_bits[lane]);
                double gd = DBL_MAX;
                if (tn > 0 && env.goal.shape != CTRB_SHAPE_NONE) {
                    const double* pt = W.pos + (size_t)(nstart + tn - 1) * 3;  // curve[-1][-1]
                    d3 tip = {pt[0], pt[1], pt[2]};
                    const double tr = P.rad[T - 1];
                    const DevPrim& gl = env.goal;
                    if (gl.shape == CTRB_SHAPE_SPHERE) {
                        d3 dd = tip - d3{gl.center[0], gl.center[1], gl.center[2]};
                        gd = sqrt(dot(dd, dd)) - gl.dims[0] - tr;
                    } else if (gl.shape == CTRB_SHAPE_BOX) {
                        gd = point_box_dist(tip, gl.center, gl.rot, gl.dims) - tr;
                    } else if (gl.shape == CTRB_SHAPE_CYLINDER) {
                        d3 ax = {gl.rot[2], gl.rot[5], gl.rot[8]};
                        gd = point_cyl_dist(tip, d3{gl.center[0], gl.center[1], gl.center[2]}, ax, 0.5 * gl.dims[1], gl.dims[0]) - tr;
                    } else if (gl.shape == CTRB_SHAPE_MESH && env.goal_mesh.ntri > 0) {
                        gd = point_vs_mesh(env.goal_mesh, tip, wk) - tr;
                    }
                    if (gd <= 0.0) gd = -1.0;
                }
                const double om = collided ? -1.0 : (have_obstacles ? best : DBL_MAX);
                P.obstacle_min[b] = om;
                P.goal_dist[b] = gd;
                if (P.collide) P.collide[b] = collided ? 1 : 0;
                if (P.cost) {
                    double cst = CUDART_NAN;  // rrt.py:116-129
                    if (!collided) {
                        const double g0 = gd < 0.0 ? 0.0 : gd;
                        if (P.cd.cost_type == CTRB_COST_SOAPWG) {
                            const double inv = 1.0 / om;  // obstacles_and_goal.py:39-45,87-88
                            cst = g0 * P.cd.goal_weight + inv * inv;
                        } else if (P.cd.cost_type == CTRB_COST_ONLY_GOAL) {
                            cst = g0;  // only_goal_distance.py:20-24
                        } else {
                            cst = 0.0;  // FTL family: the twist term comes from ctrb_eta_ftl_batch
                        }
                    }
                    P.cost[b] = cst;
                }
            }
            __syncwarp();
            done += m;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        wk.nodes += __shfl_xor_sync(0xffffffffu, wk.nodes, off);
        wk.pre += __shfl_xor_sync(0xffffffffu, wk.pre, off);
        wk.exact += __shfl_xor_sync(0xffffffffu, wk.exact, off);
        nconf += __shfl_xor_sync(0xffffffffu, nconf, off);
    }
    if (lane == 0) {
        atomicAdd(P.counters + 0, (unsigned long long)wk.nodes);
        atomicAdd(P.counters + 1, (unsigned long long)wk.pre);
        atomicAdd(P.counters + 2, (unsigned long long)wk.exact);
        atomicAdd(P.counters + 3, (unsigned long long)nconf);
    }
}

size_t eval_smem_bytes(int max_nodes) {
    return (size_t)kTopNodes * sizeof(BvhNode) + (size_t)kCWarps * max_nodes * 3 * sizeof(double) +
           (size_t)kCWarps * sizeof(WarpCtl);
}

int launch_eval(ctrb_ctx* ctx, const EvalParams& P, bool fused) {
    if (P.B <= 0) return CTRB_OK;
    // grouped kernel (several configurations per warp) whenever a configuration fits the warp's node pool;
    // CTRB_OPT_EVAL_GROUP = 0 forces one warp per configuration
    const bool group_off = ctx->opt[CTRB_OPT_EVAL_GROUP] == 0;
    if (P.max_nodes <= kPoolMax && P.bvh_depth + 2 <= kGStack && !group_off) {
        const int pool = P.max_nodes <= kPoolMin ? kPoolMin : kPoolMax;
        const size_t smem_p = (size_t)kGWarps * (sizeof(GroupMem) + (size_t)pool * (3 * sizeof(double) + sizeof(unsigned int)));
        CTRB_CUDA(cudaMemsetAsync(P.counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
        int per_sm = 1;
        if (fused) {
            CTRB_CUDA(cudaFuncSetAttribute(chain_eval_group_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
            CTRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chain_eval_group_kernel<true>, kGWarps * 32, smem_p));
        } else {
            CTRB_CUDA(cudaFuncSetAttribute(chain_eval_group_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
            CTRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chain_eval_group_kernel<false>, kGWarps * 32, smem_p));
        }
        // a small batch (an RRT round, a simplex) is spread over all resident warps rather than packed 32 to a warp
        const long long resident = (long long)ctx->sm_count * std::max(per_sm, 1) * kGWarps;
        EvalParams Q = P;
        Q.pool = pool;
        Q.chunk = (int)std::min<long long>(32, std::max<long long>(1, P.B / resident));
        const long long chunks = (P.B + Q.chunk - 1) / Q.chunk;
        const int grid_p = (int)std::min<long long>((chunks + kGWarps - 1) / kGWarps, (long long)ctx->sm_count * std::max(per_sm, 1));
        KernelTimer kt(ctx, CTRB_KERNEL_CHAIN_EVAL);
        if (fused) chain_eval_group_kernel<true><<<grid_p, kGWarps * 32, smem_p, ctx->stream>>>(Q);
        else chain_eval_group_kernel<false><<<grid_p, kGWarps * 32, smem_p, ctx->stream>>>(Q);
        ctx->launches++;
        CTRB_CUDA(cudaGetLastError());
        return CTRB_OK;
    }
    size_t smem = eval_smem_bytes(P.max_nodes);
    if (smem > 200 * 1024) {
        set_error("configuration has too many nodes (%d) for the shared-memory staging", P.max_nodes);
        return CTRB_ERR_UNSUPPORTED;
    }
    CTRB_CUDA(cudaMemsetAsync(P.counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    long long blocks_needed = (P.B + kCWarps - 1) / kCWarps;
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
    int grid = (int)std::min<long long>(blocks_needed, (long long)ctx->sm_count * per_sm);
    if (fused) {
        CTRB_CUDA(cudaFuncSetAttribute(chain_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelTimer kt(ctx, CTRB_KERNEL_CHAIN_EVAL);
        chain_eval_kernel<true><<<grid, kCWarps * 32, smem, ctx->stream>>>(P);
    } else {
        CTRB_CUDA(cudaFuncSetAttribute(chain_eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelTimer kt(ctx, CTRB_KERNEL_CHAIN_EVAL);
        chain_eval_kernel<false><<<grid, kCWarps * 32, smem, ctx->stream>>>(P);
    }
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}


int launch_eval_fused(ctrb_ctx* ctx, const ctrb_model* m, const ctrb_env* env, const ctrb_cost_desc* cd, long long B,
                      const int32_t* idx, const double* theta, const double* rad, double* obstacle_min,
                      double* goal_dist, double* cost, uint8_t* collide) {
    EvalParams P;
    memset(&P, 0, sizeof P);
    P.mc = m->mc;
    P.tb = m->tables();
    P.env = env->d_ec;
    P.B = B;
    P.idx = idx;
    P.theta = theta;
    P.tube_num = m->mc.T;
    for (int n = 0; n < m->mc.T; ++n) P.rad[n] = rad[n];
    P.obstacle_min = obstacle_min;
    P.goal_dist = goal_dist;
    P.cost = cost;
    P.collide = collide;
    P.cd = *cd;
    P.counters = ctx->d_counters;
    P.bvh_depth = env->stats[2];
    P.max_nodes = m->mc.total_nodes;
    return launch_eval(ctx, P, true);
}

int launch_eval_curves(ctrb_ctx* ctx, const ctrb_env* env, long long B, int tube_num, const double* g,
                       const long long* offsets, int max_nodes, const double* rad, double* obstacle_min,
                       double* goal_dist) {
    EvalParams P;
    memset(&P, 0, sizeof P);
    P.env = env->d_ec;
    P.B = B;
    P.g = g;
    P.offsets = offsets;
    P.tube_num = tube_num;
    for (int n = 0; n < tube_num; ++n) P.rad[n] = rad[n];
    P.obstacle_min = obstacle_min;
    P.goal_dist = goal_dist;
    P.counters = ctx->d_counters;
    P.bvh_depth = env->stats[2];
    P.max_nodes = max_nodes;
    P.g_stride = 16;
    P.g_poff = 3;
    P.g_pstep = 4;
    return launch_eval(ctx, P, false);
}

// pos_only: g holds 3 doubles per node (the node positions) instead of 4x4 matrices
int launch_eval_curves_cost(ctrb_ctx* ctx, const ctrb_env* env, const ctrb_cost_desc* cd, long long B, int tube_num,
                            const double* g, const long long* offsets, int max_nodes, const double* rad,
                            double* obstacle_min, double* goal_dist, double* cost, uint8_t* collide, int pos_only) {
    EvalParams P;
    memset(&P, 0, sizeof P);
    P.env = env->d_ec;
    P.B = B;
    P.g = g;
    P.offsets = offsets;
    P.tube_num = tube_num;
    for (int n = 0; n < tube_num; ++n) P.rad[n] = rad[n];
    P.obstacle_min = obstacle_min;
    P.goal_dist = goal_dist;
    P.cost = cost;
    P.collide = collide;
    P.cd = *cd;
    P.counters = ctx->d_counters;
    P.bvh_depth = env->stats[2];
    P.max_nodes = max_nodes;
    P.g_stride = pos_only ? 3 : 16;
    P.g_poff = pos_only ? 0 : 3;
    P.g_pstep = pos_only ? 1 : 4;
    return launch_eval(ctx, P, false);
}

// ------------------------------------------------------------------ best node: min finite cost, then lowest index
__device__ __forceinline__ unsigned long long cost_key(double c) {
    unsigned long long b = (unsigned long long)__double_as_longlong(c);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}

__global__ void argmin_key_kernel(long long B, const double* __restrict__ cost, unsigned long long* __restrict__ key) {
    unsigned long long best = ~0ULL;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
        double c = cost[i];
        if (c == c) best = min(best, cost_key(c));
    }
    for (int off = 16; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
    if ((threadIdx.x & 31) == 0 && best != ~0ULL) atomicMin(key, best);
}

__global__ void argmin_index_kernel(long long B, const double* __restrict__ cost, unsigned long long* __restrict__ key) {
    const unsigned long long k = key[0];
    if (k == ~0ULL) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
        double c = cost[i];
        if (c == c && cost_key(c) == k) atomicMin(key + 1, (unsigned long long)i);
    }
}

// key_out (from launch_argmin) -> out[0] = signed order-preserving key, out[1] = index_base + index (INT64_MAX: none)
__global__ void argmin_finish_kernel(const unsigned long long* __restrict__ key, long long index_base, long long* __restrict__ out) {
    const unsigned long long k = key[0], i = key[1];
    const bool none = k == ~0ULL || i == ~0ULL;
    out[0] = none ? 0x7fffffffffffffffLL : (long long)(k ^ 0x8000000000000000ULL);
    out[1] = none ? 0x7fffffffffffffffLL : index_base + (long long)i;
}

int launch_argmin_finish(ctrb_ctx* ctx, const unsigned long long* key, long long index_base, long long* out) {
    argmin_finish_kernel<<<1, 1, 0, ctx->stream>>>(key, index_base, out);
    ctx->launches++;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

int launch_argmin(ctrb_ctx* ctx, long long B, const double* cost, unsigned long long* key_out) {
    CTRB_CUDA(cudaMemsetAsync(key_out, 0xff, 2 * sizeof(unsigned long long), ctx->stream));
    int grid = (int)std::min<long long>((B + 255) / 256, (long long)ctx->sm_count * 8);
    argmin_key_kernel<<<grid, 256, 0, ctx->stream>>>(B, cost, key_out);
    argmin_index_kernel<<<grid, 256, 0, ctx->stream>>>(B, cost, key_out);
    ctx->launches += 2;
    CTRB_CUDA(cudaGetLastError());
    return CTRB_OK;
}

}  // namespace ctrb
