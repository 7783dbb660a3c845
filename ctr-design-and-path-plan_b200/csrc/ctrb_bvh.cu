// ctrb_bvh.cu -- host-side BVH builder for the triangle soup of an STL environment
// (replaces the fcl.BVHModel built at init_collision.py:117-123).  Binned SAH,
// leaves of <= kLeafMax triangles, nodes emitted in BFS order so the top levels are a
// prefix that the traversal kernels stage in shared memory.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <queue>

#include "ctrb_internal.cuh"

namespace ctrb {

namespace {
#ifndef CTRB_LEAF_MAX
#define CTRB_LEAF_MAX 4
#endif
#ifndef CTRB_LEAF_FORCE
#define CTRB_LEAF_FORCE 0
#endif
constexpr int kLeafMax = CTRB_LEAF_MAX;
constexpr int kBins = 16;

struct Box {
    double lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; ++k) {
            lo[k] = DBL_MAX;
            hi[k] = -DBL_MAX;
        }
    }
    void grow(const double* p) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], p[k]);
            hi[k] = std::max(hi[k], p[k]);
        }
    }
    void grow(const Box& b) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], b.lo[k]);
            hi[k] = std::max(hi[k], b.hi[k]);
        }
    }
    double area() const {
        double d[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        if (d[0] < 0) return 0.0;
        return 2.0 * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0]);
    }
};

struct BuildNode {
    Box box;
    int first, count;  // range in `order`
    int child[2];      // -1 for leaf
};

float down32(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -INFINITY);
    return f;
}
float up32(double x) {
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, INFINITY);
    return f;
}
}  // namespace

void build_bvh(const double* tris, int ntri, HostBvh* out) {
    out->nodes.clear();
    out->tri64.clear();
    out->tri32.clear();
    out->ntri = ntri;
    out->nleaves = 0;
    out->depth = 0;
    out->slack = 0.f;
    if (ntri <= 0) return;

    std::vector<Box> tb(ntri);
    std::vector<double> cen((size_t)ntri * 3);
    double maxabs = 0.0;
    for (int i = 0; i < ntri; ++i) {
        tb[i].reset();
        for (int v = 0; v < 3; ++v) {
            tb[i].grow(tris + (size_t)i * 9 + v * 3);
            for (int k = 0; k < 3; ++k) maxabs = std::max(maxabs, std::fabs(tris[(size_t)i * 9 + v * 3 + k]));
        }
        for (int k = 0; k < 3; ++k) cen[(size_t)i * 3 + k] = 0.5 * (tb[i].lo[k] + tb[i].hi[k]);
    }
    std::vector<int> order(ntri);
    for (int i = 0; i < ntri; ++i) order[i] = i;

    std::vector<BuildNode> bn;
    bn.reserve(2 * (size_t)ntri);
    {
        BuildNode root;
        root.box.reset();
        for (int i = 0; i < ntri; ++i) root.box.grow(tb[i]);
        root.first = 0;
        root.count = ntri;
        root.child[0] = root.child[1] = -1;
        bn.push_back(root);
    }
    // iterative top-down build
    std::vector<int> todo{0};
    while (!todo.empty()) {
        int id = todo.back();
        todo.pop_back();
        BuildNode nd = bn[id];
        if (nd.count <= 1) continue;
        Box cb;
        cb.reset();
        for (int i = nd.first; i < nd.first + nd.count; ++i) cb.grow(&cen[(size_t)order[i] * 3]);
        int best_axis = -1, best_bin = -1;
        double best_cost = DBL_MAX;
        for (int ax = 0; ax < 3; ++ax) {
            double ext = cb.hi[ax] - cb.lo[ax];
            if (!(ext > 0)) continue;
            Box bb[kBins];
            int bc[kBins] = {0};
            for (auto& b : bb) b.reset();
            for (int i = nd.first; i < nd.first + nd.count; ++i) {
                int t = order[i];
                int b = (int)(kBins * (cen[(size_t)t * 3 + ax] - cb.lo[ax]) / ext);
                b = std::min(std::max(b, 0), kBins - 1);
                bb[b].grow(tb[t]);
                bc[b]++;
            }
            double la[kBins], ra[kBins];
            int lc[kBins], rc[kBins];
            Box acc;
            acc.reset();
            int c = 0;
            for (int b = 0; b < kBins; ++b) {
                if (bc[b]) acc.grow(bb[b]);
                c += bc[b];
                la[b] = acc.area();
                lc[b] = c;
            }
            acc.reset();
            c = 0;
            for (int b = kBins - 1; b >= 0; --b) {
                if (bc[b]) acc.grow(bb[b]);
                c += bc[b];
                ra[b] = acc.area();
                rc[b] = c;
            }
            for (int b = 0; b + 1 < kBins; ++b) {
                if (lc[b] == 0 || rc[b + 1] == 0) continue;
                double cost = la[b] * lc[b] + ra[b + 1] * rc[b + 1];
                if (cost < best_cost) {
                    best_cost = cost;
                    best_axis = ax;
                    best_bin = b;
                }
            }
        }
        double leaf_cost = nd.box.area() * nd.count;
        bool make_leaf = (nd.count <= kLeafMax) && (CTRB_LEAF_FORCE || best_axis < 0 || best_cost >= leaf_cost);
        if (make_leaf) continue;
        int mid;
        if (best_axis < 0) {  // all centroids coincide: split evenly
            mid = nd.first + nd.count / 2;
        } else {
            double ext = cb.hi[best_axis] - cb.lo[best_axis];
            auto it = std::partition(order.begin() + nd.first, order.begin() + nd.first + nd.count, [&](int t) {
                int b = (int)(kBins * (cen[(size_t)t * 3 + best_axis] - cb.lo[best_axis]) / ext);
                b = std::min(std::max(b, 0), kBins - 1);
                return b <= best_bin;
            });
            mid = (int)(it - order.begin());
            if (mid == nd.first || mid == nd.first + nd.count) mid = nd.first + nd.count / 2;
        }
        BuildNode l, r;
        l.first = nd.first;
        l.count = mid - nd.first;
        r.first = mid;
        r.count = nd.first + nd.count - mid;
        l.child[0] = l.child[1] = r.child[0] = r.child[1] = -1;
        l.box.reset();
        r.box.reset();
        for (int i = l.first; i < l.first + l.count; ++i) l.box.grow(tb[order[i]]);
        for (int i = r.first; i < r.first + r.count; ++i) r.box.grow(tb[order[i]]);
        int li = (int)bn.size();
        bn.push_back(l);
        int ri = (int)bn.size();
        bn.push_back(r);
        bn[id].child[0] = li;
        bn[id].child[1] = ri;
        todo.push_back(li);
        todo.push_back(ri);
    }

    // reorder triangles into leaf order (the `order` permutation is already leaf-contiguous)
    out->tri64.resize((size_t)ntri * 9);
    out->tri32.assign((size_t)ntri * 12, 0.f);
    for (int i = 0; i < ntri; ++i) {
        const double* s = tris + (size_t)order[i] * 9;
        std::memcpy(&out->tri64[(size_t)i * 9], s, 9 * sizeof(double));
        for (int v = 0; v < 3; ++v)
            for (int k = 0; k < 3; ++k) out->tri32[(size_t)i * 12 + v * 4 + k] = (float)s[v * 3 + k];
    }

    auto leaf_link = [&](const BuildNode& b) { return ~((b.first << 3) | (b.count - 1)); };

    // emit inner nodes in BFS order
    if (bn[0].child[0] < 0) {  // single leaf mesh: one inner node whose second child is empty
        BvhNode n;
        std::memset(&n, 0, sizeof n);
        for (int k = 0; k < 3; ++k) {
            n.lo0[k] = down32(bn[0].box.lo[k]);
            n.hi0[k] = up32(bn[0].box.hi[k]);
            n.lo1[k] = FLT_MAX;  // empty box: infinitely far
            n.hi1[k] = FLT_MAX;
        }
        n.c0 = leaf_link(bn[0]);
        n.c1 = ~((0 << 3) | 0);
        // mark the second child as an empty leaf by giving it an unreachable box
        out->nodes.push_back(n);
        out->nleaves = 1;
        out->depth = 1;
    } else {
        std::vector<int> inner_index(bn.size(), -1);
        std::vector<int> bfs;
        std::vector<int> depth_of(bn.size(), 0);
        std::queue<int> qu;
        qu.push(0);
        while (!qu.empty()) {
            int id = qu.front();
            qu.pop();
            inner_index[id] = (int)bfs.size();
            bfs.push_back(id);
            for (int c = 0; c < 2; ++c) {
                int ch = bn[id].child[c];
                depth_of[ch] = depth_of[id] + 1;
                out->depth = std::max(out->depth, depth_of[ch] + 1);
                if (bn[ch].child[0] >= 0) qu.push(ch);
                else out->nleaves++;
            }
        }
        out->nodes.resize(bfs.size());
        for (size_t i = 0; i < bfs.size(); ++i) {
            const BuildNode& b = bn[bfs[i]];
            BvhNode n;
            std::memset(&n, 0, sizeof n);
            const BuildNode& l = bn[b.child[0]];
            const BuildNode& r = bn[b.child[1]];
            for (int k = 0; k < 3; ++k) {
                n.lo0[k] = down32(l.box.lo[k]);
                n.hi0[k] = up32(l.box.hi[k]);
                n.lo1[k] = down32(r.box.lo[k]);
                n.hi1[k] = up32(r.box.hi[k]);
            }
            n.c0 = l.child[0] >= 0 ? inner_index[b.child[0]] : leaf_link(l);
            n.c1 = r.child[0] >= 0 ? inner_index[b.child[1]] : leaf_link(r);
            out->nodes[i] = n;
        }
    }
    // fp32 error budget: vertices are rounded to nearest (<= ulp/2 per coordinate), the query
    // point likewise, and a point-triangle distance is ~20 dependent fp32 operations on values
    // of magnitude <= 2*maxabs: 64 ulps of that magnitude is a generous absolute bound.
    float mag = (float)(2.0 * maxabs + 1024.0);
    out->slack = 64.f * (std::nextafterf(mag, INFINITY) - mag);
}

}  // namespace ctrb
