// ctrb_se3.cuh -- SE(3) arithmetic in registers (3x4 blocks; the constant last row is never stored).
#pragma once

#include <cuda_runtime.h>

#include "ctrb_internal.cuh"

namespace ctrb {

struct SE3 {
    double r[9];  // row-major rotation
    double p[3];
};

__device__ __forceinline__ void se3_identity(SE3& g) {
#pragma unroll
    for (int i = 0; i < 9; ++i) g.r[i] = (i % 4 == 0) ? 1.0 : 0.0;
    g.p[0] = g.p[1] = g.p[2] = 0.0;
}

// one column of a model table (SoA, see ModelTables)
__device__ __forceinline__ SE3 load_tab(const double* __restrict__ tab, int tot, int col) {
    SE3 g;
#pragma unroll
    for (int k = 0; k < 9; ++k) g.r[k] = tab[k * tot + col];
#pragma unroll
    for (int k = 0; k < 3; ++k) g.p[k] = tab[(9 + k) * tot + col];
    return g;
}

// out = a * b
__device__ __forceinline__ SE3 se3_mul(const SE3& a, const SE3& b) {
    SE3 o;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
            o.r[i * 3 + j] = fma(a.r[i * 3 + 0], b.r[0 * 3 + j], fma(a.r[i * 3 + 1], b.r[1 * 3 + j], a.r[i * 3 + 2] * b.r[2 * 3 + j]));
        o.p[i] = fma(a.r[i * 3 + 0], b.p[0], fma(a.r[i * 3 + 1], b.p[1], fma(a.r[i * 3 + 2], b.p[2], a.p[i])));
    }
    return o;
}

// (R, p)^-1 = (R^T, -R^T p): replaces numpy.linalg.inv on SE(3) matrices (kinematic.py:124,191)
__device__ __forceinline__ SE3 se3_inv(const SE3& g) {
    SE3 o;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o.r[i * 3 + j] = g.r[j * 3 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) o.p[i] = -(o.r[i * 3 + 0] * g.p[0] + o.r[i * 3 + 1] * g.p[1] + o.r[i * 3 + 2] * g.p[2]);
    return o;
}

// g <- g * Rx(theta): what exponential_map(theta, hat(theta e_x)) multiplies in at kinematic.py:95-97
__device__ __forceinline__ void se3_rot_x(SE3& g, double theta) {
    double s, c;
    sincos(theta, &s, &c);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double y = g.r[i * 3 + 1], z = g.r[i * 3 + 2];
        g.r[i * 3 + 1] = fma(c, y, s * z);
        g.r[i * 3 + 2] = fma(c, z, -s * y);
    }
}

// strain field omega(x) = (B(x) q)[0:3] for every base of strain_bases.py:7-224; the linear
// part of the twist is the bias [1,0,0] (kinematic.py:49) because no base touches rows 3..5.
__device__ __forceinline__ void strain_omega(const ModelConst& mc, int n, double x, double w[3]) {
    const double* q = mc.q[n];
    const int dof = mc.dof[n];
    w[0] = w[1] = w[2] = 0.0;
    switch (mc.kind[n]) {
        case CTRB_BASE_CONSTANT:
            w[1] = q[0];
            break;
        case CTRB_BASE_LINEAR:
            w[1] = dof > 2 ? fma(x, q[1], q[0]) : q[0];
            w[2] = x * q[dof - 1];
            break;
        case CTRB_BASE_QUADRATIC:
            w[1] = dof > 2 ? fma(x * x, q[1], q[0]) : q[0];
            w[2] = (x * x) * q[dof - 1];
            break;
        case CTRB_BASE_PURE_HELIX:
            w[1] = sin(x / 10) * q[0];
            w[2] = cos(x / 10) * q[1];
            break;
        case CTRB_BASE_LINEAR_HELIX:
            w[1] = sin(x / 10) * q[0];
            w[2] = (x * cos(x / 10)) * q[1];
            break;
        case CTRB_BASE_TORSION_HELIX:
            w[0] = q[0];
            w[1] = q[1];
            w[2] = q[2];
            break;
        case CTRB_BASE_TORSION_LINEAR_HELIX:
            w[0] = q[0];
            w[1] = q[1];
            w[2] = x * q[2];
            break;
        case CTRB_BASE_FULL: {
            int bs = (dof & 1) ? 1 : 2;
            w[0] = (dof & 1) ? q[0] : fma(x, q[1], q[0]);
            if (dof <= 6) {
                w[1] = fma(x, q[bs + 1], q[bs]);
                w[2] = fma(x, q[bs + 3], q[bs + 2]);
            } else {
                w[1] = fma(x * x, q[bs + 2], fma(x, q[bs + 1], q[bs]));
                w[2] = fma(x * x, q[bs + 5], fma(x, q[bs + 4], q[bs + 3]));
            }
            break;
        }
        default:
            break;
    }
}

// exp of the se(3) element dx * hat([w; v]) in closed form (matrix_utils.py:157-177):
// R = I + (sin phi/phi) W + a W^2,  p = (I + a W + c W^2) V, W = dx w^, V = dx v, phi = |dx w|,
// a = (1-cos phi)/phi^2, c = (phi - sin phi)/phi^3; the phi == 0 branch gives I + Gamma exactly.
__device__ __forceinline__ SE3 se3_exp(const double w_in[3], const double v_in[3], double dx) {
    double w[3] = {dx * w_in[0], dx * w_in[1], dx * w_in[2]};
    double v[3] = {dx * v_in[0], dx * v_in[1], dx * v_in[2]};
    double phi2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double phi = sqrt(phi2);
    SE3 g;
    if (phi == 0.0) {
        se3_identity(g);
        g.p[0] = v[0];
        g.p[1] = v[1];
        g.p[2] = v[2];
        return g;
    }
    double sh = sin(0.5 * phi);
    double a = 2.0 * sh * sh / phi2;          // (1 - cos phi)/phi^2 without cancellation
    double b = sin(phi) / phi;                // 1 - c phi^2
    double c;
    if (phi < 1e-2) {
        c = (1.0 / 6.0) * (1.0 - phi2 * (1.0 / 20.0) * (1.0 - phi2 * (1.0 / 42.0) * (1.0 - phi2 / 72.0)));
    } else {
        c = (phi - sin(phi)) / (phi2 * phi);
    }
    // W^2 = w w^T - phi^2 I
    double ww[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) ww[i * 3 + j] = w[i] * w[j] - (i == j ? phi2 : 0.0);
    double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
#pragma unroll
    for (int i = 0; i < 9; ++i) g.r[i] = ((i % 4 == 0) ? 1.0 : 0.0) + b * W[i] + a * ww[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double wv = W[i * 3 + 0] * v[0] + W[i * 3 + 1] * v[1] + W[i * 3 + 2] * v[2];
        double wwv = ww[i * 3 + 0] * v[0] + ww[i * 3 + 1] * v[1] + ww[i * 3 + 2] * v[2];
        g.p[i] = v[i] + a * wv + c * wwv;
    }
    return g;
}

}  // namespace ctrb
