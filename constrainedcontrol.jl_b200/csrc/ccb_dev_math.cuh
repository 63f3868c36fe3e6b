// csrc/ccb_dev_math.cuh -- small FP64 vector / quaternion helpers for the device kernels (sm_100a)
#pragma once
#include <cuda_runtime.h>

namespace ccb {

struct V3 { double x, y, z; };
struct Q4 { double s, x, y, z; };

__device__ __forceinline__ V3 v3(double x, double y, double z) { return {x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
    return {a.s * b.s - a.x * b.x - a.y * b.y - a.z * b.z,
            a.s * b.x + b.s * a.x + (a.y * b.z - a.z * b.y),
            a.s * b.y + b.s * a.y + (a.z * b.x - a.x * b.z),
            a.s * b.z + b.s * a.z + (a.x * b.y - a.y * b.x)};
}
__device__ __forceinline__ Q4 qconj(Q4 a) { return {a.s, -a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 qvec(Q4 a) { return {a.x, a.y, a.z}; }
// R(q) p
__device__ __forceinline__ V3 qrot(Q4 q, V3 p) {
    V3 v = qvec(q);
    V3 t = 2.0 * cross(v, p);
    return p + q.s * t + cross(v, t);
}
// R(q)^T p
__device__ __forceinline__ V3 qrotT(Q4 q, V3 p) {
    V3 v = {-q.x, -q.y, -q.z};
    V3 t = 2.0 * cross(v, p);
    return p + q.s * t + cross(v, t);
}
__device__ __forceinline__ V3 ld3(const double* p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ Q4 ld4(const double* p) { return {p[0], p[1], p[2], p[3]}; }
__device__ __forceinline__ void st3(double* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ void st4(double* p, Q4 a) { p[0] = a.s; p[1] = a.x; p[2] = a.y; p[3] = a.z; }
__device__ __forceinline__ V3 ldg3(const double* p) { return {__ldg(p), __ldg(p + 1), __ldg(p + 2)}; }
__device__ __forceinline__ Q4 ldg4(const double* p) { return {__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3)}; }

// y = M x for a row-major 3x3 in memory
__device__ __forceinline__ V3 mat3_mul(const double* M, V3 x) {
    return {fma(M[0], x.x, fma(M[1], x.y, M[2] * x.z)), fma(M[3], x.x, fma(M[4], x.y, M[5] * x.z)), fma(M[6], x.x, fma(M[7], x.y, M[8] * x.z))};
}
// y^T = x^T M
__device__ __forceinline__ V3 vec_mat3(V3 x, const double* M) {
    return {fma(x.x, M[0], fma(x.y, M[3], x.z * M[6])), fma(x.x, M[1], fma(x.y, M[4], x.z * M[7])), fma(x.x, M[2], fma(x.y, M[5], x.z * M[8]))};
}

// One scalar constraint row: value g and its tangent Jacobian rows w.r.t. (x, phi) of the
// parent (Ga) and the child (Gb).  kind 0: translational row pr . (R_a^T (x_b + R_b p2 - x_a) - p1);
// kind 1: rotational row pr . vec(q_a^-1 q_b q_off^-1).
__device__ __forceinline__ void row_eval(int kind, V3 pr, V3 p1, V3 p2, Q4 qoff, V3 xa, Q4 qa, V3 xb, Q4 qb, bool jac, double& g, V3& Gax,
                                         V3& Gap, V3& Gbx, V3& Gbp) {
    if (kind == 0) {
        V3 w = xb + qrot(qb, p2) - xa;
        V3 c = qrotT(qa, w);
        g = dot(pr, c - p1);
        if (jac) {
            V3 n = qrot(qa, pr);        // world direction of this row
            V3 m = qrotT(qb, n);
            Gbx = n;
            Gbp = -2.0 * cross(m, p2);
            Gax = -n;
            Gap = 2.0 * cross(pr, c);
        }
    } else {
        Q4 qab = qmul(qconj(qa), qb);
        Q4 qe = qmul(qab, qconj(qoff));
        g = pr.x * qe.x + pr.y * qe.y + pr.z * qe.z;
        if (jac) {
            Q4 e = {0.0, pr.x, pr.y, pr.z};
            Gbx = {0.0, 0.0, 0.0};
            Gax = {0.0, 0.0, 0.0};
            Gbp = qvec(qmul(qmul(qconj(qab), e), qoff));
            Gap = -qvec(qmul(e, qconj(qe)));
        }
    }
}

// The same row, split for kernels whose lanes hold rows of different kinds (K1r, K1m): the pose-only part that every row of a
// joint shares is formed by all lanes without divergence, then each lane finishes its own row.
struct JointFrame { V3 c; Q4 qab, qe; };   // c = R_a^T (x_b + R_b p2 - x_a),  qab = q_a^-1 q_b,  qe = qab q_off^-1
__device__ __forceinline__ JointFrame joint_frame(V3 p2, Q4 qoff, V3 xa, Q4 qa, V3 xb, Q4 qb) {
    JointFrame f;
    f.c = qrotT(qa, xb + qrot(qb, p2) - xa);
    f.qab = qmul(qconj(qa), qb);
    f.qe = qmul(f.qab, qconj(qoff));
    return f;
}
__device__ __forceinline__ void row_finish(int kind, V3 pr, V3 p1, V3 p2, Q4 qoff, Q4 qa, Q4 qb, const JointFrame& f, bool jac, double& g, V3& Gax,
                                           V3& Gap, V3& Gbx, V3& Gbp) {
    if (kind == 0) {
        g = dot(pr, f.c - p1);
        if (jac) {
            V3 n = qrot(qa, pr);
            V3 m = qrotT(qb, n);
            Gbx = n;
            Gbp = -2.0 * cross(m, p2);
            Gax = -n;
            Gap = 2.0 * cross(pr, f.c);
        }
    } else {
        g = pr.x * f.qe.x + pr.y * f.qe.y + pr.z * f.qe.z;
        if (jac) {
            Q4 e = {0.0, pr.x, pr.y, pr.z};
            Gbx = {0.0, 0.0, 0.0};
            Gax = {0.0, 0.0, 0.0};
            Gbp = qvec(qmul(qmul(qconj(f.qab), e), qoff));
            Gap = -qvec(qmul(e, qconj(f.qe)));
        }
    }
}

// inverse of a row-major 3x3 by the adjugate
__device__ __forceinline__ void inv3(const double* A, double* Ai) {
    double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    double det = fma(A[0], c00, fma(A[1], c01, A[2] * c02));
    double id = 1.0 / det;
    Ai[0] = c00 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    Ai[3] = c01 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    Ai[6] = c02 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

template <int G> __device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ccb
