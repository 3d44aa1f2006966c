// kp_umeyama.cuh -- rigid alignment of matched point sets from their sums (shared by the point-to-point ICP
// pass, kp_icp.cu, and the hypothesis kernel of the global registration, kp_global.cu)
#pragma once
#include <math.h>

// TransformationEstimationPointToPoint::ComputeTransformation = Eigen::umeyama without scaling
// (manual_pointcloud_registration.py:90-98): R = U diag(1, 1, det(U) det(V)) V^T from the SVD of the
// cross-covariance, t = mean_t - R mean_s.  The 3x3 SVD is built from the Jacobi eigenvectors of C^T C:
// with v3 := v1 x v2 and u3 := u1 x u2 both factors are proper rotations and R = [u1 u2 u3][v1 v2 v3]^T is
// that product for either sign of det C.  tot: sum s (0..2), sum t (3..5), sum t_i s_j (6 + 3 i + j).
__device__ __noinline__ static void kp_umeyama(const double *tot, double n, double *Un)
{
    double ms[3], mt[3], Cm[3][3];
    for (int i = 0; i < 3; ++i) { ms[i] = tot[i] / n; mt[i] = tot[3 + i] / n; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Cm[i][j] = tot[6 + 3 * i + j] / n - mt[i] * ms[j];
    // A = C^T C, Jacobi eigen-decomposition (cyclic sweeps)
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += Cm[k][i] * Cm[k][j]; A[i][j] = a; }
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300 || off < 1e-22 * (fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]))) break;
        for (int pq = 0; pq < 3; ++pq) {
            const int pi = pq == 2 ? 1 : 0, qi = pq == 0 ? 1 : 2;
            if (A[pi][qi] == 0.0) continue;
            const double theta = (A[qi][qi] - A[pi][pi]) / (2.0 * A[pi][qi]);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
            for (int k = 0; k < 3; ++k) { const double akp = A[k][pi], akq = A[k][qi]; A[k][pi] = c * akp - sn * akq; A[k][qi] = sn * akp + c * akq; }
            for (int k = 0; k < 3; ++k) { const double apk = A[pi][k], aqk = A[qi][k]; A[pi][k] = c * apk - sn * aqk; A[qi][k] = sn * apk + c * aqk; }
            for (int k = 0; k < 3; ++k) { const double vkp = V[k][pi], vkq = V[k][qi]; V[k][pi] = c * vkp - sn * vkq; V[k][qi] = sn * vkp + c * vkq; }
        }
    }
    // order the eigenpairs by descending eigenvalue
    int o[3] = {0, 1, 2};
    for (int a = 0; a < 2; ++a) for (int b = a + 1; b < 3; ++b) if (A[o[b]][o[b]] > A[o[a]][o[a]]) { const int t = o[a]; o[a] = o[b]; o[b] = t; }
    double v1[3], v2[3], v3[3], u1[3], u2[3], u3[3];
    for (int k = 0; k < 3; ++k) { v1[k] = V[k][o[0]]; v2[k] = V[k][o[1]]; }
    v3[0] = v1[1] * v2[2] - v1[2] * v2[1]; v3[1] = v1[2] * v2[0] - v1[0] * v2[2]; v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
    for (int i = 0; i < 3; ++i) { u1[i] = Cm[i][0] * v1[0] + Cm[i][1] * v1[1] + Cm[i][2] * v1[2]; u2[i] = Cm[i][0] * v2[0] + Cm[i][1] * v2[1] + Cm[i][2] * v2[2]; }
    double l1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    if (!(l1 > 1e-300)) return;                             // zero cross-covariance: keep the identity
    for (int i = 0; i < 3; ++i) u1[i] /= l1;
    double d12 = u2[0] * u1[0] + u2[1] * u1[1] + u2[2] * u1[2];
    for (int i = 0; i < 3; ++i) u2[i] -= d12 * u1[i];
    double l2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    if (!(l2 > 1e-12 * l1)) {                               // rank 1: any unit vector orthogonal to u1
        const int ax = fabs(u1[0]) <= fabs(u1[1]) && fabs(u1[0]) <= fabs(u1[2]) ? 0 : (fabs(u1[1]) <= fabs(u1[2]) ? 1 : 2);
        double e[3] = {0, 0, 0}; e[ax] = 1.0;
        const double d = u1[ax];
        for (int i = 0; i < 3; ++i) u2[i] = e[i] - d * u1[i];
        l2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    }
    for (int i = 0; i < 3; ++i) u2[i] /= l2;
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1]; u3[1] = u1[2] * u2[0] - u1[0] * u2[2]; u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    double Rm[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rm[i][j] = u1[i] * v1[j] + u2[i] * v2[j] + u3[i] * v3[j];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Un[4 * i + j] = Rm[i][j];
        Un[4 * i + 3] = mt[i] - (Rm[i][0] * ms[0] + Rm[i][1] * ms[1] + Rm[i][2] * ms[2]);
    }
}

