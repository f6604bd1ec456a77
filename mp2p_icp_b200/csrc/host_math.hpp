// Small fixed-size host math of the solvers' closing steps (product code, no oracle dependency):
// SE(3) exp / compose (mrpt::poses::Lie::SE<3>::exp, CPose3D::operator+), 6x6 LDL^T solve
// (Eigen `H.ldlt().solve(g)`, optimal_tf_gauss_newton.cpp:351), symmetric 4x4 eigen-solve
// (CMatrixFixed::eig_symmetric, optimal_tf_horn.cpp:156-160) and quaternion -> rotation
// (CPose3D(CQuaternionDouble,...), optimal_tf_horn.cpp:238). These run once per solver call on
// 6..16 numbers; the O(N) work is in solve.cu.
#pragma once
#include <cmath>
#include <cstring>

#ifdef __CUDACC__
#define MP2P_HD __host__ __device__
#else
#define MP2P_HD
#endif

namespace mp2p
{
namespace hm
{
struct Pose34
{
    double m[12];
};

MP2P_HD inline Pose34 compose(const Pose34& a, const Pose34& b)
{
    Pose34 o;
    for (int r = 0; r < 3; r++)
    {
        for (int c = 0; c < 3; c++)
            o.m[4 * r + c] = a.m[4 * r] * b.m[c] + a.m[4 * r + 1] * b.m[4 + c] + a.m[4 * r + 2] * b.m[8 + c];
        o.m[4 * r + 3] = a.m[4 * r] * b.m[3] + a.m[4 * r + 1] * b.m[7] + a.m[4 * r + 2] * b.m[11] + a.m[4 * r + 3];
    }
    return o;
}

// xi = (v, w): R = exp([w]x), t = V(w) v
MP2P_HD inline Pose34 se3_exp(const double xi[6])
{
    const double wx = xi[3], wy = xi[4], wz = xi[5];
    const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
    double       A, B, C;
    if (th < 1e-6)
        A = 1.0 - th2 / 6.0, B = 0.5 - th2 / 24.0, C = 1.0 / 6.0 - th2 / 120.0;
    else
        A = sin(th) / th, B = (1.0 - cos(th)) / th2, C = (th - sin(th)) / (th2 * th);
    const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double       W2[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) W2[3 * r + c] = W[3 * r] * W[c] + W[3 * r + 1] * W[3 + c] + W[3 * r + 2] * W[6 + c];
    Pose34 o;
    for (int r = 0; r < 3; r++)
    {
        double t = 0;
        for (int c = 0; c < 3; c++)
        {
            const double I = (r == c) ? 1.0 : 0.0;
            o.m[4 * r + c] = I + A * W[3 * r + c] + B * W2[3 * r + c];
            t += (I + B * W[3 * r + c] + C * W2[3 * r + c]) * xi[c];
        }
        o.m[4 * r + 3] = t;
    }
    return o;
}

// x = H^{-1} b for symmetric 6x6 H, LDL^T with diagonal pivoting; null pivots contribute zero.
MP2P_HD inline void ldlt_solve6(const double Hin[36], const double b[6], double x[6])
{
    const int N = 6;
    double    A[36];
    for (int i = 0; i < 36; i++) A[i] = Hin[i];
    int perm[6] = {0, 1, 2, 3, 4, 5};
    double maxdiag = 0;
    for (int i = 0; i < N; i++) maxdiag = fmax(maxdiag, fabs(A[i * N + i]));
    const double tol = maxdiag * 2.220446049250313e-16 * N;
    for (int k = 0; k < N; k++)
    {
        int    piv = k;
        double best = fabs(A[k * N + k]);
        for (int i = k + 1; i < N; i++)
            if (fabs(A[i * N + i]) > best) best = fabs(A[i * N + i]), piv = i;
        if (piv != k)
        {
            for (int c = 0; c < N; c++)
            {
                const double t = A[k * N + c];
                A[k * N + c]   = A[piv * N + c];
                A[piv * N + c] = t;
            }
            for (int r = 0; r < N; r++)
            {
                const double t = A[r * N + k];
                A[r * N + k]   = A[r * N + piv];
                A[r * N + piv] = t;
            }
            const int tp = perm[k];
            perm[k]      = perm[piv];
            perm[piv]    = tp;
        }
        const double d = A[k * N + k];
        if (fabs(d) <= tol) continue;
        double col[6];
        for (int i = k + 1; i < N; i++) col[i] = A[i * N + k];
        for (int i = k + 1; i < N; i++)
        {
            const double l = col[i] / d;
            for (int j = k + 1; j <= i; j++)
            {
                A[i * N + j] -= l * col[j];
                A[j * N + i] = A[i * N + j];
            }
            A[i * N + k] = l;
        }
    }
    double y[6];
    for (int i = 0; i < N; i++) y[i] = b[perm[i]];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < i; j++) y[i] -= A[i * N + j] * y[j];
    for (int i = 0; i < N; i++)
    {
        const double d = A[i * N + i];
        y[i]           = (fabs(d) > tol) ? y[i] / d : 0.0;
    }
    for (int i = N - 1; i >= 0; i--)
        for (int j = i + 1; j < N; j++) y[i] -= A[j * N + i] * y[j];
    for (int i = 0; i < N; i++) x[perm[i]] = y[i];
}

// Same solve without pivoting, all loops fully unrolled on static indices (registers only — the
// device-side GN step is a single thread, where the pivoted version's dynamically indexed arrays
// live in local memory). Returns false if a pivot is numerically null; callers then fall back to
// the pivoted solve. For a positive definite H the two agree to rounding.
MP2P_HD inline bool ldlt_solve6_nopivot(const double Hin[36], const double b[6], double x[6])
{
    double A[36];
#pragma unroll
    for (int i = 0; i < 36; i++) A[i] = Hin[i];
    double maxdiag = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) maxdiag = fmax(maxdiag, fabs(A[i * 6 + i]));
    const double tol = maxdiag * 2.220446049250313e-16 * 6 * 1e3;
    bool         ok  = true;
#pragma unroll
    for (int k = 0; k < 6; k++)
    {
        const double d = A[k * 6 + k];
        if (!(fabs(d) > tol)) ok = false;
        const double inv = 1.0 / d;
#pragma unroll
        for (int i = k + 1; i < 6; i++)
        {
            const double l = A[i * 6 + k] * inv;
#pragma unroll
            for (int j = k + 1; j <= i; j++) A[i * 6 + j] -= l * A[j * 6 + k];
            A[k * 6 + i] = l;  // L^T above the diagonal; column k below it stays the original a_ik
        }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++)
    {
        double v = b[i];
#pragma unroll
        for (int j = 0; j < i; j++) v -= A[j * 6 + i] * y[j];
        y[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] = y[i] / A[i * 6 + i];
#pragma unroll
    for (int i = 5; i >= 0; i--)
    {
        double v = y[i];
#pragma unroll
        for (int j = i + 1; j < 6; j++) v -= A[i * 6 + j] * x[j];
        x[i] = v;
    }
    return ok;
}

// eigenvector (unit) of the LARGEST eigenvalue of a symmetric 4x4, by cyclic Jacobi.
inline void eig_sym4_largest(const double Nin[16], double q[4])
{
    double A[16], V[16];
    std::memcpy(A, Nin, sizeof(A));
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) V[4 * r + c] = (r == c) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; sweep++)
    {
        double off = 0, diag = 0;
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) (r == c ? diag : off) += A[4 * r + c] * A[4 * r + c];
        if (off <= 1e-300 || off <= 1e-32 * diag) break;
        for (int p = 0; p < 3; p++)
            for (int qq = p + 1; qq < 4; qq++)
            {
                const double apq = A[4 * p + qq];
                if (apq == 0.0) continue;
                const double tau = (A[4 * qq + qq] - A[4 * p + p]) / (2.0 * apq);
                const double t   = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < 4; k++)
                {
                    const double akp = A[4 * k + p], akq = A[4 * k + qq];
                    A[4 * k + p] = c * akp - s * akq, A[4 * k + qq] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; k++)
                {
                    const double apk = A[4 * p + k], aqk = A[4 * qq + k];
                    A[4 * p + k] = c * apk - s * aqk, A[4 * qq + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; k++)
                {
                    const double vkp = V[4 * k + p], vkq = V[4 * k + qq];
                    V[4 * k + p] = c * vkp - s * vkq, V[4 * k + qq] = s * vkp + c * vkq;
                }
            }
    }
    int m = 0;
    for (int i = 1; i < 4; i++)
        if (A[5 * i] > A[5 * m]) m = i;
    for (int k = 0; k < 4; k++) q[k] = V[4 * k + m];
}

inline Pose34 pose_from_quat(const double q[4])
{
    const double r = q[0], x = q[1], y = q[2], z = q[3];
    Pose34       p{};
    p.m[0] = r * r + x * x - y * y - z * z, p.m[1] = 2 * (x * y - r * z), p.m[2] = 2 * (z * x + r * y);
    p.m[4] = 2 * (x * y + r * z), p.m[5] = r * r - x * x + y * y - z * z, p.m[6] = 2 * (y * z - r * x);
    p.m[8] = 2 * (z * x - r * y), p.m[9] = 2 * (y * z + r * x), p.m[10] = r * r - x * x - y * y + z * z;
    return p;
}
}  // namespace hm
}  // namespace mp2p
