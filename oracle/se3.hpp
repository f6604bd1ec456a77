// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into, imported by, or called from the product
// path (mp2p_icp_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it.
//
// Minimal SE(3) algebra restating the MRPT semantics that the reference hot path depends on.
// MRPT itself is NOT under /root/reference (external dependency, "MRPT >= 2.11.5",
// CMakeLists.txt:53-59), so these follow the public MRPT 2.x definitions; each item names the
// reference call site that relies on it.
#pragma once
#include <cmath>
#include <cstring>

namespace orc
{
// Pose = 3x4 row-major [R | t]:  T[4*r + c], c<3 rotation, c==3 translation.
struct Pose
{
    double m[12];
    double R(int r, int c) const { return m[4 * r + c]; }
    double t(int r) const { return m[4 * r + 3]; }
};

inline Pose pose_identity()
{
    Pose p;
    std::memset(p.m, 0, sizeof(p.m));
    p.m[0] = p.m[5] = p.m[10] = 1.0;
    return p;
}

// mrpt::poses::CPose3D(x,y,z,yaw,pitch,roll): R = Rz(yaw)*Ry(pitch)*Rx(roll).
// Used by apps/icp-run/main.cpp:98-103,262 and tests/test-mp2p_matcher_pt2pt.cpp:103.
inline Pose pose_from_xyzypr(double x, double y, double z, double yaw, double pitch, double roll)
{
    const double cy = std::cos(yaw), sy = std::sin(yaw);
    const double cp = std::cos(pitch), sp = std::sin(pitch);
    const double cr = std::cos(roll), sr = std::sin(roll);
    Pose         p;
    p.m[0]  = cy * cp;
    p.m[1]  = cy * sp * sr - sy * cr;
    p.m[2]  = cy * sp * cr + sy * sr;
    p.m[3]  = x;
    p.m[4]  = sy * cp;
    p.m[5]  = sy * sp * sr + cy * cr;
    p.m[6]  = sy * sp * cr - cy * sr;
    p.m[7]  = y;
    p.m[8]  = -sp;
    p.m[9]  = cp * sr;
    p.m[10] = cp * cr;
    p.m[11] = z;
    return p;
}

// CPose3D::composePoint(double...) : g = R*l + t, evaluated left to right in double.
// Depended on by mp2p_icp/src/errorTerms.cpp:44,125.
inline void compose_point(const Pose& T, double lx, double ly, double lz, double& gx, double& gy,
                          double& gz)
{
    gx = T.m[0] * lx + T.m[1] * ly + T.m[2] * lz + T.m[3];
    gy = T.m[4] * lx + T.m[5] * ly + T.m[6] * lz + T.m[7];
    gz = T.m[8] * lx + T.m[9] * ly + T.m[10] * lz + T.m[11];
}

// CPose3D::composePoint(float lx.., float& gx..): computed in double, rounded to float.
// Defines the BITS of the transformed query coordinates (mp2p_icp/src/Matcher_Points_Base.cpp:216).
inline void compose_point_f(const Pose& T, float lx, float ly, float lz, float& gx, float& gy,
                            float& gz)
{
    double x, y, z;
    compose_point(T, lx, ly, lz, x, y, z);
    gx = static_cast<float>(x);
    gy = static_cast<float>(y);
    gz = static_cast<float>(z);
}

// a + b  (pose composition a∘b).  ICP.cpp:166; optimal_tf_gauss_newton.cpp:356.
inline Pose compose(const Pose& a, const Pose& b)
{
    Pose o;
    for (int r = 0; r < 3; r++)
    {
        for (int c = 0; c < 3; c++)
            o.m[4 * r + c] = a.R(r, 0) * b.R(0, c) + a.R(r, 1) * b.R(1, c) + a.R(r, 2) * b.R(2, c);
        o.m[4 * r + 3] = a.R(r, 0) * b.t(0) + a.R(r, 1) * b.t(1) + a.R(r, 2) * b.t(2) + a.t(r);
    }
    return o;
}

inline Pose inverse(const Pose& a)
{
    Pose o;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) o.m[4 * r + c] = a.R(c, r);
    for (int r = 0; r < 3; r++)
        o.m[4 * r + 3] = -(a.R(0, r) * a.t(0) + a.R(1, r) * a.t(1) + a.R(2, r) * a.t(2));
    return o;
}

// a - b = inverse(b) ∘ a.   ICP.cpp:166,203.
inline Pose inverse_compose(const Pose& a, const Pose& b) { return compose(inverse(b), a); }

// CPose3D::inverseComposePoint: l = R^T (g - t)
inline void inverse_compose_point(const Pose& T, double gx, double gy, double gz, double& lx,
                                  double& ly, double& lz)
{
    const double dx = gx - T.t(0), dy = gy - T.t(1), dz = gz - T.t(2);
    lx              = T.R(0, 0) * dx + T.R(1, 0) * dy + T.R(2, 0) * dz;
    ly              = T.R(0, 1) * dx + T.R(1, 1) * dy + T.R(2, 1) * dz;
    lz              = T.R(0, 2) * dx + T.R(1, 2) * dy + T.R(2, 2) * dz;
}

// mrpt::poses::Lie::SE<3>::exp, tangent order (v, w):  R = exp_SO3(w),  t = V(w) v.
// optimal_tf_gauss_newton.cpp:354; ICP.cpp:194-196.
inline Pose se3_exp(const double xi[6])
{
    const double vx = xi[0], vy = xi[1], vz = xi[2];
    const double wx = xi[3], wy = xi[4], wz = xi[5];
    const double th2 = wx * wx + wy * wy + wz * wz;
    const double th  = std::sqrt(th2);
    double       A, B, C;  // sin(th)/th, (1-cos th)/th^2, (th - sin th)/th^3
    if (th < 1e-6)
    {
        A = 1.0 - th2 / 6.0;
        B = 0.5 - th2 / 24.0;
        C = 1.0 / 6.0 - th2 / 120.0;
    }
    else
    {
        A = std::sin(th) / th;
        B = (1.0 - std::cos(th)) / th2;
        C = (th - std::sin(th)) / (th2 * th);
    }
    // W = [w]x ; W2 = W*W
    const double W[9]  = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double       W2[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            W2[3 * r + c] = W[3 * r + 0] * W[0 + c] + W[3 * r + 1] * W[3 + c] + W[3 * r + 2] * W[6 + c];
    Pose o;
    for (int r = 0; r < 3; r++)
    {
        double Vr[3];
        for (int c = 0; c < 3; c++)
        {
            const double I = (r == c) ? 1.0 : 0.0;
            o.m[4 * r + c] = I + A * W[3 * r + c] + B * W2[3 * r + c];
            Vr[c]          = I + B * W[3 * r + c] + C * W2[3 * r + c];
        }
        o.m[4 * r + 3] = Vr[0] * vx + Vr[1] * vy + Vr[2] * vz;
    }
    return o;
}

// SO(3) log -> rotation vector.
inline void so3_log(const Pose& T, double w[3])
{
    const double tr = T.R(0, 0) + T.R(1, 1) + T.R(2, 2);
    double       c  = 0.5 * (tr - 1.0);
    if (c > 1.0) c = 1.0;
    if (c < -1.0) c = -1.0;
    const double th = std::acos(c);
    const double ax = T.R(2, 1) - T.R(1, 2), ay = T.R(0, 2) - T.R(2, 0), az = T.R(1, 0) - T.R(0, 1);
    if (th < 1e-7)
    {
        const double k = 0.5 * (1.0 + th * th / 6.0);
        w[0] = k * ax, w[1] = k * ay, w[2] = k * az;
        return;
    }
    if (M_PI - th < 1e-6)
    {
        // near pi: use diagonal to recover axis
        double xx = std::sqrt(std::fmax(0.0, 0.5 * (T.R(0, 0) + 1.0)));
        double yy = std::sqrt(std::fmax(0.0, 0.5 * (T.R(1, 1) + 1.0)));
        double zz = std::sqrt(std::fmax(0.0, 0.5 * (T.R(2, 2) + 1.0)));
        // fix signs using off-diagonals
        if (xx >= yy && xx >= zz)
        {
            yy = std::copysign(yy, T.R(0, 1) + T.R(1, 0));
            zz = std::copysign(zz, T.R(0, 2) + T.R(2, 0));
        }
        else if (yy >= xx && yy >= zz)
        {
            xx = std::copysign(xx, T.R(0, 1) + T.R(1, 0));
            zz = std::copysign(zz, T.R(1, 2) + T.R(2, 1));
        }
        else
        {
            xx = std::copysign(xx, T.R(0, 2) + T.R(2, 0));
            yy = std::copysign(yy, T.R(1, 2) + T.R(2, 1));
        }
        // orientation of the axis from the (small) antisymmetric part
        if (ax * xx + ay * yy + az * zz < 0) xx = -xx, yy = -yy, zz = -zz;
        w[0] = th * xx, w[1] = th * yy, w[2] = th * zz;
        return;
    }
    const double k = th / (2.0 * std::sin(th));
    w[0] = k * ax, w[1] = k * ay, w[2] = k * az;
}

// mrpt::poses::Lie::SE<3>::log -> (v, w) with v = V^{-1} t.   ICP.cpp:194-196.
inline void se3_log(const Pose& T, double xi[6])
{
    double w[3];
    so3_log(T, w);
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double th  = std::sqrt(th2);
    // V^{-1} = I - 1/2 W + D W^2,   D = (1 - A/(2B)) / th^2
    double D;
    if (th < 1e-6)
        D = 1.0 / 12.0 + th2 / 720.0;
    else
    {
        const double A = std::sin(th) / th, B = (1.0 - std::cos(th)) / th2;
        D              = (1.0 - A / (2.0 * B)) / th2;
    }
    const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double       W2[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            W2[3 * r + c] = W[3 * r + 0] * W[0 + c] + W[3 * r + 1] * W[3 + c] + W[3 * r + 2] * W[6 + c];
    for (int r = 0; r < 3; r++)
    {
        double v = 0;
        for (int c = 0; c < 3; c++)
        {
            const double I = (r == c) ? 1.0 : 0.0;
            v += (I - 0.5 * W[3 * r + c] + D * W2[3 * r + c]) * T.t(c);
        }
        xi[r] = v;
    }
    xi[3] = w[0], xi[4] = w[1], xi[5] = w[2];
}

// CPose3D(CQuaternionDouble q=(r,x,y,z), x,y,z) -> rotation.  optimal_tf_horn.cpp:238.
inline Pose pose_from_quat(double r, double x, double y, double z)
{
    Pose p = pose_identity();
    p.m[0]  = r * r + x * x - y * y - z * z;
    p.m[1]  = 2 * (x * y - r * z);
    p.m[2]  = 2 * (z * x + r * y);
    p.m[4]  = 2 * (x * y + r * z);
    p.m[5]  = r * r - x * x + y * y - z * z;
    p.m[6]  = 2 * (y * z - r * x);
    p.m[8]  = 2 * (z * x - r * y);
    p.m[9]  = 2 * (y * z + r * x);
    p.m[10] = r * r - x * x - y * y + z * z;
    return p;
}

// Cyclic Jacobi for small symmetric matrices (N = 3 or 4). Restates what
// CMatrixFixed::eig_symmetric(V, vals, sorted=true) delivers: eigenvalues ASCENDING, matching
// eigenvectors in the COLUMNS of V (estimate_points_eigen.cpp:106-118; optimal_tf_horn.cpp:156-160).
// Reads the lower triangle like MRPT/Eigen's SelfAdjointEigenSolver.
template <int N>
inline void eig_symmetric(const double* Ain /*NxN row-major*/, double* V /*NxN row-major*/,
                          double* vals /*N*/)
{
    double A[N * N];
    for (int r = 0; r < N; r++)
        for (int c = 0; c < N; c++) A[r * N + c] = (c <= r) ? Ain[r * N + c] : Ain[c * N + r];
    for (int r = 0; r < N; r++)
        for (int c = 0; c < N; c++) V[r * N + c] = (r == c) ? 1.0 : 0.0;

    for (int sweep = 0; sweep < 64; sweep++)
    {
        double off = 0, diag = 0;
        for (int r = 0; r < N; r++)
            for (int c = 0; c < N; c++)
                if (r != c)
                    off += A[r * N + c] * A[r * N + c];
                else
                    diag += A[r * N + c] * A[r * N + c];
        if (off <= 1e-300 || off <= 1e-32 * diag) break;
        for (int p = 0; p < N - 1; p++)
            for (int q = p + 1; q < N; q++)
            {
                const double apq = A[p * N + q];
                if (apq == 0.0) continue;
                const double app = A[p * N + p], aqq = A[q * N + q];
                const double tau = (aqq - app) / (2.0 * apq);
                const double t =
                    (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < N; k++)
                {
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq;
                    A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; k++)
                {
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk;
                    A[q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; k++)
                {
                    const double vkp = V[k * N + p], vkq = V[k * N + q];
                    V[k * N + p] = c * vkp - s * vkq;
                    V[k * N + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < N; i++) vals[i] = A[i * N + i];
    // selection sort ascending, permuting columns of V
    for (int i = 0; i < N - 1; i++)
    {
        int m = i;
        for (int j = i + 1; j < N; j++)
            if (vals[j] < vals[m]) m = j;
        if (m != i)
        {
            const double tv = vals[i];
            vals[i]         = vals[m];
            vals[m]         = tv;
            for (int k = 0; k < N; k++)
            {
                const double t = V[k * N + i];
                V[k * N + i]   = V[k * N + m];
                V[k * N + m]   = t;
            }
        }
    }
}

// 6x6 symmetric solve  x = H^{-1} b  through LDL^T with diagonal pivoting; pivots that are
// numerically zero contribute 0 (what Eigen's LDLT::solve does for the reference:
// optimal_tf_gauss_newton.cpp:351 `H.ldlt().solve(g)`).
inline void ldlt_solve6(const double Hin[36], const double b[6], double x[6])
{
    const int N = 6;
    double    A[36];
    std::memcpy(A, Hin, sizeof(A));
    int perm[6];
    for (int i = 0; i < N; i++) perm[i] = i;
    double maxdiag = 0;
    for (int i = 0; i < N; i++) maxdiag = std::fmax(maxdiag, std::fabs(A[i * N + i]));
    const double tol = maxdiag * 2.220446049250313e-16 * 6;
    for (int k = 0; k < N; k++)
    {
        int    piv  = k;
        double best = std::fabs(A[k * N + k]);
        for (int i = k + 1; i < N; i++)
            if (std::fabs(A[i * N + i]) > best) best = std::fabs(A[i * N + i]), piv = i;
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
            const int t = perm[k];
            perm[k]     = perm[piv];
            perm[piv]   = t;
        }
        const double d = A[k * N + k];
        if (std::fabs(d) <= tol) continue;
        double col[6];
        for (int i = k + 1; i < N; i++) col[i] = A[i * N + k];  // original a_ik
        for (int i = k + 1; i < N; i++)
        {
            const double lik = col[i] / d;
            for (int j = k + 1; j <= i; j++)
            {
                A[i * N + j] -= lik * col[j];
                A[j * N + i] = A[i * N + j];
            }
            A[i * N + k] = lik;  // store L below the diagonal
        }
    }
    // Solve P^T L D L^T P x = b
    double y[6];
    for (int i = 0; i < N; i++) y[i] = b[perm[i]];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < i; j++) y[i] -= A[i * N + j] * y[j];
    for (int i = 0; i < N; i++)
    {
        const double d = A[i * N + i];
        y[i]           = (std::fabs(d) > tol) ? y[i] / d : 0.0;
    }
    for (int i = N - 1; i >= 0; i--)
        for (int j = i + 1; j < N; j++) y[i] -= A[j * N + i] * y[j];
    for (int i = 0; i < N; i++) x[perm[i]] = y[i];
}

}  // namespace orc
