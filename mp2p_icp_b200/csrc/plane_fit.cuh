// Per-query plane fit of the pt2pl matcher (product code): estimate_points_eigen
// (mp2p_icp_map/src/estimate_points_eigen.cpp:27-123) + the planarity / distance tests of
// mp2p_icp/src/Matcher_Adaptive.cpp:229-253, i.e. NearestPlaneCapable::nn_search_pt2pl realised
// over a plain point layer. Compiled with -fmad=false: the float mean and the centred moments round exactly as in the
// CPU statement of the algorithm; the 3x3 eigen-solve uses a cheaper (mathematically identical)
// rotation formula, so plane coefficients agree to ~1e-15 relative, not bit for bit.
#pragma once
#include "common.cuh"

namespace mp2p
{
struct PlaneCandidate
{
    double coefs[4];
    double centroid[3];
};

// Cyclic Jacobi, 3x3 symmetric, eigenvalues ascending with matching eigenvector columns
// (what CMatrixFixed::eig_symmetric(V, vals) delivers).
__device__ __forceinline__ void eig_sym3(const double Ain[9], double V[9], double vals[3])
{
    double A[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) A[r * 3 + c] = (c <= r) ? Ain[r * 3 + c] : Ain[c * 3 + r];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) V[r * 3 + c] = (r == c) ? 1.0 : 0.0;

    for (int sweep = 0; sweep < 64; sweep++)
    {
        double off = 0, diag = 0;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++)
                if (r != c)
                    off += A[r * 3 + c] * A[r * 3 + c];
                else
                    diag += A[r * 3 + c] * A[r * 3 + c];
        // |off| <= 1e-13 |diag|: the sweep after that would bring it to ~1e-26 (quadratic convergence),
        // far below what the planarity / distance tests or the 1e-9 parity bound can see; a tighter
        // bound sits at the rounding level and makes a few lanes spin for extra sweeps
        if (off <= 1e-300 || off <= 1e-26 * diag) break;
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int q = p + 1; q < 3; q++)
            {
                const double apq = A[p * 3 + q];
                if (apq == 0.0) continue;
                // Jacobi rotation angle: t = sgn(tau) / (|tau| + sqrt(1 + tau^2)), tau = (aqq-app)/(2 apq),
                // evaluated without forming tau (one sqrt, one division, one rsqrt instead of two
                // sqrt and three divisions — fp64 division and sqrt are ~50-instruction sequences)
                const double d  = A[q * 3 + q] - A[p * 3 + p], a2 = 2.0 * apq;
                const double sg = (d == 0.0) ? 1.0 : ((d > 0.0) == (apq > 0.0) ? 1.0 : -1.0);
                const double t  = sg * fabs(a2) / (fabs(d) + sqrt(d * d + a2 * a2));
                const double c = rsqrt(1.0 + t * t), s = t * c;
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    const double akp = A[k * 3 + p], akq = A[k * 3 + q];
                    A[k * 3 + p] = c * akp - s * akq;
                    A[k * 3 + q] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
                    A[p * 3 + k] = c * apk - s * aqk;
                    A[q * 3 + k] = s * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
                    V[k * 3 + p] = c * vkp - s * vkq;
                    V[k * 3 + q] = s * vkp + c * vkq;
                }
            }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) vals[i] = A[i * 3 + i];
#pragma unroll
    for (int i = 0; i < 2; i++)
    {
        int m = i;
#pragma unroll
        for (int j = i + 1; j < 3; j++)
            if (vals[j] < vals[m]) m = j;
        if (m != i)
        {
            const double tv = vals[i];
            vals[i]         = vals[m];
            vals[m]         = tv;
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
                const double t = V[k * 3 + i];
                V[k * 3 + i]   = V[k * 3 + m];
                V[k * 3 + m]   = t;
            }
        }
    }
}

// neighbours px/py/pz[0..cnt) in ascending (d2, index) order; query (qx,qy,qz) already transformed.
template <int KT>
__device__ __forceinline__ bool fit_plane(const float (&px)[KT], const float (&py)[KT],
                                          const float (&pz)[KT], int cnt, float qx, float qy, float qz,
                                          double planeEigenThreshold, float distThr,
                                          PlaneCandidate& out)
{
    float mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll
    for (int k = 0; k < KT; k++)
        if (k < cnt) mx += px[k], my += py[k], mz += pz[k];
    const float inv_n = 1.0f / (float)cnt;
    mx *= inv_n, my *= inv_n, mz *= inv_n;
    double a00 = 0, a10 = 0, a20 = 0, a11 = 0, a21 = 0, a22 = 0;
#pragma unroll
    for (int k = 0; k < KT; k++)
        if (k < cnt)
        {
            const float ax = px[k] - mx, ay = py[k] - my, az = pz[k] - mz;
            a00 += (double)(ax * ax), a10 += (double)(ax * ay), a20 += (double)(ax * az);
            a11 += (double)(ay * ay), a21 += (double)(ay * az), a22 += (double)(az * az);
        }
    const double dn = (double)inv_n;
    a00 *= dn, a10 *= dn, a20 *= dn, a11 *= dn, a21 *= dn, a22 *= dn;
    const double A[9] = {a00, a10, a20, a10, a11, a21, a20, a21, a22};
    double       V[9], vals[3];
    eig_sym3(A, V, vals);
    if (!(vals[0] < planeEigenThreshold * vals[2] && vals[0] < planeEigenThreshold * vals[1])) return false;
    const double cx = mx, cy = my, cz = mz;
    double       nx = V[0], ny = V[3], nz = V[6];
    const double inv_nn = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
    nx *= inv_nn, ny *= inv_nn, nz *= inv_nn;
    const double D    = -nx * cx - ny * cy - nz * cz;
    const double ev   = nx * (double)qx + ny * (double)qy + nz * (double)qz + D;
    const double dist = fabs(ev) / sqrt(nx * nx + ny * ny + nz * nz);
    if ((float)dist > distThr) return false;  // Matcher_Point2Plane.cpp:101
    out.coefs[0] = nx, out.coefs[1] = ny, out.coefs[2] = nz, out.coefs[3] = D;
    out.centroid[0] = cx, out.centroid[1] = cy, out.centroid[2] = cz;
    return true;
}

// Plane test of Matcher_Adaptive (mp2p_icp/src/Matcher_Adaptive.cpp:222-268): the same fit over the cnt
// kept neighbours, planarity test :240-241, TPlane(centroid, eigenvector 0 AS IS), and the distance of
// `(ex,ey,ez)` — the caller passes the LOCAL-frame point, as upstream does (:247) — strictly below
// planeMinimumDistance (:249).
template <int KT>
__device__ __forceinline__ bool fit_plane_adaptive(const float (&px)[KT], const float (&py)[KT], const float (&pz)[KT], int cnt,
                                                   float ex, float ey, float ez, double planeEigenThreshold,
                                                   double planeMinimumDistance, PlaneCandidate& out)
{
    float mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll
    for (int k = 0; k < KT; k++)
        if (k < cnt) mx += px[k], my += py[k], mz += pz[k];
    const float inv_n = 1.0f / (float)cnt;
    mx *= inv_n, my *= inv_n, mz *= inv_n;
    double a00 = 0, a10 = 0, a20 = 0, a11 = 0, a21 = 0, a22 = 0;
#pragma unroll
    for (int k = 0; k < KT; k++)
        if (k < cnt)
        {
            const float ax = px[k] - mx, ay = py[k] - my, az = pz[k] - mz;
            a00 += (double)(ax * ax), a10 += (double)(ax * ay), a20 += (double)(ax * az);
            a11 += (double)(ay * ay), a21 += (double)(ay * az), a22 += (double)(az * az);
        }
    const double dn = (double)inv_n;
    a00 *= dn, a10 *= dn, a20 *= dn, a11 *= dn, a21 *= dn, a22 *= dn;
    const double A[9] = {a00, a10, a20, a10, a11, a21, a20, a21, a22};
    double       V[9], vals[3];
    eig_sym3(A, V, vals);
    if (!(vals[0] < planeEigenThreshold * vals[2] && vals[0] < planeEigenThreshold * vals[1])) return false;
    const double cx = mx, cy = my, cz = mz;
    const double nx = V[0], ny = V[3], nz = V[6];
    const double D    = -(nx * cx + ny * cy + nz * cz);
    const double ev   = nx * (double)ex + ny * (double)ey + nz * (double)ez + D;
    const double dist = fabs(fabs(ev) / sqrt(nx * nx + ny * ny + nz * nz));
    if (!(dist < planeMinimumDistance)) return false;
    out.coefs[0] = nx, out.coefs[1] = ny, out.coefs[2] = nz, out.coefs[3] = D;
    out.centroid[0] = cx, out.centroid[1] = cy, out.centroid[2] = cz;
    return true;
}

// Line fit of the pt2ln matcher (mp2p_icp/src/Matcher_Point2Line.cpp:132-156): the same moments over
// the cnt neighbours, line test e0 <= thr*e2 && e1 <= thr*e2 (:148-149), director = eigenvector of
// the largest eigenvalue, unitarized; out.coefs[0..2] = director, out.centroid = pBase (the mean).
template <int KT>
__device__ __forceinline__ bool fit_line(const float (&px)[KT], const float (&py)[KT], const float (&pz)[KT], int cnt,
                                         double lineEigenThreshold, PlaneCandidate& out)
{
    float mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll
    for (int k = 0; k < KT; k++)
        if (k < cnt) mx += px[k], my += py[k], mz += pz[k];
    const float inv_n = 1.0f / (float)cnt;
    mx *= inv_n, my *= inv_n, mz *= inv_n;
    double a00 = 0, a10 = 0, a20 = 0, a11 = 0, a21 = 0, a22 = 0;
#pragma unroll
    for (int k = 0; k < KT; k++)
        if (k < cnt)
        {
            const float ax = px[k] - mx, ay = py[k] - my, az = pz[k] - mz;
            a00 += (double)(ax * ax), a10 += (double)(ax * ay), a20 += (double)(ax * az);
            a11 += (double)(ay * ay), a21 += (double)(ay * az), a22 += (double)(az * az);
        }
    const double dn = (double)inv_n;
    a00 *= dn, a10 *= dn, a20 *= dn, a11 *= dn, a21 *= dn, a22 *= dn;
    const double A[9] = {a00, a10, a20, a10, a11, a21, a20, a21, a22};
    double       V[9], vals[3];
    eig_sym3(A, V, vals);
    if (vals[0] > lineEigenThreshold * vals[2]) return false;
    if (vals[1] > lineEigenThreshold * vals[2]) return false;
    const double ux = V[2], uy = V[5], uz = V[8];
    const double inv = 1.0 / sqrt(ux * ux + uy * uy + uz * uz);
    out.coefs[0] = ux * inv, out.coefs[1] = uy * inv, out.coefs[2] = uz * inv, out.coefs[3] = 0.0;
    out.centroid[0] = mx, out.centroid[1] = my, out.centroid[2] = mz;
    return true;
}
}  // namespace mp2p
