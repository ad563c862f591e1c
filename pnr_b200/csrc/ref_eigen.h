// ref_eigen.h -- the reference's symmetric 3x3 eigen-solver (frangi.cpp:1230-1495: EISPACK tred2 / tql2 as in JAMA, then
// the ordering by magnitude with the reference's tie rules), written once for two users:
//   * the host members Frangi::tred2 / tql2 / eigen_decomposition(_static) of the drop-in class (frangi_shim_host.cpp,
//     T = double, compiled with -ffp-contract=off), bit-identical to the compiled reference (tests/test_cpp_shim.py);
//   * the device pass of FRANGI_GPU_FLAG_REFERENCE_DIRECTION (frangi_kernels.cuh), T = rdouble: a double whose
//     + - * / and sqrt are the round-to-nearest intrinsics, which the compiler never contracts into fused multiply-adds,
//     so the device executes the operation sequence the reference's x86-64 build executes and returns the same bits --
//     in particular the same SIGN of every eigenvector, which is otherwise arbitrary.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define RE_HD __host__ __device__ __forceinline__
#define RE_FN __host__ __device__
#else
#define RE_HD inline
#define RE_FN inline
#endif

#if defined(__CUDACC__)
struct rdouble {
    double v;
    RE_HD rdouble() {}
    RE_HD rdouble(double x) : v(x) {}
};
#if defined(__CUDA_ARCH__)
RE_HD rdouble operator+(rdouble a, rdouble b) { return rdouble(__dadd_rn(a.v, b.v)); }
RE_HD rdouble operator-(rdouble a, rdouble b) { return rdouble(__dsub_rn(a.v, b.v)); }
RE_HD rdouble operator*(rdouble a, rdouble b) { return rdouble(__dmul_rn(a.v, b.v)); }
RE_HD rdouble operator/(rdouble a, rdouble b) { return rdouble(__ddiv_rn(a.v, b.v)); }
RE_HD rdouble re_sqrt(rdouble a) { return rdouble(__dsqrt_rn(a.v)); }
#else
RE_HD rdouble operator+(rdouble a, rdouble b) { return rdouble(a.v + b.v); }
RE_HD rdouble operator-(rdouble a, rdouble b) { return rdouble(a.v - b.v); }
RE_HD rdouble operator*(rdouble a, rdouble b) { return rdouble(a.v * b.v); }
RE_HD rdouble operator/(rdouble a, rdouble b) { return rdouble(a.v / b.v); }
RE_HD rdouble re_sqrt(rdouble a) { return rdouble(sqrt(a.v)); }
#endif
RE_HD rdouble operator-(rdouble a) { return rdouble(-a.v); }
RE_HD rdouble& operator+=(rdouble& a, rdouble b) { a = a + b; return a; }
RE_HD rdouble& operator-=(rdouble& a, rdouble b) { a = a - b; return a; }
RE_HD rdouble& operator/=(rdouble& a, rdouble b) { a = a / b; return a; }
RE_HD bool operator<(rdouble a, rdouble b) { return a.v < b.v; }
RE_HD bool operator>(rdouble a, rdouble b) { return a.v > b.v; }
RE_HD bool operator<=(rdouble a, rdouble b) { return a.v <= b.v; }
RE_HD bool operator>=(rdouble a, rdouble b) { return a.v >= b.v; }
RE_HD bool operator==(rdouble a, rdouble b) { return a.v == b.v; }
RE_HD bool operator!=(rdouble a, rdouble b) { return a.v != b.v; }
RE_HD rdouble re_abs(rdouble a) { return rdouble(fabs(a.v)); }
#endif
RE_HD double re_sqrt(double a) { return sqrt(a); }
RE_HD double re_abs(double a) { return fabs(a); }
template <class T> RE_HD T re_hypot(T x, T y) { return re_sqrt(x * x + y * y); }     // frangi.cpp hypot2

// Householder reduction of the symmetric matrix held in V to tridiagonal form (EISPACK tred2): on return d holds
// the diagonal, e[1..2] the sub-diagonal and V the accumulated orthogonal transformation.
template <class T>
RE_FN void ref_tred2(T V[3][3], T d[3], T e[3])
{
    const int N = 3;
    for (int c = 0; c < N; ++c) d[c] = V[N - 1][c];
    for (int i = N - 1; i >= 1; --i) {
        T norm1 = 0.0, hsum = 0.0;
        for (int k = 0; k < i; ++k) norm1 = norm1 + re_abs(d[k]);
        if (norm1 == 0.0) {
            e[i] = d[i - 1];
            for (int c = 0; c < i; ++c) {
                d[c] = V[i - 1][c];
                V[i][c] = 0.0;
                V[c][i] = 0.0;
            }
        } else {
            for (int k = 0; k < i; ++k) {
                d[k] /= norm1;
                hsum += d[k] * d[k];
            }
            const T last = d[i - 1];
            T root = re_sqrt(hsum);
            if (last > 0) root = -root;
            e[i] = norm1 * root;
            hsum = hsum - last * root;
            d[i - 1] = last - root;
            for (int c = 0; c < i; ++c) e[c] = 0.0;
            for (int c = 0; c < i; ++c) {            // similarity transform of the leading block
                const T dc = d[c];
                V[c][i] = dc;
                T acc = e[c] + V[c][c] * dc;
                for (int k = c + 1; k <= i - 1; ++k) {
                    acc += V[k][c] * d[k];
                    e[k] += V[k][c] * dc;
                }
                e[c] = acc;
            }
            T dot = 0.0;
            for (int c = 0; c < i; ++c) {
                e[c] /= hsum;
                dot += e[c] * d[c];
            }
            const T half = dot / (hsum + hsum);
            for (int c = 0; c < i; ++c) e[c] -= half * d[c];
            for (int c = 0; c < i; ++c) {
                const T dc = d[c], ec = e[c];
                for (int k = c; k <= i - 1; ++k) V[k][c] -= (dc * e[k] + ec * d[k]);
                d[c] = V[i - 1][c];
                V[i][c] = 0.0;
            }
        }
        d[i] = hsum;
    }
    for (int i = 0; i < N - 1; ++i) {                // accumulate the transformations
        V[N - 1][i] = V[i][i];
        V[i][i] = 1.0;
        const T hh = d[i + 1];
        if (hh != 0.0) {
            for (int k = 0; k <= i; ++k) d[k] = V[k][i + 1] / hh;
            for (int c = 0; c <= i; ++c) {
                T acc = 0.0;
                for (int k = 0; k <= i; ++k) acc += V[k][i + 1] * V[k][c];
                for (int k = 0; k <= i; ++k) V[k][c] -= acc * d[k];
            }
        }
        for (int k = 0; k <= i; ++k) V[k][i + 1] = 0.0;
    }
    for (int c = 0; c < N; ++c) {
        d[c] = V[N - 1][c];
        V[N - 1][c] = 0.0;
    }
    V[N - 1][N - 1] = 1.0;
    e[0] = 0.0;
}

// Implicit-shift QL on the tridiagonal matrix (EISPACK tql2), eigenvectors accumulated in V, then eigenvalues and
// vectors sorted ascending by value (first minimum wins).
template <class T>
RE_FN void ref_tql2(T V[3][3], T d[3], T e[3])
{
    const int N = 3;
    for (int i = 1; i < N; ++i) e[i - 1] = e[i];
    e[N - 1] = 0.0;
    T shift_total = 0.0, scale_ref = 0.0;
    const T eps = T(0x1p-52);
    for (int lo = 0; lo < N; ++lo) {
        const T cand = re_abs(d[lo]) + re_abs(e[lo]);
        scale_ref = scale_ref > cand ? scale_ref : cand;
        int m = lo;
        while (m < N) {
            if (re_abs(e[m]) <= eps * scale_ref) break;
            ++m;
        }
        if (m > lo) {
            do {
                T g = d[lo];
                T p = (d[lo + 1] - g) / (2.0 * e[lo]);
                T r = re_hypot(p, T(1.0));
                if (p < 0) r = -r;
                d[lo] = e[lo] / (p + r);
                d[lo + 1] = e[lo] * (p + r);
                const T dl1 = d[lo + 1];
                T h = g - d[lo];
                for (int i = lo + 2; i < N; ++i) d[i] -= h;
                shift_total = shift_total + h;
                p = d[m];
                T c = 1.0, c2 = c, c3 = c;
                const T el1 = e[lo + 1];
                T s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= lo; --i) {
                    c3 = c2;
                    c2 = c;
                    s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = re_hypot(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < N; ++k) {
                        h = V[k][i + 1];
                        V[k][i + 1] = s * V[k][i] + c * h;
                        V[k][i] = c * V[k][i] - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[lo] / dl1;
                e[lo] = s * p;
                d[lo] = c * p;
            } while (re_abs(e[lo]) > eps * scale_ref);
        }
        d[lo] = d[lo] + shift_total;
        e[lo] = 0.0;
    }
    for (int i = 0; i < N - 1; ++i) {
        int best = i;
        T pv = d[i];
        for (int j = i + 1; j < N; ++j)
            if (d[j] < pv) { best = j; pv = d[j]; }
        if (best != i) {
            d[best] = d[i];
            d[i] = pv;
            for (int r = 0; r < N; ++r) {
                const T t = V[r][i];
                V[r][i] = V[r][best];
                V[r][best] = t;
            }
        }
    }
}

template <class T>
RE_HD void ref_swap_pair(T V[3][3], T d[3], T mag[3], int a, int b)
{
    T t = d[a]; d[a] = d[b]; d[b] = t;
    t = mag[a]; mag[a] = mag[b]; mag[b] = t;
    for (int r = 0; r < 3; ++r) { t = V[r][a]; V[r][a] = V[r][b]; V[r][b] = t; }
}

// A symmetric -> columns of V = unit eigenvectors, d ordered |d0| <= |d1| <= |d2| with the reference's tie rules
// (frangi.cpp:1284-1304): the largest magnitude goes last (`>=` against the other candidate, `>` against the last
// slot), then the first two are swapped on a strict `>`.  The magnitude is the reference's absd (frangi.h:58).
template <class T>
RE_FN void ref_eigen_decomposition(const T A[3][3], T V[3][3], T d[3])
{
    T e[3], mag[3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) V[r][c] = A[r][c];
    ref_tred2(V, d, e);
    ref_tql2(V, d, e);
    for (int k = 0; k < 3; ++k) mag[k] = d[k] > T(0.0) ? d[k] : -d[k];
    if (mag[0] >= mag[1] && mag[0] > mag[2]) ref_swap_pair(V, d, mag, 0, 2);
    else if (mag[1] >= mag[0] && mag[1] > mag[2]) ref_swap_pair(V, d, mag, 1, 2);
    if (mag[0] > mag[1]) ref_swap_pair(V, d, mag, 0, 1);
}
