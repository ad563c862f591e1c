/*
 * oracle/frangi_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement, in plain C, of the reference's multi-scale 3-D Frangi
 * tubularity filter (miroslavradojevic/pnr, pnr-vaa3d/frangi.cpp).  It exists
 * only so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * can check the CUDA path in pnr_b200/ against the reference's arithmetic.
 * Nothing under pnr_b200/ may import, link or call it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors
 * (SURVEY.md section 4), so the pin is (a) the unmodified reference compiled in
 * place into oracle/_ref/ (oracle/Makefile) and compared bit-for-bit with this
 * file in tests/test_oracle_golden.py (test_port_equals_reference_where_built), (b) golden fixtures generated from
 * oracle/_ref and committed under tests/golden/, and (c) the known-answer
 * vectors recorded from the reference in SURVEY.md section 8c.
 *
 * Every function cites the reference lines it follows.  The arithmetic order
 * (float32 accumulation order of the smoothing, double-precision eigen solver,
 * tie rules) is kept because the results are compared bit-for-bit; the code
 * structure is our own.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* Gaussian taps.  frangi.cpp:651-680: L = ceil(3*sigma) (float product),    */
/* w[k] = expf(-(k*k)/(2*sigma*sigma)) with the integer numerator negated    */
/* before the float division, sequential float32 sum, then per-tap division. */
/* ------------------------------------------------------------------------ */
ORACLE_API int oracle_gauss_radius(float sigma)
{
    float t = 3.0f * sigma;
    return (int)ceilf(t);
}

ORACLE_API void oracle_gauss_taps(float sigma, int radius, float *taps /* 2*radius+1 */)
{
    float total = 0.0f;
    float denom = 2.0f * sigma * sigma;
    for (int k = -radius; k <= radius; ++k) {
        int neg_sq = -(k * k);
        float v = expf((float)neg_sq / denom);
        taps[k + radius] = v;
        total += v;
    }
    for (int k = 0; k < 2 * radius + 1; ++k) taps[k] /= total;
}

static inline int64_t clamp64(int64_t v, int64_t lo, int64_t hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}

/* ------------------------------------------------------------------------ */
/* Separable truncated Gaussian, x -> y -> z, float32, replicate borders.    */
/* frangi.cpp:647-784.  The reference splits each axis into "clamped head /  */
/* unclamped body / clamped tail" ranges; together they visit every sample   */
/* once with acc = 0; acc += in[clamp(p+k)] * w[k] for ascending k, which is */
/* what is restated here (clamping is a no-op in the body range).            */
/* z uses sigma/zdist and its own radius (frangi.cpp:649,669-680).           */
/* 64-bit indexing (the reference's int indexing overflows at 2^31 voxels).  */
/* ------------------------------------------------------------------------ */
ORACLE_API int oracle_imgaussian(const uint8_t *I, int w, int h, int l,
                                 float sigma, float zdist, float *F)
{
    const int64_t W = w, H = h, plane = W * H, total = plane * (int64_t)l;
    float sigma_z = sigma / zdist;
    int Lxy = oracle_gauss_radius(sigma);
    int Lz = oracle_gauss_radius(sigma_z);
    float *gxy = (float *)malloc(sizeof(float) * (size_t)(2 * Lxy + 1));
    float *gz = (float *)malloc(sizeof(float) * (size_t)(2 * Lz + 1));
    float *K = (float *)malloc(sizeof(float) * (size_t)total);
    if (!gxy || !gz || !K) { free(gxy); free(gz); free(K); return 1; }
    oracle_gauss_taps(sigma, Lxy, gxy);
    oracle_gauss_taps(sigma_z, Lz, gz);

    /* x pass: u8 -> F (frangi.cpp:683-714) */
    for (int64_t z = 0; z < l; ++z)
        for (int64_t y = 0; y < h; ++y) {
            const uint8_t *row = I + z * plane + y * W;
            float *out = F + z * plane + y * W;
            for (int64_t x = 0; x < w; ++x) {
                float acc = 0.0f;
                for (int k = -Lxy; k <= Lxy; ++k) {
                    int64_t xs = clamp64(x + k, 0, W - 1);
                    acc += (float)(int)row[xs] * gxy[k + Lxy];
                }
                out[x] = acc;
            }
        }
    /* y pass: F -> K (frangi.cpp:717-748) */
    for (int64_t z = 0; z < l; ++z)
        for (int64_t y = 0; y < h; ++y) {
            float *out = K + z * plane + y * W;
            for (int64_t x = 0; x < w; ++x) out[x] = 0.0f;
            for (int k = -Lxy; k <= Lxy; ++k) {
                int64_t ys = clamp64(y + k, 0, H - 1);
                const float *in = F + z * plane + ys * W;
                float g = gxy[k + Lxy];
                for (int64_t x = 0; x < w; ++x) out[x] += in[x] * g;
            }
        }
    /* z pass: K -> F (frangi.cpp:751-782) */
    for (int64_t z = 0; z < l; ++z) {
        float *out = F + z * plane;
        for (int64_t i = 0; i < plane; ++i) out[i] = 0.0f;
        for (int k = -Lz; k <= Lz; ++k) {
            int64_t zs = clamp64(z + k, 0, (int64_t)l - 1);
            const float *in = K + zs * plane;
            float g = gz[k + Lz];
            for (int64_t i = 0; i < plane; ++i) out[i] += in[i] * g;
        }
    }
    free(gxy); free(gz); free(K);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* First difference along one axis with the reference's face rule            */
/* (frangi.cpp:308-310, 327-329, 354-356): interior 0.5*(f[+1]-f[-1]), low   */
/* face f[+1]-f[0], high face f[0]-f[-1].  The float subtraction is rounded  */
/* first; the 0.5 (a double literal in the reference) is then exact.         */
/* axis stride s, coordinate c in [0,n).                                     */
/* ------------------------------------------------------------------------ */
static inline float face_diff(const float *f, int64_t i, int64_t s, int c, int n)
{
    if (c == 0) return f[i + s] - f[i];
    if (c < n - 1) return (float)(0.5 * (double)(f[i + s] - f[i - s]));
    return f[i] - f[i - s];
}

/* hessian3d, frangi.cpp:291-390: F = imgaussian; first difference along an  */
/* axis; the SAME face rule applied to that first-difference volume along    */
/* the second axis; product with sigma*sigma (float) (frangi.cpp:319,339,    */
/* 345,368,374,380).  Dxy = d/dy(d/dx F), Dxz = d/dz(d/dx F), Dyz = d/dz(d/dy F). */
ORACLE_API int oracle_hessian3d(const uint8_t *I, int w, int h, int l, float sigma, float zdist,
                                float *Dzz, float *Dyy, float *Dyz,
                                float *Dxx, float *Dxy, float *Dxz)
{
    if (w < 2 || h < 2 || l < 2) return 2;
    const int64_t W = w, plane = W * (int64_t)h, total = plane * (int64_t)l;
    float *F = (float *)malloc(sizeof(float) * (size_t)total);
    float *DD = (float *)malloc(sizeof(float) * (size_t)total);
    if (!F || !DD) { free(F); free(DD); return 1; }
    int rc = oracle_imgaussian(I, w, h, l, sigma, zdist, F);
    if (rc) { free(F); free(DD); return rc; }
    const float s2 = sigma * sigma;

#define FOR_VOXELS                                                        \
    for (int z = 0; z < l; ++z) for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x)
#define IDX ((int64_t)z * plane + (int64_t)y * W + x)

    /* d/dz, then Dzz (frangi.cpp:306-320) */
    FOR_VOXELS { int64_t i = IDX; DD[i] = face_diff(F, i, plane, z, l); }
    FOR_VOXELS { int64_t i = IDX; Dzz[i] = face_diff(DD, i, plane, z, l) * s2; }
    /* d/dy, then Dyy, Dyz (frangi.cpp:325-346) */
    FOR_VOXELS { int64_t i = IDX; DD[i] = face_diff(F, i, W, y, h); }
    FOR_VOXELS {
        int64_t i = IDX;
        Dyy[i] = face_diff(DD, i, W, y, h) * s2;
        Dyz[i] = face_diff(DD, i, plane, z, l) * s2;
    }
    /* d/dx, then Dxx, Dxy, Dxz (frangi.cpp:352-381) */
    FOR_VOXELS { int64_t i = IDX; DD[i] = face_diff(F, i, 1, x, w); }
    FOR_VOXELS {
        int64_t i = IDX;
        Dxx[i] = face_diff(DD, i, 1, x, w) * s2;
        Dxy[i] = face_diff(DD, i, W, y, h) * s2;
        Dxz[i] = face_diff(DD, i, plane, z, l) * s2;
    }
#undef FOR_VOXELS
#undef IDX
    free(F); free(DD);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Symmetric 3x3 eigen-decomposition in double: Householder reduction to     */
/* tridiagonal form followed by implicit-shift QL (the EISPACK tred2/tql2    */
/* pair as published in JAMA), frangi.cpp:1309-1387 and 1390-1493, then the  */
/* reference's re-ordering by absolute eigenvalue, frangi.cpp:1284-1304.     */
/* Operation order is preserved so that results are bit-identical.           */
/* ------------------------------------------------------------------------ */
#define N3 3

static void householder_tridiag(double V[N3][N3], double d[N3], double e[N3])
{
    /* frangi.cpp:1309-1387 */
    for (int c = 0; c < N3; ++c) d[c] = V[N3 - 1][c];

    for (int i = N3 - 1; i >= 1; --i) {
        double norm1 = 0.0, hsum = 0.0;
        for (int k = 0; k < i; ++k) norm1 = norm1 + fabs(d[k]);
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
            double last = d[i - 1];
            double root = sqrt(hsum);
            if (last > 0) root = -root;
            e[i] = norm1 * root;
            hsum = hsum - last * root;
            d[i - 1] = last - root;
            for (int c = 0; c < i; ++c) e[c] = 0.0;

            /* similarity transform of the leading block */
            for (int c = 0; c < i; ++c) {
                double dc = d[c];
                V[c][i] = dc;
                double acc = e[c] + V[c][c] * dc;
                for (int k = c + 1; k <= i - 1; ++k) {
                    acc += V[k][c] * d[k];
                    e[k] += V[k][c] * dc;
                }
                e[c] = acc;
            }
            double dot = 0.0;
            for (int c = 0; c < i; ++c) {
                e[c] /= hsum;
                dot += e[c] * d[c];
            }
            double half = dot / (hsum + hsum);
            for (int c = 0; c < i; ++c) e[c] -= half * d[c];
            for (int c = 0; c < i; ++c) {
                double dc = d[c], ec = e[c];
                for (int k = c; k <= i - 1; ++k) V[k][c] -= (dc * e[k] + ec * d[k]);
                d[c] = V[i - 1][c];
                V[i][c] = 0.0;
            }
        }
        d[i] = hsum;
    }

    /* accumulate the transformations */
    for (int i = 0; i < N3 - 1; ++i) {
        V[N3 - 1][i] = V[i][i];
        V[i][i] = 1.0;
        double hh = d[i + 1];
        if (hh != 0.0) {
            for (int k = 0; k <= i; ++k) d[k] = V[k][i + 1] / hh;
            for (int c = 0; c <= i; ++c) {
                double acc = 0.0;
                for (int k = 0; k <= i; ++k) acc += V[k][i + 1] * V[k][c];
                for (int k = 0; k <= i; ++k) V[k][c] -= acc * d[k];
            }
        }
        for (int k = 0; k <= i; ++k) V[k][i + 1] = 0.0;
    }
    for (int c = 0; c < N3; ++c) {
        d[c] = V[N3 - 1][c];
        V[N3 - 1][c] = 0.0;
    }
    V[N3 - 1][N3 - 1] = 1.0;
    e[0] = 0.0;
}

static inline double pythag(double a, double b) { return sqrt(a * a + b * b); } /* frangi.cpp:1495 */

static void ql_implicit(double V[N3][N3], double d[N3], double e[N3])
{
    /* frangi.cpp:1390-1493 */
    for (int i = 1; i < N3; ++i) e[i - 1] = e[i];
    e[N3 - 1] = 0.0;

    double shift_total = 0.0, scale_ref = 0.0;
    const double eps = pow(2.0, -52.0);
    for (int lo = 0; lo < N3; ++lo) {
        double cand = fabs(d[lo]) + fabs(e[lo]);
        scale_ref = scale_ref > cand ? scale_ref : cand; /* MAX(a,b) = a>b?a:b, frangi.cpp:16 */
        int m = lo;
        while (m < N3) {
            if (fabs(e[m]) <= eps * scale_ref) break;
            ++m;
        }
        if (m > lo) {
            do {
                double g = d[lo];
                double p = (d[lo + 1] - g) / (2.0 * e[lo]);
                double r = pythag(p, 1.0);
                if (p < 0) r = -r;
                d[lo] = e[lo] / (p + r);
                d[lo + 1] = e[lo] * (p + r);
                double dl1 = d[lo + 1];
                double h = g - d[lo];
                for (int i = lo + 2; i < N3; ++i) d[i] -= h;
                shift_total = shift_total + h;

                p = d[m];
                double c = 1.0, c2 = c, c3 = c;
                double el1 = e[lo + 1];
                double s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= lo; --i) {
                    c3 = c2;
                    c2 = c;
                    s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = pythag(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < N3; ++k) {
                        h = V[k][i + 1];
                        V[k][i + 1] = s * V[k][i] + c * h;
                        V[k][i] = c * V[k][i] - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[lo] / dl1;
                e[lo] = s * p;
                d[lo] = c * p;
            } while (fabs(e[lo]) > eps * scale_ref);
        }
        d[lo] = d[lo] + shift_total;
        e[lo] = 0.0;
    }

    /* ascending selection sort of eigenvalues with their columns */
    for (int i = 0; i < N3 - 1; ++i) {
        int best = i;
        double pv = d[i];
        for (int j = i + 1; j < N3; ++j)
            if (d[j] < pv) { best = j; pv = d[j]; }
        if (best != i) {
            d[best] = d[i];
            d[i] = pv;
            for (int r = 0; r < N3; ++r) {
                double t = V[r][i];
                V[r][i] = V[r][best];
                V[r][best] = t;
            }
        }
    }
}

static inline void swap_pair(double V[N3][N3], double d[N3], double da[N3], int a, int b)
{
    double t = d[a]; d[a] = d[b]; d[b] = t;
    t = da[a]; da[a] = da[b]; da[b] = t;
    for (int r = 0; r < N3; ++r) { t = V[r][a]; V[r][a] = V[r][b]; V[r][b] = t; }
}

/* A symmetric 3x3 (row-major 9 doubles) -> V (columns = eigenvectors,       */
/* row-major 9 doubles), d sorted so that |d0| <= |d1| <= |d2| with the      */
/* reference's tie rules (frangi.cpp:1284-1304):                             */
/*   if (|d0| >= |d1| && |d0| > |d2|) swap(0,2)                              */
/*   else if (|d1| >= |d0| && |d1| > |d2|) swap(1,2)                         */
/*   then if (|d0| > |d1|) swap(0,1)                                         */
ORACLE_API void oracle_eigen3(const double *A, double *Vout, double *dout)
{
    double V[N3][N3], d[N3], e[N3], da[N3];
    for (int r = 0; r < N3; ++r)
        for (int c = 0; c < N3; ++c) V[r][c] = A[r * N3 + c];
    householder_tridiag(V, d, e);
    ql_implicit(V, d, e);
    for (int k = 0; k < N3; ++k) da[k] = d[k] > 0 ? d[k] : -d[k]; /* absd, frangi.h:58 */
    if (da[0] >= da[1] && da[0] > da[2]) swap_pair(V, d, da, 0, 2);
    else if (da[1] >= da[0] && da[1] > da[2]) swap_pair(V, d, da, 1, 2);
    if (da[0] > da[1]) swap_pair(V, d, da, 0, 1);
    for (int r = 0; r < N3; ++r)
        for (int c = 0; c < N3; ++c) Vout[r * N3 + c] = V[r][c];
    for (int k = 0; k < N3; ++k) dout[k] = d[k];
}

/* round-half-away-from-zero then clamp to a byte, frangi.cpp:240-250 */
static inline uint8_t dir_code(double component)
{
    int v = (int)round(((component + 1) / 2) * 255);
    if (v < 0) v = 0; else if (v > 255) v = 255;
    return (uint8_t)v;
}

/* Per-voxel Frangi measure from one scale's Hessian, frangi.cpp:194-231.    */
/* Returns the double vesselness; writes the unit eigenvector of the         */
/* smallest-|lambda| eigenvalue (column 0) to dir[3].                        */
static inline double voxel_vesselness(float dxx, float dxy, float dxz, float dyy, float dyz,
                                      float dzz, float alpha, float beta, float C,
                                      int blackwhite, double dir[3])
{
    double A[9] = { dxx, dxy, dxz, dxy, dyy, dyz, dxz, dyz, dzz };
    double V[9], lam[3];
    oracle_eigen3(A, V, lam);
    double a1 = fabs(lam[0]), a2 = fabs(lam[1]), a3 = fabs(lam[2]);
    double Ra = a2 / a3;
    double Rb = a1 / sqrt(a2 * a3);
    double S = sqrt(a1 * a1 + a2 * a2 + a3 * a3);
    /* 2*alpha*alpha etc. are float products in the reference (frangi.cpp:215-217) */
    float two_a2 = 2 * alpha * alpha, two_b2 = 2 * beta * beta, two_c2 = 2 * C * C;
    double tRa = 1 - exp(-((Ra * Ra) / two_a2));
    double tRb = exp(-((Rb * Rb) / two_b2));
    double tS = 1 - exp(-(S * S) / two_c2);
    double v = tRa * tRb * tS;
    if (blackwhite) {
        if (lam[1] < 0) v = 0;
        if (lam[2] < 0) v = 0;
    } else {
        if (lam[1] > 0) v = 0;
        if (lam[2] > 0) v = 0;
    }
    if (isnan(v)) v = 0;
    dir[0] = V[0]; dir[1] = V[3]; dir[2] = V[6];
    return v;
}

/* ------------------------------------------------------------------------ */
/* frangi3d, frangi.cpp:152-289: scale loop, first scale stores, later       */
/* scales overwrite only on strictly greater response (double v compared     */
/* with the stored float, frangi.cpp:254), Jmin/Jmax updated only where J    */
/* is assigned (frangi.cpp:237-238, 257-258).                                */
/* scale_idx (nullable) and dir_xyz (nullable, 3 planar float32 volumes:     */
/* x then y then z) are extra outputs of the new C-ABI; they record which    */
/* scale made the last assignment and its unquantised direction.             */
/* ------------------------------------------------------------------------ */
ORACLE_API int oracle_frangi3d(const uint8_t *I, int w, int h, int l,
                               const float *sigmas, int nsig, float zdist,
                               float alpha, float beta, float C, int blackwhite,
                               float *J, float *Jmin, float *Jmax,
                               uint8_t *Vx, uint8_t *Vy, uint8_t *Vz,
                               uint8_t *scale_idx, float *dir_xyz)
{
    if (w < 2 || h < 2 || l < 2 || nsig < 1) return 2;
    const int64_t total = (int64_t)w * h * l;
    float *D[6];
    for (int k = 0; k < 6; ++k) {
        D[k] = (float *)malloc(sizeof(float) * (size_t)total);
        if (!D[k]) { for (int q = 0; q < k; ++q) free(D[q]); return 1; }
    }
    float *Dzz = D[0], *Dyy = D[1], *Dyz = D[2], *Dxx = D[3], *Dxy = D[4], *Dxz = D[5];
    float lo = FLT_MAX, hi = -FLT_MAX;
    int rc = 0;
    for (int si = 0; si < nsig && !rc; ++si) {
        rc = oracle_hessian3d(I, w, h, l, sigmas[si], zdist, Dzz, Dyy, Dyz, Dxx, Dxy, Dxz);
        if (rc) break;
        for (int64_t i = 0; i < total; ++i) {
            double dir[3];
            double v = voxel_vesselness(Dxx[i], Dxy[i], Dxz[i], Dyy[i], Dyz[i], Dzz[i],
                                        alpha, beta, C, blackwhite, dir);
            if (si == 0 || v > J[i]) {
                J[i] = (float)v;
                if (J[i] < lo) lo = J[i];
                if (J[i] > hi) hi = J[i];
                Vx[i] = dir_code(dir[0]);
                Vy[i] = dir_code(dir[1]);
                Vz[i] = dir_code(dir[2]);
                if (scale_idx) scale_idx[i] = (uint8_t)si;
                if (dir_xyz) {
                    dir_xyz[i] = (float)dir[0];
                    dir_xyz[total + i] = (float)dir[1];
                    dir_xyz[2 * total + i] = (float)dir[2];
                }
            }
        }
    }
    for (int k = 0; k < 6; ++k) free(D[k]);
    *Jmin = lo; *Jmax = hi;
    return rc;
}

/* Single-scale per-voxel stage, for stage-level parity tests: vesselness    */
/* (float) and direction from six Hessian volumes.                           */
ORACLE_API void oracle_vesselness_stage(const float *Dxx, const float *Dxy, const float *Dxz,
                                        const float *Dyy, const float *Dyz, const float *Dzz,
                                        int64_t total, float alpha, float beta, float C,
                                        int blackwhite, float *v_out, float *dir_xyz,
                                        double *lambda_out /* nullable, 3 per voxel */)
{
    for (int64_t i = 0; i < total; ++i) {
        double dir[3];
        double v = voxel_vesselness(Dxx[i], Dxy[i], Dxz[i], Dyy[i], Dyz[i], Dzz[i],
                                    alpha, beta, C, blackwhite, dir);
        v_out[i] = (float)v;
        if (dir_xyz) {
            dir_xyz[i] = (float)dir[0];
            dir_xyz[total + i] = (float)dir[1];
            dir_xyz[2 * total + i] = (float)dir[2];
        }
        if (lambda_out) {
            double A[9] = { Dxx[i], Dxy[i], Dxz[i], Dxy[i], Dyy[i], Dyz[i], Dxz[i], Dyz[i], Dzz[i] };
            double V[9];
            oracle_eigen3(A, V, lambda_out + 3 * i);
        }
    }
}

/* ------------------------------------------------------------------------ */
/* J -> J8 min-max normalisation done by the caller right after frangi3d,    */
/* Advantra_plugin.cpp:2499-2512 with round() from :120-123                  */
/* (r > 0 ? floor(r+0.5) : ceil(r-0.5)); all-zero when |Jmax-Jmin|<=FLT_MIN. */
/* The quotient and the product with 255 are float32 there.                  */
/* ------------------------------------------------------------------------ */
ORACLE_API void oracle_j_to_j8(const float *J, int64_t total, float Jmin, float Jmax, uint8_t *J8)
{
    if (fabsf(Jmax - Jmin) <= FLT_MIN) {
        memset(J8, 0, (size_t)total);
        return;
    }
    for (int64_t i = 0; i < total; ++i) {
        float q = ((J[i] - Jmin) / (Jmax - Jmin)) * 255;
        double r = q;
        int v = (int)((r > 0.0) ? floor(r + 0.5) : ceil(r - 0.5));
        if (v < 0) v = 0; else if (v > 255) v = 255;
        J8[i] = (uint8_t)v;
    }
}

/* ------------------------------------------------------------------------ */
/* 2-D path: Frangi::frangi2d (frangi.cpp:392-505) over    */
/* hessian2d (:507-560) over the 2-D imgaussian (:562-645).  The reference's  */
/* mixed float / double arithmetic is kept operation by operation: pow(x, 2)  */
/* on a float is a double square, `.5 * float` is a double product stored to  */
/* float, exp / sqrt / abs on floats are the float overloads.                 */
/* ------------------------------------------------------------------------ */
static void smooth2d(const uint8_t *I, int w, int h, float sigma, float *F)
{
    /* frangi.cpp:562-645: same taps as the 3-D filter (:566-577), x pass u8 -> K, y pass K -> F,
     * float32 accumulation from zero in ascending tap order, replicate clamp */
    const int L = oracle_gauss_radius(sigma);
    float *g = (float *)malloc(sizeof(float) * (size_t)(2 * L + 1));
    float *K = (float *)malloc(sizeof(float) * (size_t)w * h);
    oracle_gauss_taps(sigma, L, g);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float acc = 0.0f;
            for (int k = -L; k <= L; ++k) {
                int xs = x + k < 0 ? 0 : (x + k > w - 1 ? w - 1 : x + k);
                acc += (float)(int)I[(size_t)y * w + xs] * g[k + L];
            }
            K[(size_t)y * w + x] = acc;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float acc = 0.0f;
            for (int k = -L; k <= L; ++k) {
                int ys = y + k < 0 ? 0 : (y + k > h - 1 ? h - 1 : y + k);
                acc += K[(size_t)ys * w + x] * g[k + L];
            }
            F[(size_t)y * w + x] = acc;
        }
    free(g); free(K);
}

/* first difference of V along an axis with the reference's face rules (frangi.cpp:516-520 etc.) */
static float diff2d(const float *V, int w, int h, int x, int y, int along_y)
{
    if (along_y) {
        if (y == 0) return V[(size_t)(y + 1) * w + x] - V[(size_t)y * w + x];
        if (y < h - 1) return (float)(.5 * (V[(size_t)(y + 1) * w + x] - V[(size_t)(y - 1) * w + x]));
        return V[(size_t)y * w + x] - V[(size_t)(y - 1) * w + x];
    }
    if (x == 0) return V[(size_t)y * w + x + 1] - V[(size_t)y * w + x];
    if (x < w - 1) return (float)(.5 * (V[(size_t)y * w + x + 1] - V[(size_t)y * w + x - 1]));
    return V[(size_t)y * w + x] - V[(size_t)y * w + x - 1];
}

/* Dyy, Dxy, Dxx of one scale (frangi.cpp:507-560); F_out (nullable) receives the smoothed image */
ORACLE_API int oracle_hessian2d(const uint8_t *I, int w, int h, float sigma, float *Dyy, float *Dxy, float *Dxx, float *F_out)
{
    const size_t n = (size_t)w * h;
    float *F = (float *)malloc(sizeof(float) * n), *DD = (float *)malloc(sizeof(float) * n);
    if (!F || !DD) { free(F); free(DD); return -1; }
    smooth2d(I, w, h, sigma, F);
    const float s2 = sigma * sigma;
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) DD[(size_t)y * w + x] = diff2d(F, w, h, x, y, 1);
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) Dyy[(size_t)y * w + x] = diff2d(DD, w, h, x, y, 1) * s2;
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) DD[(size_t)y * w + x] = diff2d(F, w, h, x, y, 0);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            Dxx[(size_t)y * w + x] = diff2d(DD, w, h, x, y, 0) * s2;
            Dxy[(size_t)y * w + x] = diff2d(DD, w, h, x, y, 1) * s2;
        }
    if (F_out) memcpy(F_out, F, sizeof(float) * n);
    free(F); free(DD);
    return 0;
}

static uint8_t dir_code2d(float c, float n)
{
    /* frangi.cpp:457-459: round((((c/n)+1)/2)*255.0), clamped; a 0/0 direction converts to a negative int -> 0 */
    float q = ((c / n) + 1) / 2;
    double r = round((double)q * 255.0);
    if (!(r == r)) return 0;
    int val = (int)r;
    return (uint8_t)(val < 0 ? 0 : (val > 255 ? 255 : val));
}

ORACLE_API int oracle_frangi2d(const uint8_t *I, int w, int h, const float *sigmas, int nsig, float beta_one, float beta_two,
                               int blackwhite, float *J, float *Jmin_out, float *Jmax_out, uint8_t *Vx, uint8_t *Vy, uint8_t *Vz)
{
    const size_t n = (size_t)w * h;
    float *Dxx = (float *)malloc(sizeof(float) * n), *Dxy = (float *)malloc(sizeof(float) * n), *Dyy = (float *)malloc(sizeof(float) * n);
    if (!Dxx || !Dxy || !Dyy) { free(Dxx); free(Dxy); free(Dyy); return -1; }
    const float beta = (float)(2 * ((double)beta_one * (double)beta_one));   /* :411-412 */
    const float c = (float)(2 * ((double)beta_two * (double)beta_two));
    float Jmin = FLT_MAX, Jmax = -FLT_MAX;
    for (int si = 0; si < nsig; ++si) {
        oracle_hessian2d(I, w, h, sigmas[si], Dyy, Dxy, Dxx, NULL);
        for (size_t i = 0; i < n; ++i) {
            float d = Dxx[i] - Dyy[i];
            float tmp = (float)sqrt((double)d * (double)d + 4 * ((double)Dxy[i] * (double)Dxy[i]));      /* :425 */
            float v2x = 2 * Dxy[i];
            float v2y = Dyy[i] - Dxx[i] + tmp;
            float mag = (float)sqrt((double)v2x * (double)v2x + (double)v2y * (double)v2y);            /* :430 */
            if (mag > 0) { v2x /= mag; v2y /= mag; }
            float v1x = -v2y, v1y = v2x;
            float mu1 = (float)(0.5 * (double)(Dxx[i] + Dyy[i] + tmp));                                 /* :440-441 */
            float mu2 = (float)(0.5 * (double)(Dxx[i] + Dyy[i] - tmp));
            int check = fabsf(mu1) < fabsf(mu2);                                                        /* :444 */
            float L1 = check ? mu2 : mu1, L2 = check ? mu1 : mu2;
            float Vecx = check ? v2x : v1x, Vecy = check ? v2y : v1y;
            L1 = (L1 == 0) ? FLT_MIN : L1;
            float q = L2 / L1;
            float Rb = (float)((double)q * (double)q);
            float S2 = (float)((double)L1 * (double)L1 + (double)L2 * (double)L2);
            float v = expf(-Rb / beta) * (1 - expf(-S2 / c));                                           /* :454 */
            if (blackwhite) v = (L1 < 0) ? 0 : v; else v = (L1 > 0) ? 0 : v;
            if (si == 0 || v > J[i]) {                                                                  /* :462-503 */
                J[i] = v;
                if (J[i] < Jmin) Jmin = J[i];
                if (J[i] > Jmax) Jmax = J[i];
                float Vecn = sqrtf(Vecx * Vecx + Vecy * Vecy);
                Vx[i] = dir_code2d(Vecx, Vecn);
                Vy[i] = dir_code2d(Vecy, Vecn);
                Vz[i] = 0;
            }
        }
    }
    *Jmin_out = Jmin; *Jmax_out = Jmax;
    free(Dxx); free(Dxy); free(Dyy);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Soma helpers (second part; Advantra_plugin.cpp:2426-2440). */
/* ------------------------------------------------------------------------ */
/* Frangi::imerode(I,w,h,l,rad,E) (frangi.cpp:879-969) / Frangi::imdilate(I,w,h,l,rad) (:1110-1199): separable xy
 * minimum / maximum over [c - ceil(rad), c + ceil(rad)], indices clamped (the three x ranges of the reference are the
 * same clamped formula), x pass into a scratch volume, y pass from it. */
ORACLE_API int oracle_morph_xy(const uint8_t *I, int w, int h, int l, float rad, int is_min, uint8_t *out)
{
    const int L = (int)ceilf(rad);
    const size_t n = (size_t)w * h * l;
    uint8_t *K = (uint8_t *)malloc(n);
    if (!K) return -1;
    for (int z = 0; z < l; ++z)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const uint8_t *row = I + ((size_t)z * h + y) * w;
                uint8_t v = row[x];
                for (int k = -L; k <= L; ++k) {
                    int xs = x + k < 0 ? 0 : (x + k > w - 1 ? w - 1 : x + k);
                    if (is_min ? row[xs] < v : row[xs] > v) v = row[xs];
                }
                K[((size_t)z * h + y) * w + x] = v;
            }
    for (int z = 0; z < l; ++z)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const uint8_t *pl = K + (size_t)z * h * w;
                uint8_t v = pl[(size_t)y * w + x];
                for (int k = -L; k <= L; ++k) {
                    int ys = y + k < 0 ? 0 : (y + k > h - 1 ? h - 1 : y + k);
                    uint8_t q = pl[(size_t)ys * w + x];
                    if (is_min ? q < v : q > v) v = q;
                }
                out[((size_t)z * h + y) * w + x] = v;
            }
    free(K);
    return 0;
}

/* Frangi::imerode(I,w,h,l,rad,zdist,E) (frangi.cpp:971-1108): the xy erosion above followed by the minimum along z
 * over [z - ceil(rad/zdist), z + ceil(rad/zdist)] clamped; a single plane skips the z pass (:1062-1064). */
ORACLE_API int oracle_imerode_z(const uint8_t *I, int w, int h, int l, float rad, float zdist, uint8_t *out)
{
    const size_t n = (size_t)w * h * l;
    uint8_t *K = (uint8_t *)malloc(n);
    if (!K) return -1;
    if (oracle_morph_xy(I, w, h, l, rad, 1, K)) { free(K); return -1; }
    if (l == 1) {
        memcpy(out, K, n);
    } else {
        const int Lz = (int)ceilf(rad / zdist);
        const size_t plane = (size_t)w * h;
        for (int z = 0; z < l; ++z)
            for (size_t q = 0; q < plane; ++q) {
                uint8_t v = K[(size_t)z * plane + q];
                for (int k = -Lz; k <= Lz; ++k) {
                    int zs = z + k < 0 ? 0 : (z + k > l - 1 ? l - 1 : z + k);
                    uint8_t c = K[(size_t)zs * plane + q];
                    if (c < v) v = c;
                }
                out[(size_t)z * plane + q] = v;
            }
    }
    free(K);
    return 0;
}

/* Frangi::imgaussian(I,w,h,l,sig) (frangi.cpp:786-877): xy Gaussian of a uint8 volume in place.  The x pass
 * accumulates in float32 (:806-838); the y pass accumulates INTO the unsigned char -- `I[i0] = 0; I[i0] += K*G` --
 * so every tap is (unsigned char)((float)I[i0] + K*G) (:841-873). */
ORACLE_API int oracle_imgaussian_xy_u8(uint8_t *I, int w, int h, int l, float sigma)
{
    const int L = oracle_gauss_radius(sigma);
    const size_t n = (size_t)w * h * l;
    float *g = (float *)malloc(sizeof(float) * (size_t)(2 * L + 1));
    float *K = (float *)malloc(sizeof(float) * n);
    if (!g || !K) { free(g); free(K); return -1; }
    oracle_gauss_taps(sigma, L, g);
    for (int z = 0; z < l; ++z)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const uint8_t *row = I + ((size_t)z * h + y) * w;
                float acc = 0.0f;
                for (int k = -L; k <= L; ++k) {
                    int xs = x + k < 0 ? 0 : (x + k > w - 1 ? w - 1 : x + k);
                    acc += (float)(int)row[xs] * g[k + L];
                }
                K[((size_t)z * h + y) * w + x] = acc;
            }
    for (int z = 0; z < l; ++z)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const float *pl = K + (size_t)z * h * w;
                uint8_t acc = 0;
                for (int k = -L; k <= L; ++k) {
                    int ys = y + k < 0 ? 0 : (y + k > h - 1 ? h - 1 : y + k);
                    acc = (uint8_t)((float)(int)acc + pl[(size_t)ys * w + x] * g[k + L]);
                }
                I[((size_t)z * h + y) * w + x] = acc;
            }
    free(g); free(K);
    return 0;
}
