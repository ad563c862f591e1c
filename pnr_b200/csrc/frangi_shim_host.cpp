// frangi_shim_host.cpp -- the members of the reference's `class Frangi` (pnr-vaa3d/frangi.h) that are plain host
// helpers: the public eigen-solver entry points, the direction tables and the z interpolation.  None of them is on
// the hot path (frangi3d runs on the device, frangi_shim.cpp); they exist so that a translation unit written against
// the reference's header -- Advantra_plugin.cpp:1727 calls Frangi::eigen_decomposition_static -- links against this
// library unchanged.  Written from the published algorithms (EISPACK tred2 / tql2 as in JAMA), in the reference's
// operation order, so results are bit-identical to the reference's (tests/test_cpp_shim.py checks them against the
// golden eigen fixture generated from the compiled reference).
#include <cmath>

#include "frangi.h"

unsigned char Frangi::ndirs2d = 30;   // frangi.cpp:24
unsigned char Frangi::ndirs3d = 90;   // frangi.cpp:25

// ---- direction tables (frangi.cpp:60-150) -------------------------------------------------------------------
// Ndir points on the unit sphere along a spiral: heights uniform in [-1, 1], azimuth advanced by
// 3.6 / sqrt(Ndir (1 - h^2)) per point (Saff & Kuijlaars), both poles at azimuth 0.
void Frangi::generate_3d_unit_directions(unsigned char Ndir, std::vector<std::vector<float> >& dirs)
{
    dirs.clear();
    double azimuth = 0.0;
    for (int k = 0; k < Ndir; ++k) {
        const double height = -1.0 + 2.0 * (double)k / (Ndir - 1);
        const double polar = std::acos(height);
        if (k == 0 || k == Ndir - 1) azimuth = 0.0;
        else azimuth = azimuth + 3.6 / (std::sqrt((double)Ndir) * std::sqrt(1.0 - height * height));
        std::vector<float> v(3);
        v[0] = (float)(std::sin(polar) * std::cos(azimuth));
        v[1] = (float)(std::sin(polar) * std::sin(azimuth));
        v[2] = (float)std::cos(polar);
        dirs.push_back(v);
    }
}

// Ndir points on the unit circle, step (2 * 3.14) / Ndir in float (the reference's constant, frangi.cpp:100)
void Frangi::generate_2d_unit_directions(unsigned char Ndir, std::vector<std::vector<float> >& dirs)
{
    dirs.clear();
    for (int k = 0; k < Ndir; ++k) {
        const float ang = (float)(k * ((2.0 * 3.14) / Ndir));
        std::vector<float> v(3);
        v[0] = (float)std::cos(ang);
        v[1] = (float)std::sin(ang);
        v[2] = 0.0f;
        dirs.push_back(v);
    }
}

// Index of a table entry "closer" than entry 0.  As in the reference (frangi.cpp:113-131) the running minimum is
// never updated, so the result is the LAST entry whose squared chord distance is below that of entry 0, with the
// reference's early exits (partial sums are compared against the same bound).
unsigned char Frangi::get_direction_idx(float vx, float vy, float vz, std::vector<std::vector<float> > dirs)
{
    unsigned char best = 0;
    const float bound = (vx - dirs[0][0]) * (vx - dirs[0][0]) + (vy - dirs[0][1]) * (vy - dirs[0][1]) +
                        (vz - dirs[0][2]) * (vz - dirs[0][2]);
    for (size_t k = 1; k < dirs.size(); ++k) {
        float d2 = (vx - dirs[k][0]) * (vx - dirs[k][0]);
        if (!(d2 < bound)) continue;
        d2 += (vy - dirs[k][1]) * (vy - dirs[k][1]);
        if (!(d2 < bound)) continue;
        d2 += (vz - dirs[k][2]) * (vz - dirs[k][2]);
        if (d2 < bound) best = (unsigned char)k;
    }
    return best;
}

unsigned char Frangi::get_direction_idx(float vx, float vy, std::vector<std::vector<float> > dirs)
{
    unsigned char best = 0;
    const float bound = (vx - dirs[0][0]) * (vx - dirs[0][0]) + (vy - dirs[0][1]) * (vy - dirs[0][1]);
    for (size_t k = 1; k < dirs.size(); ++k) {
        float d2 = (vx - dirs[k][0]) * (vx - dirs[k][0]);
        if (!(d2 < bound)) continue;
        d2 += (vy - dirs[k][1]) * (vy - dirs[k][1]);
        if (d2 < bound) best = (unsigned char)k;
    }
    return best;
}

// Linear interpolation along z only (frangi.cpp:1201-1228): plane pair [z1, z1+1] with z1 clamped to [0, l-2],
// fraction clamped to [0, 1]; a single plane returns the pixel.
float Frangi::interpz(int x, int y, float z, float* img, int w, int h, int l)
{
    if (l == 1) return img[y * w + x];
    int lower = (int)z;
    if (lower < 0) lower = 0;
    if (lower > l - 2) lower = l - 2;
    float t = z - lower;
    if (t < 0.0) t = 0.0;
    if (t > 1.0) t = 1.0;
    const float v_lo = img[lower * w * h + y * w + x];
    const float v_hi = img[(lower + 1) * w * h + y * w + x];
    return (1 - t) * v_lo + t * v_hi;
}

// ---- symmetric 3x3 eigen-decomposition in double (frangi.cpp:1230-1495) -----------------------------------------
double Frangi::hypot2(double x, double y) { return std::sqrt(x * x + y * y); }

// Householder reduction of the symmetric matrix held in V to tridiagonal form (EISPACK tred2): on return d holds
// the diagonal, e[1..2] the sub-diagonal and V the accumulated orthogonal transformation.
void Frangi::tred2(double V[3][3], double d[3], double e[3])
{
    const int N = 3;
    for (int c = 0; c < N; ++c) d[c] = V[N - 1][c];
    for (int i = N - 1; i >= 1; --i) {
        double norm1 = 0.0, hsum = 0.0;
        for (int k = 0; k < i; ++k) norm1 = norm1 + std::fabs(d[k]);
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
            const double last = d[i - 1];
            double root = std::sqrt(hsum);
            if (last > 0) root = -root;
            e[i] = norm1 * root;
            hsum = hsum - last * root;
            d[i - 1] = last - root;
            for (int c = 0; c < i; ++c) e[c] = 0.0;
            for (int c = 0; c < i; ++c) {            // similarity transform of the leading block
                const double dc = d[c];
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
            const double half = dot / (hsum + hsum);
            for (int c = 0; c < i; ++c) e[c] -= half * d[c];
            for (int c = 0; c < i; ++c) {
                const double dc = d[c], ec = e[c];
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
        const double hh = d[i + 1];
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
    for (int c = 0; c < N; ++c) {
        d[c] = V[N - 1][c];
        V[N - 1][c] = 0.0;
    }
    V[N - 1][N - 1] = 1.0;
    e[0] = 0.0;
}

// Implicit-shift QL on the tridiagonal matrix (EISPACK tql2), eigenvectors accumulated in V, then eigenvalues and
// vectors sorted ascending by value (first minimum wins).
void Frangi::tql2(double V[3][3], double d[3], double e[3])
{
    const int N = 3;
    for (int i = 1; i < N; ++i) e[i - 1] = e[i];
    e[N - 1] = 0.0;
    double shift_total = 0.0, scale_ref = 0.0;
    const double eps = std::pow(2.0, -52.0);
    for (int lo = 0; lo < N; ++lo) {
        const double cand = std::fabs(d[lo]) + std::fabs(e[lo]);
        scale_ref = scale_ref > cand ? scale_ref : cand;
        int m = lo;
        while (m < N) {
            if (std::fabs(e[m]) <= eps * scale_ref) break;
            ++m;
        }
        if (m > lo) {
            do {
                double g = d[lo];
                double p = (d[lo + 1] - g) / (2.0 * e[lo]);
                double r = hypot2(p, 1.0);
                if (p < 0) r = -r;
                d[lo] = e[lo] / (p + r);
                d[lo + 1] = e[lo] * (p + r);
                const double dl1 = d[lo + 1];
                double h = g - d[lo];
                for (int i = lo + 2; i < N; ++i) d[i] -= h;
                shift_total = shift_total + h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c;
                const double el1 = e[lo + 1];
                double s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= lo; --i) {
                    c3 = c2;
                    c2 = c;
                    s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = hypot2(p, e[i]);
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
            } while (std::fabs(e[lo]) > eps * scale_ref);
        }
        d[lo] = d[lo] + shift_total;
        e[lo] = 0.0;
    }
    for (int i = 0; i < N - 1; ++i) {
        int best = i;
        double pv = d[i];
        for (int j = i + 1; j < N; ++j)
            if (d[j] < pv) { best = j; pv = d[j]; }
        if (best != i) {
            d[best] = d[i];
            d[i] = pv;
            for (int r = 0; r < N; ++r) {
                const double t = V[r][i];
                V[r][i] = V[r][best];
                V[r][best] = t;
            }
        }
    }
}

namespace {
void swap_pair(double V[3][3], double d[3], double mag[3], int a, int b)
{
    double t = d[a]; d[a] = d[b]; d[b] = t;
    t = mag[a]; mag[a] = mag[b]; mag[b] = t;
    for (int r = 0; r < 3; ++r) { t = V[r][a]; V[r][a] = V[r][b]; V[r][b] = t; }
}
}  // namespace

// A symmetric -> columns of V = unit eigenvectors, d ordered |d0| <= |d1| <= |d2| with the reference's tie rules
// (frangi.cpp:1284-1304): the largest magnitude goes last (`>=` against the other candidate, `>` against the last
// slot), then the first two are swapped on a strict `>`.
void Frangi::eigen_decomposition_static(double A[3][3], double V[3][3], double d[3])
{
    double e[3], mag[3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) V[r][c] = A[r][c];
    tred2(V, d, e);
    tql2(V, d, e);
    for (int k = 0; k < 3; ++k) mag[k] = absd(d[k]);
    if (mag[0] >= mag[1] && mag[0] > mag[2]) swap_pair(V, d, mag, 0, 2);
    else if (mag[1] >= mag[0] && mag[1] > mag[2]) swap_pair(V, d, mag, 1, 2);
    if (mag[0] > mag[1]) swap_pair(V, d, mag, 0, 1);
}

void Frangi::eigen_decomposition(double A[3][3], double V[3][3], double d[3]) { eigen_decomposition_static(A, V, d); }
