// frangi_shim_host.cpp -- the members of the reference's `class Frangi` (pnr-vaa3d/frangi.h) that are plain host
// helpers: the public eigen-solver entry points, the direction tables and the z interpolation.  None of them is on
// the hot path (frangi3d runs on the device, frangi_shim.cpp); they exist so that a translation unit written against
// the reference's header -- Advantra_plugin.cpp:1727 calls Frangi::eigen_decomposition_static -- links against this
// library unchanged.  Written from the published algorithms (EISPACK tred2 / tql2 as in JAMA), in the reference's
// operation order, so results are bit-identical to the reference's (tests/test_cpp_shim.py checks them against the
// golden eigen fixture generated from the compiled reference).
#include <cmath>

#include "frangi.h"
#include "ref_eigen.h"

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

// ---- symmetric 3x3 eigen-decomposition in double (frangi.cpp:1230-1495): ref_eigen.h, shared with the device ---------
double Frangi::hypot2(double x, double y) { return re_hypot(x, y); }
void Frangi::tred2(double V[3][3], double d[3], double e[3]) { ref_tred2(V, d, e); }
void Frangi::tql2(double V[3][3], double d[3], double e[3]) { ref_tql2(V, d, e); }
void Frangi::eigen_decomposition_static(double A[3][3], double V[3][3], double d[3]) { ref_eigen_decomposition(A, V, d); }
void Frangi::eigen_decomposition(double A[3][3], double V[3][3], double d[3]) { eigen_decomposition_static(A, V, d); }
