// tests/cpp/frangi_members.cpp -- every use of `Frangi` in the reference's Advantra_plugin.cpp, written against the
// shim header pnr_b200/csrc/frangi.h, so that the test proves the shim is link-complete for the plugin:
//   :1727        Frangi::eigen_decomposition_static(cov, vec, eig)          (static, double[3][3])
//   :2346        Frangi::imgaussian(data1d, N, M, P, sig, zdist, G)         (static, 3-D)
//   :2432        Frangi::imerode(data1d, N, M, P, somaradius, E8)           (static)
//   :2438        Frangi::imgaussian(E8, N, M, P, somaradius)                (static, in place)
//   :2488-2497   Frangi frangiflt(...); frangiflt.frangi3d(...) / frangi2d(...)
//   :2521        frangiflt.Vxyz                                            (in a comment there; the field exists)
// plus the remaining public members of the reference's header (frangi.h:18-58), so nothing of the class is missing.
// The device-backed members are only NAMED here (their addresses are taken); the host helpers are executed:
//   frangi_members eigen <n>      reads n 3x3 double matrices from stdin, writes V (9) and d (3) per matrix to stdout
//   frangi_members dirs3 <n> | dirs2 <n>     writes the n x 3 float table
//   frangi_members idx3 | idx2    reads a table size n, n x 3 floats, then query vectors until EOF; writes indices
//   frangi_members interp <w> <h> <l>   reads the float volume then (x, y, z) triples until EOF; writes values
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "frangi.h"

namespace {
// the exact member-pointer types the plugin's calls resolve to; a missing or mistyped member fails to compile
void (*const p_eigen)(double[3][3], double[3][3], double[3]) = &Frangi::eigen_decomposition_static;
void (*const p_smooth3)(unsigned char*, int, int, int, float, float, float*) = &Frangi::imgaussian;
void (*const p_smooth_xy)(unsigned char*, int, int, int, float) = &Frangi::imgaussian;
void (*const p_smooth2)(unsigned char*, int, int, float, float*) = &Frangi::imgaussian;
void (*const p_erode)(unsigned char*, int, int, int, float, unsigned char*) = &Frangi::imerode;
void (*const p_erode_z)(unsigned char*, int, int, int, float, float, unsigned char*) = &Frangi::imerode;
void (*const p_dilate)(unsigned char*, int, int, int, float) = &Frangi::imdilate;
void (Frangi::*const p_f3)(unsigned char*, int, int, int, float*, float&, float&, unsigned char*, unsigned char*,
                           unsigned char*) = &Frangi::frangi3d;
void (Frangi::*const p_f2)(unsigned char*, int, int, int, float*, float&, float&, unsigned char*, unsigned char*,
                           unsigned char*) = &Frangi::frangi2d;
void (Frangi::*const p_h3)(unsigned char*, int, int, int, float, float, float*, float*, float*, float*, float*,
                           float*) = &Frangi::hessian3d;
void (Frangi::*const p_h2)(unsigned char*, int, int, float, float*, float*, float*) = &Frangi::hessian2d;
void (Frangi::*const p_eig)(double[3][3], double[3][3], double[3]) = &Frangi::eigen_decomposition;
void (*const p_tred2)(double[3][3], double[3], double[3]) = &Frangi::tred2;
void (*const p_tql2)(double[3][3], double[3], double[3]) = &Frangi::tql2;
double (*const p_hypot2)(double, double) = &Frangi::hypot2;
double (*const p_absd)(double) = &Frangi::absd;
}  // namespace

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    // the addresses are read through volatile objects so that every symbol is really referenced at link time
    void (*volatile named[])() = { (void (*)())p_eigen, (void (*)())p_smooth3, (void (*)())p_smooth_xy, (void (*)())p_smooth2,
                                   (void (*)())p_erode, (void (*)())p_erode_z, (void (*)())p_dilate, (void (*)())p_tred2,
                                   (void (*)())p_tql2, (void (*)())p_hypot2, (void (*)())p_absd };
    int present = 0;
    for (auto q : named) present += q != nullptr;
    volatile auto m1 = p_f3; volatile auto m2 = p_f2; volatile auto m3 = p_h3; volatile auto m4 = p_h2; volatile auto m5 = p_eig;
    present += (m1 != nullptr) + (m2 != nullptr) + (m3 != nullptr) + (m4 != nullptr) + (m5 != nullptr);
    if (present != 16) return 5;
    std::vector<float> sigs(1, 2.0f);
    Frangi frangiflt(sigs, 2.0f, .5f, .5f, 500.f, .5f, 15.f);              // Advantra_plugin.cpp:2488
    if (!frangiflt.Vxyz.empty() || Frangi::ndirs2d != 30 || Frangi::ndirs3d != 90 || frangiflt.blackwhite) return 6;
    const std::string mode = argv[1];
    if (mode == "eigen") {
        const int n = atoi(argv[2]);
        for (int i = 0; i < n; ++i) {
            double cov[3][3], vec[3][3], eig[3];
            if (fread(cov, sizeof(double), 9, stdin) != 9) return 2;
            Frangi::eigen_decomposition_static(cov, vec, eig);             // Advantra_plugin.cpp:1727
            double v2[3][3], e2[3];
            frangiflt.eigen_decomposition(cov, v2, e2);                    // the member form, frangi.cpp:198
            if (memcmp(vec, v2, sizeof vec) || memcmp(eig, e2, sizeof eig)) return 7;
            fwrite(vec, sizeof(double), 9, stdout);
            fwrite(eig, sizeof(double), 3, stdout);
        }
        return 0;
    }
    if (mode == "dirs3" || mode == "dirs2") {
        std::vector<std::vector<float> > t;
        if (mode == "dirs3") frangiflt.generate_3d_unit_directions((unsigned char)atoi(argv[2]), t);
        else frangiflt.generate_2d_unit_directions((unsigned char)atoi(argv[2]), t);
        for (auto& v : t) fwrite(v.data(), sizeof(float), 3, stdout);
        return 0;
    }
    if (mode == "idx3" || mode == "idx2") {
        int n = 0;
        if (fread(&n, sizeof n, 1, stdin) != 1) return 2;
        std::vector<std::vector<float> > t(n, std::vector<float>(3));
        for (auto& v : t) if (fread(v.data(), sizeof(float), 3, stdin) != 3) return 2;
        float q[3];
        while (fread(q, sizeof(float), 3, stdin) == 3) {
            const unsigned char k = mode == "idx3" ? frangiflt.get_direction_idx(q[0], q[1], q[2], t)
                                                   : frangiflt.get_direction_idx(q[0], q[1], t);
            fwrite(&k, 1, 1, stdout);
        }
        return 0;
    }
    if (mode == "interp") {
        const int w = atoi(argv[2]), h = atoi(argv[3]), l = atoi(argv[4]);
        std::vector<float> img((size_t)w * h * l);
        if (fread(img.data(), sizeof(float), img.size(), stdin) != img.size()) return 2;
        float q[3];
        while (fread(q, sizeof(float), 3, stdin) == 3) {
            const float v = frangiflt.interpz((int)q[0], (int)q[1], q[2], img.data(), w, h, l);
            fwrite(&v, sizeof v, 1, stdout);
        }
        return 0;
    }
    return 2;
}
