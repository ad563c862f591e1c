// tests/cpp/callsite.cpp -- the reference's only call site of the hot path
// (Advantra_plugin.cpp:2488-2497, 2499-2512) compiled against the shim class of
// pnr_b200/csrc/frangi.h.  Reads a raw uint8 volume, runs Frangi::frangi3d the way
// reconstruction_func does, writes J / Vx / Vy / Vz / J8 raw files next to it.
//   callsite <in.u8> <w> <h> <l> <out_prefix> <sigma,sigma,...>
// Exit code 0 on success, 3 when the GPU library reports an error (message on stderr).
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "frangi.h"

static bool dump(const std::string& path, const void* p, size_t n)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(p, 1, n, f) == n;
    fclose(f);
    return ok;
}

int main(int argc, char** argv)
{
    if (argc < 7) { fprintf(stderr, "usage: callsite in w h l out_prefix sigmas\n"); return 2; }
    const int N = atoi(argv[2]), M = atoi(argv[3]), P = atoi(argv[4]);
    const long size = (long)N * M * P;
    std::vector<unsigned char> data1d(size);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(data1d.data(), 1, size, f) != (size_t)size) { fprintf(stderr, "cannot read input\n"); return 2; }
    fclose(f);
    std::vector<float> sigs;
    for (char* tok = strtok(argv[6], ","); tok; tok = strtok(nullptr, ",")) sigs.push_back((float)atof(tok));
    const float zdist = 2.0f, alpha = .5f, beta = .5f, C = 500.f, beta_one = .5f, beta_two = 15.f;   // Advantra_plugin.cpp:66-70

    try {
        Frangi frangiflt(sigs, zdist, alpha, beta, C, beta_one, beta_two);                          // :2488
        float* J = new float[size];                                                                 // :2490-2493
        unsigned char* Vx = new unsigned char[size];
        unsigned char* Vy = new unsigned char[size];
        unsigned char* Vz = new unsigned char[size];
        float Jmin, Jmax;
        frangiflt.frangi3d(data1d.data(), N, M, P, J, Jmin, Jmax, Vx, Vy, Vz);                       // :2496
        unsigned char* J8 = new unsigned char[size];                                                // :2499-2512
        if (std::fabs(Jmax - Jmin) <= FLT_MIN) {
            for (long i = 0; i < size; ++i) J8[i] = 0;
        } else {
            for (long i = 0; i < size; ++i) {
                const double r = ((J[i] - Jmin) / (Jmax - Jmin)) * 255;
                int val = (int)((r > 0.0) ? std::floor(r + 0.5) : std::ceil(r - 0.5));
                val = val < 0 ? 0 : (val > 255 ? 255 : val);
                J8[i] = (unsigned char)val;
            }
        }
        // the device-side variant of the same normalisation must agree byte for byte
        std::vector<unsigned char> J8dev(size);
        float lo2, hi2;
        frangiflt.frangi3d_j8(data1d.data(), N, M, P, nullptr, lo2, hi2, Vx, Vy, Vz, J8dev.data());
        long mism = 0;
        for (long i = 0; i < size; ++i) mism += J8dev[i] != J8[i];
        const std::string pre = argv[5];
        bool ok = dump(pre + ".J", J, size * 4) && dump(pre + ".Vx", Vx, size) && dump(pre + ".Vy", Vy, size) &&
                  dump(pre + ".Vz", Vz, size) && dump(pre + ".J8", J8, size);
        printf("Jmin=%.9g Jmax=%.9g j8_device_mismatches=%ld\n", Jmin, Jmax, mism);
        delete[] J; delete[] J8; delete[] Vx; delete[] Vy; delete[] Vz;
        return ok && mism == 0 && lo2 == Jmin && hi2 == Jmax ? 0 : 4;
    } catch (const std::exception& e) {
        fprintf(stderr, "frangi3d: %s\n", e.what());
        return 3;
    }
}
