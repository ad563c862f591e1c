// tests/cpp/ref_eigen_rdouble.cu -- the instantiation of pnr_b200/csrc/ref_eigen.h that the device pass of
// FRANGI_GPU_FLAG_REFERENCE_DIRECTION runs (T = rdouble), compiled by nvcc for the HOST, where rdouble's operators are the
// plain IEEE double operations (on the device they are the __d*_rn intrinsics: the same roundings).  Reads n 3x3 double
// matrices from stdin, writes V (9) and d (3) per matrix to stdout; tests/test_cpp_shim.py holds the output bit for bit
// against the fixture generated from the compiled reference (tests/golden/case_c_eigen.npz).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ref_eigen.h"

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 0;
    std::vector<double> in(9 * (size_t)n), out(12 * (size_t)n);
    if (fread(in.data(), sizeof(double), in.size(), stdin) != in.size()) return 2;
    for (int k = 0; k < n; ++k) {
        rdouble A[3][3], V[3][3], d[3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) A[r][c] = rdouble(in[9 * k + 3 * r + c]);
        ref_eigen_decomposition(A, V, d);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) out[12 * k + 3 * r + c] = V[r][c].v;
        for (int c = 0; c < 3; ++c) out[12 * k + 9 + c] = d[c].v;
    }
    fwrite(out.data(), sizeof(double), out.size(), stdout);
    return 0;
}
