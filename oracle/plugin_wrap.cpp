// oracle/plugin_wrap.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's WHOLE plugin translation unit, Advantra_plugin.cpp, compiled UNMODIFIED where it lies (found through
// -I$(REF)), against stand-ins (ours, oracle/stubs/) for Qt and for the Vaa3D plugin API, so that its batch entry point
// Advantra::dofunc("advantra_func", ...) -- load image, soma extraction, Frangi, seed extraction, the SMC tracker,
// reconstruct(), SWC export (Advantra_plugin.cpp:274-335, 2183-2830, 2096-2181, 480-523) -- runs without Qt or Vaa3D.
// Built twice by oracle/Makefile:
//   _ref/libpnr_plugin_ref.so   with the reference's own frangi.h / frangi.cpp: the all-reference arm;
//   _ref/libpnr_plugin_gpu.so   with -DPNR_PLUGIN_GPU_FRANGI: the include guard of the reference's frangi.h is defined
//                               up front and the `Frangi` the plugin names is the drop-in class of pnr_b200/csrc/frangi.h
//                               (libfrangi_shim.so -> the C-ABI -> the CUDA kernels).  Nothing else differs: this is the
//                               swap INTEGRATION.md describes, made on the unchanged call site.
//   _ref/libpnr_plugin_replay.so  with -DPNR_PLUGIN_REPLAY_FRANGI: the reference's frangi.h / frangi.cpp under the class
//                               name RefFrangi (a -DFrangi=RefFrangi compile of the unmodified files), and a `Frangi`
//                               derived from it whose frangi3d hands back arrays the test supplied (plugin_set_replay):
//                               filter outputs captured on the GPU box, or the reference's own outputs with some
//                               eigenvector signs turned, go through the unmodified plugin on a CPU-only host.
// tests/test_plugin_e2e.py runs them on the same volume and compares the SWC files they write.
//
// The "image file" the plugin loads is a volume the test handed over (plugin_run below); SWC files are written as
// plain text (one "n type x y z r parent" row per record, the Vaa3D layout); images the plugin saves are dropped.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include <QtGui>
#define PNR_PLUGIN_STUBS
#define NF_DIALOG_H                       // nf_dialog.h (Qt GUI parameter dialog): skipped, CommonDialog is a stand-in
#include <v3d_interface.h>

#ifdef PNR_PLUGIN_GPU_FRANGI
#include "../pnr_b200/csrc/frangi.h"      // class Frangi = the drop-in
#define FRANGI3D_H                        // the reference's frangi.h (same directory as the plugin source): skipped
#endif

#ifdef PNR_PLUGIN_REPLAY_FRANGI
#define Frangi RefFrangi
#include "frangi.h"                       // the reference's header (-I$(REF)); its include guard is now set
#undef Frangi
namespace { const unsigned char* g_replay[4] = { 0, 0, 0, 0 }; }     // J8, Vx, Vy, Vz of the volume plugin_run is given
class Frangi : public RefFrangi {
public:
    Frangi(std::vector<float> s, float zd, float a, float b, float c, float b1, float b2) : RefFrangi(s, zd, a, b, c, b1, b2) {}
    // The plugin keeps of J only its 8-bit form round((J-Jmin)/(Jmax-Jmin)*255) (Advantra_plugin.cpp:2499-2512), then
    // deletes J: with Jmin = 0, Jmax = 255 and J = the supplied byte that conversion returns the byte.
    void frangi3d(unsigned char*, int w, int h, int l, float* J, float& Jmin, float& Jmax, unsigned char* Vx,
                  unsigned char* Vy, unsigned char* Vz)
    {
        if (!g_replay[0]) { fprintf(stderr, "plugin_set_replay was not called\n"); abort(); }
        const long long n = (long long)w * h * l;
        for (long long i = 0; i < n; ++i) J[i] = (float)g_replay[0][i];
        memcpy(Vx, g_replay[1], (size_t)n); memcpy(Vy, g_replay[2], (size_t)n); memcpy(Vz, g_replay[3], (size_t)n);
        Jmin = 0.f; Jmax = 255.f;
    }
};
extern "C" __attribute__((visibility("default")))
void plugin_set_replay(const unsigned char* J8, const unsigned char* Vx, const unsigned char* Vy, const unsigned char* Vz)
{
    g_replay[0] = J8; g_replay[1] = Vx; g_replay[2] = Vy; g_replay[3] = Vz;
}
#endif

#include "Advantra_plugin.cpp"

// The tracker reseeds the C generator with srand(time(NULL)) on every resampling step (tracker.cpp:655,808,1003,1098).
// As in ref_wrap.cpp the library carries its own time() and is linked with -Bsymbolic-functions, so every run -- and
// both arms of the comparison -- draws the same numbers.
extern "C" time_t time(time_t* t) noexcept
{
    if (t) *t = (time_t)1;
    return (time_t)1;
}

namespace {
const unsigned char* g_vol = 0;
long long g_dim[3] = { 0, 0, 0 };
int g_swc_written = 0;
}

void v3d_msg(const QString& msg, bool) { fprintf(stderr, "[v3d_msg] %s\n", msg.toStdString().c_str()); }

bool writeSWC_file(const QString& filename, const NeuronTree& nt)
{
    FILE* f = fopen(filename.toStdString().c_str(), "w");
    if (!f) return false;
    fprintf(f, "#name %s\n#comment %s\n##n,type,x,y,z,radius,parent\n", nt.name.toStdString().c_str(), nt.comment.toStdString().c_str());
    for (int i = 0; i < nt.listNeuron.size(); ++i) {
        const NeuronSWC& n = nt.listNeuron[i];
        fprintf(f, "%lld %d %.9g %.9g %.9g %.9g %lld\n", n.n, n.type, n.x, n.y, n.z, n.r, n.parent);
    }
    fclose(f);
    ++g_swc_written;
    return true;
}

bool simple_loadimage_wrapper(V3DPluginCallback2&, const char*, unsigned char*& data1d, V3DLONG sz[4], int& datatype)
{
    if (!g_vol) return false;
    const long long n = g_dim[0] * g_dim[1] * g_dim[2];
    data1d = new unsigned char[n];
    memcpy(data1d, g_vol, (size_t)n);
    sz[0] = g_dim[0]; sz[1] = g_dim[1]; sz[2] = g_dim[2]; sz[3] = 1;
    datatype = V3D_UINT8;
    return true;
}

bool simple_saveimage_wrapper(V3DPluginCallback2&, const char*, unsigned char*, V3DLONG*, int) { return true; }

// Runs the plugin's batch function on vol[l][h][w] (uint8).  `prefix` is what the plugin takes for the image file name:
// every output path starts with it.  params = the eleven strings of the command line (neuritesigmas, somaradius,
// tolerance, znccth, kappa, step, ni, np, zdist, nodepervol, vol; Advantra_plugin.cpp:301-312).  The three switches are
// file-scope variables of the plugin (:61, :81, :72): intermediate SWC files on, the single-tree export on (with the
// defaults reconstruct() writes no final SWC at all), a bound on the number of traces.
// Returns the number of SWC files written, or -1 when dofunc refused the arguments.
extern "C" __attribute__((visibility("default")))
int plugin_run(const unsigned char* vol, int w, int h, int l, const char* prefix, const char* const* params, int nparams,
               int save_midres, int enforce_single_tree, int max_trace_count)
{
    g_vol = vol; g_dim[0] = w; g_dim[1] = h; g_dim[2] = l;
    g_swc_written = 0;
    saveMidres = save_midres != 0;
    ENFORCE_SINGLE_TREE = enforce_single_tree != 0;
    if (max_trace_count > 0) MAX_TRACE_COUNT = max_trace_count;
    std::vector<char*> infiles(1, const_cast<char*>(prefix));
    std::vector<char*> paras;
    for (int i = 0; i < nparams; ++i) paras.push_back(const_cast<char*>(params[i]));
    V3DPluginArgItem a0, a1;
    a0.type = ""; a0.p = &infiles;
    a1.type = ""; a1.p = &paras;
    V3DPluginArgList input, output;
    input << a0 << a1;
    V3DPluginCallback2 cb;
    Advantra plugin;
    // the plugin prints progress to std::cout from every stage: parked in a failed state for the call
    const std::ios_base::iostate saved = std::cout.rdstate();
    std::cout.setstate(std::ios_base::failbit);
    const bool ok = plugin.dofunc(QString("advantra_func"), input, output, cb, 0);
    std::cout.clear(saved);
    g_vol = 0;
    return ok ? g_swc_written : -1;
}
