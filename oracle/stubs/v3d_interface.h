// oracle/stubs/v3d_interface.h -- TEST INFRASTRUCTURE.  Stand-in (ours) for the Vaa3D plugin header that the
// reference's tracker.h:11 and Advantra_plugin.h:9 include.  tracker.cpp uses nothing from Vaa3D itself, only names
// that the real header pulls in transitively (INT_MAX at tracker.cpp:359,524, printf, std::string).  When the Qt
// stand-in (oracle/stubs/QtGui) has been included first -- i.e. in the translation unit of Advantra_plugin.cpp -- this
// header also declares the small part of the Vaa3D plugin API that file names: the image handle, the callback, the
// plugin interface, the SWC records and the load / save wrappers.  The wrappers are DEFINED in oracle/plugin_wrap.cpp:
// the "image file" is a volume the test put in memory, SWC files are written as plain text, saved images are dropped.
#pragma once
#include <climits>
#include <cstdio>
#include <ctime>
#include <string>

#ifdef PNR_PLUGIN_STUBS       // set by oracle/plugin_wrap.cpp
#include <QtGui>
#include <vector>
typedef long long V3DLONG;
typedef void* v3dhandle;
enum ImagePixelType { V3D_UNKNOWN = 0, V3D_UINT8 = 1, V3D_UINT16 = 2, V3D_FLOAT32 = 4 };

class Image4DSimple {
public:
    unsigned char* data; V3DLONG sz[4]; QString name;
    unsigned char* getRawData() { return data; }
    V3DLONG getXDim() const { return sz[0]; }
    V3DLONG getYDim() const { return sz[1]; }
    V3DLONG getZDim() const { return sz[2]; }
    V3DLONG getCDim() const { return sz[3]; }
    ImagePixelType getDatatype() const { return V3D_UINT8; }
    QString getFileName() const { return name; }
};

class V3DPluginCallback2 {
public:
    v3dhandle currentImageWindow() { return 0; }      // no GUI: the menu path finds no open image
    Image4DSimple* getImage(v3dhandle) { return 0; }
};

struct V3DPluginArgItem { QString type; void* p; };
typedef QList<V3DPluginArgItem> V3DPluginArgList;

class V3DPluginInterface2_1 {
public:
    virtual ~V3DPluginInterface2_1() {}
    virtual float getPluginVersion() const = 0;
    virtual QStringList menulist() const = 0;
    virtual void domenu(const QString& menu_name, V3DPluginCallback2& callback, QWidget* parent) = 0;
    virtual QStringList funclist() const = 0;
    virtual bool dofunc(const QString& func_name, const V3DPluginArgList& input, V3DPluginArgList& output,
                        V3DPluginCallback2& callback, QWidget* parent) = 0;
};

struct NeuronSWC {
    V3DLONG n; int type; float x, y, z, r; V3DLONG pn, parent; V3DLONG seg_id, nodeinseg_id;
    NeuronSWC() : n(0), type(0), x(0), y(0), z(0), r(0), pn(-1), parent(-1), seg_id(-1), nodeinseg_id(-1) {}
};
struct NeuronTree {
    QList<NeuronSWC> listNeuron;
    QHash<int, int> hashNeuron;
    QString name, comment;
};

void v3d_msg(const QString& msg, bool display = true);
bool writeSWC_file(const QString& filename, const NeuronTree& nt);
bool simple_loadimage_wrapper(V3DPluginCallback2& cb, const char* filename, unsigned char*& data1d, V3DLONG sz[4], int& datatype);
bool simple_saveimage_wrapper(V3DPluginCallback2& cb, const char* filename, unsigned char* data1d, V3DLONG sz[4], int datatype);

// nf_dialog.h is Qt GUI code (a parameter dialog); plugin_wrap.cpp defines its include guard so that it is skipped,
// and this is what Advantra::domenu (the menu path, never run here) sees instead
class CommonDialog : public QDialog {
public:
    CommonDialog(const std::vector<std::string>&, const std::vector<std::string>&, QWidget* = 0) {}
    std::string get_para(const std::string&) { return std::string(); }
    template <class T> void get_num(const std::string&, T&) {}
};
#endif
