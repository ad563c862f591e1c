// oracle/stubs/basic_surf_objs.h -- TEST INFRASTRUCTURE.  Stand-in (ours) for the Vaa3D header of that name
// (Advantra_plugin.cpp:18); the SWC records it would declare are in oracle/stubs/v3d_interface.h.
#pragma once
