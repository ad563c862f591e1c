// oracle/stubs/v3d_interface.h -- TEST INFRASTRUCTURE.  Stand-in (ours) for the Vaa3D plugin header that
// the reference's tracker.h:11 includes; tracker.cpp uses nothing from Vaa3D itself, only names that the
// real header pulls in transitively (INT_MAX at tracker.cpp:359,524, printf, std::string).
#pragma once
#include <climits>
#include <cstdio>
#include <ctime>
#include <string>
