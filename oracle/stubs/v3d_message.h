// oracle/stubs/v3d_message.h -- TEST INFRASTRUCTURE.  Empty stand-in (ours) for the Vaa3D header of that name.
#pragma once
