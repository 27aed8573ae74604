// Compatibility header: lets a translation unit written for the reference (`#include "NEST.hh"`) compile
// against the nest-b200 host mirror unchanged (include/NEST/NEST.hh of NESTCollaboration/nest).  Put
// nest_b200/host/compat ahead of the user's detector directory on the include path.
#pragma once
#include "NEST_b200.hh"
#ifndef VDetector_hh
#define VDetector_hh 1  // a detector header's own `#include "VDetector.hh"` finds the mirror's class already declared
#endif
