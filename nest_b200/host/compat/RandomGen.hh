// Compatibility header (see compat/NEST.hh): the reference's RandomGen.hh is part of the one mirror header here.
#pragma once
#include "NEST.hh"
