// Compatibility header (see compat/NEST.hh): the reference's execNEST.hh is part of the one mirror header here.
#pragma once
#include "NEST.hh"
