// shim: see oracle/ref_shim/ref_shim.h
#include "ref_shim.h"
