// compiled with --fmad=false -DVILTRUM_B200_EXACT (see viltrum_b200/build.py)
#include "builtin_registry.cuh"
