#include "builtin_registry.cuh"
