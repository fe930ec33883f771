// constant_medium.h — drop-in for the reference header of the same name: the scene-description
// classes live in rtx.h (namespace rtx) and are exported to the global namespace here so
// that existing scene code (`#include "constant_medium.h"`, -Iinclude/rtx) compiles unchanged.
#ifndef RTX_COMPAT_CONSTANT_MEDIUM_H
#define RTX_COMPAT_CONSTANT_MEDIUM_H
#include "rtx.h"
using namespace rtx;
#endif
