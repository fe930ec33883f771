// box.h — drop-in for the reference header of the same name (see rtx.h).
#ifndef RTX_COMPAT_BOX_H
#define RTX_COMPAT_BOX_H
#include "rtx.h"
using namespace rtx;
#endif
