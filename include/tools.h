/* tools.h -- the reference's header name (reference src/tools.h declares nothing beyond assist.h). */
#ifndef _TOOLS_H
#define _TOOLS_H
#include "assist.h"
#endif
