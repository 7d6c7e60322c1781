/* forces.h -- the reference's header name (reference src/forces.h:30-38): assist_additional_forces and
 * assist_all_ephem are declared in assist.h here. */
#ifndef _ASSIST_FORCES_H
#define _ASSIST_FORCES_H
#include "assist.h"
#endif
