// stand-in: see ../../mini_cv.h
#include "../../mini_cv.h"
