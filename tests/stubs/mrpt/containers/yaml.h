#pragma once
#include <mrpt_shape.h>  // shape stub (tests/stubs/mrpt_shape.h)
