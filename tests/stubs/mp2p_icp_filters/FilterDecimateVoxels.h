#pragma once
#include <mp2p_icp_filters_shape.h>  // shape stub (tests/stubs/mp2p_icp_filters_shape.h)
