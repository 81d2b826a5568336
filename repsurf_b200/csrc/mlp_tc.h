// operand / epilogue descriptors of the tensor-core GEMMs — defined once, in the public C header
#pragma once
#include "../../include/repsurf_b200.h"
