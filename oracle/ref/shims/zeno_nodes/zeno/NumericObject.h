// TEST INFRASTRUCTURE ONLY (legacy include path of the reference nodes)
#pragma once
#include <zeno/types/NumericObject.h>
