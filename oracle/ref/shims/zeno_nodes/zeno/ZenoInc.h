// TEST INFRASTRUCTURE ONLY (zeno/include/zeno/ZenoInc.h pulls the core headers in)
#pragma once
#include <zeno/zeno.h>
#include <zeno/types/NumericObject.h>
