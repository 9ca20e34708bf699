// TEST INFRASTRUCTURE ONLY (the FastFLIP hot-path nodes include it and use nothing from it)
#pragma once
#include <zeno/zeno.h>
