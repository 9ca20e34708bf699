// TEST INFRASTRUCTURE ONLY. Stand-in for projects/zenvdb/include/zeno/VDBGrid.h used when building
// oracle/_ref: FLIP_vdb.h:7 includes it only to get packed_FloatGrid3 (the VDBGridWrapper node
// objects belong to the Zeno runtime, which oracle/_ref does not link). The real packed3grids.h and
// the real zeno/core/IObject.h (header-only with ZENO_APIFREE) are used from the reference tree.
#pragma once
#ifndef ZENO_APIFREE
#define ZENO_APIFREE
#endif
#include <optional>
#include <vector>
#include <iostream>
#include <openvdb/points/PointCount.h>
#include <openvdb/tree/LeafManager.h>
#include <openvdb/points/PointAdvect.h>
#include <openvdb/tools/Morphology.h>
#include <openvdb/tools/MeshToVolume.h>
#include <openvdb/openvdb.h>
#include <string.h>
#include <zeno/packed3grids.h>
