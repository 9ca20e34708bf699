// TEST INFRASTRUCTURE ONLY. The socket object types of projects/zenvdb/include/zeno/VDBGrid.h:80-86,316-320
// (VDBGridWrapper<GridT>{ GridT::Ptr m_grid; }) over the minimal IObject of this shim.
#pragma once
#include <zeno/zeno.h>
#include <openvdb/openvdb.h>
#include <openvdb/points/PointDataGrid.h>
#if __has_include(<zeno/packed3grids.h>)
#include <optional>
#include <iostream>
#include <openvdb/points/PointCount.h>
#include <openvdb/tree/LeafManager.h>
#include <openvdb/points/PointAdvect.h>
#include <openvdb/tools/Morphology.h>
#include <openvdb/tools/MeshToVolume.h>
#include <zeno/packed3grids.h>   // FF/FLIP_vdb.h gets packed_FloatGrid3 through this header
#endif
namespace zeno {
// the type-erased base some zenvdb nodes ask for (VDBGrid.h:52-78); only what projects/zenvdb/VDBRenormalize.cpp touches
struct VDBGrid : IObject {
    virtual std::string getType() const { return {}; }
    virtual void dilateTopo(int) {}
};
template <typename GridT>
struct VDBGridWrapper : VDBGrid {
    typename GridT::Ptr m_grid;
    VDBGridWrapper() = default;
    explicit VDBGridWrapper(typename GridT::Ptr g) : m_grid(std::move(g)) {}
    std::string getType() const override {
        if (std::is_same<GridT, openvdb::FloatGrid>::value) return "FloatGrid";
        if (std::is_same<GridT, openvdb::Vec3fGrid>::value) return "Vec3fGrid";
        return "PointDataGrid";
    }
};
using VDBFloatGrid = VDBGridWrapper<openvdb::FloatGrid>;
using VDBFloat3Grid = VDBGridWrapper<openvdb::Vec3fGrid>;
using VDBPointsGrid = VDBGridWrapper<openvdb::points::PointDataGrid>;
}  // namespace zeno
