#!/bin/bash
# TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/libflipref.so: the REAL reference FastFLIP CPU path
# (projects/FastFLIP/{FLIP_vdb,simd_vdb_poisson_uaamg,vdb_velocity_extrapolator,levelset_util}.cpp,
# projects/zenvdb/packed3grids.cpp) compiled from the sources where they lie under /root/reference,
# against the reference's vendored OpenVDB 9.0.1 and vendored TBB 2020, behind the flat C API of
# oracle/ref/ref_driver.cpp. Nothing from /root/reference is copied into the repo: third-party
# objects are staged under $BUILD (default /tmp/flipref_build), only the final shared objects land in
# oracle/_ref/ (git-ignored, travels to the GPU box).
#
# Not in the image: Boost, Eigen, TBB. Boost -> oracle/ref/shims/boost (14 tiny headers);
# Eigen -> oracle/ref/shims/Eigen (just the API the path touches; coarsest-level CG restated from
# Eigen's documented algorithm); TBB -> built from the reference's vendored copy with its own
# Makefile.old (the one thing we run that is not "gcc on a few files").
set -e
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../_ref"
BUILD=${BUILD:-/tmp/flipref_build}
JOBS=${JOBS:-8}
# the image exports CXX=/opt/gcc/bin/g++, a wrapper that links libstdc++ statically (two libstdc++ copies in one
# python process crash); use the distro g++ unless REF_CXX says otherwise
CXX=${REF_CXX:-$(test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)}
VDB="$REF/projects/zenvdb/openvdb/openvdb"
FF="$REF/projects/FastFLIP"
[ -d "$REF" ] || { echo "no $REF: keeping the prebuilt oracle/_ref"; exit 0; }
PLUGIN_SRC="$HERE/../../zeno_b200/plugin/flipb200_nodes.cpp"
if [ -f "$OUT/libflipref.so" ] && [ "$OUT/libflipref.so" -nt "$HERE/ref_driver.cpp" ] && [ "$OUT/libflipref.so" -nt "$HERE/shims/Eigen/Eigen" ] && [ "$OUT/libflipref.so" -nt "$PLUGIN_SRC" ] && [ "$OUT/libflipref.so" -nt "$HERE/plugin_nodes_test.cpp" ] && [ "$OUT/libflipref.so" -nt "$HERE/ref_nodes_test.cpp" ] && [ "$OUT/libflipref.so" -nt "$HERE/node_harness.inc" ] && [ -f "$OUT/libflipplugin_gpu.so" ] && [ "$OUT/libflipplugin_gpu.so" -nt "$HERE/plugin_nodes_gpu.cpp" ] && [ "$OUT/libflipplugin_gpu.so" -nt "$PLUGIN_SRC" ] && [ -z "$FORCE" ]; then
  echo "oracle/_ref/libflipref.so is up to date"; exit 0
fi
mkdir -p "$OUT" "$BUILD/gen/openvdb" "$BUILD/vdbobj" "$BUILD/ffobj"

# ---- 1. TBB 2020 from the vendored copy (out-of-tree copy: the reference tree is read-only)
if [ ! -f "$BUILD/libtbb.so.2" ]; then
  rm -rf "$BUILD/tbb" && cp -r "$REF/projects/Geometry/instant_meshes/ext/tbb" "$BUILD/tbb" && chmod -R u+w "$BUILD/tbb"
  ( cd "$BUILD/tbb" && make -f Makefile.old tbb compiler=gcc stdver=c++14 CXXFLAGS="-fpermissive -w" -j$JOBS > "$BUILD/tbb.log" 2>&1 ) || { tail -30 "$BUILD/tbb.log"; exit 1; }
  cp "$(find "$BUILD/tbb/build" -name 'libtbb.so.2' | grep release | head -1)" "$BUILD/libtbb.so.2"
  ln -sf libtbb.so.2 "$BUILD/libtbb.so"
fi
TBBINC="$BUILD/tbb/include"

# ---- 2. openvdb/version.h from version.h.in (9.0.1, ABI 9, no blosc / zlib / imath half / explicit instantiation)
python3 - "$VDB/openvdb/version.h.in" "$BUILD/gen/openvdb/version.h" <<'EOF'
import re, sys
s = open(sys.argv[1]).read()
sub = {"OpenVDB_MAJOR_VERSION": "9", "OpenVDB_MINOR_VERSION": "0", "OpenVDB_PATCH_VERSION": "1",
       "OPENVDB_ABI_VERSION_NUMBER": "9", "OPENVDB_PACKED_VERSION": "0x09000001", "OPENVDB_NAMESPACE_SUFFIX": ""}
for k, v in sub.items():
    s = s.replace("${%s}" % k, v)
s = re.sub(r"#cmakedefine (\w+)", r"/* #undef \1 */", s)
open(sys.argv[2], "w").write(s)
EOF

COMMON="-std=c++17 -O2 -fPIC -w -include cstring -include $HERE/shims/boost_compat.h -I$HERE/shims -I$BUILD/gen -I$BUILD/gen/openvdb -I$VDB -I$VDB/openvdb -I$TBBINC -DOPENVDB_PRIVATE"

# ---- 3. the OpenVDB library objects (23 of 26 .cc; io/{File,Stream,TempFile}.cc are file IO only)
VDBSRC="Grid MetaMap Metadata Platform openvdb io/Archive io/Compression io/DelayedLoadMetadata io/GridDescriptor io/Queue math/Half math/Maps math/Proximity math/QuantizedUnitVec math/Transform points/AttributeArray points/AttributeArrayString points/AttributeGroup points/AttributeSet points/StreamCompression points/points util/Formats util/Util"
: > "$BUILD/cmds.txt"
for s in $VDBSRC; do
  o="$BUILD/vdbobj/$(echo $s | tr '/' '_').o"
  [ -f "$o" ] || echo "$CXX $COMMON -c $VDB/openvdb/$s.cc -o $o 2> $o.log || { tail -20 $o.log; exit 255; }" >> "$BUILD/cmds.txt"
done
xargs -P "$JOBS" -d '\n' -I{} bash -c {} < "$BUILD/cmds.txt" || { echo "openvdb compile failed"; exit 1; }

# ---- 4. the reference FastFLIP sources, unmodified, with the reference's flags (FF/CMakeLists.txt:56: -mavx -mfma)
FFFLAGS="$COMMON -mavx -mfma -DZENO_APIFREE -I$HERE/shims/zeno_min -I$REF/projects/zenvdb/include -I$REF/zeno/include -I$FF -I$HERE/../../include"
: > "$BUILD/cmds.txt"
# plugin_nodes_test.cpp compiles the drop-in's NODES against a minimal stand-in of the Zeno node runtime (shims/zeno_nodes,
# searched before the real zeno headers) and without FLIP_vdb.h
NODEFLAGS="$COMMON -DZENO_APIFREE -I$HERE/shims/zeno_nodes -I$REF/zeno/include -I$HERE/../../include"
for s in "$FF/FLIP_vdb.cpp" "$FF/simd_vdb_poisson_uaamg.cpp" "$FF/vdb_velocity_extrapolator.cpp" "$FF/levelset_util.cpp" "$REF/projects/zenvdb/include/zeno/packed3grids.cpp" "$HERE/ref_driver.cpp" "$HERE/ref_stubs.cpp" "$HERE/plugin_nodes_test.cpp" "$HERE/ref_nodes_test.cpp" "$HERE/zeno_error_impl.cpp"; do
  o="$BUILD/ffobj/$(basename $s .cpp).o"
  FL="$FFFLAGS"
  USES_PLUGIN=""
  case "$(basename $s)" in
    FLIP_vdb.cpp) FL="$FFFLAGS -include random -include $HERE/shims/seeded_random.h"; [ "$HERE/shims/seeded_random.h" -nt "$o" ] && rm -f "$o" ;;   # std::random_device -> a seeded stand-in (reseed / emitter)
    plugin_nodes_test.cpp) FL="$NODEFLAGS"; USES_PLUGIN=1 ;;
    ref_nodes_test.cpp) FL="$COMMON -mavx -mfma -DZENO_APIFREE -I$HERE/shims/zeno_nodes -I$REF/projects/zenvdb/include -I$REF/zeno/include -I$FF -I$REF/projects/zenvdb -I$HERE/../../include" ;;
    ref_driver.cpp) USES_PLUGIN=1 ;;
  esac
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ "$HERE/shims/Eigen/Eigen" -nt "$o" ] || { [ -n "$USES_PLUGIN" ] && [ "$PLUGIN_SRC" -nt "$o" ]; } || { case "$(basename $s)" in plugin_nodes_test.cpp|ref_nodes_test.cpp) [ "$HERE/shims/zeno_nodes/zeno/zeno.h" -nt "$o" ] || [ "$HERE/node_harness.inc" -nt "$o" ] ;; *) false ;; esac; }; then
    echo "$CXX $FL -c $s -o $o 2> $o.log || { grep -m 30 -E 'error|Error' $o.log; exit 255; }" >> "$BUILD/cmds.txt"
  fi
done
xargs -P "$JOBS" -d '\n' -I{} bash -c {} < "$BUILD/cmds.txt" || { echo "FastFLIP compile failed"; exit 1; }

# ---- 5. link
cp "$BUILD/libtbb.so.2" "$OUT/libtbb.so.2"
$CXX -shared -o "$OUT/libflipref.so" "$BUILD"/ffobj/*.o "$BUILD"/vdbobj/*.o -L"$BUILD" -ltbb -lpthread -ldl -Wl,-rpath,'$ORIGIN' -Wl,-z,defs 2> "$BUILD/link.log" || { head -40 "$BUILD/link.log"; exit 1; }
echo "built $OUT/libflipref.so"

# ---- 5b. the drop-in's node classes against the PRODUCT (libflipb200.so): oracle/_ref/libflipplugin_gpu.so
LIBB200="$HERE/../../zeno_b200/libflipb200.so"
if [ -f "$LIBB200" ]; then
  $CXX $NODEFLAGS -fvisibility=hidden -fvisibility-inlines-hidden -c "$HERE/plugin_nodes_gpu.cpp" -o "$BUILD/plugin_nodes_gpu.o" 2> "$BUILD/plugin_nodes_gpu.log" \
      || { grep -m 30 -E 'error|Error' "$BUILD/plugin_nodes_gpu.log"; exit 1; }
  $CXX -shared -o "$OUT/libflipplugin_gpu.so" "$BUILD/plugin_nodes_gpu.o" -L"$OUT" -lflipref -L"$(dirname "$LIBB200")" -lflipb200 -L"$BUILD" -ltbb -lpthread -ldl \
      -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../zeno_b200' 2> "$BUILD/link_gpu.log" || { head -40 "$BUILD/link_gpu.log"; exit 1; }
  echo "built $OUT/libflipplugin_gpu.so"
fi

# ---- 6. compile check of the Zeno-side drop-in (zeno_b200/plugin/flipb200_nodes.cpp) against the reference's own
#         headers (zeno core, zenvdb VDBGrid.h, OpenVDB): it is built into the zeno target, so here only -fsyntax-only
PLUGIN="$HERE/../../zeno_b200/plugin/flipb200_nodes.cpp"
if [ -f "$PLUGIN" ]; then
  $CXX -std=c++17 -O1 -fPIC -w -fsyntax-only -fopenmp -include cstring -include $HERE/shims/boost_compat.h -I$HERE/shims -I$BUILD/gen -I$BUILD/gen/openvdb \
      -I$VDB -I$VDB/openvdb -I$TBBINC -I$REF/projects/zenvdb/include -I$REF/zeno/include -I$HERE/../../include "$PLUGIN" \
      && echo "plugin compile check ok" > "$OUT/plugin_check.txt" || { echo "plugin compile check FAILED"; exit 1; }
fi
