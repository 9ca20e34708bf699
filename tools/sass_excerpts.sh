#!/bin/bash
# SASS evidence for the hot kernels, from the built libflipb200.so (run here, no GPU): per kernel the ptxas resource line, the
# instruction mix and the instructions that show HOW data moves (cluster barriers, DSMEM stores, L1-bypassing loads, named barriers).
# usage: bash tools/sass_excerpts.sh r02   -> profiles/r02_sass_<kernel>.txt
TAG=${1:-r02}
SO=zeno_b200/libflipb200.so
cuobjdump -sass $SO > /tmp/flipb200.sass 2>/dev/null
for K in mg_cluster_kernel mg_cycle_kernel g2p_tile_kernelILb1 p2g_gather_kernel p2g_xrow_kernel; do
  OUT=profiles/${TAG}_sass_$(echo $K | sed 's/ILb1//').txt
  awk -v k="$K" '/Function :/ {on = index($0, k) > 0} on {print}' /tmp/flipb200.sass > /tmp/k.sass
  {
    echo "# $K -- cuobjdump -sass zeno_b200/libflipb200.so (sm_100a), $(grep -c -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/k.sass) instructions"
    grep -h -A3 "$(echo $K | sed 's/ILb1//')" zeno_b200/csrc/_build/*.log 2>/dev/null | grep -m1 "Used"
    echo; echo "## instruction mix (top 30 mnemonics)"
    grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/k.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed 's/;$//' | sort | uniq -c | sort -rn | head -30
    echo; echo "## data-movement / synchronisation instructions (count, first occurrence)"
    for PAT in 'UCGABAR' 'CCTL' 'MAPA' 'ST\.E\.[A-Z0-9.]*' 'STS' 'LDS' 'LDG\.E\.[A-Z0-9.]*' 'LD\.E\.[A-Z0-9.]*' 'BAR\.SYNC' 'BAR\.ARV' 'ATOM' 'RED\.' 'MEMBAR' 'FENCE' 'SHFL' 'DFMA' 'F2F' 'UTMALDG' 'UBLKCP' 'SYNCS'; do
      n=$(grep -c -E "\s$PAT" /tmp/k.sass); [ "$n" -gt 0 ] && echo "$n x $PAT   e.g. $(grep -m1 -E "\s$PAT" /tmp/k.sass | sed -E 's/^\s+//; s/\s+\/\*.*$//')"
    done
  } > $OUT
  echo "wrote $OUT"
done
