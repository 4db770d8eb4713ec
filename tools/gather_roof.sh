#!/bin/bash
# tools/gather_roof.sh -- what bounds k_small: the same kernel built twice for measurement only,
#   gather   the staging pipeline and the register loads without the AND / POPC work (-DLDPGTA_SMALL_GATHER_ONLY)
#   compute  the AND / POPC / polynomial work without any copy (-DLDPGTA_SMALL_COMPUTE_ONLY)
# timed beside the product build on the configs[1] draw pattern (tools/sweep_sizes.py).  Run under gpurun; the variant
# libraries are built on the CPU side first:  python -c "import __graft_entry__ as g; g.build_variant('gather', ['-DLDPGTA_SMALL_GATHER_ONLY']); g.build_variant('compute', ['-DLDPGTA_SMALL_COMPUTE_ONLY'])"
set -u
for model in homogeneous recent; do
  echo "## $model: product build"; python tools/sweep_sizes.py $model 2,3,4 2>&1 | grep -v "^#" | cut -c1-120
  for v in gather compute; do
    echo "## $model: $v only"; LDPGTA_LIB=$PWD/ldpgta_b200/libldpgta_$v.so python tools/sweep_sizes.py $model 2,3,4 2>&1 | grep -v "^#" | cut -c1-120
  done
done
