#!/bin/bash
# pair_dist_bwd tuning sweep: (jets per CTA, column groups per thread) -> per-kernel time from the eager step profile
for cfg in "0 0" "4 4" "4 2" "6 4" "5 4" "3 4" "2 4"; do
set -- $cfg
echo "JPB=$1 C4=$2"
GJ_PDL=0 GJ_PD_JPB=$1 GJ_PD_C4=$2 python tools/step_profile.py 30 4096 bf16 2>/dev/null | grep -i "pair_dist_bwd\|us of kernel"
done
