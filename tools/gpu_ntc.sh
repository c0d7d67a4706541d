#!/bin/bash
# node_post_bwd_tc / node_pre_bwd_tc grid sweep: CTAs per SM the grid is sized for (tiles are split evenly over the CTAs)
for cap in 0 2 3 5 7; do
echo "CAP=$cap"
GJ_PDL=0 GJ_NTC_PRE_CAP=$cap GJ_NTC_POST_CAP=$cap python tools/step_profile.py 30 4096 bf16 2>/dev/null | grep -i "bwd_tc\|reduce_steps\|us of kernel"
done
