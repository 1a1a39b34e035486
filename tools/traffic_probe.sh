#!/bin/bash
# Developer tool (GPU box): DRAM bytes, L2 hit rate and duration of the last launch of an op for a list of library builds and lags.
#   tools/traffic_probe.sh <op> <size> <frames> "<lib1> <lib2> ..." "<lag1> <lag2> ..."
op=$1; size=$2; frames=$3
for lib in $4; do for lag in $5; do
  echo "== $op $size $frames frames lib=$lib lag=$lag"
  NV12EQ_LIB=$PWD/opencv-opencl_b200/$lib timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
    --clock-control none -k regex:"clahe_kernel|equalize_kernel" -s 2 -c 1 python tools/profile_target.py --op $op --size $size --frames $frames --launches 3 --lag $lag 2>&1 \
    | grep -E "dram__|lts__|gpu__time"
done; done
