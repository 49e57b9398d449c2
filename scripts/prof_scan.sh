#!/bin/bash
# ncu --set full of the scan kernel on one workload / shard emulation.  usage: prof_scan.sh <tag> <workload> <shard> [lib]
TAG=$1; WL=$2; SH=$3; LIB=${4:-libtpq_b200.so}
mkdir -p gpurun_out
export TPQ_B200_LIB=$PWD/torchpq_b200/$LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan_kernel -s 3 -c 1 \
  -f -o gpurun_out/prof_$TAG python scripts/sweep_scan.py $WL $SH > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
