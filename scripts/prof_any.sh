#!/bin/bash
# usage: prof_any.sh <tag> <kernel regex> <command...>
TAG=$1; K=$2; shift 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/prof_$TAG "$@" > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log
