#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_host.py C2 2>&1 | head -60 > gpurun_out/r2d_host_C2.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2d_launches_C2.csv python bench.py --config C2 --steps 3 --warmup 1 --no-cpu-baseline --no-sharded-parity > /dev/null 2>&1
grep -o '"[^"]*k_[a-z_0-9<>, ]*[^"]*","1","7","([0-9, ]*)","([0-9, ]*)".*' gpurun_out/r2d_launches_C2.csv | awk -F'","' '{print $1, $4, $NF}' | tail -25
head -45 gpurun_out/r2d_host_C2.txt
