#!/bin/bash
# launch list of the host-buffer batch path (waves): which kernel holds a wave up
set -u
mkdir -p gpurun_out
T=${1:-e2encu}; G=${2:-0.25}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svb|zstd|vbz|xfer|words' -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
   python tools/e2e_probe.py $G full > gpurun_out/${T}_probe.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${T}_launches.csv gpurun_out/${T}_launches.txt; head -24 gpurun_out/${T}_launches.txt
