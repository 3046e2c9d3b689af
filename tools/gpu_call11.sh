#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vocos.csv python tools/vocos_probe.py 32 1275 > gpurun_out/call11_vocos.log 2>&1
tail -3 gpurun_out/call11_vocos.log
