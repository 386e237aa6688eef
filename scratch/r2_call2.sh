#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 400 python scratch/r2_dos_ab.py 2>&1 | tail -12 | tee gpurun_out/r2_dos_ab.txt
echo "ab done at $(( $(date +%s) - S )) s"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dos_compact -c 1 -o gpurun_out/r2_k_dos_compact_cfg3 -f python scratch/r2_dos_ab.py quick > gpurun_out/ncu_dos.log 2>&1
tail -3 gpurun_out/ncu_dos.log
echo "ncu done at $(( $(date +%s) - S )) s"
