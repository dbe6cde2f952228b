#!/bin/bash
export PYTHONUNBUFFERED=1 HBTU_WALK_MASKED=1
SEC="--section SourceCounters --section WarpStateStats --section SchedulerStats --section InstructionStats --section LaunchStats --section Occupancy --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis"
for b in 5; do
  HBTU_WALK_MASKED_BLOCKS=$b timeout 120 python scratch/probe_potential.py 4e6 count 2>&1 | tail -1
  HBTU_WALK_MASKED_BLOCKS=$b timeout 400 ncu $SEC --clock-control none --import-source on -k regex:walk_masked -c 1 -o gpurun_out/r01_masked_v3_b$b python scratch/probe_potential.py 4e6 > gpurun_out/ncu_b$b.log 2>&1
  tail -2 gpurun_out/ncu_b$b.log
  ncu -i gpurun_out/r01_masked_v3_b$b.ncu-rep --page raw --csv > gpurun_out/r01_masked_v3_b${b}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r01_masked_v3_b$b.ncu-rep --page source --csv --print-source sass > gpurun_out/r01_masked_v3_b${b}_sass.csv 2>/dev/null
done
