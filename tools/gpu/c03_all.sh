#!/bin/bash
# combined call: c02 (tests + A/B sweeps) then the ncu evidence for the default masked kernel
bash tools/gpu/c02_tests_ab.sh
echo "== ncu (default kernel)"
timeout 700 bash tools/ncu_top.sh walk_masked_kernel r02_walk_masked_default_top; echo "ncu rc=$?"
ls -la gpurun_out | head -40
