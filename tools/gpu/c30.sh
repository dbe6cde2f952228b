#!/bin/bash
# call 30: last sanity of the final binary (budget: two minutes): smoke() and the pipeline tests
export PYTHONUNBUFFERED=1
timeout 50 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 45 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -x 2>&1 | tail -2
