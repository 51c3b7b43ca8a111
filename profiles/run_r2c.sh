#!/bin/bash
# round-2 GPU visit c: device ingest + st_match tests first (fast feedback), then the whole GPU suite
T=${1:-r2c}
timeout 600 python -m pytest tests/test_ingest_gpu.py tests/test_st_match.py -x -q 2>&1 | tail -30 > gpurun_out/${T}_pytest_ingest.log; cat gpurun_out/${T}_pytest_ingest.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
