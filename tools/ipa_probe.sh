#!/bin/bash
python -m pytest tests/test_gpu_msm.py tests/test_gpu_fullsize.py tests/test_poseidon_ro.py tests/test_gpu_multigpu_splits.py tests/test_gpu_snark.py -m gpu -x -q --timeout 900 2>&1 | tail -2
python tools/perf_probe.py 2>&1 | grep -i "hyrax"
