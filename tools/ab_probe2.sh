#!/bin/bash
# A/B of the TMA-staged first accumulation pass + NIFS parity (run under gpurun)
mkdir -p gpurun_out
TAG=${1:-ab2}
python -m pytest tests/test_gpu_nifs.py tests/test_gpu_snark.py -m gpu -x -q --timeout 900 > gpurun_out/${TAG}_pytest_nifs.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_nifs.log
REEF_MSM_TMA=1 python -m pytest tests/test_gpu_msm.py -m gpu -x -q --timeout 900 > gpurun_out/${TAG}_pytest_msm_tma.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_msm_tma.log
REEF_MSM_TMA=0 python tools/msm_probe.py 15 18 20 > gpurun_out/${TAG}_msm_plain.txt 2>&1
REEF_MSM_TMA=1 python tools/msm_probe.py 15 18 20 > gpurun_out/${TAG}_msm_tma.txt 2>&1
cat gpurun_out/${TAG}_msm_plain.txt; echo ---; cat gpurun_out/${TAG}_msm_tma.txt
