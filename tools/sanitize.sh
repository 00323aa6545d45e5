#!/bin/bash
# memcheck / racecheck of the kernels added in round 2 (run under gpurun)
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_witness.py tests/test_poseidon_ro.py tests/test_gpu_nifs.py "tests/test_gpu_snark.py::test_ipa_over_registered_generators_gives_the_same_proof" -m gpu -x -q --timeout 480 > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_poseidon_ro.py -m gpu -x -q --timeout 380 -k "matches_the_oracle" > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r2_racecheck.log
