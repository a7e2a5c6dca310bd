#!/bin/bash
TAG=b1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --timeout 240 --timeout-method thread --deselect tests/test_gpu_parity.py::test_get_bboxes_matches_oracle --deselect tests/test_gpu_parity.py::test_get_bboxes_matches_reference_golden -k "not full_size and not run_to_run" > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -25 gpurun_out/pytest_gpu_$TAG.log
timeout 100 python tools/prof_timeline.py csp608_sparse 64 > gpurun_out/timeline_608_quad_$TAG.txt 2>&1; head -12 gpurun_out/timeline_608_quad_$TAG.txt
YPP_PROF_LIB=$PWD/tools/libyolopp_prof_noquad.so timeout 100 python tools/prof_timeline.py csp608_sparse 64 > gpurun_out/timeline_608_noquad_$TAG.txt 2>&1; head -12 gpurun_out/timeline_608_noquad_$TAG.txt
timeout 100 python tools/prof_phases.py csp608_sparse 64 > gpurun_out/phases_608_$TAG.txt 2>&1; head -34 gpurun_out/phases_608_$TAG.txt
timeout 100 python tools/prof_phases.py csp1280_sparse 64 > gpurun_out/phases_1280_$TAG.txt 2>&1; head -12 gpurun_out/phases_1280_$TAG.txt
timeout 100 python tools/pipe_timeline.py 3 30 > gpurun_out/pipe_timeline_$TAG.txt 2>&1; tail -22 gpurun_out/pipe_timeline_$TAG.txt
YOLOPP_LIB=$PWD/tools/var/lib_noquad.so timeout 100 python tools/pipe_timeline.py 3 30 > gpurun_out/pipe_timeline_noquad_$TAG.txt 2>&1; tail -4 gpurun_out/pipe_timeline_noquad_$TAG.txt
