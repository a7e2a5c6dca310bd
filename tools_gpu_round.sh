#!/bin/bash
# usage: tools_gpu_round.sh <tag>   — runs on the GPU box (via gpurun): tests, smoke, bench lines of every workload,
# ncu launch list + full captures of the three kernels, timelines, phase clocks, sanitizer. tools/make_profiles.py
# turns the scratch output (gpurun_out/*_<tag>.*) into the tracked summaries under profiles/.
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu_$TAG.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 240 --timeout-method thread > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
B="timeout 600 python bench.py"
Q="--no-cpu-baseline --no-e2e"
$B --steps 200 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 400 gpurun_out/bench_$TAG.json
$B --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_driver_$TAG.json 2>> gpurun_out/bench_$TAG.err      # the driver's own command line
$B --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
$B --workload yolov4_608_b64_dense --steps 100 $Q > gpurun_out/bench_dense_$TAG.json 2>> gpurun_out/bench_$TAG.err
$B --pipeline-depth 1 --steps 100 $Q > gpurun_out/bench_depth1_$TAG.json 2>> gpurun_out/bench_$TAG.err
$B --layout nhwc --steps 200 $Q > gpurun_out/bench_nhwc_$TAG.json 2>> gpurun_out/bench_$TAG.err
for w in yolov5_640_b128_sparse yolov3_640_b128_sparse yolov4_1280_b128_sparse; do
  $B --workload $w --steps 50 --warmup 3 $Q > gpurun_out/bench_${w}_$TAG.json 2>> gpurun_out/bench_$TAG.err
done
$B --workload yolov4_1280_b1024_sparse --steps 5 --warmup 3 $Q > gpurun_out/bench_1280_b1024_n1_$TAG.json 2>> gpurun_out/bench_$TAG.err
timeout 200 python tools/stock_gpu.py 64 5 > gpurun_out/stock_gpu_$TAG.txt 2>> gpurun_out/bench_$TAG.err
timeout 200 python tools/mish_bench.py gpurun_out/mish_$TAG.json > gpurun_out/mish_$TAG.txt 2>> gpurun_out/bench_$TAG.err
timeout 100 python tools/host_submit_probe.py > gpurun_out/host_submit_$TAG.txt 2>> gpurun_out/bench_$TAG.err
timeout 100 python tools/prof_timeline.py csp608_sparse 64 > gpurun_out/timeline_608_$TAG.txt 2>> gpurun_out/bench_$TAG.err
timeout 100 python tools/prof_timeline.py csp640_sparse 128 > gpurun_out/timeline_640_$TAG.txt 2>> gpurun_out/bench_$TAG.err
timeout 100 python tools/prof_timeline.py v3_640_sparse 128 > gpurun_out/timeline_640v3_$TAG.txt 2>> gpurun_out/bench_$TAG.err
timeout 100 python tools/prof_phases.py csp608_sparse 64 > gpurun_out/phases_608_$TAG.txt 2>> gpurun_out/bench_$TAG.err
# compute-sanitizer on small end-to-end runs (every kernel, both decode paths, the class-parallel NMS pass)
for tool in memcheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize.py csp_tiny csp_odd csp608_sparse v3_tiny_nopre csp608_crowd v3_crowd csp_blobs_heavy 2>&1 | grep -E "ERROR SUMMARY|count|Error" | sort | uniq -c | head -20 > gpurun_out/sanitizer_${tool}_$TAG.txt
done
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize.py csp608_sparse csp608_crowd v3_crowd csp_blobs_heavy csp_tiny v3_tiny_nopre csp_force_global 2>&1 | grep -E "RACECHECK SUMMARY|Race reported|and (Read|Write)|count ok|MISMATCH|kept" | sed -E "s/\(int, int.*\)\+/(...)+/" | sort | uniq -c | sort -rn | head -40 > gpurun_out/sanitizer_racecheck_$TAG.txt
# launch list of the bench command (per-launch durations, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 $Q --no-verify > /dev/null 2>> gpurun_out/bench_$TAG.err
# DRAM traffic of the three kernels in their natural cache state (single pass, no replay)
timeout 300 ncu --clock-control none --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:"select_kernel|decode_tma|nms_image" -s 6 -c 6 --csv --log-file gpurun_out/pipe_traffic_$TAG.csv python bench.py --pipeline-depth 1 --steps 4 --warmup 3 $Q --no-verify > /dev/null 2>> gpurun_out/bench_$TAG.err
for k in decode_tma nms_image select_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${k}_$TAG python bench.py --pipeline-depth 1 --steps 3 --warmup 3 $Q --no-verify > /dev/null 2>> gpurun_out/bench_$TAG.err
done
ls -la gpurun_out | grep $TAG | wc -l
