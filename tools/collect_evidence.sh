#!/bin/bash
# GPU box: everything profiles/ cites for one round, written to gpurun_out/<tag>_*.  Usage: bash tools/collect_evidence.sh r02 [quick]
# (ncu passes run only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values)
TAG=${1:-r02}; OUT=gpurun_out; mkdir -p $OUT
TL=lidar_odometry_b200/libb2lo_tl.so
timeout 900 python -m pytest tests -q -m gpu > $OUT/${TAG}_gpu_tests.log 2>&1; tail -2 $OUT/${TAG}_gpu_tests.log
timeout 900 python bench.py > $OUT/${TAG}_bench_full_n1.json 2> $OUT/${TAG}_bench_full_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_reference_arm.err; echo "reference arm rc=$?"
if [ -f $TL ]; then
  B2LO_LIB=$TL timeout 300 python tools/gpu_timeline.py --scans 40 > $OUT/${TAG}_timeline_stream.txt 2>&1
  B2LO_LIB=$TL timeout 300 python tools/gpu_timeline.py --scans 40 --lookahead > $OUT/${TAG}_timeline_lookahead.txt 2>&1
  B2LO_LIB=$TL timeout 300 python tools/gpu_timeline_lockstep.py 128 > $OUT/${TAG}_timeline_lockstep_128.txt 2>&1
fi
SHORT="bench.py --steps 20 --warmup 3 --no-stress --no-cpu-baseline --concurrent --lockstep"
timeout 600 python $SHORT > $OUT/${TAG}_short.json 2> $OUT/${TAG}_short.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 1500 --csv --log-file $OUT/${TAG}_launches.csv python $SHORT > $OUT/${TAG}_launches_run.log 2>&1
python tools/launch_summary.py $OUT/${TAG}_launches.csv $OUT/${TAG}_launches_summary.csv
if [ "$2" != "quick" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_icp_corr -s 40 -c 8 -o $OUT/${TAG}_k2_kitti -f python $SHORT > $OUT/${TAG}_k2_kitti_run.log 2>&1
  python tools/ncu_summary.py $OUT/${TAG}_k2_kitti.ncu-rep $OUT/${TAG}_ncu_k2_kitti.csv "# ncu --set full --clock-control none --kernel-name-base demangled -k regex:k_icp_corr -s 40 -c 8 python $SHORT" > /dev/null
fi
ls -la $OUT | grep ${TAG}_
