#!/bin/bash
# One GPU visit.  Usage (repo root, under gpurun):  bash tools/gpu_round.sh <tag> <mode>...
#   modes: tc (tcgen05 self-tests, guarded by a hard timeout)  test (pytest -m gpu + smoke)  bench
#          micro (tools/microbench.py)  launches (ncu time-only launch list)  full:<regex>:<count> (ncu --set full)
TAG=${1:-r01}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
for MODE in "$@"; do
  case $MODE in
    tc)
      timeout -s KILL 240 python -m pytest tests/test_gpu_tcgen05.py -m gpu -q > $OUT/pytest_tc.log 2>&1; RC=$?; echo "tc rc=$RC" | tee -a $OUT/status.txt
      tail -40 $OUT/pytest_tc.log
      if [ $RC -ne 0 ]; then export RIDERS_LOFTR_IMPL=simt; echo "tcgen05 tests failed -> remaining modes run with RIDERS_LOFTR_IMPL=simt"; fi ;;
    tailtc)
      timeout -s KILL 150 python tools/tail_tc_check.py > $OUT/tail_tc_check.txt 2>&1; echo "tailtc rc=$?" | tee -a $OUT/status.txt
      tail -40 $OUT/tail_tc_check.txt ;;
    test)
      timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
      tail -15 $OUT/pytest_gpu.log
      timeout -s KILL 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/status.txt
      tail -3 $OUT/smoke.log ;;
    testsel:*)
      SEL=$(echo $MODE | cut -d: -f2-)
      timeout -s KILL 420 python -m pytest tests -m gpu -x -q -k "$SEL" > $OUT/pytest_sel.log 2>&1; echo "pytest-sel($SEL) rc=$?" | tee -a $OUT/status.txt
      tail -15 $OUT/pytest_sel.log ;;
    bench)
      timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/status.txt
      tail -c 4000 $OUT/bench.json; tail -5 $OUT/bench.err ;;
    benchstrict)
      timeout -s KILL 900 python bench.py --steps 3 --warmup 3 --strict-fp32 > $OUT/bench_strict.json 2> $OUT/bench_strict.err; echo "bench-strict rc=$?" | tee -a $OUT/status.txt
      tail -c 2500 $OUT/bench_strict.json; tail -5 $OUT/bench_strict.err ;;
    benchs)
      timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --align s > $OUT/bench_align_s.json 2> $OUT/bench_align_s.err; echo "bench-align-s rc=$?" | tee -a $OUT/status.txt
      tail -c 2500 $OUT/bench_align_s.json; tail -5 $OUT/bench_align_s.err ;;
    config1)
      timeout -s KILL 600 python bench.py --workload config1 > $OUT/bench_config1.json 2> $OUT/bench_config1.err; echo "config1 rc=$?" | tee -a $OUT/status.txt
      cat $OUT/bench_config1.json; tail -5 $OUT/bench_config1.err
      timeout -s KILL 600 python bench.py --workload config1 --align s > $OUT/bench_config1_s.json 2> $OUT/bench_config1_s.err; echo "config1-s rc=$?" | tee -a $OUT/status.txt
      cat $OUT/bench_config1_s.json; tail -5 $OUT/bench_config1_s.err ;;
    eager)
      timeout -s KILL 900 python bench.py --impl eager --steps 8 --warmup 2 > $OUT/bench_eager.json 2> $OUT/bench_eager.err; echo "eager rc=$?" | tee -a $OUT/status.txt
      cat $OUT/bench_eager.json; tail -5 $OUT/bench_eager.err ;;
    train)
      timeout -s KILL 900 python bench.py --workload train --steps 10 --warmup 3 > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "train rc=$?" | tee -a $OUT/status.txt
      cat $OUT/bench_train.json; tail -5 $OUT/bench_train.err ;;
    strong64)
      timeout -s KILL 900 python bench.py --global-frames 64 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_strong64_1gpu.json 2> $OUT/bench_strong64_1gpu.err; echo "strong64 rc=$?" | tee -a $OUT/status.txt
      tail -c 700 $OUT/bench_strong64_1gpu.json; tail -3 $OUT/bench_strong64_1gpu.err ;;
    trainprof)
      timeout -s KILL 600 python tools/train_profile.py > $OUT/train_profile.txt 2> $OUT/train_profile.err; echo "trainprof rc=$?" | tee -a $OUT/status.txt
      head -32 $OUT/train_profile.txt; tail -3 $OUT/train_profile.err ;;
    ncuall)
      timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:loftr_layer_tc|lsq_|extract_knots|delaunay_star|raster_kernel|confidence_reduce|roi_gather_tma|decoder_tail_tc|upconv_tc" -c 48 -f -o $OUT/prof_r02_hotpath \
          python bench.py --steps 1 --profile > $OUT/ncu_all.log 2>&1; echo "ncu-all rc=$?" | tee -a $OUT/status.txt
      ncu -i $OUT/prof_r02_hotpath.ncu-rep --page raw --csv > $OUT/prof_r02_hotpath_raw.csv 2>/dev/null
      python tools/ncu_summary.py $OUT/prof_r02_hotpath_raw.csv > $OUT/ncu_summary_hotpath.txt 2>&1; head -60 $OUT/ncu_summary_hotpath.txt
      timeout -s KILL 900 ncu --set full --clock-control none -k "regex:compact_valid|l1_brent|l1_apply" -c 6 -f -o $OUT/prof_r02_l1 \
          python bench.py --steps 1 --profile --align s > $OUT/ncu_l1.log 2>&1; echo "ncu-l1 rc=$?" | tee -a $OUT/status.txt
      ncu -i $OUT/prof_r02_l1.ncu-rep --page raw --csv > $OUT/prof_r02_l1_raw.csv 2>/dev/null
      python tools/ncu_summary.py $OUT/prof_r02_l1_raw.csv > $OUT/ncu_summary_l1.txt 2>&1; head -40 $OUT/ncu_summary_l1.txt
      for f in $OUT/prof_r02_hotpath.ncu-rep $OUT/prof_r02_l1.ncu-rep; do SZ=$(stat -c %s $f 2>/dev/null || echo 0); if [ "$SZ" -gt 30000000 ]; then rm -f $f; echo "dropped $f ($SZ bytes)"; fi; done ;;
    benchnotail)
      RIDERS_FUSED_TAIL=0 timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_notail.json 2> $OUT/bench_notail.err; echo "bench-notail rc=$?" | tee -a $OUT/status.txt
      tail -c 1500 $OUT/bench_notail.json ;;
    benchnotc)
      RIDERS_SUBCONV_TC=0 timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_notc.json 2> $OUT/bench_notc.err; echo "bench-notc rc=$?" | tee -a $OUT/status.txt
      tail -c 1500 $OUT/bench_notc.json ;;
    benchsub)
      RIDERS_FUSED_SUBCONV=1 timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_sub.json 2> $OUT/bench_sub.err; echo "bench-sub rc=$?" | tee -a $OUT/status.txt
      tail -c 1500 $OUT/bench_sub.json ;;
    bench3)
      RIDERS_SUBPIXEL_TAPS=3 timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_taps3.json 2> $OUT/bench_taps3.err; echo "bench-taps3 rc=$?" | tee -a $OUT/status.txt
      tail -c 1500 $OUT/bench_taps3.json ;;
    benchsimt)
      RIDERS_LOFTR_IMPL=simt timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_simt.json 2> $OUT/bench_simt.err; echo "bench-simt rc=$?" | tee -a $OUT/status.txt
      tail -c 1500 $OUT/bench_simt.json ;;
    stress)
      timeout -s KILL 600 python tools/stress_sweep.py > $OUT/stress_sweep.jsonl 2> $OUT/stress_sweep.err; echo "stress rc=$?" | tee -a $OUT/status.txt
      cat $OUT/stress_sweep.jsonl; tail -3 $OUT/stress_sweep.err ;;
    trace)
      timeout -s KILL 300 python tools/loftr_trace.py 2048 > $OUT/loftr_trace.txt 2>&1; echo "trace rc=$?" | tee -a $OUT/status.txt
      cat $OUT/loftr_trace.txt ;;
    micro:*)
      WHICH=$(echo $MODE | cut -d: -f2 | tr ',' ' ')
      timeout -s KILL 600 python tools/microbench.py $WHICH > $OUT/microbench_$(echo $WHICH | tr " " "_").jsonl 2> $OUT/microbench_part.err; echo "microbench($WHICH) rc=$?" | tee -a $OUT/status.txt
      cat $OUT/microbench_$(echo $WHICH | tr " " "_").jsonl ;;
    micro)
      timeout -s KILL 600 python tools/microbench.py > $OUT/microbench.jsonl 2> $OUT/microbench.err; echo "microbench rc=$?" | tee -a $OUT/status.txt
      cat $OUT/microbench.jsonl ;;
    launches)
      timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
          python bench.py --steps 1 --profile > $OUT/bench_under_ncu.log 2>&1; echo "ncu-launches rc=$?" | tee -a $OUT/status.txt
      python tools/summarize_launches.py $OUT/launches.csv prologue_flag > $OUT/launches_summary.txt 2>&1; head -40 $OUT/launches_summary.txt
      gzip -f $OUT/launches.csv ;;
    traffic)
      timeout -s KILL 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:roi_gather|loftr_|confidence_reduce|lsq_|extract_knots|delaunay|interp_fill|scaffold|upsample_nearest|bias_act|leaky_concat|d2s_|conv3x3_head|metrics_|l1_|decoder_tail|subpixel_conv|upconv_tc|prologue' --csv --log-file $OUT/launches_traffic.csv \
          python bench.py --steps 1 --profile > $OUT/bench_under_ncu_traffic.log 2>&1; echo "ncu-traffic rc=$?" | tee -a $OUT/status.txt
      python tools/traffic_from_ncu.py $OUT/launches_traffic.csv > $OUT/traffic.json 2> $OUT/traffic.err; cat $OUT/traffic.json; tail -3 $OUT/traffic.err
      gzip -f $OUT/launches_traffic.csv ;;
    full:*)
      REGEX=$(echo $MODE | cut -d: -f2); COUNT=$(echo $MODE | cut -d: -f3); NAME=$(echo $REGEX | tr -c 'a-zA-Z0-9\n' '_' | cut -c1-40)
      timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:$REGEX" -c ${COUNT:-6} -f -o $OUT/prof_$NAME \
          python bench.py --steps 1 --profile > $OUT/ncu_full_$NAME.log 2>&1; echo "ncu-full $NAME rc=$?" | tee -a $OUT/status.txt
      ncu -i $OUT/prof_$NAME.ncu-rep --page raw --csv > $OUT/prof_${NAME}_raw.csv 2>/dev/null
      ncu -i $OUT/prof_$NAME.ncu-rep --page details --csv > $OUT/prof_${NAME}_details.csv 2>/dev/null
      SZ=$(stat -c %s $OUT/prof_$NAME.ncu-rep 2>/dev/null || echo 0)
      if [ "$SZ" -gt 12000000 ]; then rm -f $OUT/prof_$NAME.ncu-rep; echo "dropped $NAME.ncu-rep ($SZ bytes)"; fi ;;
  esac
done
du -sh $OUT; cat $OUT/status.txt
