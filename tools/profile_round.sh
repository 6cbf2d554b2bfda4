#!/usr/bin/env bash
# ncu evidence of one round, on the GPU box (each profiler run only after the same command exited 0 without it):
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh'
# Outputs under gpurun_out/ (copy the summaries you keep into profiles/).
set -u
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --reps 1 --e2e-reps 1 --parity-seqs 0"
timeout 200 $B > gpurun_out/plain.log 2>&1 || { echo 'plain bench failed'; tail -5 gpurun_out/plain.log; exit 1; }
# 1. launch list of the bench command (kernel shares of a step)
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k 'regex:^k[0-9]' -c 1200 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
python tools/ncu_summary.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt
# 2. K2 score pass alone (per-case medians: tools/k2_cases.py)
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k 'regex:k2_score' --csv \
    --log-file gpurun_out/k2_score_shipped.csv python tools/bench_k2.py > gpurun_out/ncu_k2.log 2>&1
python tools/k2_cases.py gpurun_out/k2_score_shipped.csv | tee gpurun_out/k2_score_cases.txt
# 3. --set full captures of the kernels this round changed (one launch each, steady state)
full() {  # name regex skip command...
  local name=$1 re=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$re" -s "$skip" -c 1 -f -o "gpurun_out/full_$name" "$@" > "gpurun_out/ncu_full_$name.log" 2>&1
  ncu -i "gpurun_out/full_$name.ncu-rep" --page raw --csv > "gpurun_out/full_${name}_metrics.csv" 2>/dev/null
  ls -la "gpurun_out/full_$name.ncu-rep"
}
full k4_track_step2 k4_track_step2 30 python tools/bench_k4.py 40
full k3r_box_depth_med k3r_box_depth_med 20 python tools/bench_k3.py 6 20
full k3r_box_depth_hist k3r_box_depth_hist 20 python tools/bench_k3.py 6 20
full k3r_box_depth_warp 'k3r_box_depth\(' 20 python tools/bench_k3.py 6 20
full k1b_pack_image_v4 k1b_pack_image_v4 5 python tools/bench_k1b.py
full k1_disp2depth_flat k1_disp2depth_flat 10 python tools/bench_k1.py
full k2_score_flat k2_score_flat 8 python tools/bench_k2.py big
python tools/ncu_hot_lines.py gpurun_out/full_k4_track_step2.ncu-rep k4_track_step2 k4_track2 40 > gpurun_out/k4_hot_lines.txt 2>&1; head -5 gpurun_out/k4_hot_lines.txt
