#!/bin/bash
# The training-step part of profiles/ (cfg 5), each ncu pass only after the same command exited 0 without ncu:
#   gpurun --timeout 1500 -- 'bash tools/run_profiles_train.sh'
set -u
O=gpurun_out
mkdir -p $O
RT="regex:^(lstm_bwd|lstm_tc|lstm_bptt|atb_|gemm_|col_sum|chunk_reduce|gn_bwd|norm_residual|train_loss|clip_adam|sqnorm|shift_rows|prelu_bwd|axpy|mul_|gated|act_bwd|decoder_bwd|convw2|utt_col|bcast|bn_|pool3|channel_stats|cast_|affine)"
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis"
# launch list of the bench (one timed step after 3 warm-up steps)
timeout 300 python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline > $O/p_plain_cfg5.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4600 --csv --log-file $O/p_launches_cfg5.csv \
    python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline > $O/p_ncu_l5.log 2>&1
# per-kernel table (light sections) + full capture of the top kernels, 1-block model (same tensor sizes)
timeout 200 python tools/profile_train_step.py > $O/p_plain_train.log 2>&1 && \
timeout 600 ncu $SECS --clock-control none -k "$RT" -o /tmp/p_train -f python tools/profile_train_step.py > $O/p_ncu_train.log 2>&1
ncu -i /tmp/p_train.ncu-rep --page raw --csv > $O/p_train_raw.csv 2>/dev/null
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:^(lstm_bptt|atb_dual|lstm_tc|gemm_kdeep)" -c 14 \
    -o /tmp/p_train_top -f python tools/profile_train_step.py > $O/p_ncu_train_top.log 2>&1
ncu -i /tmp/p_train_top.ncu-rep --page raw --csv > $O/p_train_top_raw.csv 2>/dev/null
tail -2 $O/p_plain_train.log $O/p_ncu_train.log $O/p_ncu_train_top.log
ls -la $O/p_* | tail -12
