#!/bin/bash
# Everything profiles/ is built from, in one GPU session (each ncu pass only after the same command exited 0 without ncu).
#   gpurun --timeout 2400 -- 'bash tools/run_profiles.sh'
set -u
O=gpurun_out
mkdir -p $O
RI="regex:^(affine|att_|bn_|cast_|encoder|fold|gemm_|linear_|lstm_|mask_|norm_|row_stats|small_linear|softmax_rows|time_sum|unfold|utt_stats|channel|prologue|resample)"
RT="regex:^(lstm_bwd|lstm_tc|lstm_bptt|atb_tc|gemm_tc|gemm_atb|col_sum|chunk_reduce|gn_bwd|norm_residual|train_loss|clip_adam|sqnorm|shift_rows|prelu_bwd|axpy|mul_|gated|act_bwd|decoder_bwd|convw2|utt_col|bcast|bn_bwd|pool3)"
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis"
Q="--no-cpu-baseline --modes 0 --gpu-reference 0 --cfg5 0"

# 1. launch list of the headline bench (inference, cfg 2, default fp16 mode), single stream so launches do not overlap
timeout 300 python bench.py --steps 2 --warmup 3 --streams 1 $Q > $O/p_plain_cfg2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/p_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --streams 1 $Q > $O/p_ncu_l2.log 2>&1
# 2. per-kernel table (light sections, all kernels) + full capture of the top kernels, 1-block forward (fp16 mode)
timeout 200 python tools/profile_step.py --fusion att --precision fp16 > $O/p_plain_step.log 2>&1 && \
timeout 900 ncu $SECS --clock-control none -k "$RI" -o /tmp/p_all -f python tools/profile_step.py --fusion att --precision fp16 > $O/p_ncu_all.log 2>&1
ncu -i /tmp/p_all.ncu-rep --page raw --csv > $O/p_all_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^(lstm_tc|linear_persist|norm_residual)" -c 6 \
    -o /tmp/p_top3 -f python tools/profile_step.py --fusion att --precision fp16 > $O/p_ncu_top3.log 2>&1
ncu -i /tmp/p_top3.ncu-rep --page raw --csv > $O/p_top3_raw.csv 2>/dev/null
ncu -i /tmp/p_top3.ncu-rep --page details > $O/p_top3_details.txt 2>/dev/null
# 2b. the same with the persistent time-sliced LSTM kernel, and the bf16-pair split 1x1 convolution next to its TF32 form
timeout 200 python tools/profile_step.py --fusion att --precision fp16 --lstm-slices 0 > $O/p_plain_step_sliced.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^(lstm_tc_sliced)" -c 2 \
    -o /tmp/p_sliced -f python tools/profile_step.py --fusion att --precision fp16 --lstm-slices 0 > $O/p_ncu_sliced.log 2>&1
ncu -i /tmp/p_sliced.ncu-rep --page raw --csv > $O/p_sliced_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k "regex:^(gemm_persist)" -c 8 \
    -o /tmp/p_gp -f python tools/profile_step.py --fusion att --precision fp16 > $O/p_ncu_gp.log 2>&1
ncu -i /tmp/p_gp.ncu-rep --page raw --csv > $O/p_gp_raw.csv 2>/dev/null
# 3. training step (cfg 5): launch list of the bench + per-kernel table + full capture of the top kernels, 1-block model
timeout 300 python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline > $O/p_plain_cfg5.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/p_launches_cfg5.csv \
    python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline > $O/p_ncu_l5.log 2>&1
timeout 200 python tools/profile_train_step.py > $O/p_plain_train.log 2>&1 && \
timeout 900 ncu $SECS --clock-control none -k "$RT" -o /tmp/p_train -f python tools/profile_train_step.py > $O/p_ncu_train.log 2>&1
ncu -i /tmp/p_train.ncu-rep --page raw --csv > $O/p_train_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^(lstm_bptt|atb_tc|lstm_tc)" -c 14 \
    -o /tmp/p_train_top -f python tools/profile_train_step.py > $O/p_ncu_train_top.log 2>&1
ncu -i /tmp/p_train_top.ncu-rep --page raw --csv > $O/p_train_top_raw.csv 2>/dev/null
tail -2 $O/p_plain_step.log $O/p_plain_train.log $O/p_ncu_all.log $O/p_ncu_train.log $O/p_ncu_sliced.log
ls -la $O/p_* | tail -24
