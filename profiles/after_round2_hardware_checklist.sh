#!/usr/bin/env bash
# First GPU call after round 2: everything written after the round's GPU budget ran out, each in its own pytest process (a
# CUDA fault in one must not poison the others), with --runxfail so that the non-strict xfail markers report real results.
#   gpurun --timeout 1800 -- 'bash profiles/after_round2_hardware_checklist.sh'
# Order: default paths first (strict tests), then each opt-in path alone, then the bench legs that use them.
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 900 python -m pytest "$@" -q --runxfail -x -p no:cacheprovider 2>&1 | tail -15 | cut -c1-220; }
# 1. PPO+LSTM default path (one launch per recurrence step, K-major weight copies) and FastSAC default path
run tests/test_gpu_zzz_ppo_lstm.py -m gpu -k "fwdbwd or trains_on_synthetic"
run tests/test_gpu_zzzz_fastsac.py -m gpu -k "golden or plugin_runs"
# 2. FiLM / shared encoder, CUDA-graph replay of the update
run tests/test_gpu_zzz_ppo_lstm.py -m gpu -k "options or cuda_graph"
# 3. dense layers on the tcgen05 engine (subprocess inside the tests), then the one-launch recurrence
run tests/test_gpu_zzzz_fastsac.py -m gpu -k "tensor_engine"
run tests/test_gpu_zzz_ppo_lstm.py -m gpu -k "tensor_engine"
run tests/test_gpu_zzz_ppo_lstm.py -m gpu -k "one_launch_recurrence"
# 4. the two workloads with whatever the verdict allows, and without the opt-in paths for the A/B
for w in fastsac ppo_lstm; do
  python bench.py --workload $w --no-cpu > gpurun_out/after_r2_${w}_optin.json 2> gpurun_out/after_r2_${w}_optin.log
  python bench.py --workload $w --no-cpu --no-aux-engine > gpurun_out/after_r2_${w}_default.json 2> gpurun_out/after_r2_${w}_default.log
  tail -c 1500 gpurun_out/after_r2_${w}_optin.json; echo; tail -c 600 gpurun_out/after_r2_${w}_default.json; echo
done
# 5. launch list of one PPO+LSTM iteration with the one-launch recurrence (how much of the step the recurrence still is)
RLX_LSTM_PERSISTENT=1 RLX_AUX_GEMM_ENGINE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/after_r2_lstm_launches.csv \
  python bench.py --workload ppo_lstm --no-cpu --no-aux-engine --steps 2 > gpurun_out/after_r2_lstm_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/after_r2_lstm_launches.csv 2>/dev/null | head -30
