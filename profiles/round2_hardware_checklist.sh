#!/usr/bin/env bash
# First GPU call of the next round: everything that was written or changed after round 1's GPU budget ran out, each in its own pytest
# process (a CUDA fault in one must not poison the others), with --runxfail so that the non-strict xfail markers report real results.
#   gpurun --timeout 1500 -- 'bash profiles/round2_hardware_checklist.sh'            (1 GPU)
#   gpurun --gpus 2 --timeout 900 -- 'bash profiles/round2_hardware_checklist.sh 2'  (adds the sharded-PPO parity cases)
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 600 python -m pytest "$@" -q --runxfail --timeout 300 -x 2>&1 | tail -15 | cut -c1-220; }
run tests/test_gpu_train_espo.py -m gpu
run tests/test_gpu_zz_fastsac_replay.py -m gpu
run tests/test_gpu_zz_head_gemm.py -m gpu
run tests/test_gpu_zzzz_fastsac.py -m gpu
run tests/test_gpu_zzz_ppo_lstm.py -m gpu
if [ "${1:-1}" -ge 2 ]; then
  run tests/test_gpu_train.py -m gpu -k "two_gpu or peer_allreduce"
fi
