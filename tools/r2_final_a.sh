cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rollout.py tests/test_gpu_configs.py -q -m gpu -k "swings or deltabot_closed" > gpurun_out/newtests.log 2>&1; echo "pytest rc $?" >> gpurun_out/newtests.log
tail -5 gpurun_out/newtests.log
python tools/persist_probe2.py > gpurun_out/persist_probe2.log 2>&1; cat gpurun_out/persist_probe2.log | tail -15
