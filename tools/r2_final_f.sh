cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CCB_ERRLOG=$GRAFT_REPO_ROOT/gpurun_out/test_errors.txt timeout 200 python -m pytest tests -q -m gpu -x > gpurun_out/gpu_tests.txt 2>&1; echo "pytest rc $?" >> gpurun_out/gpu_tests.txt; tail -4 gpurun_out/gpu_tests.txt
timeout 60 python tests/fuzz_parity.py 80 41 > gpurun_out/fuzz_final.jsonl 2>gpurun_out/fuzz_final.err; tail -1 gpurun_out/fuzz_final.jsonl
