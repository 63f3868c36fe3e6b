cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/variants.log
python tools/persist_probe3.py >> gpurun_out/variants.log 2>gpurun_out/variants.err
for v in $(ls build_var/libccb_*.so); do CCB_LIB=$GRAFT_REPO_ROOT/$v python tools/persist_probe3.py >> gpurun_out/variants.log 2>>gpurun_out/variants.err; done
cat gpurun_out/variants.log
best=$(grep VARIANT gpurun_out/variants.log | grep -v current | sort -k8 -n | head -1 | awk '{print $2}')
echo "best $best" | tee -a gpurun_out/variants.log
CCB_LIB=$GRAFT_REPO_ROOT/build_var/$best python -m pytest tests/test_gpu_edges.py tests/test_gpu_rollout.py -q -m gpu -x > gpurun_out/best_tests.log 2>&1; echo "rc $?" >> gpurun_out/best_tests.log; tail -3 gpurun_out/best_tests.log
