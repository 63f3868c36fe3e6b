cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export CCB_LIB=$GRAFT_REPO_ROOT/build_var/libccb_stagger.so
timeout 90 python tools/persist_probe5.py > gpurun_out/persist_probe5.log 2>gpurun_out/persist_probe5.err; cat gpurun_out/persist_probe5.log; tail -2 gpurun_out/persist_probe5.err
timeout 100 python -m pytest tests/test_gpu_edges.py tests/test_gpu_configs.py -q -m gpu -x -k "time_sliced or whole_batch or launch_selection or c2_full or record_pipeline" > gpurun_out/stagger_tests.txt 2>&1; echo "rc $?" >> gpurun_out/stagger_tests.txt; tail -3 gpurun_out/stagger_tests.txt
