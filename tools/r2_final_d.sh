cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/persist_probe4.py > gpurun_out/persist_probe4.log 2>gpurun_out/persist_probe4.err; cat gpurun_out/persist_probe4.log; tail -3 gpurun_out/persist_probe4.err
