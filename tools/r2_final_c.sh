cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/variants2.log
for v in $(ls build_var/libccb_*.so); do CCB_LIB=$GRAFT_REPO_ROOT/$v python tools/persist_probe3.py >> gpurun_out/variants2.log 2>>gpurun_out/variants2.err; done
cat gpurun_out/variants2.log
