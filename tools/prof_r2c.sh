cd $GRAFT_REPO_ROOT
P="python tools/profile_case.py"
N="ncu --set full --clock-control none --import-source on"
mkdir -p gpurun_out/sum
S() { python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/sum/$1.txt 2>/dev/null; echo >> gpurun_out/sum/$1.txt; python tools/ncu_lines.py gpurun_out/$1.ncu-rep 30 >> gpurun_out/sum/$1.txt 2>/dev/null; echo >> gpurun_out/sum/$1.txt; python tools/ncu_opmix.py gpurun_out/$1.ncu-rep 16 >> gpurun_out/sum/$1.txt 2>/dev/null; [ "$2" = keep ] || rm -f gpurun_out/$1.ncu-rep; }
$P sawyer_lqr > gpurun_out/plain_sawyer_lqr.log 2>&1 || echo PLAIN FAIL
$N -k regex:project_lu -c 1 -o gpurun_out/r2_proj_lu -f $P sawyer_lqr > gpurun_out/ncu_a.log 2>&1
S r2_proj_lu
$N -k regex:project_apply -c 1 -o gpurun_out/r2_proj_apply -f $P sawyer_lqr > gpurun_out/ncu_b.log 2>&1
S r2_proj_apply
