cd $GRAFT_REPO_ROOT
P="python tools/profile_case.py"
N="ncu --set full --clock-control none --import-source on"
mkdir -p gpurun_out/sum
S() { python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/sum/$1.txt 2>/dev/null; echo >> gpurun_out/sum/$1.txt; python tools/ncu_lines.py gpurun_out/$1.ncu-rep 25 >> gpurun_out/sum/$1.txt 2>/dev/null; echo >> gpurun_out/sum/$1.txt; python tools/ncu_opmix.py gpurun_out/$1.ncu-rep 16 >> gpurun_out/sum/$1.txt 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
$P k1r_record > gpurun_out/plain_k1r_record.log 2>&1 || echo "PLAIN FAIL"
$N -k regex:rollout_rows -s 1 -c 1 -o gpurun_out/r2_k1r_record -f $P k1r_record > gpurun_out/ncu_r.log 2>&1; S r2_k1r_record
