cd $GRAFT_REPO_ROOT
P="python tools/profile_case.py"
N="ncu --set full --clock-control none --import-source on"
mkdir -p gpurun_out/sum
S() { python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/sum/$1.txt 2>/dev/null; echo >> gpurun_out/sum/$1.txt; python tools/ncu_lines.py gpurun_out/$1.ncu-rep 25 >> gpurun_out/sum/$1.txt 2>/dev/null; echo >> gpurun_out/sum/$1.txt; python tools/ncu_opmix.py gpurun_out/$1.ncu-rep 16 >> gpurun_out/sum/$1.txt 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
for c in sawyer_lqr k1m_sawyer tracking; do $P $c > gpurun_out/plain_$c.log 2>&1 || echo "PLAIN FAIL $c"; done
$N -k regex:project_apply -c 1 -o gpurun_out/r2_proj_apply_final -f $P sawyer_lqr > gpurun_out/ncu_a.log 2>&1; S r2_proj_apply_final
$N -k regex:rollout_mw -s 1 -c 1 -o gpurun_out/r2_k1m_sawyer1024_persistent -f $P k1m_sawyer > gpurun_out/ncu_b.log 2>&1; S r2_k1m_sawyer1024_persistent
$N -k regex:riccati_sweep_dmma -s 1 -c 1 -o gpurun_out/r2_k4_tracking_cpasync -f $P tracking > gpurun_out/ncu_c.log 2>&1; S r2_k4_tracking_cpasync
