set -x
cd $GRAFT_REPO_ROOT
P="python tools/profile_case.py"
N="ncu --set full --clock-control none --import-source on"
for c in sawyer_lqr tracking k1m_sawyer k1m_sawyer128 k1_deltabot k1r_record; do $P $c > gpurun_out/plain_$c.log 2>&1 || echo "PLAIN FAIL $c"; done
$N -k regex:riccati_project_mech -c 1 -o gpurun_out/r2_proj_mech -f $P sawyer_lqr > gpurun_out/ncu_a.log 2>&1
$N -k regex:linearize_kernel -c 1 -o gpurun_out/r2_linearize -f $P sawyer_lqr > gpurun_out/ncu_b.log 2>&1
$N -k regex:rollout_mw -c 1 -o gpurun_out/r2_k1m_sawyer1024 -f $P k1m_sawyer > gpurun_out/ncu_c.log 2>&1
$N -k regex:rollout_mw -c 1 -o gpurun_out/r2_k1m_sawyer128 -f $P k1m_sawyer128 > gpurun_out/ncu_d.log 2>&1
$N -k regex:rollout_kernel -c 1 -o gpurun_out/r2_k1_deltabot -f $P k1_deltabot > gpurun_out/ncu_e.log 2>&1
$N -k regex:rollout_rows -c 1 -o gpurun_out/r2_k1r_record -f $P k1r_record > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/plain_*.log
