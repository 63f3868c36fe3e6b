# round-2 ncu captures (one gpurun call): every kernel a BASELINE config runs.  -s skips the knot-solve launches of the LQR constructors.
cd $GRAFT_REPO_ROOT
P="python tools/profile_case.py"
N="ncu --set full --clock-control none --import-source on"
mkdir -p gpurun_out/sum
# summarise a report on the box (raw metrics, hottest source lines, opcode mix) and drop it unless asked to keep: gpurun copies back <= 64 MiB
S() { python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/sum/$1.txt 2>/dev/null; echo >> gpurun_out/sum/$1.txt; python tools/ncu_lines.py gpurun_out/$1.ncu-rep 25 >> gpurun_out/sum/$1.txt 2>/dev/null; echo >> gpurun_out/sum/$1.txt; python tools/ncu_opmix.py gpurun_out/$1.ncu-rep 16 >> gpurun_out/sum/$1.txt 2>/dev/null; [ "$2" = keep ] || rm -f gpurun_out/$1.ncu-rep; }
for c in k1r k1m_sawyer k1m_sawyer128 k1m_npend8 k1m_npend32 k1_deltabot tracking sawyer_lqr are k3 k4b pid; do $P $c > gpurun_out/plain_$c.log 2>&1 || echo "PLAIN FAIL $c"; done
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 || echo "PLAIN FAIL bench"
$N -k regex:rollout_rows -s 1 -c 1 -o gpurun_out/r2_k1r_bench -f $P k1r > gpurun_out/ncu_1.log 2>&1
S r2_k1r_bench keep
$N -k regex:rollout_mw -s 1 -c 1 -o gpurun_out/r2_k1m_sawyer1024 -f $P k1m_sawyer > gpurun_out/ncu_2.log 2>&1
S r2_k1m_sawyer1024
$N -k regex:rollout_mw -s 1 -c 1 -o gpurun_out/r2_k1m_sawyer128 -f $P k1m_sawyer128 > gpurun_out/ncu_3.log 2>&1
S r2_k1m_sawyer128 keep
$N -k regex:rollout_mw -c 1 -o gpurun_out/r2_k1m_npend8 -f $P k1m_npend8 > gpurun_out/ncu_4.log 2>&1
S r2_k1m_npend8
$N -k regex:rollout_mw -c 1 -o gpurun_out/r2_k1m_npend32 -f $P k1m_npend32 > gpurun_out/ncu_5.log 2>&1
S r2_k1m_npend32
$N -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/r2_k1_deltabot -f $P k1_deltabot > gpurun_out/ncu_6.log 2>&1
S r2_k1_deltabot
$N -k regex:riccati_sweep_dmma -s 1 -c 1 -o gpurun_out/r2_k4_tracking -f $P tracking > gpurun_out/ncu_7.log 2>&1
S r2_k4_tracking
$N -k regex:riccati_sweep_dmma -c 1 -o gpurun_out/r2_k4_sawyer -f $P sawyer_lqr > gpurun_out/ncu_8.log 2>&1
S r2_k4_sawyer keep
$N -k regex:are_sda -c 1 -o gpurun_out/r2_k5_are -f $P are > gpurun_out/ncu_9.log 2>&1
S r2_k5_are
$N -k regex:riccati_sweep_kernel -c 1 -o gpurun_out/r2_k3 -f $P k3 > gpurun_out/ncu_10.log 2>&1
S r2_k3
$N -k regex:riccati_sweep_big -c 1 -o gpurun_out/r2_k4b -f $P k4b > gpurun_out/ncu_11.log 2>&1
S r2_k4b
$N -k regex:rollout_rows -c 1 -o gpurun_out/r2_k1r_pid -f $P pid > gpurun_out/ncu_12.log 2>&1
S r2_k1r_pid
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_13.log 2>&1
for c in k1r k1m_sawyer tracking sawyer_lqr bench; do echo "== $c"; tail -n 2 gpurun_out/plain_$c.log; done
ls -la gpurun_out/*.ncu-rep | wc -l
