cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/sum
CCB_ERRLOG=$GRAFT_REPO_ROOT/gpurun_out/test_errors.txt python -m pytest tests -q -m gpu -x > gpurun_out/gpu_tests.txt 2>&1; echo "pytest rc $?" >> gpurun_out/gpu_tests.txt; tail -4 gpurun_out/gpu_tests.txt
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['fp64'].get('frac'), d.get('e2e_record',{}).get('value'), d['strong'], d['clocks'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:rollout_rows -s 2 -c 1 -o gpurun_out/r2_k1r_final -f python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_k1r.log 2>&1
S() { python tools/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/sum/$1.txt 2>/dev/null; echo >> gpurun_out/sum/$1.txt; python tools/ncu_lines.py gpurun_out/$1.ncu-rep 25 >> gpurun_out/sum/$1.txt 2>/dev/null; echo >> gpurun_out/sum/$1.txt; python tools/ncu_opmix.py gpurun_out/$1.ncu-rep 16 >> gpurun_out/sum/$1.txt 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
S r2_k1r_final
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
head -30 gpurun_out/sum/r2_k1r_final.txt
