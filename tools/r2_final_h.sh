cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 60 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_final_quick.json 2>gpurun_out/bench_final_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final_quick.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
