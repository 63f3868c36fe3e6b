cd $GRAFT_REPO_ROOT
for v in "" expopt o2 nolineinfo; do
  if [ -z "$v" ]; then unset CCB_LIB; name=current; else export CCB_LIB=$GRAFT_REPO_ROOT/build_var/libccb_$v.so; name=$v; fi
  python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['ms_per_step'], d['value'])"
done
