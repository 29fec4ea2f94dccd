cd "$GRAFT_REPO_ROOT"
for v in "" variants/libmtgp_api2.so variants/libmtgp_api8.so ""; do
  MTGP_LIB=${v:+$PWD/$v} python bench.py --only-secondary --steps 5 > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/v.json'))
print("%-28s %.2f ms  e2e %.2f ms"%(sys.argv[1] or "default(4)", d['ms_per_step'], d['e2e']['ms_per_step']))
PY
done
