#!/bin/bash
# one GPU trip, several configurations: prints value / ms_per_step / top families for each env combination
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {
  tag="$1"; shift
  env "$@" DFD_PROFILE_OUT=gpurun_out/profile_$tag.txt timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    for line in open("gpurun_out/bench_%s.log" % tag):
        if line.startswith("{"):
            d = json.loads(line)
            fam = d["roofline"]["families"]
            print(tag, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], " | ".join("%s %.2f" % (k.replace("dfd_", ""), v["ms"]) for k, v in list(fam.items())[:7]))
            break
    else:
        print(tag, "NO JSON", open("gpurun_out/bench_%s.log" % tag).read()[-400:])
except Exception as e:
    print(tag, "ERR", e)
PY
}
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "parity" 2>&1 | tail -3
run base DFD_X=0
run nt3_128 DFD_DW_NT3=128
run nt5_256 DFD_DW_NT5=256
