#!/bin/bash
# A/B of the pipeline's knobs on the headline bench (ms per step, fraction of the measured HBM peak)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { echo "== $*"; env "$@" python bench.py --no-cpu --no-e2e --steps 20 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['roofline']['frac'])"; }
run CRYOGPU_ZP_PREFILL_L2HINT=0 CRYOGPU_ZP_EXEC_PREFETCH=0 CRYOGPU_ZP_HUF=x1
run CRYOGPU_ZP_PREFILL_L2HINT=1 CRYOGPU_ZP_EXEC_PREFETCH=0 CRYOGPU_ZP_HUF=x1
run CRYOGPU_ZP_PREFILL_L2HINT=0 CRYOGPU_ZP_EXEC_PREFETCH=1 CRYOGPU_ZP_HUF=x1
run CRYOGPU_ZP_PREFILL_L2HINT=1 CRYOGPU_ZP_EXEC_PREFETCH=1 CRYOGPU_ZP_HUF=x1
run CRYOGPU_ZP_PREFILL_L2HINT=1 CRYOGPU_ZP_EXEC_PREFETCH=1 CRYOGPU_ZP_HUF=x2
run CRYOGPU_ZP_PREFILL_L2HINT=0 CRYOGPU_ZP_EXEC_PREFETCH=0 CRYOGPU_ZP_HUF=x2
for h in x1 x2; do echo "== probe $h"; CRYOGPU_ZP_HUF=$h python tools/gpu_probe.py 1024 1:1:S:hex 1:1:M:hex 1:1:D:hex 1:1:D:lowcard 2>&1 | grep method; done
CRYOGPU_ZP_HUF=x2 python tools/zp_timeline.py 2>&1 | tail -9
