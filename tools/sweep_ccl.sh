#!/bin/bash
# CCL one-CTA kernel: occupancy / hint / thread-count sweep on 2048 config-5 maps (needs AGENDA_KNOBS=1)
export AGENDA_KNOBS=1
run() { echo "== $*"; env "$@" python tools/bench_ccl.py 2048 512 2>&1 | head -2; }
run X=0
run AGENDA_CCL_CTA_SMEM_KB=120
run AGENDA_CCL_CTA_SMEM_KB=120 AGENDA_CCL_HINTS=0
run AGENDA_CCL_CTA_SMEM_KB=150
run AGENDA_CCL_CTA_THREADS=512
run AGENDA_CCL_CTA_THREADS=512 AGENDA_CCL_CTA_SMEM_KB=56
run AGENDA_CCL_CTA_THREADS=256 AGENDA_CCL_CTA_SMEM_KB=56
run AGENDA_CCL_CTA_CLUSTER=2
run AGENDA_CCL_HINTS=0
