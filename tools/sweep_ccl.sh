#!/bin/bash
# CCL one-CTA kernel: L2 keep-fraction sweep on 2048 config-5 maps (needs AGENDA_KNOBS=1)
export AGENDA_KNOBS=1
run() { echo "== $*"; env "$@" timeout 60 python tools/bench_ccl.py 2048 512 2>&1 | head -2; }
run AGENDA_CCL_HINTS=1
run AGENDA_CCL_HINTS=15
run AGENDA_CCL_HINTS=25
run AGENDA_CCL_HINTS=35
run AGENDA_CCL_HINTS=50
run AGENDA_CCL_HINTS=0
