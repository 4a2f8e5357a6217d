#!/bin/bash
# Stage a copy of the upstream tree for ONE gpurun call (the GPU box has no /root/reference), run the command, remove it.
#   scripts/stage_upstream.sh gpurun --timeout 900 -- 'python scripts/upstream_dropin.py --gpu --out gpurun_out/x.json'
# _upstream/ is git-ignored: upstream sources never enter this repository's history.  Only the Python sources and the
# small list files travel (no logs, no experiment dumps).
set -u
cd "$(dirname "$0")/.."
SRC="${DSSMD_REFERENCE:-/root/reference}"
rm -rf _upstream
mkdir -p _upstream
(cd "$SRC" && find . -name '*.py' -not -path './experiments_logdir/*' -not -path './logs/*' | cpio -pdm --quiet "$OLDPWD/_upstream" 2>/dev/null) \
  || (cd "$SRC" && find . -name '*.py' -not -path './experiments_logdir/*' -not -path './logs/*' -exec cp --parents {} "$OLDPWD/_upstream" \;)
"$@"
rc=$?
rm -rf _upstream
exit $rc
