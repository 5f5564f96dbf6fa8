#!/bin/bash
# gpurun with retries while the pod answers "busy" (status transient / exit code 3); nothing is charged for those.
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$st" != "transient" ] && [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $attempt: busy, sleeping 150 s"
  sleep 150
done
exit 3
