#!/bin/bash
# retry wrapper around gpurun: tools/gr.sh <log> <timeout> <command...>  (retries while the pod answers "busy")
LOG=$1; TO=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $LOG || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "[gr.sh done rc=$rc]" >> $LOG
