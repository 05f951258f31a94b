#!/bin/bash
# End-to-end IPM runs on the GPU box: the UNMODIFIED reference driver (src/tt_ipm.py) with the Newton-system
# path swapped by ttipm_b200.dropin, next to the pure reference on the box's host cores.
# The GPU box has no /root/reference, so the three directories the driver needs (src/, psd_system/, configs/)
# are staged for the duration of ONE gpurun call under the git-ignored oracle/_ref/ and removed again when the
# call returns; nothing of the reference is kept in the repository or its history.
# usage: tools/e2e_gpu.sh <timeout_s> '<command run on the box with TTIPM_REF_TREE set>' [logfile]
set -e
cd "$(dirname "$0")/.."
STAGE=oracle/_ref/reftree_transient
cleanup() { rm -rf "$STAGE"; }
trap cleanup EXIT
rm -rf "$STAGE"; mkdir -p "$STAGE"
for d in src psd_system configs; do cp -r /root/reference/$d "$STAGE/$d"; done
find "$STAGE" -name '*.so' -delete
python tensor-train-interior-point-method_b200/build.py > /dev/null
mkdir -p gpurun_out
for attempt in $(seq 1 20); do      # exit code 3 = no GPU slot free right now (nothing charged): retry
  set +e
  gpurun --timeout "$1" -- "export TTIPM_REF_TREE=\$PWD/$STAGE; $2" > "${3:-gpurun_out/last_e2e.log}" 2>&1
  rc=$?
  set -e
  if [ $rc -ne 3 ]; then break; fi
  sleep 90
done
