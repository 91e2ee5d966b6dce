#!/bin/bash
# SASS evidence for profiles/: which Blackwell-specific instructions the built library contains.  $1 = tag
TAG=${1:-r03}
mkdir -p gpurun_out
SO=isaac_ros_apriltag_b200/libb200apriltags.so
{
  echo "# cuobjdump -sass $SO (built with nvcc -gencode arch=compute_100a,code=sm_100a): instruction counts per cubin"
  cuobjdump -lelf $SO
  for pat in UTMALDG UBLKCP "SYNCS" "ATOMS" "ATOMG\|RED\." "MATCH" "REDUX" "SHFL" "DSETP\|DADD\|DMUL\|DFMA" "LDGSTS" "UTC.*MMA\|LDTM\|STTM\|HMMA"; do
    echo "== $pat: $(cuobjdump -sass $SO | grep -c "$pat")"
  done
  echo "== functions containing UTMALDG (TMA bulk tensor load):"
  cuobjdump -sass $SO | awk '/Function :/ {fn=$3} /UTMALDG/ {print fn}' | sort | uniq -c
  echo "== resource usage (registers / shared memory / stack) of every kernel:"
  cuobjdump -res-usage $SO 2>/dev/null | grep -A1 "Function" | grep -o "Function [^:]*\|REG:[0-9]*\|STACK:[0-9]*\|SHARED:[0-9]*" | paste - - - - | sed 's/_ZN6b200at[0-9]*//' | cut -c1-150
} > gpurun_out/${TAG}_sass_evidence.txt 2>&1
head -30 gpurun_out/${TAG}_sass_evidence.txt
