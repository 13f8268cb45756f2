#!/bin/bash
# SASS opcode histogram of one kernel: tools/sass_hist.sh <object> <mangled-name-substring>
obj=$1; pat=$2
cuobjdump -sass "$obj" 2>/dev/null | awk -v pat="$pat" '
  /Function :/ { on = index($0, pat) > 0 }
  on && /^[ \t]+\/\*[0-9a-f]{4}\*\// { op=$2; sub(/;$/,"",op); if (op ~ /^@/) op=$3; split(op,a,"."); print a[1] }' | sort | uniq -c | sort -rn
