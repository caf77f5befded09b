#!/bin/bash
# usage: tools/sassmix.sh <object under csrc/build> <mangled-name fragment>: SASS of one kernel to /tmp/k.clean + instruction mix
O=/root/repo/cuda_fortran_mc_simulation_spin_b200/csrc/build/$1
L=$(cuobjdump -sass $O | grep -n "Function :" | grep -A1 "$2" | head -2 | cut -d: -f1 | tr '\n' ' '); set -- $L
E=${2:-999999}
cuobjdump -sass $O | sed -n "$1,${E}p" | sed -E 's/\/\* 0x[0-9a-f]+ \*\///' | grep -E "^\s+/\*[0-9a-f]{4}\*/" > /tmp/k.clean
wc -l < /tmp/k.clean
sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' /tmp/k.clean | sed -E 's/^@!?U?P[0-9T]+ +//' | awk '{print $1}' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn | head -${3:-14}
