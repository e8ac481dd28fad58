#!/bin/bash
# opcode histogram of one kernel of a variant library, next to the reference variant:  sass_hist.sh <ref> <variant> [mangled name]
cd "$(dirname "$0")/../build_variants"
F=${3:-_ZN2dd21policy_rollout_kernelILb1ELi16ELb0ELi3ELb1ELb1EEEvNS_5PArgsE}
for v in $1 $2; do
  cuobjdump -sass libdd_$v.so | awk -v f="$F" '/Function :/ {on = ($3==f)} on' > fn_$v.sass
  grep -E "^\s+/\*[0-9a-f]{4,5}\*/" fn_$v.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | awk '{ if ($1 ~ /^@/) print $2; else print $1}' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn > sass_$v.txt
  echo $v $(awk '{s+=$1} END {print s}' sass_$v.txt) instructions
done
join -1 2 -2 2 -a1 -a2 -e0 -o 0,1.1,2.1 <(sort -k2 sass_$1.txt) <(sort -k2 sass_$2.txt) | awk '{d=$3-$2; if (d>4||d<-4) print $0, d}' | sort -k4 -n
