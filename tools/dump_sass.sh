#!/bin/bash
# SASS listing of the two kernels of the headline step (instruction text only), with the opcode histogram on top.
# usage: tools/dump_sass.sh <round tag>   -> profiles/<tag>_sass_headline_{front,tail}.txt
cd "$(dirname "$0")/.."
R=${1:-r2}
dump() {   # $1 mangled-name pattern, $2 output
  cuobjdump -sass sdim_b200/csrc/libsdimb.so 2>/dev/null | awk -v pat="$1" '
    /Function :/ { on = ($0 ~ pat) } on' | grep -E "Function :|^\s+/\*[0-9a-f]{4,}\*/" | sed -E 's@/\* 0x[0-9a-f]+ \*/@@; s/[[:space:]]+$//' > /tmp/sass_$$.txt
  { echo "# cuobjdump -sass sdim_b200/csrc/libsdimb.so, kernel pattern $1 ($(grep -c '/\*' /tmp/sass_$$.txt) instructions)"
    echo "# opcode histogram:"
    grep -v "Function :" /tmp/sass_$$.txt | sed -E 's@^\s+/\*[0-9a-f]+\*/\s+@@; s/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -24 | awk '{printf "#   %-10s %s\n", $2, $1}'
    cat /tmp/sass_$$.txt; } > "$2"
  rm -f /tmp/sass_$$.txt
}
dump "gate_stream_kernelILi3ELb0ELb1" profiles/${R}_sass_headline_front.txt
dump "run_tail_kernelILi3ELb0ELi16" profiles/${R}_sass_headline_tail.txt
wc -l profiles/${R}_sass_headline_*.txt
