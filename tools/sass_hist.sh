#!/bin/bash
# SASS instruction histogram of one kernel of an object file / library:  tools/sass_hist.sh <file> <regex matching the mangled name> [top]
# Counts opcodes (modifiers stripped) and prints the multiplier-pipe share (IMAD.WIDE / IMAD.HI / IMAD / IMAD.X ...).
f=$1; pat=$2; top=${3:-30}
cuobjdump -sass "$f" 2>/dev/null | awk -v pat="$pat" '
  /Function :/ { on = ($3 ~ pat); if (on) name = $3 }
  on && /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+/ {
    op = $2; if (op ~ /^@/) op = $3;
    full = op; sub(/;$/, "", full);
    base = full; sub(/\..*/, "", base);
    cnt[base]++; tot++;
    if (full ~ /^IMAD\.WIDE/) wide++;
    if (full ~ /^IMAD\.HI/) hi++;
    if (full ~ /^IMAD\.MOV/ || full ~ /^IMAD\.IADD/ || full ~ /^IMAD\.SHL/ ) imadmisc++;
    if (full ~ /^IMAD/) imadall++;
    if (full ~ /^(LDL|STL)/) local++;
  }
  END {
    printf "kernel: %s\n", name;
    for (k in cnt) printf "%7d %s\n", cnt[k], k | "sort -rn | head -'"$top"'";
    close("sort -rn | head -'"$top"'");
    printf "total %d  IMAD.WIDE %d  IMAD.HI %d  IMAD(all forms) %d  IMAD.MOV/IADD/SHL %d  local LD/ST %d\n", tot, wide, hi, imadall, imadmisc, local;
  }'
