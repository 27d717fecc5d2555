#!/bin/bash
# usage: tools/sass_of.sh <mangled-substring>   -> prints the SASS of the first matching kernel (opcodes only)
LIB=/root/repo/reinmav-gym_b200/gym_reinmav_b200/lib/librmv.so
cuobjdump -sass "$LIB" | awk -v pat="$1" '
  /Function :/ { if (on) exit; on = (index($0, pat) > 0); if (on) print; next }
  on && /\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\// && !/^[ \t]*\/\* 0x/ { line=$0; sub(/^[ \t]*\/\*/, "", line); sub(/\*\/[ \t]+/, " ", line); sub(/[ \t]*\/\*.*$/, "", line); print line }'
