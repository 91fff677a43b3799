#!/bin/bash
# usage: tools/sass_of.sh <mangled-name-substring> [lib]  -> prints the SASS of the first matching kernel, addresses + opcodes only
LIB=${2:-daachorse_b200/libdaachorse_b200.so}
cuobjdump -sass "$LIB" 2>/dev/null | awk -v pat="$1" '
  /Function :/ { on = (index($0, pat) > 0 && !done); if (on) { print; done = 1 } else on = 0 }
  on && /^\s+\/\*[0-9a-f]{4}\*\// { sub(/\/\* 0x[0-9a-f]+ \*\//, ""); $1 = $1; print }'
