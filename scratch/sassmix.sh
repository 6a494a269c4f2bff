#!/bin/bash
# usage: sassmix.sh <obj> <kernel-substring>  -> static opcode histogram of the matching kernel(s)
cuobjdump -sass "$1" | awk -v pat="$2" '
/Function :/ { on = index($0, pat) > 0; if (on) print "== " $3 }
on && /^\s+\/\*[0-9a-f]+\*\// { op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); sub(/;/,"",op); c[op]++ ; n++}
END { for (k in c) printf "%6d %s\n", c[k], k | "sort -rn"; print n " total" }'
