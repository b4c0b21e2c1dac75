#!/bin/bash
# usage: sass_ops.sh <lib.so> <function-name-substring>   -> one "addr op operands" line per instruction of the first matching function
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function : /{f = index($0, pat) > 0 && !done; if (f) seen = 1; else if (seen) done = 1} f' \
  | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
