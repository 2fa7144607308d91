#!/bin/bash
# usage: sass_stats.sh <obj> <mangled-substring>   -> instruction histogram of the first matching kernel
f=$(cuobjdump -sass $1 | grep "Function :" | grep "$2" | head -1 | awk '{print $3}')
echo "$f"
cuobjdump -sass -fun "$f" $1 > /tmp/k.sass
grep -cE "^\s+/\*[0-9a-f]{4,5}\*/" /tmp/k.sass
grep -E "^\s+/\*[0-9a-f]{4,5}\*/" /tmp/k.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+ )?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?.*/\2/' | sort | uniq -c | sort -rn | head -${3:-24}
