#!/bin/bash
# usage: tools/sass_fn.sh <object> <mangled-name-substring>  -> plain SASS listing of that function
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function :/ {on = index($0, pat) > 0} on' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
