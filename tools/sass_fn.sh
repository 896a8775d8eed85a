#!/bin/bash
# usage: tools/sass_fn.sh <object> <regex on demangled-ish mangled name>   -> SASS of the first matching function, code lines only
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function : /{p=($0 ~ pat)} p' | grep -v "^\s*/\* 0x" | grep "^\s*/\*[0-9a-f]\{4\}\*/" | cut -c1-110
