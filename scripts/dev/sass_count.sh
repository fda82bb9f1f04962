#!/bin/bash
# FP64 / total instruction counts per kernel from the SASS of libmahakala_b200.so:  scripts/dev/sass_count.sh <regex>
SO=${2:-/root/repo/mahakala_b200/libmahakala_b200.so}
cuobjdump -sass "$SO" | awk -v pat="$1" '
/Function :/ { if (name != "") report(); name=$3; dfma=dmul=dadd=mufu=tot=dsetp=stl=ldl=0; keep = (name ~ pat) }
keep && /^ +\/\*[0-9a-f]+\*\// { tot++; if ($0 ~ /DFMA/) dfma++; if ($0 ~ /DMUL/) dmul++; if ($0 ~ /DADD/) dadd++; if ($0 ~ /MUFU/) mufu++; if ($0 ~ /DSETP/) dsetp++; if ($0 ~ /STL/) stl++; if ($0 ~ /LDL/) ldl++ }
function report() { if (keep) printf "%-90s total %5d  DFMA %4d DMUL %4d DADD %4d DSETP %3d MUFU %3d STL %d LDL %d\n", substr(name,1,90), tot, dfma, dmul, dadd, dsetp, mufu, stl, ldl }
END { report() }'
