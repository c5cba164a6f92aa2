#!/bin/bash
# SASS instruction count (and local-memory traffic) of every kernel in an object file:  tools/sass_count.sh build/launch_exact.o [filter]
set -e
obj=$(realpath "${1:-build/launch_exact.o}"); filt="${2:-}"
tmp=$(mktemp -d); cd "$tmp"
cuobjdump -xelf all "$obj" >/dev/null
nvdisasm -c *.cubin 2>/dev/null | awk -v f="$filt" '
  /\.section\t\.text\./ { if (name != "") print n, l, name; name=$2; n=0; l=0; next }
  /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]/ { n++; if ($0 ~ /STL|LDL/) l++ }
  END { if (name != "") print n, l, name }' | grep -- "$filt" | sed 's/\.text\.//; s/,"ax".*//'
rm -rf "$tmp"
