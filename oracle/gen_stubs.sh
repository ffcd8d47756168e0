#!/bin/bash
# TEST INFRASTRUCTURE ONLY. usage: gen_stubs.sh <lib with unresolved symbols> <out stub .so>
# Emits an aborting definition for every undefined ngp:: / tcnn:: / filesystem:: symbol of <lib> (see oracle/Makefile).
set -e
lib=$1; out=$2
tmp=$(mktemp --suffix=.c)
echo '#include <stdlib.h>' > $tmp
i=0
for sym in $(nm -D -u "$lib" | awk '{print $2}' | grep -E '^_Z' | grep -E 'ngp|tcnn|filesystem|nlohmann' | sort -u); do
	echo "void ref_stub_$i(void) __asm__(\"$sym\"); void ref_stub_$i(void) { abort(); }" >> $tmp
	i=$((i+1))
done
${ORACLE_CC:-/usr/bin/gcc} -shared -fPIC -o "$out" $tmp
rm -f $tmp
echo "gen_stubs: $i stubs -> $out"
