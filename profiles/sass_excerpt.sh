#!/bin/bash
# SASS evidence of the tcgen05 convolution kernel, from the in-tree libeig.so (cuobjdump -sass; no GPU needed).
# usage: profiles/sass_excerpt.sh > profiles/r2/sass_conv3x3_tc_kernel.txt
set -euo pipefail
cd "$(dirname "$0")/.."
SO=evolutionary_illusion_generator_b200/libeig.so
TMP=$(mktemp)
cuobjdump -sass -fun '_ZN3eig17conv3x3_tc_kernelILb0ELb1EEEv14CUtensorMap_stS1_S1_NS_8TcParamsE' $SO > $TMP 2>/dev/null || cuobjdump -sass $SO > $TMP
echo "# cuobjdump -sass of eig::conv3x3_tc_kernel<false, true> (the folded-taps instantiation; <false, false> differs only by the absence of the K-block skip, tap masks and Z add) in $SO (sm_100a), $(date -u +%Y-%m-%d), source state $(git rev-parse --short HEAD)"
echo "# instruction counts by mnemonic (tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG, tcgen05.commit = UTCBAR, TMEM alloc = UTCATOM*)"
grep -oE "\b(UTCHMMA[.A-Z0-9_]*|UTCQMMA[.A-Z0-9_]*|UTMALDG[.A-Z0-9_]*|UTMAPF[.A-Z0-9_]*|LDTM[.A-Z0-9_x]*|STTM[.A-Z0-9_x]*|UTCBAR[.A-Z0-9_]*|UTCATOMSWS[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*|UCGABAR[.A-Z0-9_]*|ELECT[.A-Z0-9_]*|HMMA[.A-Z0-9_]*)" $TMP | sort | uniq -c | sort -rn
echo
echo "# first occurrences in context"
for pat in UTCHMMA UTMALDG LDTM UTCBAR UTCATOMSWS; do
  echo "## $pat"
  grep -m 6 -E "$pat" $TMP | sed 's/^ *//' | cut -c1-200
done
rm -f $TMP
