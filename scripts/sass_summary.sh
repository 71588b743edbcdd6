#!/bin/bash
# Writes profiles/r2_sass_summary.md: Blackwell-native opcode counts of the built library (cuobjdump -sass).
lib=ivos-w_b200/lib/libivosw_b200.so
out=profiles/r2_sass_summary.md
{
echo "# SASS opcode summary of \`$lib\` (sm_100a), \`scripts/sass_summary.sh\`"
echo
echo "What proves a Blackwell-native kernel (B200_PROFILING.md): \`UTCHMMA\` = tcgen05.mma, \`LDTM\` = tcgen05.ld,"
echo "\`UTMALDG\` / \`UTMASTG\` / \`UBLKCP\` = TMA loads / stores / bulk copies, \`UTCBAR\` = tcgen05.commit, \`SYNCS\` = mbarrier."
echo "No \`HMMA\` (legacy mma.sync) anywhere."
echo
echo "| opcode | count |"
echo "|---|---|"
cuobjdump -sass $lib | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn \
  | grep -E " (UTC[A-Z]*|UTMA[A-Z]*|LDTM|STTM|UBLKCP|SYNCS|HMMA|HGMMA|REDG|FENCE)$" | awk '{print "| `" $2 "` | " $1 " |"}'
echo
echo "Per kernel (\`UTCHMMA\` / \`UTMALDG\` / \`UTMASTG\` / \`LDTM\`):"
echo
echo "| kernel | UTCHMMA | UTMALDG | UTMASTG | LDTM |"
echo "|---|---|---|---|---|"
cuobjdump -sass $lib | awk '/Function :/ {fn=$3} /UTCHMMA/ {a[fn]++} /UTMALDG/ {b[fn]++} /UTMASTG/ {c[fn]++} /LDTM/ {d[fn]++} END {for (f in a) print "| `" f "` | " a[f] " | " b[f]+0 " | " c[f]+0 " | " d[f]+0 " |"}' | sort
} > $out
cat $out
