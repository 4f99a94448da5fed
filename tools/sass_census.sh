#!/bin/bash
# Per-kernel SASS census of the shipped library -> stdout (profiles/r02_sass_census.txt): static instruction count and counts of the
# opcodes that prove / disprove TMA, mbarrier, tcgen05 and legacy MMA use.   tools/sass_census.sh > profiles/r02_sass_census.txt
LIB=${1:-ipcdnnwalk_b200/libsrbtrack_b200.so}
echo "# SASS census of $LIB (sm_100a cubins only, see the ELF list): cuobjdump -sass, per kernel: static instruction"
echo "# count, 1-D bulk copies (UBLKCP = cp.async.bulk), mbarrier ops (SYNCS), tensor-map TMA (UTMALDG/UTMASTG), tcgen05 MMA (UTC*MMA), legacy MMA."
echo "# No kernel on this path has a contraction larger than 6 x 40 per environment, so no tensor-core instruction is expected; the bulk copies are"
echo "# in the FK tile kernels and in fullbody_extract (contiguous row tiles: 1-D bulk copies, no tensor map needed)."
cuobjdump -lelf $LIB | sed 's/^/# /'
cuobjdump -sass $LIB | awk '
/Function :/ { if (name != "") print name "\tinstr=" n "\tUBLKCP=" b "\tSYNCS=" s "\tUTMALDG/STG=" t "\tUTC*MMA=" u "\tMMA=" m; name=$3; n=0; b=0; s=0; t=0; u=0; m=0; next }
/^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/ { n++; if ($0 ~ /UBLKCP/) b++; if ($0 ~ /SYNCS/) s++; if ($0 ~ /UTMALDG|UTMASTG/) t++; if ($0 ~ /UTC[A-Z]*MMA/) u++; if ($0 ~ / [HIDF]MMA|QMMA|BMMA/) m++ }
END { if (name != "") print name "\tinstr=" n "\tUBLKCP=" b "\tSYNCS=" s "\tUTMALDG/STG=" t "\tUTC*MMA=" u "\tMMA=" m }' | while IFS=$'\t' read -r nm rest; do echo -e "$(echo "$nm" | c++filt)\t$rest"; done
