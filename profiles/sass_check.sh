#!/bin/sh
# SASS evidence that the hot kernels use the Blackwell paths (B200_PROFILING.md): tcgen05 MMA (UTCHMMA, .2CTA for the SM-pair
# GEMM), TMEM loads / stores (LDTM / STTM), TMA (UTMALDG / UTMASTG, .MULTICAST), tcgen05 commit barriers (UTCBAR).
# Runs on the CPU: cuobjdump reads the in-tree library.   sh profiles/sass_check.sh > profiles/r1_sass_mnemonics.txt
LIB="$(dirname "$0")/../videoprism-mlx_b200/libvideoprism_b200.so"
echo "# cuobjdump -sass $(basename "$LIB"): instruction counts over all kernels"
cuobjdump -sass "$LIB" | grep -o "UTCHMMA[.A-Z0-9_]*\|UTMALDG[.A-Z0-9_]*\|UTMASTG[.A-Z0-9_]*\|LDTM[.A-Z0-9_x]*\|STTM[.A-Z0-9_x]*\|UTCBAR[.A-Z0-9_]*\|UTCATOMSWS[.A-Z0-9_]*\|ELECT[.A-Z0-9_]*\|HMMA[.A-Z0-9_]*\|MUFU[.A-Z0-9_]*\|FFMA2\|FMUL2\|FADD2\|SYNCS[.A-Z0-9_]*" | sort | uniq -c | sort -rn
echo
echo "# per kernel family (template instantiations merged): instructions per instantiation"
cuobjdump -sass "$LIB" | awk '/Function :/ {name=$3} /UTCHMMA/ {m[name]++} /UTMALDG/ {l[name]++} /UTMASTG/ {s[name]++} /LDTM/ {t[name]++} /STTM/ {w[name]++} /[ \t]HMMA\./ {h[name]++} /Function :/ {seen[name]=1} END {for (k in seen) printf "%s UTCHMMA=%d UTMALDG=%d UTMASTG=%d LDTM=%d STTM=%d HMMA=%d\n", k, m[k], l[k], s[k], t[k], w[k], h[k]}' \
  | c++filt | sed -e 's/vp::(anonymous namespace):://' -e 's/<[^ ]*>//' -e 's/(.*) / /' | grep -v "UTCHMMA=0 UTMALDG=0 UTMASTG=0 LDTM=0 STTM=0 HMMA=0" | sort | uniq -c | sort -k2
