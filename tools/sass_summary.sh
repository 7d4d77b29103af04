#!/bin/bash
# SASS evidence of the production library (CPU only): tcgen05 / TMEM / TMA mnemonic counts, per kernel and in total.
# usage: bash tools/sass_summary.sh > profiles/r02_sass_summary.txt
LIB=voxsrc2020_speaker_verification_b200/libsvx.so
cuobjdump -sass $LIB > /tmp/svx_sass.txt
echo "# SASS evidence, round 2 final build (cuobjdump -sass $LIB, production build; tools/sass_summary.sh)"
echo
echo "## tcgen05 / TMEM / TMA mnemonics in the whole library (count)"
grep -oE "\b(UTCHMMA(\.2CTA)?|UTCBAR(\.2CTA)?(\.MULTICAST)?|UTMALDG\.[0-9]D(\.2CTA)?|UTMASTG\.[0-9]D|LDTM\.x[0-9]+|STTM\.x[0-9]+|UBLKCP\.[A-Z.]+|UTCCP[A-Z0-9.]*|UBLKPF[A-Z0-9.]*|UTMAPF[A-Z0-9.]*|HMMA[A-Z0-9.]*|WGMMA[A-Z0-9.]*)" /tmp/svx_sass.txt | sort | uniq -c | sort -rn
echo
echo "(no HMMA / mma.sync and no WGMMA: every tensor-core instruction is a tcgen05 UTCHMMA; .2CTA = cta_group::2 forms of conv_pair_kernel)"
echo
echo "## per-kernel UTCHMMA / LDTM / STTM / UTMALDG / UTMASTG counts"
awk '/Function :/{name=$3; for(i=4;i<=NF;i++) name=name" "$i} /UTCHMMA/{a[name]++} /LDTM/{b[name]++} /STTM/{c[name]++} /UTMALDG/{d[name]++} /UTMASTG/{e[name]++} END{for (n in a) printf "%5d UTCHMMA %4d LDTM %4d STTM %4d UTMALDG %4d UTMASTG  %s\n", a[n], b[n], c[n], d[n], e[n], n}' /tmp/svx_sass.txt | while read -r line; do f=$(echo "$line" | awk '{print $NF}'); echo "$(echo "$line" | sed "s/ [^ ]*$//")  $(echo $f | c++filt | cut -c1-150)"; done | sort -k11
echo
echo "## conv_pair_kernel<__half, 1, 1>: the cta_group::2 main-loop issue sequence (first 16 UTCHMMA / UTCBAR instructions)"
awk '/Function :/{on = ($0 ~ /conv_pair_kernelI6__halfLi1ELi1E/)} on && /UTCHMMA|UTCBAR/' /tmp/svx_sass.txt | sed 's/\/\*[0-9a-f]*\*\/\s*$//' | cut -c20-140 | head -16
