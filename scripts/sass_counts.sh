#!/bin/bash
# SASS evidence that the shipped library uses the Blackwell async machinery (no GPU needed).
# Usage: bash scripts/sass_counts.sh > profiles/r02_sass_counts.md
cd "$(dirname "$0")/.." || exit 1
SO=tapnet_b200/libtapir_b200.so
TMP=$(mktemp)
cuobjdump -sass "$SO" > "$TMP"
echo "# SASS mnemonic counts of tapnet_b200/libtapir_b200.so"
echo
echo "\`cuobjdump -sass $SO | grep -c <mnemonic>\` at commit $(git rev-parse --short HEAD) ($(date -u +%Y-%m-%d)); arch: $(cuobjdump -lelf "$SO" | head -1)"
echo
echo "| mnemonic | count | meaning |"
echo "|---|---|---|"
row() { printf '| `%s` | %s | %s |\n' "$1" "$(grep -c -- "$1" "$TMP")" "$2"; }
row 'UTCHMMA' 'tcgen05.mma kind::f16 (bf16 operands, fp32 accumulate in TMEM), all forms'
row 'UTCHMMA.2CTA' 'cta_group::2 MMAs (256-row pair tiles)'
row 'UTMALDG' 'TMA tensor loads (cp.async.bulk.tensor), all forms'
row 'UTMALDG.5D' '5-D TMA loads = implicit-GEMM convolution taps (zero padding by OOB fill)'
row '2CTA' 'instructions in cta_group::2 form'
row 'LDTM' 'tcgen05.ld (TMEM -> registers, epilogues)'
row 'UTCBAR' 'tcgen05.commit (MMA completion -> mbarrier)'
row 'SYNCS' 'mbarrier operations'
row 'HMMA.16816' 'legacy mma.sync (stage-A head hid3, split bf16)'
row 'FFMA2' 'packed fp32 FMA (mixer depthwise kernel)'
row 'FMUL2' 'packed fp32 multiply'
row 'FADD2' 'packed fp32 add'
row 'MUFU.EX2' 'ex2.approx (tanh-GELU)'
row 'STG.E.ENL2.256' '256-bit global stores (GEMM epilogue: one full sector per request)'
echo
echo "Per kernel (UTCHMMA / UTMALDG / LDTM / HMMA / FFMA2):"
echo
echo '```'
awk '/Function :/ {name=$3} /UTCHMMA/ {a[name]++} /UTMALDG/ {b[name]++} /LDTM/ {c[name]++} /HMMA\.16816/ {d[name]++} /FFMA2/ {e[name]++} END {for (n in a) print n, a[n]+0, b[n]+0, c[n]+0, d[n]+0, e[n]+0; for (n in d) if (!(n in a)) print n, 0, 0, 0, d[n], e[n]+0; for (n in e) if (!(n in a) && !(n in d)) print n, 0, 0, 0, 0, e[n]}' "$TMP" | c++filt | sed 's/tapir::(anonymous namespace):://' | sort | cut -c1-150
echo '```'
rm -f "$TMP"
