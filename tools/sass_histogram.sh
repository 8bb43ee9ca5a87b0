#!/usr/bin/env bash
# SASS opcode evidence for the in-tree library: which tensor-core / TMA / TMEM instructions the product .so really contains.
# Usage: bash tools/sass_histogram.sh > profiles/rNN_sass_histogram.md   (runs in the build container; no GPU needed)
SO=visualcloze_b200/libvcb200.so
echo "# SASS opcode histogram of $SO ($(date -u +%Y-%m-%dT%H:%MZ), $(git rev-parse --short HEAD 2>/dev/null))"
echo
echo '`cuobjdump -sass visualcloze_b200/libvcb200.so | grep -c <mnemonic>` (sm_100a cubin; B200_PROFILING.md lists what each one proves)'
echo
echo "| mnemonic | count | meaning |"
echo "|---|---|---|"
SASS=$(mktemp)
cuobjdump -sass "$SO" > "$SASS"
row() { printf '| `%s` | %s | %s |\n' "${3:-$1}" "$(grep -cE -- "(^|[^A-Z.])$1" "$SASS")" "$2"; }
row "UTCHMMA" "tcgen05.mma (all kinds, incl. the .2CTA forms)"
row "UTCHMMA\.2CTA" "tcgen05.mma cta_group::2 (CTA-pair tiles)" "UTCHMMA.2CTA"
row "UTCQMMA" "tcgen05.mma kind::f8f6f4 (e4m3 operands of the opt-in fp8 projections)"
row "LDTM" "tcgen05.ld (TMEM -> registers)"
row "STTM" "tcgen05.st (registers -> TMEM: P of the attention kernel, stream-K folds)"
row "UTMALDG" "cp.async.bulk.tensor loads (TMA)"
row "UTMASTG" "cp.async.bulk.tensor stores (TMA tile stores of the sequence-parallel epilogues)"
row "UTCBAR" "tcgen05.commit -> mbarrier"
row "SYNCS" "mbarrier operations"
row "HMMA" "legacy mma.sync tensor instructions (must be 0)"
row "MUFU.EX2" "ex2.approx (softmax / GELU)"
row "MUFU.TANH" "tanh.approx (GELU of the fp8 instantiations)"
row "F2FP" "packed float conversions (bf16 / e4m3 packs)"
echo
echo "Kernels in the cubin:"
echo
cuobjdump -elf "$SO" 2>/dev/null | grep -o "\.text\._ZN3vcb[A-Za-z0-9_]*" | sed 's/\.text\.//' | c++filt | sed 's/(.*//' | sort -u | sed 's/^/    /'
rm -f "$SASS"
