#!/usr/bin/env bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell tensor / bulk-copy paths
# (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier,
# FFMA2 = fma.rn.f32x2, UCGABAR = barrier.cluster).
#   bash tools/sass_markers.sh > profiles/sass_markers.txt
set -euo pipefail
LIB="${1:-pmarlo_b200/libpmb200.so}"
echo "# cuobjdump -sass $LIB : marker counts per kernel (kernels without any marker omitted)"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { name=$3; next }
  { n=0 }
  /UTCHMMA/ { c[name,"UTCHMMA"]++; seen[name]=1 }
  /UTCHMMA\.2CTA/ { c[name,"UTCHMMA.2CTA"]++ }
  /LDTM/ { c[name,"LDTM"]++; seen[name]=1 }
  /UTCBAR/ { c[name,"UTCBAR"]++; seen[name]=1 }
  /UBLKCP/ { c[name,"UBLKCP"]++; seen[name]=1 }
  /UTCATOMSWS|UTCALLOC/ { c[name,"TMEM_ALLOC"]++ }
  /SYNCS/ { c[name,"SYNCS"]++ }
  /FFMA2/ { c[name,"FFMA2"]++; seen[name]=1 }
  /UCGABAR/ { c[name,"UCGABAR"]++; seen[name]=1 }
  END {
    for (k in seen) printf "%s UTCHMMA=%d UTCHMMA.2CTA=%d LDTM=%d UTCBAR=%d UBLKCP=%d SYNCS=%d FFMA2=%d UCGABAR=%d\n", k, c[k,"UTCHMMA"], c[k,"UTCHMMA.2CTA"], c[k,"LDTM"], c[k,"UTCBAR"], c[k,"UBLKCP"], c[k,"SYNCS"], c[k,"FFMA2"], c[k,"UCGABAR"]
  }' | sort | while read -r name rest; do echo "$(echo "$name" | c++filt | sed 's/(.*//') $rest"; done
