#!/bin/bash
# SASS mnemonic counts of the in-tree library (what proves tcgen05 / TMEM / TMA): profiles/r02_sass_mnemonics.txt
cd "$(dirname "$0")/.."
{
  echo "# cuobjdump -sass virtex_b200/libvirtex_b200.so at commit $(git rev-parse --short HEAD) (sha256 $(sha256sum virtex_b200/libvirtex_b200.so | cut -c1-16))"
  echo "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store, UTCBAR = tcgen05.commit,"
  echo "# HMMA / LDSM = mma.sync + ldmatrix of the attention cores, LDGSTS = cp.async staging, ACQBULK / PREEXIT = griddepcontrol (PDL)"
  echo "# .2CTA forms = cta_group::2 (CTA pairs): UTCHMMA.2CTA, UTMALDG.xD.2CTA, UTCBAR.2CTA.MULTICAST, UTCATOMSWS.2CTA (TMEM alloc); UTMAPF = TMA L2 prefetch"
  cuobjdump -sass virtex_b200/libvirtex_b200.so | grep -oE "\b(UTCHMMA(\.2CTA)?|UTCQMMA|LDTM|STTM|UTMALDG(\.[0-9]D)?(\.2CTA)?|UTMASTG(\.[0-9]D)?|UTMAPF(\.L2)?(\.[0-9]D)?|UTCBAR(\.2CTA)?(\.MULTICAST)?|UTCATOMSWS(\.2CTA)?|UCGABAR_ARV|HMMA\.[0-9]+|LDSM|LDGSTS|ACQBULK|PREEXIT|UBLKCP|FFMA2|FADD2|REDG|RED)\b" | sort | uniq -c
} > profiles/r02_sass_mnemonics.txt
cat profiles/r02_sass_mnemonics.txt
