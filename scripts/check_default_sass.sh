#!/bin/bash
# Is the device code of the DEFAULT library still exactly what the round-1 GPU runs validated?
# (profiles/r01_validated_sass.sha256 = sha256 of the address-stripped `cuobjdump -sass` text of libvirtex_b200.so at
#  the commit whose tests / bench ran on the B200.)  Prints MATCH or DIFFERENT.
cd "$(dirname "$0")/.."
h=$(cuobjdump -sass virtex_b200/libvirtex_b200.so 2>/dev/null | grep -v "^\s*/\* 0x" | sed 's#/\*[0-9a-f]*\*/##' | sha256sum | cut -d' ' -f1)
if [ "$h" = "$(cat profiles/r01_validated_sass.sha256)" ]; then echo MATCH; else echo "DIFFERENT ($h)"; exit 1; fi
