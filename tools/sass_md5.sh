#!/bin/bash
# md5 of the SASS listing of a library (default: the in-tree libstratusgi.so): two builds whose device code is the same instruction
# for instruction print the same sum whatever their host code or build time.  Used at the end of round 2 to show that the library
# the driver runs carries the device code that was validated and measured on the B200s (profiles/r02_review_items.md).
lib=${1:-$(dirname "$0")/../stratusgfx_b200/csrc/libstratusgi.so}
cuobjdump -sass "$lib" | grep -v '^\s*$' | md5sum | cut -d' ' -f1
