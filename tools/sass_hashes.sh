#!/bin/bash
# md5 of the SASS instruction stream (addresses and encodings stripped) of every translation unit's device code: two
# builds with the same hashes run the same kernels.  Used to show that host-side edits after the last GPU verification
# left the device code untouched (profiles/r1_verified_sass_hashes.txt).
cd "$(dirname "$0")/.." || exit 1
for o in aimnetcentral_b200/_lib/*.o; do
  s=$(basename "$o" .o)
  h=$(cuobjdump -sass "$o" | grep -v '^\s*/\* 0x' | grep -E '^\s+/\*[0-9a-f]{4}\*/|Function' | md5sum | cut -c1-16)
  echo "$s $h"
done
