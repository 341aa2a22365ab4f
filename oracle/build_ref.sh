#!/bin/bash
# Compile the reference's own CPU partition functors + murmur3 header, UNMODIFIED
# and read in place from /root/reference, against oracle/tf_stub (a ~70-line
# stand-in for the 4 TensorFlow headers they include).  Output: oracle/_ref/ only
# (git-ignored; travels to the GPU box).  No reference source is copied.
set -e
cd "$(dirname "$0")"
REF=${HB_REFERENCE_ROOT:-/root/reference}
if [ ! -d "$REF/hybridbackend" ]; then
  echo "build_ref.sh: $REF not present; keeping any prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p _ref
P=$REF/hybridbackend/tensorflow/distribute/partition
g++ -O2 -std=c++14 -fPIC -shared -w -DHYBRIDBACKEND_TENSORFLOW=1 \
    -Itf_stub -I"$REF" \
    "$P/partition_by_modulo_functors.cc" "$P/partition_by_dual_modulo_functors.cc" \
    ref_driver.cc -o _ref/libhbref.so
echo "built oracle/_ref/libhbref.so"
