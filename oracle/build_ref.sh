#!/bin/bash
# oracle/build_ref.sh -- install the UNMODIFIED reference (pure Python, simonsobs/hmvec) from /root/reference into
# oracle/_ref/ so that `bench.py --impl reference` and the cpu_baseline leg can time the real thing on the GPU box's
# host cores (oracle/_ref is git-ignored, not gpurun-ignored: it travels with the snapshot like the built .so files).
# Run in the build container only; /root/reference is read-only, so pip builds from a copy under /tmp.
# The reference imports `camb` at module level; tests/golden/camb_standin (flat-LCDM closed forms, accuracy='low' never
# calls CAMB for P(k)) is put on sys.path by the callers, exactly as tests/golden/make_golden.py does.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=${1:-/root/reference}
[ -d "$SRC/hmvec" ] || { echo "no reference at $SRC: nothing to do"; exit 0; }
TMP=$(mktemp -d /tmp/hmvec_ref.XXXXXX)
cp -r "$SRC"/. "$TMP"/
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP"
rm -rf "$TMP"
ls "$HERE/_ref"
