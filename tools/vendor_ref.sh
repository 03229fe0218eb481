#!/usr/bin/env bash
# Recipe: make the UNMODIFIED reference modules of the hot path available as the timed CPU / eager-GPU baseline arm.
# The reference is pure Python, so its "build" is a file copy: the files are taken from where they lie under
# /root/reference into the git-ignored oracle/_ref/ (outputs only -- nothing under oracle/_ref is tracked; it
# travels to the GPU box with the gpurun snapshot like the built .so does).  Nothing in the product imports it;
# only oracle/ref_harness.py (bench.py --impl reference, cpu_baseline) does.
set -euo pipefail
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")/.." && pwd)
OUT=$HERE/oracle/_ref
[ -d "$REF/models/search/darts" ] || { echo "vendor_ref: $REF not present (GPU box?) -- keeping whatever oracle/_ref holds"; exit 0; }
rm -rf "$OUT"
mkdir -p "$OUT/models/search/darts" "$OUT/models/auxiliary"
cp "$REF/models/__init__.py" "$OUT/models/"
cp "$REF/models/search/__init__.py" "$OUT/models/search/"
for f in __init__ genotypes operations node_operations node_search model_search architect model node utils; do
  cp "$REF/models/search/darts/$f.py" "$OUT/models/search/darts/"
done
cp "$REF/models/auxiliary/__init__.py" "$REF/models/auxiliary/scheduler.py" "$OUT/models/auxiliary/"
( cd "$REF" && sha256sum models/search/darts/{genotypes,operations,node_operations,node_search,model_search,architect,model,node,utils}.py \
    models/auxiliary/scheduler.py ) > "$OUT/SHA256SUMS"
echo "vendor_ref: $(wc -l < "$OUT/SHA256SUMS") reference files -> $OUT"
